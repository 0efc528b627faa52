set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_r02y.log
tail -2 gpurun_out/pytest_r02y.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02y.log 2>&1; tail -1 gpurun_out/smoke_r02y.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02y_full.log 2> gpurun_out/bench_r02y_full.err
tail -2 gpurun_out/bench_r02y_full.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02y_ref.log 2>&1
tail -1 gpurun_out/bench_r02y_ref.log | cut -c1-400
python profiles/prof_cg.py 512 8 > gpurun_out/prof_plain_y.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02y_cg512.csv python profiles/prof_cg.py 512 8 > gpurun_out/ncu_launches_y.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmv_zmarch -s 2 -c 1 -o gpurun_out/prof_zmarch_r02y -f python profiles/prof_cg.py 512 4 > gpurun_out/ncu_zm_r02y.log 2>&1
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_r02y_full.log') if l.startswith('{')][-1])
print(round(d['value'],1), round(d['e2e']['value'],1), {k: round(v['ms'],4) for k,v in d['kernels'].items()}, d['roofline']['frac'], d['cg_iteration']['frac'], d['gpu_launches'])
print('general', d['general_csr']['value'], d['general_csr']['roofline']['frac'])
c=d['configs']; print('c2', c['c2']['assembly']['ms'], c['c2']['update_values']['ms'], c['c2']['grad_values']['ms'], 'c4', c['c4']['forward']['iters_per_s'], 'c5', c['c5']['spmm']['ms'], c['c5']['spmm']['frac_gather_model'], c['c5']['sddmm']['ms'], c['c5']['sddmm']['frac_gather_model'])
print(d['cpu_baseline'])
PY
