set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_r02g.log
tail -3 gpurun_out/pytest_r02g.log
python profiles/prof_cg.py 512 4 > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:spmv_zmarch -s 2 -c 1 -o gpurun_out/prof_zmarch_r02g -f python profiles/prof_cg.py 512 4 > gpurun_out/ncu_zm_r02g.log 2>&1
tail -2 gpurun_out/ncu_zm_r02g.log
Q="--no-legs --no-cpu --no-general --no-parity"
for o in "spmv_zmarch_len=32" "spmv_zmarch_len=80" "spmv_zmarch_nb=4"; do
AB_OPTS=$o timeout 600 python bench.py --steps 2 --warmup 3 $Q > "gpurun_out/bench_r02g_$o.log" 2>&1
done
for f in gpurun_out/bench_r02g*.log; do python - <<PY
import json
try:
    d=json.loads([l for l in open('$f') if l.startswith('{')][-1])
    print('$f', round(d['value'],1), round(d['kernels']['spmv_dot']['ms'],4), d['relres_after_step'])
except Exception as e: print('$f', 'ERR', e)
PY
done
