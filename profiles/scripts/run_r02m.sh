set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_r02m.log
tail -3 gpurun_out/pytest_r02m.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02m_full.log 2> gpurun_out/bench_r02m_full.err
tail -3 gpurun_out/bench_r02m_full.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_r02m_full.log') if l.startswith('{')][-1])
print(round(d['value'],1), round(d['e2e']['value'],1), d['kernels'], d['roofline']['frac'], d['cg_iteration']['frac'])
print('general', d['general_csr']['value'], d['general_csr']['roofline']['frac'])
print(json.dumps(d['configs'])[:3000])
print(d['cpu_baseline'])
PY
