set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_r02w.log
tail -3 gpurun_out/pytest_r02w.log
Q="--no-legs --no-cpu --no-general --no-parity"
timeout 600 python bench.py --steps 2 --warmup 3 $Q > "gpurun_out/bench_r02w_512.log" 2>&1
for g in 400 300 200; do
timeout 600 python bench.py --steps 2 --warmup 3 $Q --grid $g > "gpurun_out/bench_r02w_grid$g.log" 2>&1
AB_OPTS=spmv_zmarch=0 timeout 600 python bench.py --steps 2 --warmup 3 $Q --grid $g > "gpurun_out/bench_r02w_grid${g}_nozm.log" 2>&1
done
for f in gpurun_out/bench_r02w*.log; do python - <<PY
import json
try:
    d=json.loads([l for l in open('$f') if l.startswith('{')][-1])
    print('$f', round(d['value'],1), round(d['kernels']['spmv_dot']['ms'],4), round(d['kernels']['spmv_dot']['frac'],3), d['roofline']['kernel'][:28], d['relres_after_step'])
except Exception as e: print('$f', 'ERR', e)
PY
done
