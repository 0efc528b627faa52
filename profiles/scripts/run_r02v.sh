set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_r02v.log
tail -2 gpurun_out/pytest_r02v.log
Q="--no-legs --no-cpu --no-general --no-parity"
for o in "spmv_rowmarch=0" "spmv_zmarch=0" "spmv_rowwin=0"; do
AB_OPTS=$o timeout 600 python bench.py --steps 2 --warmup 3 $Q > "gpurun_out/bench_r02v_$o.log" 2>&1
done
AB_OPTS= timeout 600 python bench.py --steps 2 --warmup 3 $Q --grid 400 > "gpurun_out/bench_r02v_grid400.log" 2>&1
for f in gpurun_out/bench_r02v*.log; do python - <<PY
import json
try:
    d=json.loads([l for l in open('$f') if l.startswith('{')][-1])
    print('$f', round(d['value'],1), round(d['kernels']['spmv_dot']['ms'],4), d['roofline']['kernel'][:40], d['relres_after_step'])
except Exception as e: print('$f', 'ERR', e)
PY
done
