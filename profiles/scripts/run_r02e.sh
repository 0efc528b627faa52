set -x
Q="--no-legs --no-cpu --no-general --no-parity"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "lossless" 2>&1 | tail -5 > gpurun_out/pytest_r02e.log
tail -3 gpurun_out/pytest_r02e.log
for o in "spmv_zmarch_chunk=4096" "spmv_zmarch_chunk=1024" "spmv_zmarch_chunk=16384" "spmv_zmarch_chunk=65536" \
         "spmv_zmarch_g=2" "spmv_zmarch_g=2,spmv_zmarch_nb=4" "spmv_zmarch_g=2,spmv_zmarch_chunk=1024" "spmv_zmarch_g=2,spmv_zmarch_chunk=65536" \
         "spmv_zmarch_g=1" "spmv_zmarch_g=1,spmv_zmarch_chunk=1024"; do
AB_OPTS=$o timeout 600 python bench.py --steps 2 --warmup 3 $Q > "gpurun_out/bench_r02e_$o.log" 2>&1
done
for f in gpurun_out/bench_r02e*.log; do python - <<PY
import json
try:
    d=json.loads([l for l in open('$f') if l.startswith('{')][-1])
    print('$f', round(d['value'],1), round(d['kernels']['spmv_dot']['ms'],4), d['relres_after_step'])
except Exception as e: print('$f', 'ERR', e)
PY
done
