set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "lossless or spmv" 2>&1 | tail -15 > gpurun_out/pytest_r02d.log
tail -5 gpurun_out/pytest_r02d.log
Q="--no-legs --no-cpu --no-general --no-parity"
timeout 600 python bench.py --steps 3 --warmup 3 $Q > gpurun_out/bench_r02d_zmarch.log 2> gpurun_out/bench_r02d_zmarch.err
tail -3 gpurun_out/bench_r02d_zmarch.err
for o in "spmv_zmarch_g=2" "spmv_zmarch_len=32" "spmv_zmarch_len=16"; do
AB_OPTS=$o timeout 600 python bench.py --steps 3 --warmup 3 $Q > "gpurun_out/bench_r02d_$o.log" 2>&1
done
for f in gpurun_out/bench_r02d*.log; do python - <<PY
import json
try:
    d=json.loads([l for l in open('$f') if l.startswith('{')][-1])
    print('$f', d['value'], d['e2e']['value'], d['kernels']['spmv_dot'], d['relres_after_step'], d['true_relres_after_step'])
except Exception as e: print('$f', 'ERR', e)
PY
done
