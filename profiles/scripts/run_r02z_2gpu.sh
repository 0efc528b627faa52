set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tests/dist_check.py 64 > gpurun_out/dist2_check_r02z.log 2>&1
echo "dist_check rc=$?"
grep -c " OK" gpurun_out/dist2_check_r02z.log; grep "FAIL" gpurun_out/dist2_check_r02z.log | head -20; tail -2 gpurun_out/dist2_check_r02z.log
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -5 > gpurun_out/pytest_dist_r02z.log
tail -3 gpurun_out/pytest_dist_r02z.log
Q="--no-legs --no-cpu --no-general"
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 $Q > gpurun_out/bench_n2_r02z.log 2> gpurun_out/bench_n2_r02z.err
tail -3 gpurun_out/bench_n2_r02z.err
for o in "dist_epilogue_reduce=0"; do
AB_OPTS=$o timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 $Q --no-parity > "gpurun_out/bench_n2_r02z_$o.log" 2>&1
done
for f in gpurun_out/bench_n2_r02z*.log; do python - <<PY
import json
try:
    d=json.loads([l for l in open('$f') if l.startswith('{')][-1])
    print('$f', round(d['value'],1), round(d['e2e']['value'],1), d.get('gpu_launches'), (d.get('parity') or {}).get('ok'))
except Exception as e: print('$f', 'ERR', e)
PY
done
