set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "unaligned or lossless" 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tests/dist_check.py 64 > gpurun_out/dist2_check_r02x.log 2>&1
echo "dist_check rc=$?"
grep -c " OK" gpurun_out/dist2_check_r02x.log; grep "FAIL" gpurun_out/dist2_check_r02x.log | head -20; tail -1 gpurun_out/dist2_check_r02x.log
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -2
Q="--no-legs --no-cpu --no-general"
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 $Q > gpurun_out/bench_n2_r02x.log 2> gpurun_out/bench_n2_r02x.err
tail -2 gpurun_out/bench_n2_r02x.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_n2_r02x.log') if l.startswith('{')][-1])
print(round(d['value'],1), round(d['e2e']['value'],1), d.get('gpu_launches'), (d.get('parity') or {}).get('ok'))
PY
