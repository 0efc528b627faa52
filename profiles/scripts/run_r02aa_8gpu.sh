set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
Q="--no-legs --no-cpu --no-general"
timeout 600 $TR tests/dist_check.py 64 > gpurun_out/dist8_check_r02aa.log 2>&1
echo "dist_check rc=$?"; grep -c " OK" gpurun_out/dist8_check_r02aa.log; grep "FAIL" gpurun_out/dist8_check_r02aa.log | head -5; tail -1 gpurun_out/dist8_check_r02aa.log
timeout 600 $TR bench.py --gpus 8 --steps 4 --warmup 3 $Q --no-parity > gpurun_out/bench_n8_r02aa.log 2> gpurun_out/bench_n8_r02aa.err
tail -2 gpurun_out/bench_n8_r02aa.err
AB_OPTS=dist_epilogue_reduce=0 timeout 600 $TR bench.py --gpus 8 --steps 4 --warmup 3 $Q --no-parity > "gpurun_out/bench_n8_r02aa_dist_epilogue_reduce=0.log" 2>&1
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534"
timeout 600 $TR4 bench.py --gpus 4 --steps 4 --warmup 3 $Q --no-parity > gpurun_out/bench_n4_r02aa.log 2>&1
for f in gpurun_out/bench_n8_r02aa*.log gpurun_out/bench_n4_r02aa.log; do python - <<PY
import json
try:
    d=json.loads([l for l in open('$f') if l.startswith('{')][-1])
    print('$f', round(d['value'],1), round(d['e2e']['value'],1), d.get('gpu_launches'), (d.get('parity') or {}).get('ok'))
except Exception as e: print('$f', 'ERR', e)
PY
done
