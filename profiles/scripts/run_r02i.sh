set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "lossless" 2>&1 | tail -5 > gpurun_out/pytest_r02i.log
tail -3 gpurun_out/pytest_r02i.log
Q="--no-legs --no-cpu --no-general --no-parity"
timeout 600 python bench.py --steps 3 --warmup 3 $Q > gpurun_out/bench_r02i.log 2>&1
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_r02i.log') if l.startswith('{')][-1])
print(round(d['value'],1), d['kernels'])
PY
python profiles/scripts/slab_floor.py 8 4 2 1 2>&1 | tee gpurun_out/slab_floor_r02i.log
AB_OPTS=spmv_zmarch_g=1 python profiles/scripts/slab_floor.py 8 2>&1 | tee -a gpurun_out/slab_floor_r02i.log
AB_OPTS=spmv_zmarch_g=2 python profiles/scripts/slab_floor.py 8 2>&1 | tee -a gpurun_out/slab_floor_r02i.log
