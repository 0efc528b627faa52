set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_r02p.log
tail -3 gpurun_out/pytest_r02p.log
for o in "pdl=1" "pdl=0"; do
echo "== $o"; AB_OPTS=$o python profiles/scripts/slab_floor.py 8 1 2>&1 | grep slab | tee -a gpurun_out/slab_floor_r02p.log
done
Q="--no-legs --no-cpu --no-general --no-parity"
timeout 600 python bench.py --steps 3 --warmup 3 $Q > gpurun_out/bench_r02p.log 2>&1
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_r02p.log') if l.startswith('{')][-1])
print(round(d['value'],1), round(d['e2e']['value'],1), {k: round(v['ms'],4) for k,v in d['kernels'].items()})
PY
