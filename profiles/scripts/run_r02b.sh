set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_r02b.log
tail -5 gpurun_out/pytest_r02b.log
Q="--no-legs --no-general --no-parity --no-cpu"
python bench.py --steps 3 --warmup 3 $Q > gpurun_out/bench_r02b_march.log 2> gpurun_out/bench_r02b_march.err
AB_OPTS=spmv_rowmarch_ctas=2 python bench.py --steps 3 --warmup 3 $Q > gpurun_out/bench_r02b_march2.log 2>&1
AB_OPTS=spmv_rowmarch=0 python bench.py --steps 3 --warmup 3 $Q > gpurun_out/bench_r02b_rowwin.log 2>&1
for f in march march2 rowwin; do python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_r02b_$f.log') if l.startswith('{')][-1])
    print('$f', d['value'], d.get('kernels'))
except Exception as e: print('$f', 'ERR', e)
PY
done
tail -3 gpurun_out/bench_r02b_march.err
timeout 600 python bench.py > gpurun_out/bench_r02b_full.log 2> gpurun_out/bench_r02b_full.err
tail -c 1500 gpurun_out/bench_r02b_full.log
tail -5 gpurun_out/bench_r02b_full.err
python bench.py --steps 1 --warmup 3 $Q > gpurun_out/plain_r02b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:spmv_rowmarch -s 30 -c 2 -o gpurun_out/prof_rowmarch_r02b python bench.py --steps 1 --warmup 3 $Q > gpurun_out/ncu_r02b.log 2>&1
tail -3 gpurun_out/ncu_r02b.log
