set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_r02a.log
tail -5 gpurun_out/pytest_r02a.log
Q="--no-legs --no-general --no-parity --no-cpu"
python bench.py --steps 3 --warmup 3 $Q > gpurun_out/bench_r02a_rowwin.log 2> gpurun_out/bench_r02a_rowwin.err
AB_OPTS=spmv_rowwin=0 python bench.py --steps 3 --warmup 3 $Q > gpurun_out/bench_r02a_rowpat.log 2> gpurun_out/bench_r02a_rowpat.err
AB_OPTS=spmv_rowwin_ctas=4 python bench.py --steps 3 --warmup 3 $Q > gpurun_out/bench_r02a_rowwin4.log 2>&1
AB_OPTS=spmv_rowwin_ctas=3 python bench.py --steps 3 --warmup 3 $Q > gpurun_out/bench_r02a_rowwin3.log 2>&1
for f in rowwin rowpat rowwin4 rowwin3; do python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_r02a_$f.log') if l.startswith('{')][-1])
    print('$f', d['value'], d.get('kernels'))
except Exception as e: print('$f', 'ERR', e)
PY
done
timeout 600 python bench.py > gpurun_out/bench_r02a_full.log 2> gpurun_out/bench_r02a_full.err
tail -c 3000 gpurun_out/bench_r02a_full.log
tail -5 gpurun_out/bench_r02a_full.err
