set -x
python profiles/prof_cg.py 512 4 > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:spmv_zmarch -s 2 -c 1 -o gpurun_out/prof_zmarch_g4_r02d -f python profiles/prof_cg.py 512 4 > gpurun_out/ncu_zm4.log 2>&1
AB_OPTS=spmv_zmarch_g=2 ncu --set full --clock-control none --import-source on -k regex:spmv_zmarch -s 2 -c 1 -o gpurun_out/prof_zmarch_g2_r02d -f python profiles/prof_cg.py 512 4 > gpurun_out/ncu_zm2.log 2>&1
tail -3 gpurun_out/ncu_zm2.log
