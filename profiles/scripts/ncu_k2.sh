set -x
python profiles/scripts/slab_floor.py 8 > gpurun_out/slab8_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:cgs_update_kernel -s 60 -c 1 -o gpurun_out/prof_k2_slab8_r02r -f python profiles/scripts/slab_floor.py 8 > gpurun_out/ncu_k2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmv_zmarch -s 60 -c 1 -o gpurun_out/prof_zm_slab8_r02r -f python profiles/scripts/slab_floor.py 8 > gpurun_out/ncu_zm8.log 2>&1
tail -2 gpurun_out/ncu_zm8.log
