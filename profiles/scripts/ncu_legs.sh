set -x
# C2 (256^3): per-triplet value gradient = sddmm_kernel (k = 1) + gather_slot_kernel; update_values = reduce_segments_kernel
python profiles/bench_configs.py gpurun_out/configs_plain_c2.json c2 > gpurun_out/legs_plain_c2.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'gather_slot|reduce_segments|sddmm_kernel' -c 6 -o gpurun_out/prof_c2_grad_r02n -f python profiles/bench_configs.py gpurun_out/configs_ncu_c2.json c2 > gpurun_out/ncu_c2.log 2>&1
tail -2 gpurun_out/ncu_c2.log
# C5 at 0.4 scale (20 M rows, 306 M entries): SpMM k = 32 and SDDMM
python profiles/bench_configs.py gpurun_out/configs_plain_c5.json c5 0.4 > gpurun_out/legs_plain_c5.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'spmm|sddmm' -c 4 -o gpurun_out/prof_c5_r02n -f python profiles/bench_configs.py gpurun_out/configs_ncu_c5.json c5 0.4 > gpurun_out/ncu_c5.log 2>&1
tail -2 gpurun_out/ncu_c5.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c2_r02n.csv python profiles/bench_configs.py gpurun_out/configs_ncu2_c2.json c2 > gpurun_out/ncu_c2_launches.log 2>&1
