mkdir -p gpurun_out
PGBA_BATCH_GROUPS=1 ncu --set full --clock-control none --import-source on -k regex:'linearize|solve_small|update_large|plan_direct' -c 7 -o gpurun_out/ncu_c5_v27 python profiles/ncu_target.py c5 1 > gpurun_out/ncu_c5_v27.log 2>&1; echo ncu c5 rc=$?
ncu --set full --clock-control none --import-source on -k regex:'linearize|solve_small|update_kernel|plan_cluster|plan_cells' -c 7 -o gpurun_out/ncu_c2_v27 python profiles/ncu_target.py c2 1 > gpurun_out/ncu_c2_v27.log 2>&1; echo ncu c2 rc=$?
ncu --set full --clock-control none --import-source on -k regex:'nd_syrk|nd_trsm' --launch-skip 24 -c 4 -o gpurun_out/ncu_c4_steps_v27 python profiles/ncu_target_c4.py 1 > gpurun_out/ncu_c4_steps_v27.log 2>&1; echo ncu c4 steps rc=$?
ncu --set full --clock-control none --import-source on -k regex:'nd_backsolve|nd_gather|nd_order|nd_stats' -c 5 -o gpurun_out/ncu_c4_misc_v27 python profiles/ncu_target_c4.py 1 > gpurun_out/ncu_c4_misc_v27.log 2>&1; echo ncu c4 misc rc=$?
ls -la gpurun_out/*v27.ncu-rep
