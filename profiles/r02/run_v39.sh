mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c4_v39.csv python profiles/ncu_target_c4.py 1 > gpurun_out/ncu_c4_v39.log 2>&1
python profiles/launch_summary.py gpurun_out/launches_c4_v39.csv
ncu --set full --clock-control none --import-source on -k regex:'nd_border|nd_backsolve|nd_gather' -c 4 -o gpurun_out/ncu_c4_border_v39 python profiles/ncu_target_c4.py 1 > gpurun_out/ncu_c4_border_v39.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:'nd_syrk|nd_trsm' --launch-skip 6 -c 4 -o gpurun_out/ncu_c4_seg_v39 python profiles/ncu_target_c4.py 1 > gpurun_out/ncu_c4_seg_v39.log 2>&1; echo rc=$?
PGBA_LIB=cdv-slam_b200/lib/libpgba_ndts.so PGBA_ND_COOP=0 python profiles/nd_timeline.py > gpurun_out/nd_timeline_v39.txt 2>&1
