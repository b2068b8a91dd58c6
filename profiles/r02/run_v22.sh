mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c4_v22.csv python profiles/ncu_target_c4.py 1 > gpurun_out/ncu_c4_v22.log 2>&1
python profiles/launch_summary.py gpurun_out/launches_c4_v22.csv
