mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_v14.log 2>&1
tail -4 gpurun_out/pytest_gpu_v14.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_v14.json 2> gpurun_out/bench_v14.err; echo bench rc=$?
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_v14.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_c2_v14.csv python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/ncu_bench_v14.log 2>&1
python profiles/launch_summary.py gpurun_out/launches_bench_c2_v14.csv
