mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_v60.log 2>&1
tail -4 gpurun_out/pytest_gpu_v60.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_v60.json 2> gpurun_out/bench_v60.err) 2>&1 | grep real; echo bench rc=$?
