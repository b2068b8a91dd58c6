mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_ba_gpu.py -m gpu -x -q -k "global") > gpurun_out/pytest_nd_v20.log 2>&1
tail -15 gpurun_out/pytest_nd_v20.log
timeout 300 python profiles/microbench/c4time.py > gpurun_out/c4time_v20.log 2>&1; tail -12 gpurun_out/c4time_v20.log
PGBA_BIG_ND=0 timeout 300 python profiles/microbench/c4time.py 2>&1 | tail -4
