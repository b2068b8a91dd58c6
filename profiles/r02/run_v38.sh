mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_ba_gpu.py -m gpu -x -q -k "global") > gpurun_out/pytest_nd_v38.log 2>&1
tail -3 gpurun_out/pytest_nd_v38.log
timeout 300 python profiles/microbench/c4time.py 2>&1 | tail -8
PGBA_ND_COOP=0 timeout 300 python profiles/microbench/c4time.py 2>&1 | tail -2
