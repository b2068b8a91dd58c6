(timeout 600 python -m pytest tests/test_ba_gpu.py -m gpu -x -q -k "global") 2>&1 | tail -2
timeout 300 python profiles/microbench/c4time.py 2>&1 | tail -2
PGBA_ND_COOP=0 timeout 300 python profiles/microbench/c4time.py 2>&1 | tail -1
