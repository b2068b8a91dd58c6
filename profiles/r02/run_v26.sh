python profiles/ab_c5.py c2
PGBA_LIN_ONE_INSTANCE=0 python profiles/ab_c5.py c2
python profiles/ab_c5.py c2
PGBA_LIN_ONE_INSTANCE=0 python profiles/ab_c5.py c2
(timeout 900 python -m pytest tests/test_ba_gpu.py -m gpu -x -q -k "not global") 2>&1 | tail -3
