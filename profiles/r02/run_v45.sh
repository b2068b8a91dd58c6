python profiles/ab_windows.py 8 12 16 24 32 64
PGBA_PC=128 python profiles/ab_windows.py 8 12 16
PGBA_PC=64 python profiles/ab_windows.py 4 8 12
python profiles/ab_windows.py 2 4
(timeout 900 python -m pytest tests/test_ba_gpu.py -m gpu -x -q -k "batch or window") 2>&1 | tail -2
