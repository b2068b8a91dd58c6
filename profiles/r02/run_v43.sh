python profiles/ab_windows.py 8 16 32 64
PGBA_PC=64 python profiles/ab_windows.py 8 16 32 64
PGBA_PC=32 python profiles/ab_windows.py 8 16
