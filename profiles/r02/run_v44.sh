PGBA_PLAN_DIRECT_CL=4 python profiles/ab_windows.py 12 16 24 32
PGBA_PLAN_DIRECT_CL=2 python profiles/ab_windows.py 12 16 24 32
PGBA_PLAN_DIRECT_CL=8 python profiles/ab_windows.py 12 16
python profiles/ab_windows.py 12 24 48
