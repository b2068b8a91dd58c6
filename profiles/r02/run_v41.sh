python profiles/ab_c5.py c2
PGBA_LIN_L2_PREFETCH=0 python profiles/ab_c5.py c2
python profiles/ab_c5.py c2
PGBA_LIN_L2_PREFETCH=0 python profiles/ab_c5.py c2
