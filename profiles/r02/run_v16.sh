for lu in timing lu6 lu8; do echo "== $lu"; PGBA_LIB=cdv-slam_b200/lib/libpgba_$lu.so python profiles/plan_timing.py c5 2>&1 | grep -E "D0|D5 verify"; done
c2() { python bench.py --steps 30 --warmup 5 --no-extra 2>/dev/null | python -c "
import json,sys; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 $1', round(b['ms_per_step']*1e3,1), 'reuse', round(b['plan_reuse']['ms_per_step']*1e3,1), {k: round(v*1e3,1) for k,v in b['stages_ms'].items()})"; }
c2 "default (cluster + cells)"
PGBA_PLAN_DIRECT=1 c2 "direct"
c2 "default again"
PGBA_PLAN_DIRECT=1 c2 "direct again"
