mkdir -p gpurun_out
(time python -m pytest tests/test_ba_gpu.py -m gpu -x -q) > gpurun_out/pytest_ba_v6.log 2>&1
tail -5 gpurun_out/pytest_ba_v6.log
c5() { python bench.py --steps 20 --warmup 5 --workload c5 --no-extra 2>/dev/null | python -c "
import json,sys; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 $1', round(b['ms_per_step']*1e3,1), 'us; reuse', round(b['plan_reuse']['ms_per_step']*1e3,1), {k: round(v*1e3,1) for k,v in b['stages_ms'].items()})"; }
for g in 1 2 4; do PGBA_BATCH_GROUPS=$g c5 "groups=$g"; done
PGBA_BATCH_GROUPS=4 PGBA_PLAN_DIRECT_CL=4 c5 "groups=4 cl=4"
PGBA_BATCH_GROUPS=4 PGBA_PLAN_DIRECT_CL=2 c5 "groups=4 cl=2"
PGBA_BATCH_GROUPS=4 PGBA_PLAN_DIRECT=0 c5 "groups=4 old-plan"
export PGBA_LIB=cdv-slam_b200/lib/libpgba_timing.so
echo "== cta_trace c5 (cache hit)"; python profiles/cta_trace.py c5 2>&1 | tail -12
