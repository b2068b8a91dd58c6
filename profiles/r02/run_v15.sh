mkdir -p gpurun_out
python -m pytest tests/test_ba_gpu.py -m gpu -x -q -k "plan_cache or plan_variants or batched" 2>&1 | tail -3
c5() { python bench.py --steps 20 --warmup 5 --workload c5 --no-extra 2>/dev/null | python -c "
import json,sys; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 $1', round(b['ms_per_step']*1e3,1), 'us; reuse', round(b['plan_reuse']['ms_per_step']*1e3,1), {k: round(v*1e3,1) for k,v in b['stages_ms'].items()})"; }
c5 "32-bit fingerprint"
PGBA_PLAN_CACHE=0 c5 "no fingerprint"
PGBA_LIB=cdv-slam_b200/lib/libpgba_timing.so python profiles/plan_timing.py c5 2>&1 | tail -15
ncu --set full --clock-control none --import-source on -k regex:'wide32|nhwc_f32' -c 2 -o gpurun_out/ncu_corr32_v15 python profiles/ncu_target.py corr128f 1 > gpurun_out/ncu_corr32_v15.log 2>&1; echo ncu rc=$?
