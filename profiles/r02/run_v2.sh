mkdir -p gpurun_out
(time python -m pytest tests/test_ba_gpu.py -m gpu -x -q) > gpurun_out/pytest_ba_v2.log 2>&1
tail -5 gpurun_out/pytest_ba_v2.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; echo bench rc=$?
for cl in 2 4 8; do PGBA_PLAN_DIRECT_CL=$cl python bench.py --steps 20 --warmup 5 --workload c5 --no-extra 2>/dev/null | python -c "
import json,sys; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 cl=$cl', b['ms_per_step'], b['stages_ms'])"; done
PGBA_PLAN_DIRECT=0 python bench.py --steps 20 --warmup 5 --no-extra 2>/dev/null | python -c "
import json,sys; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 old plan', b['ms_per_step'], b['stages_ms'])"
