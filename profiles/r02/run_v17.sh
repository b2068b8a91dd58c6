python -m pytest tests/test_ba_gpu.py -m gpu -x -q -k "forced_chunk_size and (128-1 or 64-1)" 2>&1 | tail -2
c5() { python bench.py --steps 20 --warmup 5 --workload c5 --no-extra 2>/dev/null | python -c "
import json,sys; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 $1', round(b['ms_per_step']*1e3,1), 'us', {k: round(v*1e3,1) for k,v in b['stages_ms'].items()})"; }
c5 "FFMA2 Schur"
PGBA_SCHUR_UMMA=1 c5 "tcgen05 Schur"
c5 "FFMA2 Schur again"
PGBA_SCHUR_UMMA=1 c5 "tcgen05 Schur again"
