python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
run() { echo "== $*"; env "$@" python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fps', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'cas', round(d['roofline_cascade']['ms_per_launch'],2), 'pyr', round(d['roofline_pyramid']['ms_per_launch'],2))"; }
run A=1
run WBG_CAS_TILE_SKIP=1
run WBG_CAS_TILE_SKIP=1 WBG_CAS_FLAGS=2
run WBG_CAS_TILE_SKIP=1 WBG_CAS_COMPACT_DEN=3
run WBG_CAS_TILE_SKIP=1 WBG_CAS_ROUND_FULL=8 WBG_CAS_ROUND_MID=16 WBG_CAS_ROUND_TAIL=32
run WBG_CAS_TILE_SKIP=1 WBG_CAS_ROUND_FULL=32 WBG_CAS_ROUND_MID=64 WBG_CAS_ROUND_TAIL=128
run WBG_CAS_TILE_SKIP=2
