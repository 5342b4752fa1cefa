run() { echo "== $*"; env "$@" python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fps', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'cas', round(d['roofline_cascade']['ms_per_launch'],2), 'pyr', round(d['roofline_pyramid']['ms_per_launch'],2))"; }
run WBG_CAS_FLAGS=2
run WBG_CAS_FLAGS=1
run WBG_CAS_FLAGS=2 WBG_CAS_TILE_SKIP=1
run WBG_CAS_FLAGS=2 WBG_CAS_ROUND_FULL=32 WBG_CAS_ROUND_MID=64 WBG_CAS_ROUND_TAIL=128
run WBG_CAS_FLAGS=2 WBG_CAS_ROUND_FULL=32 WBG_CAS_ROUND_MID=32 WBG_CAS_ROUND_TAIL=32
run WBG_CAS_FLAGS=2 WBG_CAS_ROUND_FULL=64 WBG_CAS_ROUND_MID=64 WBG_CAS_ROUND_TAIL=64
run WBG_CAS_FLAGS=2 WBG_CAS_COMPACT_NUM=2 WBG_CAS_COMPACT_DEN=5
