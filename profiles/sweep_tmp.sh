python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider 2>&1 | tail -4
run() { echo "== $*"; env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fps', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'cas', round(d['roofline_cascade']['ms_per_launch'],2), 'pyr', round(d['roofline_pyramid']['ms_per_launch'],2), 'e2e', round(d['e2e']['value'],1))"; }
run A=1
