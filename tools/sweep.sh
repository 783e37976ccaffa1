run() { echo "== $*"; env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('rot/s %.0f  e2e %.0f' % (d['value'], d['e2e']['value']), {k: round(1e3*v['ms_per_step']/d['config']['rotations_per_step_per_gpu'],2) for k,v in d['roofline']['kernels'].items()})"; }
run PFB_OVERLAP=0
run PFB_OVERLAP=1
run PFB_OVERLAP=1 PFB_B_THREADS=256 PFB_B_STAGE=0
run PFB_OVERLAP=1 PFB_B_THREADS=256 PFB_B_STAGE=1
run PFB_OVERLAP=1 PFB_B_THREADS=512 PFB_B_STAGE=0
run PFB_OVERLAP=0 PFB_B_THREADS=256 PFB_B_STAGE=0
PFB_OVERLAP=1 PFB_B_THREADS=256 PFB_B_STAGE=0 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
