for ppc in 1 2 3 8; do echo "PPC=$ppc"; PFB_C_PPC=$ppc python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('rot/s %.0f' % d['value'], {k: round(1e3*v['ms_per_step']/d['config']['rotations_per_step_per_gpu'],2) for k,v in d['roofline']['kernels'].items()})"; done
for b in 16 48 64; do echo "BATCH=$b"; PFB_BATCH=$b python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('rot/s %.0f' % d['value'], {k: round(1e3*v['ms_per_step']/d['config']['rotations_per_step_per_gpu'],2) for k,v in d['roofline']['kernels'].items()})"; done
