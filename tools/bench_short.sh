# Quick A/B of build variants: VARIANTS="a b" tools/bench_short.sh   (prints per-stage step and kernel times)
for v in ${VARIANTS:-""}; do NSB_VARIANT=$v python bench.py --no-cpu-baseline --no-configs --steps 60 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; k=r['kernel_ms_total']
        print('VARIANT [$v]', 'ms/step %.4f fwd %.4f bwd %.4f wgrad %.4f'%(d['ms_per_step'], k['decode_fwd']/60, k['decode_bwd']/60, k['wgrad']/23), d['ms_per_step_by_stage'], 'trk %.4f'%d['tracking']['ms_per_iter'])
"; done
