# Times the build variants (NSB_VARIANT=<name of build.py VARIANTS>) with the default bench.
for v in ${VARIANTS:-"" f20b20 f16b24 f20b24 f24b24}; do NSB_VARIANT=$v python bench.py --no-cpu-baseline --steps 60 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; k=r['kernel_ms_total']
        print('VARIANT [$v]', 'ms/step %.4f fwd %.4f bwd %.4f'%(d['ms_per_step'], k['decode_fwd']/60, k['decode_bwd']/60), d['ms_per_step_by_stage'], 'trk %.4f'%d['tracking']['ms_per_iter'])
"; done
