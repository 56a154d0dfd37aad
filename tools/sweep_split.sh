# Sweep of the static CTA split between the decoders (NSB_SPLIT_FWD / NSB_SPLIT_BWD = coarse,middle,fine,colour weights).
# Read the by-stage step times: geometry iterations run fwd without the stash and bwd without the colour decoder.
run() { env "$1=$2" python bench.py --no-cpu-baseline --steps 60 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; k=r['kernel_ms_total']
        print('$1 $2', 'ms/step %.4f fwd %.4f bwd %.4f'%(d['ms_per_step'], k['decode_fwd']/60, k['decode_bwd']/60), d['ms_per_step_by_stage'])
"; }
run NSB_NONE 0
for s in "300,732,972,732" "300,700,1000,732" "300,760,940,732" "300,700,940,800" "300,680,940,860" "300,700,900,900" "300,660,980,860"; do run NSB_SPLIT_FWD $s; done
for s in "0,480,480,1400" "0,440,520,1400" "0,500,460,1400" "0,460,500,1250" "0,460,500,1600" "0,420,540,1400"; do run NSB_SPLIT_BWD $s; done
