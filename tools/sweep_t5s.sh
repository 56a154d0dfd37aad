run() { env "$1=$2" timeout 60 python bench.py --no-cpu-baseline --no-configs --steps 60 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernel_ms_total']; print('$1 $2', '%.4f'%d['ms_per_step'], d['ms_per_step_by_stage']['geometry'], d['ms_per_step_by_stage']['color'], 'fwd %.3f bwd %.3f'%(k['decode_fwd'],k['decode_bwd']))
"; }
run NSB_NONE 0
run NSB_SPLIT_FWD_T5S 0,700,1300,1150
run NSB_SPLIT_FWD_T5S 0,700,1300,1450
