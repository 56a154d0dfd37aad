run() { env "$1=$2" python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 $2', 'trk %.4f'%d['tracking']['ms_per_iter'])
"; }
run NSB_NONE 0
for s in "0,768,768,768" "0,700,860,768" "0,680,900,740" "0,720,820,780" "0,650,950,720" "0,740,800,800"; do run NSB_SPLIT_BWD $s; done
