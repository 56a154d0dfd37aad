# Final measurement pass of a round on the GPU box: default bench line, ncu launch list of the bench itself, `--set full` captures of one
# geometry and one colour iteration summarised ON the box (the .ncu-rep files stay there: gpurun_out/ is capped at 64 MiB).
cd ${GRAFT_REPO_ROOT:-.}
T=${1:-r5}
timeout 400 python bench.py 2>/dev/null | tail -1 > gpurun_out/${T}_bench_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_bench_launches.csv python bench.py --steps 6 --warmup 4 --no-cpu-baseline --no-configs > /tmp/ncu_list.log 2>&1
timeout 200 ncu --set full --clock-control none -c 10 -o /tmp/${T}_geom python tools/prof_iter.py 0 1 > /tmp/geom.log 2>&1
timeout 300 ncu --set full --clock-control none -c 30 -o /tmp/${T}_color python tools/prof_iter.py 59 2 > /tmp/color.log 2>&1
python tools/ncu_summary.py /tmp/${T}_geom.ncu-rep > gpurun_out/${T}_ncu_summary_geometry.txt 2>&1
python tools/ncu_summary.py /tmp/${T}_color.ncu-rep > gpurun_out/${T}_ncu_summary_color.txt 2>&1
python tools/make_traffic.py /tmp/${T}_geom.ncu-rep /tmp/${T}_color.ncu-rep > /dev/null 2>&1 && cp profiles/traffic.json gpurun_out/${T}_traffic.json
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_n1.json')); print(d['value'], d['ms_per_step'], d['ms_per_step_by_stage'], d['e2e']['value'], d['roofline']['frac'], d['tracking']['ms_per_iter'])"
ls -la gpurun_out/
