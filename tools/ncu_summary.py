"""Summarise an .ncu-rep (raw page) into the handful of numbers the design discussion uses. usage: ncu_summary.py rep [launches.csv]"""
import csv, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
if len(sys.argv) > 2:
    rows = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 5]
    hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
    agg = defaultdict(list)
    for r in rows[1:]:
        try: agg[r[ki].split('(')[0][:60]].append(float(r[vi].replace(',', '')))
        except ValueError: pass
    tot = sum(sum(v) for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-60s n=%3d avg_us=%9.1f share=%5.1f%%" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__warps_eligible.avg.per_cycle_active', 'sass__inst_executed_shared_loads', 'sass__inst_executed_global_loads',
        'lts__t_sectors_op_red.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('-----', r[hdr.index('Kernel Name')][:70])
    for w in want:
        if w in hdr: print('   %-80s %s %s' % (w, r[hdr.index(w)], rows[1][hdr.index(w)]))
    st = [(hdr[i], r[i]) for i in range(len(hdr)) if 'warp_issue_stalled' in hdr[i] and hdr[i].endswith('.pct') or ('issue_stalled' in hdr[i] and 'ratio' in hdr[i])]
    vals = []
    for k, v in st:
        try: vals.append((k, float(v.replace(',', ''))))
        except ValueError: pass
    for k, v in sorted(vals, key=lambda kv: -kv[1])[:8]:
        print('   stall %-74s %.2f' % (k, v))
