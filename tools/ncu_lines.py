"""Aggregate warp-stall samples of an .ncu-rep per CUDA source line. usage: ncu_lines.py rep [kernel-index] [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
kern = -1; cur_file = None; hdr = None; agg = {}; total = 0; seen_fn = None
for row in csv.reader(io.StringIO(out)):
    if not row: continue
    if row[0] == "File Path": cur_file = row[1].split("/")[-1]; continue
    if row[0] == "Function Name":
        if row[1] != seen_fn: seen_fn = row[1]; kern += 1
        continue
    if row[0] == "Line No": hdr = row; continue
    if hdr is None or kern != want or row[0] == "": continue
    try:
        iS = hdr.index("# Samples")
        n = int(row[iS] or 0)
    except Exception: continue
    st = {}
    for c in ("stall_lg", "stall_long_sb", "stall_math", "stall_mio", "stall_short_sb", "stall_wait", "stall_barrier", "stall_not_selected", "stall_no_inst", "stall_branch_resolving"):
        if c in hdr:
            v = row[hdr.index(c)]
            if v not in ("", "0"): st[c[6:]] = int(v)
    key = (cur_file, int(row[0]))
    a = agg.setdefault(key, [0, row[1].strip()[:100], {}, 0])
    a[0] += n; a[3] += int(row[hdr.index("Instructions Executed")] or 0)
    for k, v in st.items(): a[2][k] = a[2].get(k, 0) + v
    total += n
print("kernel", seen_fn if kern == want else want, "total samples", total)
for (f, ln), (n, src, st, ie) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% %8d inst  %s:%d  %s  %s" % (100.0 * n / max(total, 1), ie, f, ln, src, st))
