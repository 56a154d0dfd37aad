"""Builds profiles/traffic.json (DRAM bytes per launch and kernel variant) and profiles/gather_ncu.json (L1 / L2 / DRAM bytes of
k_gather_only) from `ncu --set full` captures.  usage: make_traffic.py geometry.ncu-rep color.ncu-rep [gather.ncu-rep] [l2_gather_peak_gbs]"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    return hdr, rows[2:]


def num(r, hdr, name):
    try:
        return float(r[hdr.index(name)].replace(",", ""))
    except (ValueError, IndexError):
        return None


def short(name):
    n = name.split("(")[0].replace("void ", "").replace("nsb::", "").replace("t5::", "").replace("tc16::", "").replace("tc::", "")
    return n.split("<")[0]


traffic = {"_source": "ncu --set full --clock-control none: dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the captured launches), bytes; "
                      "captures: " + ", ".join(os.path.basename(a) for a in sys.argv[1:3])}
for rep, stage in ((sys.argv[1], "geometry"), (sys.argv[2], "color")):
    hdr, rows = rows_of(rep)
    acc = {}
    for r in rows:
        k = short(r[hdr.index("Kernel Name")])
        rd, wr = num(r, hdr, "dram__bytes_read.sum"), num(r, hdr, "dram__bytes_write.sum")
        if rd is None or wr is None:
            continue
        acc.setdefault(k, []).append(rd + wr)
    for k, v in acc.items():
        traffic["%s<%s>" % (k, stage)] = int(sum(v) / len(v))
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
if len(sys.argv) > 3:
    hdr, rows = rows_of(sys.argv[3])
    for r in rows:
        if "k_gather_only" in r[hdr.index("Kernel Name")]:
            g = {"_source": "ncu --set full of nsb_bench_gather (" + os.path.basename(sys.argv[3]) + "), per launch, bytes",
                 "l1tex_t_bytes": num(r, hdr, "l1tex__t_bytes.sum") or 32.0 * (num(r, hdr, "l1tex__t_sectors.sum") or num(r, hdr, "SM_B.TriageCompute.l1tex__t_sectors.sum") or 0),
                 "lts_t_bytes": num(r, hdr, "lts__t_bytes.sum") or 32.0 * (num(r, hdr, "lts__t_sectors.sum") or num(r, hdr, "lts__t_sectors_srcunit_tex.sum") or 0),
                 "l1_hit_rate_pct": num(r, hdr, "l1tex__t_sector_hit_rate.pct"),
                 "dram_bytes": (num(r, hdr, "dram__bytes_read.sum") or 0) + (num(r, hdr, "dram__bytes_write.sum") or 0),
                 "gpu_time_us": (num(r, hdr, "gpu__time_duration.sum") or 0) / 1e3,
                 "l2_gather_peak_gbs": float(sys.argv[4]) if len(sys.argv) > 4 else None}
            json.dump(g, open(os.path.join(ROOT, "profiles", "gather_ncu.json"), "w"), indent=1)
            print(json.dumps(g, indent=1))
            break
