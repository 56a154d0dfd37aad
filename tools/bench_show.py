"""Print the interesting parts of bench.py JSON lines. usage: bench_show.py file..."""
import json, sys
for f in sys.argv[1:]:
    for l in open(f):
        if not l.startswith("{"): continue
        d = json.loads(l)
        print(f, "value %.0f  ms %.4f  e2e %.0f  launches %s clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches"), d.get("clocks")))
        r = d.get("roofline")
        if r:
            print("   dom", r["kernel"], "ach %.1f frac %.4f" % (r["achieved"], r["frac"]), {k: round(v["avg_launch_ms"], 4) for k, v in r.get("kernels", {}).items()})
            print("   kms/step", {k: round(v / d["steps"], 4) for k, v in r["kernel_ms_total"].items()}, "gather ms %.4f" % r["grid_sampling"]["ms"])
        if d.get("tracking"): print("   tracking", {k: v for k, v in d["tracking"].items() if k != "note"})
        print("   cpu", d.get("cpu_baseline"), "loss", d.get("loss_first_last"))
