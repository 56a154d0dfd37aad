"""Cycle breakdown of the tcgen05 forward kernel (needs the tctiming build variant). usage: NSB_TCGEN05=1 python tools/tc_timing.py"""
import ctypes as C, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
nsb = importlib.import_module("nice-slam-cpp_b200"); syn = nsb.synthetic
cfg = nsb.default_config("tctiming"); cfg.mapping_pixels = 5000; cfg.max_rays = 5000; cfg.frustum_feature_selection = 0
e = nsb.Engine(cfg, variant="tctiming")
e.set_model(syn.make_grids(0), syn.make_decoders(0))
d, c, p = syn.make_frames(5, 0)
for f in range(5): e.set_frame(f, d[f], c[f], p[f])
e.seed(0); e.mapping_begin(list(range(5)), 60, 1.0)
for _ in range(3): e.mapping_iter(0, sync=False)
e.synchronize()
buf = (C.c_uint64 * 32)(); e.lib.nsb_debug_counters(e.h, buf)
e.mapping_iter(0, sync=False); e.synchronize()
e.lib.nsb_debug_counters(e.h, buf)
names = ["gather", "sin", "wait_E", "wait_layers", "epilogue", "tiles"]
for dec in (1, 2, 3):
    v = [buf[8 * dec + k] for k in range(6)]
    tiles = max(v[5], 1)
    print("decoder", dec, {n: round(x / tiles) for n, x in zip(names[:5], v[:5])}, "tiles(sum over CTAs, group 0)", v[5], "cycles/tile", round(sum(v[:5]) / tiles))
