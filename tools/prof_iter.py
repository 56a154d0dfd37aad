"""Short profiling target: a few mapping iterations (5000 rays) for ncu.  usage: prof_iter.py [stage_iter] [n]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nsb = importlib.import_module("nice-slam-cpp_b200")
syn = nsb.synthetic
it = int(sys.argv[1]) if len(sys.argv) > 1 else 59
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cfg = nsb.default_config(); cfg.mapping_pixels = 5000; cfg.max_rays = 5000; cfg.frustum_feature_selection = 0
cfg.precision = int(os.environ.get("NSB_PRECISION", "0"))
e = nsb.Engine(cfg)
e.set_model(syn.make_grids(0), syn.make_decoders(0))
d, c, p = syn.make_frames(5, 0)
for f in range(5):
    e.set_frame(f, d[f], c[f], p[f])
e.seed(0)
e.mapping_begin(list(range(5)), 60, 1.0)
for _ in range(n):
    e.mapping_iter(it, sync=False)
e.synchronize()
print("ok", e.launch_count())
