"""Profiling target for the grid-sampling kernel alone (k_gather_only on the rays of one mapping iteration).  usage: prof_gather.py"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nsb = importlib.import_module("nice-slam-cpp_b200")
syn = nsb.synthetic
cfg = nsb.default_config(); cfg.mapping_pixels = 5000; cfg.max_rays = 5000; cfg.frustum_feature_selection = 0
e = nsb.Engine(cfg)
e.set_model(syn.make_grids(0), syn.make_decoders(0))
d, c, p = syn.make_frames(5, 0)
for f in range(5):
    e.set_frame(f, d[f], c[f], p[f])
e.seed(0)
e.mapping_begin(list(range(5)), 60, 1.0)
e.mapping_iter(0, sync=True)
print("gather ms", e.bench_gather(3))
