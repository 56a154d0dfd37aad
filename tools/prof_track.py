"""Short profiling target: a few tracking iterations (5000 rays) for ncu.  usage: prof_track.py [n_iters]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nsb = importlib.import_module("nice-slam-cpp_b200")
syn = nsb.synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = nsb.default_config(); cfg.tracking_pixels = 5000; cfg.max_rays = 5000
e = nsb.Engine(cfg)
e.set_model(syn.make_grids(0), syn.make_decoders(0))
d, c, p = syn.make_frames(1, 0)
e.set_frame(0, d[0], c[0], p[0])
e.seed(0)
cam = nsb.get_tensor_from_camera(p[0]); cam[4:] += 0.01
e.tracking_begin(0, cam)
for _ in range(n):
    e.tracking_iter(None, want_grad=False)
e.synchronize()
print("ok", e.launch_count())
