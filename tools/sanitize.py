"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck): render, vjp, mapping (colour + BA), tracking, dense strip.
usage: compute-sanitizer --tool memcheck python tools/sanitize.py   (closed on the round-1 GPU pool; also useful as a plain smoke pass)"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nsb = importlib.import_module("nice-slam-cpp_b200")
syn = nsb.synthetic
cfg = nsb.default_config(); cfg.mapping_pixels = 96; cfg.tracking_pixels = 64; cfg.max_rays = 128; cfg.frustum_feature_selection = 1
e = nsb.Engine(cfg)
e.set_model(syn.make_grids(0), syn.make_decoders(0, bias_scale=0.05))
d, c, p = syn.make_frames(3, 0)
for f in range(3):
    e.set_frame(f, d[f], c[f], p[f])
e.seed(0)
ro, rd, gd, gc, ins, _ = e.get_samples(0, 0, 480, 0, 640, 100)
ro, rd, gd = ro[ins], rd[ins], gd[ins]
n = ro.shape[0]
e.render_batch_ray(rd, ro, "color", gd)
e.render_batch_ray(rd, ro, "coarse", None)
e.render_vjp(rd, ro, "color", gd, np.ones((n, 3), np.float32), np.ones(n, np.float32), np.zeros(n, np.float32))
e.eval_points(np.random.RandomState(0).uniform(-4, 3, (37, 3)).astype(np.float32), "color")
e.mapping_begin([0, 1, 2], 60, 1.0, ba_mask=0b110)
for it in (0, 30, 59, 59):
    e.mapping_iter(it)
e.mapping_end()
e.keyframe_selection_overlap(0, [p[1], p[2]], 1)
cam = nsb.get_tensor_from_camera(p[0]); e.tracking_begin(0, cam)
for _ in range(2):
    e.tracking_iter()
e.close()
print("sanitize pass ok")
