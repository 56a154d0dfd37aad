"""Diagnostic: vjp at 5000 rays, GPU vs the oracle in fp32 and fp64 (which side carries the error?)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import nice_oracle as O
nsb = importlib.import_module("nice-slam-cpp_b200")
syn = nsb.synthetic
CAM = syn.CAM
grids = syn.make_grids(0); decs = syn.make_decoders(0, bias_scale=0.05)
depths, colors, poses = syn.make_frames(5, 0)
idx = syn.mt19937_indices(91, 5800, 480 * 640)
ro, rd, gd, gc = O.ray_sampler(0, 480, 0, 640, idx, 360.0, 360.0, 320.0, 240.0, torch.tensor(depths[3]), torch.tensor(colors[3]), torch.tensor(poses[3]), "reference")
m = O.inside_mask(ro, rd, gd, torch.tensor(O.BOUND)).numpy()
ro, rd, gd = ro.numpy()[m][:5000], rd.numpy()[m][:5000], gd.numpy()[m][:5000]
n = ro.shape[0]
cfg = nsb.default_config(); cfg.max_rays = 8192
e = nsb.Engine(cfg); e.set_model(grids, decs)
tt, ts = O.t_tables(); e.set_ttables(tt.numpy(), ts.numpy())
rel = lambda a, b: float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / np.abs(np.asarray(b, np.float64)).max())
rs = np.random.RandomState(4)
g_rgb = rs.randn(n, 3).astype(np.float32); g_depth = rs.randn(n).astype(np.float32); g_var0 = (0.3 * rs.randn(n)).astype(np.float32)
for tag, g_var in (("gvar", g_var0), ("gvar=0", np.zeros(n, np.float32))):
    got = e.render_vjp(rd, ro, "color", gd, g_rgb, g_depth, g_var)
    res = {}
    for dt in (torch.float32, torch.float64):
        mm = O.Model(grids, decs, dtype=dt)
        for k in ("middle", "fine", "color"):
            mm.grids[k].requires_grad_(True)
        mm.flat["color"].requires_grad_(True)
        tro = torch.tensor(ro).to(dt).requires_grad_(True); trd = torch.tensor(rd).to(dt).requires_grad_(True)
        t2, s2 = O.t_tables(dt)
        ref = O.render_batch_ray(mm, trd, tro, "color", torch.tensor(gd).to(dt), t2, s2)
        ((ref[0] * torch.tensor(g_rgb).to(dt)).sum() + (ref[1] * torch.tensor(g_depth).to(dt)).sum() + (ref[2] * torch.tensor(g_var).to(dt)).sum()).backward()
        res[dt] = {"grid_middle": mm.grids["middle"].grad.numpy(), "grid_fine": mm.grids["fine"].grad.numpy(), "grid_color": mm.grids["color"].grad.numpy(),
                   "dec_color": mm.flat["color"].grad.numpy(), "rays_o": tro.grad.numpy(), "rays_d": trd.grad.numpy()}
    for k in ("grid_middle", "grid_fine", "grid_color", "dec_color", "rays_o", "rays_d"):
        print(tag, k, "gpu-vs-f32 %.2e  gpu-vs-f64 %.2e  f32-vs-f64 %.2e" % (rel(got[k], res[torch.float32][k]), rel(got[k], res[torch.float64][k]), rel(res[torch.float32][k], res[torch.float64][k])), flush=True)
# loss-like cotangents (what the L1 losses of Mapper.cpp:435-442 produce): g_depth = +-1, g_rgb = +-0.5, g_var = 0
g_rgb = (0.5 * np.sign(rs.randn(n, 3))).astype(np.float32); g_depth = np.sign(rs.randn(n)).astype(np.float32); g_var = np.zeros(n, np.float32)
got = e.render_vjp(rd, ro, "color", gd, g_rgb, g_depth, g_var)
mm = O.Model(grids, decs)
for k in ("middle", "fine", "color"):
    mm.grids[k].requires_grad_(True)
mm.flat["color"].requires_grad_(True)
tro = torch.tensor(ro, requires_grad=True); trd = torch.tensor(rd, requires_grad=True)
ref = O.render_batch_ray(mm, trd, tro, "color", torch.tensor(gd))
((ref[0] * torch.tensor(g_rgb)).sum() + (ref[1] * torch.tensor(g_depth)).sum()).backward()
for k, r in (("grid_middle", mm.grids["middle"].grad), ("grid_fine", mm.grids["fine"].grad), ("grid_color", mm.grids["color"].grad), ("dec_color", mm.flat["color"].grad), ("rays_o", tro.grad), ("rays_d", trd.grad)):
    print("sign-cotangents", k, "gpu-vs-f32 %.2e" % rel(got[k], r.numpy()), flush=True)
# one mapping iteration
cfg2 = nsb.default_config(); cfg2.max_rays = 8192; cfg2.mapping_pixels = 5000; cfg2.frustum_feature_selection = 0
e2 = nsb.Engine(cfg2); e2.set_model(grids, decs); e2.set_ttables(tt.numpy(), ts.numpy())
for f in range(5):
    e2.set_frame(f, depths[f], colors[f], poses[f])
for dt in (torch.float32, torch.float64):
    m2 = O.Model(grids, decs, dtype=dt); go = {}
    ref_losses, _ = O.mapping_iters(m2, depths[:5], colors[:5], poses[:5], syn.CAM, 5000, ["color"], seed=17, raydir="pinhole", grads_out=go)
    e2.set_model(grids, decs); e2.seed(17); e2.mapping_capture_grads(True); e2.mapping_begin(list(range(5)), 60, 1.0)
    loss = e2.mapping_iter(59); cg = e2.captured_grads()
    print(dt, "loss", loss, ref_losses[0], abs(loss - ref_losses[0]) / abs(ref_losses[0]))
    for lv in ("middle", "fine", "color"):
        print(dt, "mapping grad", lv, "%.2e" % rel(cg["grid_" + lv], go[lv].numpy()))
    print(dt, "mapping grad dec_color %.2e" % rel(cg["dec_color"], go["dec_color"].numpy()), flush=True)
