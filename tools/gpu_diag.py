"""GPU diagnostic run: compares every stage of the CUDA path with the oracle and prints one JSON line per
check (also written to gpurun_out/diag.json).  Development aid; the assertions live in tests/.

    gpurun -- python tools/gpu_diag.py [--quick] [--variant precise_sin]
"""
import importlib
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch

nsb = importlib.import_module("nice-slam-cpp_b200")
syn = nsb.synthetic
import nice_oracle as O

OUT = []
VARIANT = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else ""
QUICK = "--quick" in sys.argv


def rec(name, **kw):
    d = {"check": name}
    d.update(kw)
    OUT.append(d)
    print(json.dumps(d), flush=True)


def relerr(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def section(fn):
    try:
        fn()
    except Exception as e:  # keep going: one gpurun call should tell as much as possible
        rec(fn.__name__, error=repr(e), tb=traceback.format_exc()[-1500:])


grids = syn.make_grids(0)
decs = syn.make_decoders(0, bias_scale=0.05)
depths, colors, poses = syn.make_frames(5, 0)
cam = syn.CAM
tt, ts = O.t_tables()


def make_engine(precision=0, max_rays=8192, **kw):
    cfg = nsb.default_config(VARIANT)
    cfg.precision = precision
    cfg.max_rays = max_rays
    for k, v in kw.items():
        setattr(cfg, k, v)
    e = nsb.Engine(cfg, variant=VARIANT)
    e.set_model(grids, decs)
    e.set_ttables(tt.numpy(), ts.numpy())
    for f in range(5):
        e.set_frame(f, depths[f], colors[f], poses[f])
    return e


def filtered_rays(n, seed, f=0):
    idx = syn.mt19937_indices(seed, n, cam["H"] * cam["W"])
    ro, rd, gd, gc = O.ray_sampler(0, cam["H"], 0, cam["W"], idx, cam["fx"], cam["fy"], cam["cx"], cam["cy"],
                                   torch.tensor(depths[f]), torch.tensor(colors[f]), torch.tensor(poses[f]), "reference")
    m = O.inside_mask(ro, rd, gd, torch.tensor(syn.BOUND))
    return ro[m].numpy(), rd[m].numpy(), gd[m].numpy(), gc[m].numpy(), idx, m.numpy()


def check_sampling():
    e = make_engine()
    n = 2000
    idx = syn.mt19937_indices(11, n, cam["H"] * cam["W"])
    ro, rd, gd, gc, ins, io = e.get_samples(1, 0, cam["H"], 0, cam["W"], n, idx=idx)
    a, b, c_, d = O.ray_sampler(0, cam["H"], 0, cam["W"], idx, cam["fx"], cam["fy"], cam["cx"], cam["cy"],
                                torch.tensor(depths[1]), torch.tensor(colors[1]), torch.tensor(poses[1]), "reference")
    m = O.inside_mask(a, b, c_, torch.tensor(syn.BOUND)).numpy()
    rec("sampling", rays_o_exact=bool(np.array_equal(ro, a.numpy())), rays_d_exact=bool(np.array_equal(rd, b.numpy())),
        rays_d_maxabs=float(np.abs(rd - b.numpy()).max()), gt_depth_exact=bool(np.array_equal(gd, c_.numpy())),
        gt_color_exact=bool(np.array_equal(gc, d.numpy())), inside_exact=bool(np.array_equal(ins, m)), n_inside=int(m.sum()))
    e.seed(123)
    _, _, _, _, _, io = e.get_samples(0, 20, 460, 20, 620, 500)
    rec("sampling_rng", idx_exact=bool(np.array_equal(io, syn.mt19937_indices(123, 500, 440 * 600))))
    e.close()


def check_forward(precision=0):
    e = make_engine(precision)
    ro, rd, gd, gc, _, _ = filtered_rays(500, 7)
    model = O.Model(grids, decs)
    for stage in ("color", "fine", "middle", "coarse"):
        rgb, depth, var, w = e.render_batch_ray(rd, ro, stage, gd)
        z = e.last_zvals(rd.shape[0])
        with torch.no_grad():
            o = O.render_batch_ray(model, torch.tensor(rd), torch.tensor(ro), stage, torch.tensor(gd), tt, ts, return_aux=True)
        zo = o[4].numpy()
        rec("forward_" + stage, precision=precision, n=int(rd.shape[0]), z_exact=bool(np.array_equal(z, zo)), z_mismatch=int((z != zo).sum()),
            z_maxabs=float(np.abs(z - zo).max()), rgb=relerr(rgb, o[0].numpy()), depth=relerr(depth, o[1].numpy()),
            var=relerr(var, o[2].numpy()), weights=relerr(w, o[3].numpy()), depth_mean=float(depth.mean()))
    rgb, depth, var, w = e.render_batch_ray(rd, ro, "coarse", None)
    with torch.no_grad():
        o = O.render_batch_ray(model, torch.tensor(rd), torch.tensor(ro), "coarse", None, tt, ts, return_aux=True)
    z = e.last_zvals(rd.shape[0], 32)
    rec("forward_nodepth_coarse", z_exact=bool(np.array_equal(z, o[4].numpy())), depth=relerr(depth, o[1].numpy()), var=relerr(var, o[2].numpy()),
        weights=relerr(w, o[3].numpy()))
    # eval_points, inside and outside the bound
    rs = np.random.RandomState(3)
    pts = np.concatenate([rs.uniform(syn.BOUND[:, 0], syn.BOUND[:, 1], (1000, 3)), rs.uniform(-6, 6, (200, 3))]).astype(np.float32)
    for stage in ("color", "coarse"):
        raw = e.eval_points(pts, stage)
        with torch.no_grad():
            ro_ = model.eval_points(torch.tensor(pts), stage).numpy()
        rec("eval_points_" + stage, precision=precision, maxabs=[float(np.abs(raw[:, k] - ro_[:, k]).max()) for k in range(4)],
            scale=[float(np.abs(ro_[:, k]).max()) for k in range(4)])
    e.close()


def check_reference_norm():
    e = make_engine(0, dist_norm=1)
    ro, rd, gd, gc, _, _ = filtered_rays(300, 9)
    rd = rd.copy(); rd[rd == 0] = 1e-3
    model = O.Model(grids, decs)
    rgb, depth, var, w = e.render_batch_ray(rd, ro, "color", gd)
    with torch.no_grad():
        o = O.render_batch_ray(model, torch.tensor(rd), torch.tensor(ro), "color", torch.tensor(gd), tt, ts, dist_norm="reference")
    rec("forward_reference_distnorm", rgb=relerr(rgb, o[0].numpy()), depth=relerr(depth, o[1].numpy()), var=relerr(var, o[2].numpy()), weights=relerr(w, o[3].numpy()))
    e.close()


def oracle_vjp(ro, rd, gd, g_rgb, g_depth, g_var, stage="color"):
    model = O.Model(grids, decs)
    for k in ("middle", "fine", "color"):
        model.grids[k].requires_grad_(True)
    model.flat["color"].requires_grad_(True)
    tro = torch.tensor(ro, requires_grad=True); trd = torch.tensor(rd, requires_grad=True)
    rgb, depth, var, _ = O.render_batch_ray(model, trd, tro, stage, torch.tensor(gd), tt, ts)
    L = (rgb * torch.tensor(g_rgb)).sum() + (depth * torch.tensor(g_depth)).sum() + (var * torch.tensor(g_var)).sum()
    L.backward()
    out = {"rays_o": tro.grad.numpy(), "rays_d": trd.grad.numpy(), "dec_color": model.flat["color"].grad.numpy()}
    for k in ("middle", "fine", "color"):
        out["grid_" + k] = model.grids[k].grad.numpy()
    return out


def check_vjp(precision=0):
    e = make_engine(precision)
    ro, rd, gd, gc, _, _ = filtered_rays(300, 21)
    rs = np.random.RandomState(5)
    n = ro.shape[0]
    g_rgb = rs.randn(n, 3).astype(np.float32); g_depth = rs.randn(n).astype(np.float32); g_var = (0.3 * rs.randn(n)).astype(np.float32)
    ref = oracle_vjp(ro, rd, gd, g_rgb, g_depth, g_var)
    got = e.render_vjp(rd, ro, "color", gd, g_rgb, g_depth, g_var)
    res = {k: relerr(got[k], ref[k]) for k in ref}
    f = nsb_dec_offsets()
    dec = {name: relerr(got["dec_color"][a:b], ref["dec_color"][a:b]) for name, (a, b) in f.items()}
    rec("vjp_color", precision=precision, n=int(n), **res, dec_parts=dec, ref_norms={k: float(np.abs(v).max()) for k, v in ref.items()})
    # grids only / rays only paths (different kernel instantiations)
    got = e.render_vjp(rd, ro, "color", gd, g_rgb, g_depth, g_var, flags=nsb.F_GRID)
    rec("vjp_color_gridonly", **{k: relerr(got[k], ref[k]) for k in ("grid_middle", "grid_fine", "grid_color")})
    got = e.render_vjp(rd, ro, "color", gd, g_rgb, g_depth, g_var, flags=nsb.F_RAY)
    rec("vjp_color_rayonly", rays_o=relerr(got["rays_o"], ref["rays_o"]), rays_d=relerr(got["rays_d"], ref["rays_d"]))
    got = e.render_vjp(rd, ro, "color", gd, g_rgb, g_depth, g_var, flags=nsb.F_GRID | nsb.F_WGRAD)
    rec("vjp_color_grid_wgrad", dec_color=relerr(got["dec_color"], ref["dec_color"]), grid_color=relerr(got["grid_color"], ref["grid_color"]))
    e.close()


def nsb_dec_offsets():
    E, H, C_ = 93, 32, 32
    K = [E, H, H, E + H, H]
    off = 0; out = {}
    out["B"] = (off, off + 3 * E); off += 3 * E
    for i in range(5):
        out["W%d" % i] = (off, off + H * K[i]); off += H * K[i]
        out["b%d" % i] = (off, off + H); off += H
    for i in range(5):
        out["Fc%d" % i] = (off, off + H * C_); off += H * C_
        out["bc%d" % i] = (off, off + H); off += H
    out["Wo"] = (off, off + 4 * H); off += 4 * H
    out["bo"] = (off, off + 4)
    return out


def check_mapping(precision=0):
    n_frames, pixels = 5, 1000
    stages = ["middle", "middle", "color", "color", "color"]
    e = make_engine(precision, mapping_pixels=pixels, frustum_feature_selection=0)
    model = O.Model(grids, decs)
    lo, nin = O.mapping_iters(model, depths[:n_frames], colors[:n_frames], poses[:n_frames], cam, pixels, stages, seed=3, tt=tt, ts=ts)
    # drive the CUDA path with the same stage sequence: iteration numbers of a 60-iteration schedule
    it_of = {"middle": 0, "color": 59}
    e.seed(3)
    e.mapping_begin(list(range(n_frames)), 60, 1.0)
    lg = [e.mapping_iter(it_of[s]) for s in stages]
    _, nin_g = e.mapping_losses(59, 1)
    res = {"loss_oracle": lo, "loss_cuda": lg, "n_inside_oracle": nin, "loss_rel": [abs(a - b) / abs(a) for a, b in zip(lo, lg)]}
    for k in ("middle", "fine", "color"):
        g = e.get_grid(k); o = model.grids[k].numpy()
        res["grid_" + k] = relerr(g, o); res["grid_%s_moved" % k] = float(np.abs(o - grids[k]).max())
        res["grid_%s_rms" % k] = float(np.sqrt(((g - o) ** 2).mean()) / max(np.sqrt(((o - grids[k]) ** 2).mean()), 1e-30))
    d = e.get_decoder("color"); o = model.flat["color"].numpy()
    res["dec_color"] = relerr(d, o); res["dec_color_moved"] = float(np.abs(o - decs["color"]).max())
    rec("mapping_iters", precision=precision, **res)
    e.close()


def check_tracking(precision=0):
    e = make_engine(precision, tracking_pixels=1000, tracking_lr=1e-3)
    model = O.Model(grids, decs)
    cam7 = O.get_tensor_from_camera(syn.yaw_pose(10.0))
    c2, l2, g2, n2 = O.tracking_iters(model, depths[0], colors[0], cam7, cam, 1000, 3, 1e-3, seed=5, tt=tt, ts=ts)
    e.seed(5)
    e.tracking_begin(0, cam7)
    lg, gg = [], []
    for _ in range(3):
        l, g = e.tracking_iter()
        lg.append(l); gg.append(g.tolist())
    rec("tracking_iters", precision=precision, loss_oracle=l2, loss_cuda=lg, grad_oracle=g2.numpy().tolist(), grad_cuda=gg[0],
        grad_rel=relerr(gg[0], g2.numpy()), cam_oracle=c2.numpy().tolist(), cam_cuda=e.tracking_camera().tolist(), n_inside=n2)
    e.close()


def check_timing(precision=0, pixels=5000):
    e = make_engine(precision, mapping_pixels=pixels, frustum_feature_selection=0)
    e.mapping_begin(list(range(5)), 60, 1.0)
    for it_name, it in (("middle", 0), ("color", 59)):
        for _ in range(3):
            e.mapping_iter(it, sync=False)
        e.synchronize()
        t0 = time.time()
        K = 20
        for _ in range(K):
            e.mapping_iter(it, sync=False)
        e.synchronize()
        dt = (time.time() - t0) / K
        e.set_profiling(True)
        e.mapping_iter(it, sync=False)
        ms = e.kernel_ms()
        e.set_profiling(False)
        rec("timing_mapping_" + it_name, precision=precision, pixels=pixels, ms_per_iter=dt * 1e3, rays_per_s=pixels / dt, kernel_ms=ms)
    e.close()
    e = make_engine(precision, max_rays=65536)
    ro, rd, gd, gc, _, _ = filtered_rays(60000, 5)
    e.render_batch_ray(rd, ro, "color", gd, want_weights=False)
    t0 = time.time()
    e.render_batch_ray(rd, ro, "color", gd, want_weights=False)
    dt = time.time() - t0
    rec("timing_render_fwd", precision=precision, n=int(rd.shape[0]), ms=dt * 1e3, rays_per_s=rd.shape[0] / dt)
    e.close()


if __name__ == "__main__":
    rec("env", build=nsb.load_library(VARIANT).nsb_build_info().decode(), gpu=torch.cuda.get_device_name(0) if torch.cuda.is_available() else None,
        variant=VARIANT)
    if "--timing-only" in sys.argv:
        section(check_timing)
        sys.exit(0)
    section(check_sampling)
    section(check_forward)
    section(check_reference_norm)
    section(check_vjp)
    section(check_mapping)
    section(check_tracking)
    if not QUICK:
        section(lambda: check_forward(1))
        section(lambda: check_vjp(1))
        section(check_timing)
        section(lambda: check_timing(1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "diag%s.json" % ("_" + VARIANT if VARIANT else "")), "w") as f:
        json.dump(OUT, f, indent=1)
