"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz from oracle/_ref, i.e. from the reference's own
Renderer.cpp + utils.h compiled against libtorch CPU (oracle/Makefile).  Run it in the container that has
/root/reference:

    make -C oracle && python oracle/make_golden.py

The fixtures carry the inputs that cannot be regenerated bit-exactly elsewhere (rays, cotangents, t-tables of
this host's ATen) plus sha256 digests of the synthetic grids / decoders they were produced with, and the
reference outputs.  Gradients of the 5 MB grids are stored at a fixed set of sampled positions.
"""
import hashlib
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch

syn = importlib.import_module("nice-slam-cpp_b200.synthetic")
import nice_oracle as O
import refbind as R

GOLD = os.path.join(ROOT, "tests", "golden")
CAM = syn.CAM
STAGE_ID = O.STAGE_ID


def digest(grids, decs):
    h = hashlib.sha256()
    for k in syn.LEVELS:
        h.update(np.ascontiguousarray(grids[k]).tobytes()); h.update(np.ascontiguousarray(decs[k]).tobytes())
    return h.hexdigest()


def sample_positions(arr, k, seed):
    """Fixed sampled positions of a big gradient: the k/2 largest-magnitude entries + k/2 random ones."""
    flat = arr.reshape(-1)
    top = np.argsort(-np.abs(flat))[:k // 2]
    rnd = np.random.RandomState(seed).randint(0, flat.size, k // 2)
    pos = np.unique(np.concatenate([top, rnd])).astype(np.int64)
    return pos, flat[pos].copy()


def main():
    os.makedirs(GOLD, exist_ok=True)
    assert R.available() and R.available(verbatim=True), "build oracle/_ref first: make -C oracle"
    grids = syn.make_grids(0); decs = syn.make_decoders(0, bias_scale=0.05)
    dg = digest(grids, decs)
    depths, colors, poses = syn.make_frames(5, 0)
    tt, ts = O.t_tables()
    ref = R.Ref(grids, decs); refv = R.Ref(grids, decs, verbatim=True)

    # ---- get_samples / raySampler (utils.h:13-55) + quad2rotation / get_camera_from_tensor (utils.h:174-210)
    ro, rd, gd, gc, idx = R.get_samples(20, 460, 20, 620, 256, CAM, poses[1], depths[1], colors[1], seed=7)
    quats = np.array([[1, 0, 0, 0], [0.9961947, 0, 0.08715574, 0], [0.7, -0.1, 0.5, 0.3], [0.2, 0.9, -0.4, 0.1], [2.0, 0.5, -1.0, 0.25]], np.float32)
    rots = np.stack([R.quad2rotation(q) for q in quats])
    cam7 = np.array([0.7, -0.1, 0.5, 0.3, 1.5, -2.0, 0.25], np.float32)
    np.savez_compressed(os.path.join(GOLD, "sampling.npz"), H0=20, H1=460, W0=20, W1=620, frame=1, seed=7, idx=idx, rays_o=ro, rays_d=rd,
                        gt_depth=gd, gt_color=gc, quats=quats, rots=rots, cam7=cam7, RT=R.get_camera_from_tensor(cam7))

    # ---- raw2outputs_nerf_color (utils.h:148-172): patched (per-ray norm) and verbatim (p = -1 batch norm)
    rs = np.random.RandomState(1)
    raw = rs.randn(16, 48, 4).astype(np.float32); z = np.sort(rs.uniform(0.1, 3, (16, 48)).astype(np.float32), 1); d = rs.randn(16, 3).astype(np.float32)
    out_p = R.raw2outputs(raw, z, d)
    libv = R.C.CDLL(R.lib_path(True))
    rgb = np.empty((16, 3), np.float32); dep = np.empty(16, np.float32); var = np.empty(16, np.float32); w = np.empty((16, 48), np.float32)
    assert libv.ref_raw2outputs(16, 48, R._p(raw), R._p(z), R._p(d), R._p(rgb), R._p(dep), R._p(var), R._p(w)) == 0
    np.savez_compressed(os.path.join(GOLD, "raw2outputs.npz"), raw=raw, z=z, rays_d=d, rgb=out_p[0], depth=out_p[1], var=out_p[2], weights=out_p[3],
                        v_rgb=rgb, v_depth=dep, v_var=var, v_weights=w)

    # ---- render_batch_ray forward (Renderer.cpp:44-125), all stages + the no-depth path
    idx = syn.mt19937_indices(7, 96, CAM["H"] * CAM["W"])
    a, b, c_, d_ = O.ray_sampler(0, CAM["H"], 0, CAM["W"], idx, CAM["fx"], CAM["fy"], CAM["cx"], CAM["cy"], torch.tensor(depths[0]), torch.tensor(colors[0]), torch.tensor(poses[0]))
    m = O.inside_mask(a, b, c_, torch.tensor(syn.BOUND)).numpy()
    ro, rd, gd = a.numpy()[m], b.numpy()[m], c_.numpy()[m]
    fx = dict(digest=dg, rays_o=ro, rays_d=rd, gt_depth=gd, t_samples=tt.numpy(), t_surface=ts.numpy())
    for st in ("color", "fine", "middle", "coarse"):
        o = ref.render_batch_ray(rd, ro, st, gd)
        ov = refv.render_batch_ray(rd, ro, st, gd)
        for nm, x, y in zip(("rgb", "depth", "var", "weights"), o, ov):
            fx["%s_%s" % (st, nm)] = x
            if st == "color":
                fx["verbatim_%s" % nm] = y          # Renderer.cpp untouched, utils.h:153 literal norm
    o = ref.render_batch_ray(rd, ro, "coarse", None, n_out=32)
    for nm, x in zip(("rgb", "depth", "var", "weights"), o):
        fx["nodepth_coarse_%s" % nm] = x
    pts = np.concatenate([np.random.RandomState(3).uniform(syn.BOUND[:, 0], syn.BOUND[:, 1], (192, 3)), np.random.RandomState(4).uniform(-6, 6, (64, 3))]).astype(np.float32)
    fx["pts"] = pts
    for st in ("color", "coarse"):
        fx["eval_%s" % st] = ref.eval_points(pts, st)
    np.savez_compressed(os.path.join(GOLD, "render_forward.npz"), **fx)

    # ---- vjp through render_batch_ray (patched build: autograd on), sampled gradient entries
    rs = np.random.RandomState(5)
    n = ro.shape[0]
    g_rgb = rs.randn(n, 3).astype(np.float32); g_depth = rs.randn(n).astype(np.float32); g_var = (0.3 * rs.randn(n)).astype(np.float32)
    g = ref.render_vjp(rd, ro, "color", gd, g_rgb, g_depth, g_var)
    fx = dict(digest=dg, rays_o=ro, rays_d=rd, gt_depth=gd, t_samples=tt.numpy(), t_surface=ts.numpy(), g_rgb=g_rgb, g_depth=g_depth, g_var=g_var,
              d_rays_o=g["rays_o"], d_rays_d=g["rays_d"], d_dec_color=g["dec_color"])
    for lv in ("middle", "fine", "color"):
        pos, val = sample_positions(g["grid_" + lv], 4096, 17)
        fx["grid_%s_pos" % lv] = pos; fx["grid_%s_val" % lv] = val
        fx["grid_%s_l2" % lv] = np.float64(np.sqrt((g["grid_" + lv].astype(np.float64) ** 2).sum()))
        fx["grid_%s_max" % lv] = np.float32(np.abs(g["grid_" + lv]).max())
    np.savez_compressed(os.path.join(GOLD, "render_vjp.npz"), **fx)

    # ---- mapping iterations (Mapper.cpp:330-465) and tracking iterations (Tracker.cpp:41-113)
    stages = ["middle", "middle", "color", "color"]
    lr = np.array([O.DEFAULT_LR[k] for k in ("coarse", "middle", "fine", "color")], np.float32)
    ref2 = R.Ref(grids, decs)
    g_it0 = ref2.capture_grads(0)          # gradients libtorch autograd leaves at Mapper.cpp:444 in the first (geometry) iteration
    losses, n_in, _ = ref2.mapping_iters(depths[:2], colors[:2], poses[:2], CAM, 200, [STAGE_ID[s] for s in stages], lr, seed=3)
    fx = dict(digest=dg, stages=np.array([STAGE_ID[s] for s in stages]), losses=losses, n_inside=n_in, pixels=200, n_frames=2, seed=3,
              t_samples=tt.numpy(), t_surface=ts.numpy(), dec_color=ref2.get_decoder("color"))
    for lv in ("middle", "fine", "color"):
        gnew = ref2.get_grid(lv)
        pos, _ = sample_positions(gnew - grids[lv], 4096, 23)
        fx["grid_%s_pos" % lv] = pos; fx["grid_%s_val" % lv] = gnew.reshape(-1)[pos]
    # a colour-stage iteration from the SAME initial parameters (fresh context, own seed): colour grid + colour-decoder gradients
    ref2c = R.Ref(grids, decs)
    g_c0 = ref2c.capture_grads(0)
    lc, nc, _ = ref2c.mapping_iters(depths[:2], colors[:2], poses[:2], CAM, 200, [STAGE_ID["color"]], lr, seed=4)
    fx["c0_loss"] = lc; fx["c0_seed"] = 4
    for tag, g in (("it0", g_it0), ("c0", g_c0)):
        for lv in ("middle", "fine", "color"):
            a = g["grid_" + lv]
            pos, val = sample_positions(a, 4096, 29)
            fx["%s_grad_%s_pos" % (tag, lv)] = pos; fx["%s_grad_%s_val" % (tag, lv)] = val
            fx["%s_grad_%s_l2" % (tag, lv)] = np.float64(np.sqrt((a.astype(np.float64) ** 2).sum()))
            fx["%s_grad_%s_max" % (tag, lv)] = np.float32(np.abs(a).max())
        fx["%s_grad_dec_color" % tag] = g["dec_color"]
    np.savez_compressed(os.path.join(GOLD, "mapping_iters.npz"), **fx)

    ref3 = R.Ref(grids, decs)
    cam7 = O.get_tensor_from_camera(syn.yaw_pose(10.0))
    c1, l1, g1, n1, _ = ref3.tracking_iters(depths[0], colors[0], cam7, CAM, 300, 3, 1e-3, seed=5)
    np.savez_compressed(os.path.join(GOLD, "tracking_iters.npz"), digest=dg, cam7_in=cam7, cam7_out=c1, losses=l1, grad_first=g1, n_inside=n1,
                        pixels=300, lr=1e-3, seed=5, t_samples=tt.numpy(), t_surface=ts.numpy())
    tot = sum(os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD))
    print("golden fixtures written to", GOLD, "total bytes", tot)


if __name__ == "__main__":
    main()
