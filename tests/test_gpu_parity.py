"""GPU parity tests (pytest -m gpu, run on a B200): every check calls the CUDA path through the C ABI
(libnsb.so via the ctypes binding) and compares it with
  * the golden vectors of oracle/_ref (the reference's own Renderer.cpp + utils.h), tests/golden/*.npz,
  * the oracle restatement (oracle/nice_oracle.py) evaluated live on the same seeded inputs,
  * size-independent properties at the BASELINE.json batch size (5000 rays x 48 samples).

Tolerances (BASELINE.json north_star): pixel indices, rays and sample placement (z values) bit-exact;
rgb / depth / var / weights within 1e-4 relative (max-norm); grid / decoder / pose gradients within 1e-3.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, relerr
import nice_oracle as O

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-4
GRAD_TOL = 1e-3
CAM = dict(H=480, W=640, fx=360.0, fy=360.0, cx=320.0, cy=240.0)


@pytest.fixture(scope="module")
def engine_factory(nsb, model_inputs, frames):
    grids, decs, _ = model_inputs
    depths, colors, poses = frames
    made = []

    def make(t_samples=None, t_surface=None, **kw):
        cfg = nsb.default_config()
        cfg.max_rays = kw.pop("max_rays", 8192)
        for k, v in kw.items():
            setattr(cfg, k, v)
        e = nsb.Engine(cfg)
        e.set_model(grids, decs)
        if t_samples is None:
            tt, ts = O.t_tables()
            t_samples, t_surface = tt.numpy(), ts.numpy()
        e.set_ttables(t_samples, t_surface)
        for f in range(5):
            e.set_frame(f, depths[f], colors[f], poses[f])
        made.append(e)
        return e

    yield make
    for e in made:
        e.close()


def filtered_rays(syn, frames, n, seed, f=0):
    depths, colors, poses = frames
    idx = syn.mt19937_indices(seed, n, CAM["H"] * CAM["W"])
    ro, rd, gd, gc = O.ray_sampler(0, CAM["H"], 0, CAM["W"], idx, CAM["fx"], CAM["fy"], CAM["cx"], CAM["cy"],
                                   torch.tensor(depths[f]), torch.tensor(colors[f]), torch.tensor(poses[f]), "reference")
    m = O.inside_mask(ro, rd, gd, torch.tensor(O.BOUND)).numpy()
    return ro.numpy()[m], rd.numpy()[m], gd.numpy()[m], gc.numpy()[m]


# ------------------------------------------------------------------------------------------------ sampling
def test_get_samples_bit_exact_vs_reference(engine_factory, frames):
    """get_samples / raySampler (utils.h:13-55,141-146) against oracle/_ref's output: indices, rays, gt gathers."""
    g = load_golden("sampling.npz")
    e = engine_factory(raydir=0)        # utils.h:44-47 as written: the golden rays come from the reference's own raySampler
    ro, rd, gd, gc, ins, idx = e.get_samples(int(g["frame"]), int(g["H0"]), int(g["H1"]), int(g["W0"]), int(g["W1"]), 256, idx=g["idx"])
    assert np.array_equal(ro, g["rays_o"]) and np.array_equal(rd, g["rays_d"])
    assert np.array_equal(gd, g["gt_depth"]) and np.array_equal(gc, g["gt_color"])
    want = O.inside_mask(torch.tensor(g["rays_o"]), torch.tensor(g["rays_d"]), torch.tensor(g["gt_depth"]), torch.tensor(O.BOUND)).numpy()
    assert np.array_equal(ins, want)                                     # Mapper.cpp:416-427
    e.seed(int(g["seed"]))                                               # utils.h:32: the library's own mt19937 stream
    *_, idx2 = e.get_samples(int(g["frame"]), int(g["H0"]), int(g["H1"]), int(g["W0"]), int(g["W1"]), 256)
    assert np.array_equal(idx2, g["idx"])


def test_get_samples_pinhole_mode(engine_factory, frames, syn, nsb):
    e = engine_factory(raydir=1)
    depths, colors, poses = frames
    idx = syn.mt19937_indices(5, 512, 440 * 600)
    ro, rd, gd, gc, ins, _ = e.get_samples(3, 20, 460, 20, 620, 512, idx=idx)
    a, b, c_, d = O.ray_sampler(20, 460, 20, 620, idx, 360.0, 360.0, 320.0, 240.0, torch.tensor(depths[3]), torch.tensor(colors[3]), torch.tensor(poses[3]), "pinhole")
    assert np.array_equal(rd, b.numpy()) and np.array_equal(ro, a.numpy()) and np.array_equal(gd, c_.numpy())


# ------------------------------------------------------------------------------------------------- forward
def test_render_forward_vs_reference_golden(engine_factory):
    """Renderer::render_batch_ray (Renderer.cpp:44-125) against oracle/_ref: all stages + no-depth path."""
    g = load_golden("render_forward.npz")
    e = engine_factory(g["t_samples"], g["t_surface"])
    for st in ("color", "fine", "middle", "coarse"):
        o = e.render_batch_ray(g["rays_d"], g["rays_o"], st, g["gt_depth"])
        for nm, x in zip(("rgb", "depth", "var", "weights"), o):
            assert relerr(x, g["%s_%s" % (st, nm)]) < FWD_TOL, (st, nm)
    o = e.render_batch_ray(g["rays_d"], g["rays_o"], "coarse", None)
    for nm, x in zip(("rgb", "depth", "var", "weights"), o):
        assert relerr(x, g["nodepth_coarse_%s" % nm]) < FWD_TOL, nm
    for st in ("color", "coarse"):                                       # Renderer::eval_points (Renderer.cpp:19-42)
        raw = e.eval_points(g["pts"], st)
        ref = g["eval_%s" % st]
        assert np.abs(raw - ref).max() < FWD_TOL * np.abs(ref[:, :3]).max() + 1e-4, st
        assert np.array_equal(raw[:, 3] == 100.0, ref[:, 3] == 100.0)    # out-of-bound mask, Renderer.cpp:26-36


def test_render_forward_vs_verbatim_reference(engine_factory):
    """dist_norm = REFERENCE reproduces the UNPATCHED reference binary (utils.h:153 literal p=-1 norm)."""
    g = load_golden("render_forward.npz")
    e = engine_factory(g["t_samples"], g["t_surface"], dist_norm=1)
    o = e.render_batch_ray(g["rays_d"], g["rays_o"], "color", g["gt_depth"])
    for nm, x in zip(("rgb", "depth", "var", "weights"), o):
        ref = g["verbatim_%s" % nm]
        assert np.abs(x - ref).max() < FWD_TOL * max(1.0, np.abs(ref).max()), nm


def test_sample_placement_bit_exact(engine_factory, frames, syn):
    """z values of Renderer.cpp:61-119 (stratified + near-surface + sort), incl. zero-depth rays, vs the oracle."""
    e = engine_factory()
    ro, rd, gd, gc = filtered_rays(syn, frames, 1500, 13, f=2)
    assert (gd == 0).sum() > 0
    e.render_batch_ray(rd, ro, "middle", gd, want_weights=False)
    z = e.last_zvals(rd.shape[0])
    tt, ts = O.t_tables()
    zo = O.z_values(torch.tensor(ro), torch.tensor(rd), torch.tensor(gd), torch.tensor(O.BOUND), tt, ts).numpy()
    assert np.array_equal(z, zo)
    assert np.all(np.diff(z, axis=1) >= 0)


def test_render_forward_vs_oracle_live(engine_factory, frames, syn, model_inputs):
    grids, decs, _ = model_inputs
    e = engine_factory()
    m = O.Model(grids, decs)
    ro, rd, gd, gc = filtered_rays(syn, frames, 700, 3, f=4)
    with torch.no_grad():
        ref = O.render_batch_ray(m, torch.tensor(rd), torch.tensor(ro), "color", torch.tensor(gd))
    got = e.render_batch_ray(rd, ro, "color", gd)
    for nm, x, y in zip(("rgb", "depth", "var", "weights"), got, ref):
        assert relerr(x, y.numpy()) < FWD_TOL, nm


def test_occupancy_branch(engine_factory, frames, syn, model_inputs):
    """Upstream's occupancy branch alpha = sigmoid(10 raw), behind the flag."""
    grids, decs, _ = model_inputs
    e = engine_factory(occupancy=1)
    m = O.Model(grids, decs)
    ro, rd, gd, gc = filtered_rays(syn, frames, 200, 4)
    with torch.no_grad():
        ref = O.render_batch_ray(m, torch.tensor(rd), torch.tensor(ro), "color", torch.tensor(gd), occupancy=True)
    got = e.render_batch_ray(rd, ro, "color", gd)
    for nm, x, y in zip(("rgb", "depth", "var", "weights"), got, ref):
        assert relerr(x, y.numpy()) < 5e-4, nm   # sigmoid(10 x) amplifies the fp32 noise of raw tenfold


# ------------------------------------------------------------------------------------------------ backward
def test_render_vjp_vs_reference_golden(engine_factory, nsb):
    """loss.backward() through render_batch_ray against oracle/_ref's libtorch autograd."""
    g = load_golden("render_vjp.npz")
    e = engine_factory(g["t_samples"], g["t_surface"])
    got = e.render_vjp(g["rays_d"], g["rays_o"], "color", g["gt_depth"], g["g_rgb"], g["g_depth"], g["g_var"])
    assert relerr(got["rays_o"], g["d_rays_o"]) < GRAD_TOL and relerr(got["rays_d"], g["d_rays_d"]) < GRAD_TOL
    assert relerr(got["dec_color"], g["d_dec_color"]) < GRAD_TOL
    for lv in ("middle", "fine", "color"):
        full = got["grid_" + lv]
        assert np.abs(full.reshape(-1)[g["grid_%s_pos" % lv]] - g["grid_%s_val" % lv]).max() < GRAD_TOL * float(g["grid_%s_max" % lv]), lv
        l2 = np.sqrt((full.astype(np.float64) ** 2).sum())
        assert abs(l2 - float(g["grid_%s_l2" % lv])) < GRAD_TOL * float(g["grid_%s_l2" % lv]), lv   # nothing scattered elsewhere
    for flags in (nsb.F_GRID, nsb.F_RAY, nsb.F_GRID | nsb.F_WGRAD, nsb.F_GRID | nsb.F_RAY):   # the other kernel instantiations
        part = e.render_vjp(g["rays_d"], g["rays_o"], "color", g["gt_depth"], g["g_rgb"], g["g_depth"], g["g_var"], flags=flags)
        if flags & nsb.F_GRID:
            assert relerr(part["grid_fine"], got["grid_fine"]) < 1e-5
        if flags & nsb.F_RAY:
            assert relerr(part["rays_d"], got["rays_d"]) < 1e-5
        if flags & nsb.F_WGRAD:
            assert relerr(part["dec_color"], got["dec_color"]) < 1e-5


def test_render_vjp_middle_and_fine_stage(engine_factory, frames, syn, model_inputs):
    grids, decs, _ = model_inputs
    e = engine_factory()
    ro, rd, gd, gc = filtered_rays(syn, frames, 150, 8)
    n = ro.shape[0]
    rs = np.random.RandomState(2)
    g_rgb = np.zeros((n, 3), np.float32); g_depth = rs.randn(n).astype(np.float32); g_var = (0.2 * rs.randn(n)).astype(np.float32)
    for stage in ("middle", "fine"):
        m = O.Model(grids, decs)
        for k in ("middle", "fine"):
            m.grids[k].requires_grad_(True)
        tro = torch.tensor(ro, requires_grad=True); trd = torch.tensor(rd, requires_grad=True)
        rgb, depth, var, _ = O.render_batch_ray(m, trd, tro, stage, torch.tensor(gd))
        ((depth * torch.tensor(g_depth)).sum() + (var * torch.tensor(g_var)).sum()).backward()
        got = e.render_vjp(rd, ro, stage, gd, g_rgb, g_depth, g_var, flags=1 | 4)
        assert relerr(got["grid_middle"], m.grids["middle"].grad.numpy()) < GRAD_TOL
        if stage == "fine":
            assert relerr(got["grid_fine"], m.grids["fine"].grad.numpy()) < GRAD_TOL
        else:
            assert np.abs(got["grid_fine"]).max() == 0.0
        assert relerr(got["rays_o"], tro.grad.numpy()) < GRAD_TOL and relerr(got["rays_d"], trd.grad.numpy()) < GRAD_TOL


def test_stash_free_weight_gradient_kernel(nsb, model_inputs, frames, monkeypatch):
    """NSB_WGRAD_STASH=0: the colour-decoder weight gradient from k_wgrad_fused (per-tile recomputation, operands exchanged through
    shared memory, no HBM stash) against oracle/_ref's libtorch autograd -- the same golden vectors as the default (stash) path:
    the vjp golden and the whole-iteration gradient golden -- and against the default path itself."""
    grids, decs, _ = model_inputs
    depths, colors, poses = frames
    g = load_golden("render_vjp.npz"); gm = load_golden("mapping_iters.npz")
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("NSB_WGRAD_STASH", mode)
        cfg = nsb.default_config(); cfg.max_rays = 8192; cfg.mapping_pixels = int(gm["pixels"]); cfg.frustum_feature_selection = 0; cfg.raydir = 0
        e = nsb.Engine(cfg)
        e.set_model(grids, decs); e.set_ttables(g["t_samples"], g["t_surface"])
        for f in range(2):
            e.set_frame(f, depths[f], colors[f], poses[f])
        got = e.render_vjp(g["rays_d"], g["rays_o"], "color", g["gt_depth"], g["g_rgb"], g["g_depth"], g["g_var"])
        assert relerr(got["dec_color"], g["d_dec_color"]) < GRAD_TOL, mode
        only = e.render_vjp(g["rays_d"], g["rays_o"], "color", g["gt_depth"], g["g_rgb"], g["g_depth"], g["g_var"], flags=nsb.F_WGRAD)
        assert relerr(only["dec_color"], got["dec_color"]) < 1e-5, mode
        e.mapping_capture_grads(True)
        e.seed(int(gm["c0_seed"]))
        e.mapping_begin(list(range(int(gm["n_frames"]))), 60, 1.0)
        loss = e.mapping_iter(59)
        cg = e.captured_grads()
        assert np.allclose(loss, gm["c0_loss"], rtol=1e-4)
        assert relerr(cg["dec_color"], gm["c0_grad_dec_color"]) < GRAD_TOL, mode
        res[mode] = (got["dec_color"], cg["dec_color"], e.get_decoder("color"))
        e.close()
    assert relerr(res["0"][0], res["1"][0]) < 3e-4 and relerr(res["0"][1], res["1"][1]) < 3e-4      # G operand is single fp16 in the fused kernel
    assert np.abs(res["0"][2] - decs["color"]).max() > 1e-4                                             # the decoder was stepped


def test_kernel_variants_agree(nsb, model_inputs, frames, monkeypatch):
    """The tcgen05 forward (default), the warp-MMA forward (NSB_TCGEN05=0), the tcgen05 data-gradient kernel (NSB_BWD_T5=1), the
    forward without ray compaction (NSB_COMPACT_RAYS=0) and the two warp-MMA stash forwards of the colour iteration (NSB_T5_STASH=0, with
    and without the SM split) are the same
    arithmetic up to fp32 rounding: the same render outputs and the same whole-iteration gradients in a geometry and in a colour
    iteration, which in turn match oracle/_ref's libtorch autograd (golden) on the colour-decoder gradient."""
    grids, decs, _ = model_inputs
    depths, colors, poses = frames
    gm = load_golden("mapping_iters.npz"); gv = load_golden("render_vjp.npz")
    variants = {"default": {}, "warp_mma_fwd": {"NSB_TCGEN05": "0"}, "t5_bwd": {"NSB_BWD_T5": "1"}, "no_compaction": {"NSB_COMPACT_RAYS": "0"},
                "split_color_fwd": {"NSB_T5_STASH": "0"}, "one_launch_warp_mma_color_fwd": {"NSB_T5_STASH": "0", "NSB_SPLIT_COLOR_SMS": "0"}}
    knobs = ("NSB_TCGEN05", "NSB_BWD_T5", "NSB_COMPACT_RAYS", "NSB_SPLIT_COLOR_SMS", "NSB_T5_STASH")
    res = {}
    for name, env in variants.items():
        for k in knobs:
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        cfg = nsb.default_config(); cfg.max_rays = 8192; cfg.mapping_pixels = int(gm["pixels"]); cfg.frustum_feature_selection = 0; cfg.raydir = 0
        e = nsb.Engine(cfg)
        e.set_model(grids, decs); e.set_ttables(gv["t_samples"], gv["t_surface"])
        for f in range(int(gm["n_frames"])):
            e.set_frame(f, depths[f], colors[f], poses[f])
        rgb, depth, var, w = e.render_batch_ray(gv["rays_d"], gv["rays_o"], "color", gv["gt_depth"])
        out = {"rgb": rgb, "depth": depth, "var": var}
        e.mapping_capture_grads(True)
        for tag, it, seed in (("geom", 0, int(gm["seed"])), ("color", 59, int(gm["c0_seed"]))):
            e.set_model(grids, decs)
            e.seed(seed)
            e.mapping_begin(list(range(int(gm["n_frames"]))), 60, 1.0)
            out[tag + "_loss"] = np.float64(e.mapping_iter(it))
            cg = e.captured_grads()
            for lv in ("middle", "fine") + (("color",) if tag == "color" else ()):
                out[tag + "_grid_" + lv] = cg["grid_" + lv]
            if tag == "color":
                assert relerr(cg["dec_color"], gm["c0_grad_dec_color"]) < GRAD_TOL, name
                assert np.allclose(out["color_loss"], gm["c0_loss"], rtol=1e-4), name
                out["color_dec"] = cg["dec_color"]
        res[name] = out
        e.close()
    for k in knobs:
        monkeypatch.delenv(k, raising=False)
    ref = res["default"]
    for name, out in res.items():
        for k, v in out.items():
            assert relerr(np.asarray(v, np.float64), np.asarray(ref[k], np.float64)) < 2e-5, (name, k)


# --------------------------------------------------------------------------------- mapping / tracking loops
def test_mapping_iterations_vs_reference(engine_factory, frames, syn, model_inputs):
    """Mapper.cpp:330-465: sampling stream, inside filter, render, loss, backward, fused Adam -- four iterations
    (two geometry, two colour) against oracle/_ref's libtorch autograd + torch::optim::Adam."""
    grids, decs, _ = model_inputs
    g = load_golden("mapping_iters.npz")
    e = engine_factory(g["t_samples"], g["t_surface"], mapping_pixels=int(g["pixels"]), frustum_feature_selection=0, raydir=0)
    it_of = {1: 0, 3: 59}          # iteration numbers of a 60-iteration schedule that select these stages
    e.mapping_capture_grads(True)

    def check_grads(tag, got):
        """the gradient loss.backward() leaves at Mapper.cpp:444, against libtorch autograd at the golden's sampled positions"""
        for lv in ("middle", "fine", "color"):
            full = got["grid_" + lv]
            gmax = float(g["%s_grad_%s_max" % (tag, lv)])
            if gmax == 0.0:
                assert np.abs(full).max() == 0.0, (tag, lv)          # geometry iteration: exact zeros in the colour grid
                continue
            assert np.abs(full.reshape(-1)[g["%s_grad_%s_pos" % (tag, lv)]] - g["%s_grad_%s_val" % (tag, lv)]).max() < GRAD_TOL * gmax, (tag, lv)
            l2 = np.sqrt((full.astype(np.float64) ** 2).sum()); ref_l2 = float(g["%s_grad_%s_l2" % (tag, lv)])
            assert abs(l2 - ref_l2) < GRAD_TOL * ref_l2, (tag, lv)
        ref_dec = g["%s_grad_dec_color" % tag]
        if np.abs(ref_dec).max() == 0.0:
            assert np.abs(got["dec_color"]).max() == 0.0, tag
        else:
            assert relerr(got["dec_color"], ref_dec) < GRAD_TOL, tag

    # a colour-stage iteration from the initial parameters: colour grid + colour-decoder weight gradients at iteration 0
    e.seed(int(g["c0_seed"]))
    e.mapping_begin(list(range(int(g["n_frames"]))), 60, 1.0)
    l0 = e.mapping_iter(59)
    assert np.allclose(l0, g["c0_loss"], rtol=1e-4), (l0, g["c0_loss"])
    check_grads("c0", e.captured_grads())
    e.set_model(grids, decs)                                             # back to the initial map
    e.seed(int(g["seed"]))
    e.mapping_begin(list(range(int(g["n_frames"]))), 60, 1.0)
    losses = []
    for k, s in enumerate(g["stages"]):
        losses.append(e.mapping_iter(it_of[int(s)]))
        if k == 0:
            check_grads("it0", e.captured_grads())
    e.mapping_capture_grads(False)
    assert np.allclose(losses, g["losses"], rtol=1e-3), (losses, g["losses"])
    dec = e.get_decoder("color")
    moved = np.abs(g["dec_color"] - decs["color"]).max()
    assert moved > 1e-3 and np.abs(dec - g["dec_color"]).max() < 2e-2 * moved
    for lv in ("middle", "fine", "color"):
        pos = g["grid_%s_pos" % lv]
        got = e.get_grid(lv).reshape(-1)[pos]; ref = g["grid_%s_val" % lv]; init = grids[lv].reshape(-1)[pos]
        move_rms = np.sqrt(((ref - init) ** 2).mean())
        # Adam normalises by sqrt(v): voxels whose gradient is at the fp32 noise level take +-lr steps of random
        # sign in ANY implementation, so the comparison is in RMS relative to the RMS update
        assert move_rms > 0 and np.sqrt(((got - ref) ** 2).mean()) < 5e-2 * move_rms, lv


def test_tracking_iterations_vs_reference(engine_factory):
    """Tracker.cpp:41-113: pose -> rays -> render -> median mask -> loss -> pose gradient -> Adam on 7 floats."""
    t = load_golden("tracking_iters.npz")
    e = engine_factory(t["t_samples"], t["t_surface"], tracking_pixels=int(t["pixels"]), tracking_lr=float(t["lr"]), raydir=0)
    e.seed(int(t["seed"]))
    e.tracking_begin(0, t["cam7_in"])
    losses, g0 = [], None
    for i in range(3):
        l, g = e.tracking_iter()
        losses.append(l)
        if i == 0:
            g0 = g
    assert np.allclose(losses, t["losses"], rtol=1e-3)
    assert relerr(g0, t["grad_first"]) < GRAD_TOL
    assert np.abs(e.tracking_camera() - t["cam7_out"]).max() < 1e-5


# ------------------------------------------------------------------ BASELINE.json batch size (5000 rays x 48 samples) vs the oracle
def test_full_size_vs_oracle(engine_factory, frames, syn, model_inputs):
    """BASELINE configs[0..1] at full size (5000 rays x 48 samples): render_batch_ray forward (1e-4) and the loss + every gradient of
    whole mapping iterations -- sampling -> filter -> render -> loss -> backward, Mapper.cpp:376-444, geometry and colour stage --
    (1e-3) against the autograd oracle on the same pixels.
    The vjp for RANDOM cotangents is checked differently: with random signs the per-voxel sums cancel and the fp32 reference itself
    is only within ~1e-2 of an fp64 evaluation (1 - exp(-x) at small x, measured with tools/diag_fullsize.py), so two fp32
    implementations cannot agree to 1e-3 there; the CUDA path must be as close to the fp64 truth as the fp32 reference is."""
    grids, decs, _ = model_inputs
    depths, colors, poses = frames
    e = engine_factory(max_rays=8192, mapping_pixels=5000, frustum_feature_selection=0)
    ro, rd, gd, gc = filtered_rays(syn, frames, 5800, 91, f=3)
    ro, rd, gd = ro[:5000], rd[:5000], gd[:5000]
    assert ro.shape[0] == 5000
    n = 5000
    rs = np.random.RandomState(4)
    g_rgb = rs.randn(n, 3).astype(np.float32); g_depth = rs.randn(n).astype(np.float32); g_var = (0.3 * rs.randn(n)).astype(np.float32)
    res = {}
    for dt in (torch.float32, torch.float64):
        m = O.Model(grids, decs, dtype=dt)
        for k in ("middle", "fine", "color"):
            m.grids[k].requires_grad_(True)
        m.flat["color"].requires_grad_(True)
        tro = torch.tensor(ro).to(dt).requires_grad_(True); trd = torch.tensor(rd).to(dt).requires_grad_(True)
        tt, ts = O.t_tables(dt)
        ref = O.render_batch_ray(m, trd, tro, "color", torch.tensor(gd).to(dt), tt, ts)
        if dt == torch.float32:
            got = e.render_batch_ray(rd, ro, "color", gd)
            for nm, x, y in zip(("rgb", "depth", "var", "weights"), got, ref):
                assert relerr(x, y.detach().numpy()) < FWD_TOL, nm
        ((ref[0] * torch.tensor(g_rgb).to(dt)).sum() + (ref[1] * torch.tensor(g_depth).to(dt)).sum() + (ref[2] * torch.tensor(g_var).to(dt)).sum()).backward()
        res[dt] = {"grid_middle": m.grids["middle"].grad.numpy(), "grid_fine": m.grids["fine"].grad.numpy(), "grid_color": m.grids["color"].grad.numpy(),
                   "dec_color": m.flat["color"].grad.numpy(), "rays_o": tro.grad.numpy(), "rays_d": trd.grad.numpy()}
    vjp = e.render_vjp(rd, ro, "color", gd, g_rgb, g_depth, g_var)
    for k in ("grid_middle", "grid_fine", "grid_color", "dec_color", "rays_o", "rays_d"):
        err_gpu, err_ref = relerr(vjp[k], res[torch.float64][k]), relerr(res[torch.float32][k], res[torch.float64][k])
        assert err_gpu < 1.5 * err_ref + 1e-4, (k, err_gpu, err_ref)
        assert relerr(vjp[k], res[torch.float32][k]) < 1e-2, k
    # whole mapping iterations on 5 frames x 1000 pixels: loss and every gradient before the optimiser step
    e.mapping_capture_grads(True)
    for stage, it in (("middle", 0), ("color", 59)):
        m2 = O.Model(grids, decs)
        go = {}
        ref_losses, _ = O.mapping_iters(m2, depths[:5], colors[:5], poses[:5], syn.CAM, 5000, [stage], seed=17, raydir="pinhole", grads_out=go)
        e.set_model(grids, decs)
        e.seed(17)
        e.mapping_begin(list(range(5)), 60, 1.0)
        loss = e.mapping_iter(it)
        cg = e.captured_grads()
        assert abs(loss - ref_losses[0]) < 1e-4 * abs(ref_losses[0]), (stage, loss, ref_losses)
        for lv in ("middle", "fine", "color"):
            if go[lv] is None or float(go[lv].abs().max()) == 0.0:
                assert np.abs(cg["grid_" + lv]).max() == 0.0, (stage, lv)
            else:
                assert relerr(cg["grid_" + lv], go[lv].numpy()) < GRAD_TOL, (stage, lv)
        if stage == "color":
            assert relerr(cg["dec_color"], go["dec_color"].numpy()) < GRAD_TOL
        else:
            assert np.abs(cg["dec_color"]).max() == 0.0
    e.mapping_capture_grads(False)


def test_graph_replay_equals_eager_launches(nsb, model_inputs, frames, monkeypatch):
    """The captured-graph iteration (one cudaGraphLaunch, device-resident iteration state) and the kernel-by-kernel path
    (NSB_GRAPH=0) are the same arithmetic: the same losses and parameters after a mixed schedule over two optimize_map calls."""
    grids, decs, _ = model_inputs
    depths, colors, poses = frames
    outs = []
    for graph in ("1", "0"):
        monkeypatch.setenv("NSB_GRAPH", graph)
        cfg = nsb.default_config(); cfg.mapping_pixels = 1500; cfg.max_rays = 2048; cfg.frustum_feature_selection = 0
        e = nsb.Engine(cfg)
        e.set_model(grids, decs)
        for f in range(3):
            e.set_frame(f, depths[f], colors[f], poses[f])
        e.seed(9)
        losses = []
        for rep_ in range(2):                                   # two optimize_map calls: the second reuses the captured graphs
            e.mapping_begin([0, 1, 2], 60, 1.0)
            for it in (0, 1, 30, 58, 59):
                e.mapping_iter(it, sync=False)
            l, k = e.mapping_losses(0, 5)
            losses.append(l); assert (k > 0).all()
        outs.append((np.concatenate(losses), {lv: e.get_grid(lv) for lv in ("middle", "fine", "color")}, e.get_decoder("color")))
        e.close()
    # same kernels, same launch parameters: the two runs differ only by the order of the fp32 atomics (loss and grid reductions)
    assert np.allclose(outs[0][0], outs[1][0], rtol=2e-5), (outs[0][0], outs[1][0])
    for lv in ("middle", "fine", "color"):
        move = np.sqrt(((outs[1][1][lv] - grids[lv]) ** 2).mean())
        assert move > 0 and np.sqrt(((outs[0][1][lv] - outs[1][1][lv]) ** 2).mean()) < 2e-2 * move, lv     # Adam sign noise on ~zero gradients
    dmove = np.sqrt(((outs[1][2] - decs["color"]) ** 2).mean())
    assert dmove > 0 and np.sqrt(((outs[0][2] - outs[1][2]) ** 2).mean()) < 5e-2 * dmove


# ------------------------------------------------------------------ properties at the BASELINE.json batch size
def test_full_size_properties(engine_factory, frames, syn, nsb):
    e = engine_factory(max_rays=4096)      # 5000 rays > max_rays: also exercises the chunked render path
    ro, rd, gd, gc = filtered_rays(syn, frames, 5700, 77, f=1)
    ro, rd, gd = ro[:5000], rd[:5000], gd[:5000]
    assert ro.shape[0] == 5000
    rgb, depth, var, w = e.render_batch_ray(rd, ro, "color", gd)
    z = None
    assert np.all(np.isfinite(rgb)) and np.all(np.isfinite(depth)) and np.all(var >= -1e-6)
    assert np.all(w >= 0) and np.all(w.sum(1) <= 1 + 1e-5)
    # determinism and permutation equivariance (the batch statistics max(gt_depth) are order independent)
    rgb2, depth2, var2, w2 = e.render_batch_ray(rd, ro, "color", gd)
    assert np.array_equal(depth, depth2) and np.array_equal(w, w2)
    perm = np.random.RandomState(0).permutation(5000)
    rgb3, depth3, var3, w3 = e.render_batch_ray(rd[perm], ro[perm], "color", gd[perm])
    assert np.array_equal(depth3, depth[perm]) and np.array_equal(rgb3, rgb[perm])
    # one call == two half calls when the halves share the batch maximum
    big = engine_factory(max_rays=8192)
    rgb4, depth4, _, _ = big.render_batch_ray(rd, ro, "color", gd)
    assert np.array_equal(depth4, depth) and np.array_equal(rgb4, rgb)
    # linearity of the vjp in the cotangent, and zero cotangent -> zero gradient
    rs = np.random.RandomState(1)
    g_rgb = rs.randn(5000, 3).astype(np.float32); g_depth = rs.randn(5000).astype(np.float32); g_var = np.zeros(5000, np.float32)
    g1 = big.render_vjp(rd, ro, "color", gd, g_rgb, g_depth, g_var)
    g2 = big.render_vjp(rd, ro, "color", gd, 2 * g_rgb, 2 * g_depth, g_var)
    for k in ("grid_middle", "grid_fine", "grid_color", "dec_color", "rays_o"):
        assert relerr(g2[k], 2 * g1[k]) < 1e-4, k
    g0 = big.render_vjp(rd, ro, "color", gd, 0 * g_rgb, 0 * g_depth, g_var)
    assert all(np.abs(g0[k]).max() == 0.0 for k in ("grid_middle", "grid_fine", "grid_color", "dec_color"))


def test_voxel_mask_limits_adam(engine_factory, model_inputs):
    """frustum_feature_selection intent (Mapper.cpp:260-290,333-350): only masked voxels are optimised."""
    grids, decs, _ = model_inputs
    e = engine_factory(mapping_pixels=1000, frustum_feature_selection=0)   # explicit masks instead of the computed frustum
    masks = {}
    rs = np.random.RandomState(0)
    for lv in ("middle", "fine", "color"):
        masks[lv] = (rs.uniform(size=grids[lv].shape[2:]) < 0.5).astype(np.uint8)
        e.set_voxel_mask(lv, masks[lv])
    e.seed(1)
    e.mapping_begin([0, 1], 60, 1.0)
    for it in (0, 59):
        e.mapping_iter(it)
    for lv in ("middle", "fine", "color"):
        d = np.abs(e.get_grid(lv) - grids[lv]).max(axis=(0, 1))
        assert d[masks[lv] == 0].max() == 0.0 and d[masks[lv] == 1].max() > 0.0


# ------------------------------------------------------------------ frustum voxel mask on the GPU
def test_frustum_mask_vs_upstream_semantics(engine_factory, frames, syn):
    """Mapper::get_mask_from_c2w (Mapper.cpp:42-130) restated with upstream's semantics + cv2.remap (oracle) vs the GPU
    kernels.  A voxel whose projection lands within float rounding of a threshold may flip: <= 0.2 % mismatches."""
    depths, colors, poses = frames
    e = engine_factory()
    # a piecewise-smooth depth image makes the depth test non-trivial
    yy, xx = np.mgrid[0:480, 0:640].astype(np.float32)
    depth = (1.5 + 0.8 * np.sin(xx / 90.0) * np.cos(yy / 70.0)).astype(np.float32)
    depth[100:140, 200:260] = 0.0
    for slot, pose in ((0, poses[1]), (1, poses[3])):
        e.set_frame(slot, depth, colors[0], pose)
        for lv in ("middle", "fine", "color"):
            got = e.frustum_mask(slot, lv)
            ref = O.get_mask_from_c2w(pose, depth, syn.grid_dims(lv), CAM)
            assert got.shape == ref.shape and ref.sum() > 50
            assert (got != ref).sum() <= max(2, int(0.002 * ref.size)), (lv, int((got != ref).sum()), int(ref.sum()))
    assert e.frustum_mask(0, "coarse").all()                     # Mapper.cpp:54-59


def test_mapping_with_frustum_feature_selection(engine_factory, model_inputs, frames):
    """frustum_feature_selection = True (nice_slam.yaml:79): nsb_mapping_begin builds the masks from the current frame and
    Adam leaves every voxel outside them untouched (Mapper.cpp:260-290, 333-350, 448-464)."""
    grids, decs, _ = model_inputs
    e = engine_factory(mapping_pixels=1000, frustum_feature_selection=1)
    e.seed(2)
    e.mapping_begin([1, 0], 60, 1.0)          # the last slot is the current frame
    for it in (0, 59):
        e.mapping_iter(it)
    for lv in ("middle", "fine", "color"):
        m = e.frustum_mask(0, lv)
        d = np.abs(e.get_grid(lv) - grids[lv]).max(axis=(0, 1))
        assert d[~m].max() == 0.0 and d[m].max() > 0.0


def test_bundle_adjustment_vs_oracle(engine_factory, model_inputs, frames, syn):
    """Mapper.cpp:305-329,366-368,382-399,467-489: the poses of the non-oldest frames are optimised with the map as
    7-vectors.  Camera gradients of the first iteration, the loss trajectory, the optimised vectors and the pose
    write-back against the autograd oracle on the same pixel stream."""
    import torch
    import nice_oracle as O
    grids, decs, _ = model_inputs
    depths, colors, poses = frames
    nf, pix = 3, 600
    e = engine_factory(mapping_pixels=pix, frustum_feature_selection=0, BA_cam_lr=0.001)
    stages = ["middle", "color", "color"]
    it_of = {"middle": 0, "color": 59}
    e.seed(11)
    e.mapping_begin(list(range(nf)), 60, 1.0, ba_mask=0b110)      # frame 0 = the oldest keyframe stays fixed
    losses, g_first = [], None
    for st in stages:
        losses.append(e.mapping_iter(it_of[st]))
        if g_first is None:
            g_first = e.mapping_cam_grads()
    cams = e.mapping_end()
    m = O.Model(grids, decs)
    go, co = {}, []
    ref_losses, _ = O.mapping_iters(m, depths[:nf], colors[:nf], poses[:nf], syn.CAM, pix, stages, seed=11, ba_frames=[1, 2],
                                    ba_cam_lr=0.001, grads_out=go, cams_out=co, raydir="pinhole")     # the library's default directions
    assert np.allclose(losses, ref_losses, rtol=1e-3), (losses, ref_losses)
    assert np.all(g_first[0] == 0)
    for f in (1, 2):
        assert relerr(g_first[f], go["cam_%d" % f].numpy()) < GRAD_TOL, f
    ref_cams = co[0].numpy()
    init = np.stack([O.get_tensor_from_camera(poses[f]) for f in range(nf)])
    moved = np.abs(ref_cams - init).max()
    assert moved > 1e-3 and np.abs(cams - ref_cams).max() < 5e-2 * moved, (cams, ref_cams)
    assert np.array_equal(cams[0], init[0])
    for f in (1, 2):   # est_c2w <- get_camera_from_tensor(camera_tensor)
        assert np.abs(e.get_frame_pose(f) - O.get_camera_from_tensor(torch.tensor(cams[f])).numpy()).max() < 1e-6
    assert np.abs(e.get_frame_pose(0) - poses[0][:3]).max() == 0


def test_coarse_mapper_vs_oracle(engine_factory, model_inputs, frames, syn, nsb):
    """The coarse mapper (Mapper(ns, cf, coarse_mapper=true); Mapper.cpp:335-338,351-352,450-453, upstream's intent: stage "coarse"
    rendered without depth guidance, only grid_coarse in the optimiser): losses, the coarse-grid gradient of the first iteration
    (backward through MLP_no_xyz, k_coarse_bwd) and the optimised grid against the autograd oracle; every other parameter untouched."""
    grids, decs, _ = model_inputs
    depths, colors, poses = frames
    nf, pix = 2, 800
    e = engine_factory(mapping_pixels=pix, frustum_feature_selection=0)
    e.seed(23)
    e.mapping_capture_grads(True)
    e.mapping_begin(list(range(nf)), 60, 1.0, flags=nsb.MAP_COARSE)
    losses, g0 = [], None
    for it in (0, 30, 59):
        losses.append(e.mapping_iter(it))
        if g0 is None:
            g0 = e.captured_grads()
    e.mapping_capture_grads(False)
    m = O.Model(grids, decs)
    go = {}
    ref_losses, _ = O.mapping_iters(m, depths[:nf], colors[:nf], poses[:nf], syn.CAM, pix, ["coarse"] * 3, seed=23, raydir="pinhole", grads_out=go, coarse_mapper=True)
    assert np.allclose(losses, ref_losses, rtol=1e-3), (losses, ref_losses)
    assert relerr(g0["grid_coarse"], go["coarse"].numpy()) < GRAD_TOL
    for lv in ("middle", "fine", "color"):
        assert np.abs(g0["grid_" + lv]).max() == 0.0 and np.array_equal(e.get_grid(lv), grids[lv]), lv
    ref = m.grids["coarse"].numpy(); got = e.get_grid("coarse")
    move = np.sqrt(((ref - grids["coarse"]) ** 2).mean())
    assert move > 0 and np.sqrt(((got - ref) ** 2).mean()) < 5e-2 * move
    assert np.array_equal(e.get_decoder("color"), decs["color"])


def test_color_refine_settings(engine_factory, model_inputs, frames, syn, nsb):
    """color_refine (Mapper.cpp:505-513): middle_iter_ratio = fine_iter_ratio = 0 (iteration 0 is still "middle": it <= int(n * 0)),
    fix_color = true (the colour decoder stays fixed), frustum_feature_selection = false (installed masks are ignored)."""
    grids, decs, _ = model_inputs
    depths, colors, poses = frames
    nf, pix = 2, 600
    e = engine_factory(mapping_pixels=pix, frustum_feature_selection=1)
    for lv in ("middle", "fine", "color"):
        e.set_voxel_mask(lv, np.zeros(grids[lv].shape[2:], np.uint8))        # would freeze every voxel if it were honoured
    e.seed(31)
    e.mapping_begin(list(range(nf)), 60, 1.0, flags=nsb.MAP_COLOR_REFINE)
    losses = [e.mapping_iter(it) for it in (0, 1, 2)]
    m = O.Model(grids, decs)
    ref_losses, _ = O.mapping_iters(m, depths[:nf], colors[:nf], poses[:nf], syn.CAM, pix, ["middle", "color", "color"], seed=31, raydir="pinhole", fix_color=True)
    assert np.allclose(losses, ref_losses, rtol=1e-3), (losses, ref_losses)
    assert np.array_equal(e.get_decoder("color"), decs["color"])
    for lv in ("middle", "fine", "color"):
        assert np.abs(e.get_grid(lv) - grids[lv]).max() > 0, lv
    for lv in ("middle", "fine", "color"):
        e.set_voxel_mask(lv, None)


def test_eval_lattice_and_device_eval_points(engine_factory, model_inputs, syn):
    """eval_points over a mesh lattice generated on the device (stage assembly + bound mask in-kernel) against
    the oracle's Renderer::eval_points on the same numpy.meshgrid points, incl. lattice points outside the bound (occupancy 100);
    the device-pointer form agrees bit for bit with the host form."""
    grids, decs, _ = model_inputs
    e = engine_factory(max_rays=2048)
    model = O.Model(grids, decs)
    lo = np.array([-5.0, -1.6, -3.2], np.float32); hi = np.array([4.0, 2.1, 2.9], np.float32)     # slightly larger than the bound
    nx, ny, nz = 21, 13, 17
    occ, raw = e.eval_lattice("color", nx, ny, nz, lo, hi, want_raw=True)
    x = np.linspace(lo[0], hi[0], nx, dtype=np.float32); y = np.linspace(lo[1], hi[1], ny, dtype=np.float32); z = np.linspace(lo[2], hi[2], nz, dtype=np.float32)
    xx, yy, zz = np.meshgrid(x, y, z)
    pts = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], 1).astype(np.float32)
    with torch.no_grad():
        ref = model.eval_points(torch.tensor(pts), "color").numpy()
    out_ref = ref[:, 3] == 100.0
    got = raw.reshape(-1, 4)
    assert 0.05 < out_ref.mean() < 0.9
    edge = np.zeros(len(pts), bool)                      # lattice points within an ulp of a bound face may fall on either side
    bnd = np.asarray(O.BOUND, np.float32)
    for a in range(3):
        edge |= (np.abs(pts[:, a] - bnd[a, 0]) < 1e-5) | (np.abs(pts[:, a] - bnd[a, 1]) < 1e-5)
    assert np.array_equal((got[:, 3] == 100.0)[~edge], out_ref[~edge])
    ok = ~out_ref & ~edge
    assert np.abs(got[ok] - ref[ok]).max() < 1e-4 * max(1.0, np.abs(ref[ok]).max())
    assert np.array_equal(occ.reshape(-1), got[:, 3])
    raw_h = e.eval_points(pts, "color")
    d_pts = torch.tensor(pts).cuda(); d_raw = torch.empty(len(pts), 4, device="cuda")
    torch.cuda.synchronize()
    e._ck(e.lib.nsb_eval_points_dev(e.h, 3, len(pts), d_pts.data_ptr(), d_raw.data_ptr()))
    e.synchronize()
    assert np.array_equal(d_raw.cpu().numpy(), raw_h)
    assert np.abs(raw_h[ok] - ref[ok]).max() < 1e-4 * max(1.0, np.abs(ref[ok]).max())
    occ_c = e.eval_lattice("coarse", 9, 7, 5)            # default box = the scene bound: the faces themselves are out of bound
    assert occ_c.shape == (7, 9, 5) and (occ_c[0] == 100.0).all() and np.isfinite(occ_c).all()


def test_keyframe_selection_overlap(engine_factory, frames, syn):
    """Mapper.cpp:132-196: fraction of the current frame's 100 x 16 depth-guided vertices that each keyframe sees, and the
    resulting ranking, against the numpy restatement on the same pixel indices."""
    import nice_oracle as O
    depths, colors, poses = frames
    e = engine_factory()
    rs = np.random.RandomState(5)
    kfs = []
    for k in range(7):   # keyframes around the current pose: growing yaw and a sideways offset, one looking away
        yaw = np.deg2rad([0, 10, 25, 40, 60, 90, 180][k])
        R = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]], np.float32)
        m = np.eye(4, dtype=np.float32); m[:3, :3] = poses[0][:3, :3] @ R; m[:3, 3] = poses[0][:3, 3] + rs.uniform(-0.3, 0.3, 3).astype(np.float32)
        kfs.append(m)
    idx = syn.mt19937_indices(3, 100, 480 * 640)
    sel, pct = e.keyframe_selection_overlap(0, kfs, 3, idx=idx)
    _, ts = O.t_tables()
    ref_sel, ref_pct = O.keyframe_selection_overlap(depths[0], colors[0], poses[0], kfs, 3, syn.CAM, idx, raydir="pinhole", ts=ts.numpy())
    assert np.abs(pct - ref_pct).max() <= 2.0 / 1600 + 1e-7, (pct, ref_pct)      # a vertex exactly on the 20 px border may flip
    assert ref_pct.max() > 0.3 and (ref_pct == 0).any()
    if np.abs(pct - ref_pct).max() == 0:
        assert sel == ref_sel
    assert len(sel) == 3 and all(pct[sel[i]] >= pct[sel[i + 1]] for i in range(2))


def test_dense_render_img(engine_factory, frames, model_inputs, syn):
    """BASELINE configs[3]: every pixel of a 640x480 frame through render_batch_ray (upstream render_img).  The chunked dense
    render equals single render_batch_ray calls on the same rays (batch-global scalars taken over the whole image), a strip
    of it matches the oracle, and the no-depth coarse level runs over the full image."""
    import torch
    import nice_oracle as O
    grids, decs, _ = model_inputs
    depths, colors, poses = frames
    e = engine_factory(max_rays=65536)
    rgb, depth, var = e.render_img(0, "color", True)
    assert rgb.shape == (480, 640, 3) and np.isfinite(rgb).all() and np.isfinite(depth).all() and (var >= 0).all()
    # rows 200..201 against the oracle, with the whole-image maxima of Renderer.cpp:76,93 injected through an extra ray
    rows = np.arange(200 * 640, 202 * 640)
    ro, rd, gd, _ = O.ray_sampler(0, 480, 0, 640, rows, 360.0, 360.0, 320.0, 240.0, torch.tensor(depths[0]), torch.tensor(colors[0]), torch.tensor(poses[0]), "pinhole")
    imax = int(np.argmax(depths[0]))
    ro2, rd2, gd2, _ = O.ray_sampler(0, 480, 0, 640, np.array([imax]), 360.0, 360.0, 320.0, 240.0, torch.tensor(depths[0]), torch.tensor(colors[0]), torch.tensor(poses[0]), "pinhole")
    tt, ts = O.t_tables()
    with torch.no_grad():
        ref = O.render_batch_ray(O.Model(grids, decs), torch.cat([rd, rd2]), torch.cat([ro, ro2]), "color", torch.cat([gd, gd2]), tt, ts)
    assert relerr(depth.reshape(-1)[rows], ref[1].numpy()[:-1]) < FWD_TOL and relerr(rgb.reshape(-1, 3)[rows], ref[0].numpy()[:-1]) < FWD_TOL
    # the same rays through render_batch_ray in one call: identical bits
    got = e.render_batch_ray(torch.cat([rd, rd2]).numpy(), torch.cat([ro, ro2]).numpy(), "color", torch.cat([gd, gd2]).numpy(), want_weights=False)
    assert np.array_equal(got[1][:-1], depth.reshape(-1)[rows])
    # coarse level, no depth guidance (Renderer.cpp:54-58): 32 samples per ray over all 307 200 pixels
    _, dc, vc = e.render_img(0, "coarse", False)
    assert np.isfinite(dc).all() and dc.min() > 0 and (vc >= 0).all()


def test_edge_cases(engine_factory, frames, model_inputs, syn, nsb):
    """Empty, tiny and ragged batches, zero-depth pixels, axis-parallel rays, batches larger than max_rays (chunked with
    whole-call batch scalars), point counts that are not a multiple of the 16-sample tile, and the error behaviour."""
    import torch
    grids, decs, _ = model_inputs
    depths, colors, poses = frames
    e = engine_factory(max_rays=512)
    tt, ts = O.t_tables()
    model = O.Model(grids, decs)
    idx = syn.mt19937_indices(7, 1200, 480 * 640)
    ro, rd, gd, _ = O.ray_sampler(0, 480, 0, 640, idx, 360.0, 360.0, 320.0, 240.0, torch.tensor(depths[0]), torch.tensor(colors[0]), torch.tensor(poses[0]))
    keep = O.inside_mask(ro, rd, gd, torch.tensor(O.BOUND))
    ro, rd, gd = ro[keep].numpy(), rd[keep].numpy(), gd[keep].numpy().copy()
    gd[::7] = 0.0                                             # zero-depth pixels: the 0.001 .. max(depth) surface branch (Renderer.cpp:93-98)
    # empty batch: a no-op
    rgb, depth, var, w = e.render_batch_ray(rd[:0], ro[:0], "color", gd[:0])
    assert rgb.shape == (0, 3) and depth.shape == (0,) and w.shape == (0, 48)
    for n in (1, 17, 33, 700):                                # 700 > max_rays: three chunks, batch scalars over the whole call
        with torch.no_grad():
            ref = O.render_batch_ray(model, torch.tensor(rd[:n]), torch.tensor(ro[:n]), "color", torch.tensor(gd[:n]), tt, ts)
        rgb, depth, var, w = e.render_batch_ray(rd[:n], ro[:n], "color", gd[:n])
        assert relerr(depth, ref[1].numpy()) < FWD_TOL and relerr(rgb, ref[0].numpy()) < FWD_TOL and relerr(w, ref[3].numpy()) < FWD_TOL, n
    # axis-parallel rays: a zero direction component gives +-inf in the AABB test (Renderer.cpp:69-73); results stay finite
    o = np.tile(np.array([[-0.34, 0.26, -0.12]], np.float32), (16, 1)); d = np.zeros((16, 3), np.float32); d[:, 0] = 1.0
    g = np.linspace(0.5, 3.0, 16).astype(np.float32)
    with torch.no_grad():
        ref = O.render_batch_ray(model, torch.tensor(d), torch.tensor(o), "color", torch.tensor(g), tt, ts)
    rgb, depth, var, _ = e.render_batch_ray(d, o, "color", g)
    assert np.isfinite(depth).all() and relerr(depth, ref[1].numpy()) < FWD_TOL
    # eval_points with a point count that is not a multiple of the tile, incl. points outside the bound (occupancy 100)
    pts = np.random.RandomState(3).uniform(-5, 4, (37, 3)).astype(np.float32)
    raw = e.eval_points(pts, "color")
    with torch.no_grad():
        ref = model.eval_points(torch.tensor(pts), "color").numpy()
    bnd = np.asarray(O.BOUND, np.float32)
    inside = np.all((pts < bnd[:, 1]) & (pts > bnd[:, 0]), axis=1)
    assert (raw[~inside, 3] == 100).all() and np.abs(raw - ref).max() < 1e-3
    # error behaviour: status + message, never a crash
    with pytest.raises(RuntimeError, match="stage"):
        e._ck(e.lib.nsb_render_batch_ray(e.h, 9, 1, None, None, None, None, None, None, None))
    with pytest.raises(RuntimeError, match="exceeds max_rays"):
        n = 600
        e.render_vjp(rd[:n], ro[:n], "color", gd[:n], np.ones((n, 3), np.float32), np.ones(n, np.float32), np.zeros(n, np.float32))
    with pytest.raises(RuntimeError, match="slot"):
        e.set_frame(99, depths[0], colors[0], poses[0])


def test_async_frame_ingest_and_checkpoint(engine_factory, frames, model_inputs, nsb, tmp_path):
    """a frame uploaded on the copy stream is what the next sampling call sees (device-side wait), and a
    checkpoint written after some mapping iterations restores grids and decoders bit-exactly into a fresh context."""
    grids, decs, _ = model_inputs
    depths, colors, poses = frames
    e = engine_factory(mapping_pixels=500, frustum_feature_selection=0)
    idx = np.arange(0, 480 * 640, 4801)[:60]
    ref = e.get_samples(3, 0, 480, 0, 640, 60, idx=idx)
    e.set_frame_async(1, depths[3], colors[3], poses[3])                 # slot 1 <- frame 3, no host wait
    got = e.get_samples(1, 0, 480, 0, 640, 60, idx=idx)
    e.frames_ready()
    for a, b in zip(ref[:4], got[:4]):
        assert np.array_equal(a, b)
    e.seed(4)
    e.mapping_begin([0, 1], 60, 1.0)
    for it in (0, 59, 59):
        e.mapping_iter(it)
    path = tmp_path / "map.nsbckpt"
    e.save_checkpoint(path)
    e2 = engine_factory()
    e2.load_checkpoint(path)
    for lv in ("coarse", "middle", "fine", "color"):
        assert np.array_equal(e2.get_grid(lv), e.get_grid(lv)) and np.array_equal(e2.get_decoder(lv), e.get_decoder(lv))
    assert not np.array_equal(e.get_grid("color"), grids["color"]) and not np.array_equal(e.get_decoder("color"), decs["color"])
    rd, ro, gd = got[1], got[0], got[2]
    assert np.array_equal(e2.render_batch_ray(rd, ro, "color", gd)[1], e.render_batch_ray(rd, ro, "color", gd)[1])
    with pytest.raises(RuntimeError, match="not an nsb checkpoint"):
        bad = tmp_path / "bad.bin"; bad.write_bytes(b"x" * 400)
        e2.load_checkpoint(bad)
