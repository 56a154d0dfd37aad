"""The C++ host-side mirror of the reference API (include/nsb/nsb.hpp: Renderer, NICE, Mapper, Tracker over the C ABI).
CPU: it compiles with plain g++ against include/ and links libnsb.so, and fails loudly without a GPU.
GPU: the same program renders / maps / tracks and agrees with the ctypes path and with the oracle."""
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, relerr
import nice_oracle as O


def build_cpp(nsb, tmp):
    exe = os.path.join(tmp, "test_wrappers")
    libdir = os.path.dirname(nsb.lib_path())
    cmd = ["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_wrappers.cpp"),
           "-o", exe, "-L" + libdir, "-l:libnsb.so", "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True)
    return exe


def test_wrappers_compile_and_fail_loudly_without_gpu(nsb, tmp_path):
    exe = build_cpp(nsb, str(tmp_path))
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_wrappers_match_ctypes_and_oracle(nsb, syn, model_inputs, frames, tmp_path):
    grids, decs, _ = model_inputs
    depths, colors, poses = frames
    d = str(tmp_path)
    for lv in syn.LEVELS:
        grids[lv].tofile(os.path.join(d, "grid_%s.bin" % lv)); decs[lv].tofile(os.path.join(d, "dec_%s.bin" % lv))
    tt, ts = O.t_tables()
    tt.numpy().tofile(os.path.join(d, "t_samples.bin")); ts.numpy().tofile(os.path.join(d, "t_surface.bin"))
    idx = syn.mt19937_indices(2, 300, 480 * 640)
    ro, rd, gd, gc = O.ray_sampler(0, 480, 0, 640, idx, 360.0, 360.0, 320.0, 240.0, torch.tensor(depths[0]), torch.tensor(colors[0]), torch.tensor(poses[0]))
    m = O.inside_mask(ro, rd, gd, torch.tensor(O.BOUND)).numpy()
    ro, rd, gd = ro.numpy()[m], rd.numpy()[m], gd.numpy()[m]
    ro.tofile(os.path.join(d, "rays_o.bin")); rd.tofile(os.path.join(d, "rays_d.bin")); gd.tofile(os.path.join(d, "gt_depth.bin"))
    depths[0].tofile(os.path.join(d, "frame_depth.bin")); colors[0].tofile(os.path.join(d, "frame_color.bin")); poses[0].tofile(os.path.join(d, "c2w.bin"))
    exe = build_cpp(nsb, d)
    r = subprocess.run([exe, d], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    n = gd.shape[0]
    out = {k: np.fromfile(os.path.join(d, "out_%s.bin" % k), np.float32) for k in ("rgb", "depth", "var", "weights", "map_losses", "trk_losses", "cam", "grid_middle")}
    # Renderer::render_batch_ray through the C++ classes vs the oracle (1e-4) and vs the ctypes binding (same library: identical)
    model = O.Model(grids, decs)
    with torch.no_grad():
        ref = O.render_batch_ray(model, torch.tensor(rd), torch.tensor(ro), "color", torch.tensor(gd), tt, ts)
    assert relerr(out["rgb"].reshape(n, 3), ref[0].numpy()) < 1e-4 and relerr(out["depth"], ref[1].numpy()) < 1e-4
    assert relerr(out["weights"].reshape(n, 48), ref[3].numpy()) < 1e-4
    cfg = nsb.default_config(); cfg.mapping_pixels = 400; cfg.tracking_pixels = 300; cfg.tracking_lr = 1e-3; cfg.tracking_iters = 2; cfg.frustum_feature_selection = 0
    e = nsb.Engine(cfg); e.set_model(grids, decs); e.set_ttables(tt.numpy(), ts.numpy())
    got = e.render_batch_ray(rd, ro, "color", gd)
    assert np.array_equal(got[1], out["depth"]) and np.array_equal(got[0].reshape(-1), out["rgb"])
    # Mapper::optimize_map (3 iterations on the current frame) vs the oracle's Mapper.cpp:330-465 restatement
    l_or, _ = O.mapping_iters(O.Model(grids, decs), depths[:1], colors[:1], poses[:1], syn.CAM, 400, O.stage_schedule(3), seed=3, tt=tt, ts=ts, raydir="pinhole")
    assert np.allclose(out["map_losses"], l_or, rtol=1e-3)
    assert np.abs(out["grid_middle"] - grids["middle"].reshape(-1)).max() > 0.05      # the dict got the optimised grid back
    cm = np.fromfile(os.path.join(d, "out_coarse_mapper.bin"), np.float32)       # coarse mapper: finite losses, grid_coarse moved, grid_middle untouched
    assert np.all(np.isfinite(cm[:2])) and cm[:2].min() > 0 and 1e-4 < cm[2] < 0.05 and cm[3] == 0.0, cm
    assert len(out["trk_losses"]) == 2 and np.all(np.isfinite(out["cam"])) and abs(np.linalg.norm(out["cam"][:4]) - 1) < 1e-2
    # Mapper::run over 7 frames with keyframe_every = 1: seven keyframes, BA from the sixth frame on (keyframes > 4), and the
    # bundle-adjusted current pose is written back into estimate_c2w (Mapper.cpp:530-534)
    seq = np.fromfile(os.path.join(d, "out_run_seq.bin"), np.float32)
    assert seq[0] == 7 and list(seq[1:8]) == [0, 0, 0, 0, 0, 1, 1]
    assert np.all(seq[8:13] == 0) and np.all(seq[13:15] > 1e-4) and np.all(seq[13:15] < 0.1)
    e.close()
