"""Two-GPU mapping: rays sharded over the ranks, gradients summed with NCCL (grids under the wgrad kernel in
colour iterations, the [loss | middle | fine] prefix in geometry iterations), identical Adam on every rank.  The result
must equal the one-GPU run on the same global batch up to the association of the fp32 sums.  Needs 2 GPUs (skipped else)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _run(rank, world, uid, q, ba):
    sys.path.insert(0, ROOT)
    nsb = importlib.import_module("nice-slam-cpp_b200")
    syn = nsb.synthetic
    grids = syn.make_grids(0); decs = syn.make_decoders(0, bias_scale=0.05)
    depths, colors, poses = syn.make_frames(2, 0)
    cfg = nsb.default_config(); cfg.mapping_pixels = 2000; cfg.max_rays = 2000; cfg.frustum_feature_selection = 0; cfg.BA_cam_lr = 0.001
    e = nsb.Engine(cfg, device=rank)
    e.set_model(grids, decs)
    for f in range(2):
        e.set_frame(f, depths[f], colors[f], poses[f])
    if world > 1:
        e.comm_init(uid, rank, world)
    e.seed(21)
    e.mapping_begin([0, 1], 60, 1.0, ba_mask=0b10 if ba else 0)
    losses = [e.mapping_iter(it) for it in (0, 30, 59, 59)]
    cams = e.mapping_end()
    out = {"losses": np.array(losses), "cams": cams, "dec": e.get_decoder("color")}
    for lv in ("middle", "fine", "color"):
        out[lv] = e.get_grid(lv)
    e.close()
    if q is not None:
        q.put((rank, out))
    return out


@pytest.mark.parametrize("ba", [False, True])
def test_two_gpu_mapping_equals_one_gpu(nsb, ba):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ref = _run(0, 1, None, None, ba)
    uid = nsb.comm_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_run, args=(r, 2, uid, q, ba)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    grids0 = nsb.synthetic.make_grids(0)
    for r in range(2):
        o = got[r]
        assert np.allclose(o["losses"], ref["losses"], rtol=1e-4), (o["losses"], ref["losses"])
        for lv in ("middle", "fine", "color"):
            move = np.sqrt(((ref[lv] - grids0[lv]) ** 2).mean())
            assert move > 0 and np.sqrt(((o[lv] - ref[lv]) ** 2).mean()) < 2e-2 * move, lv     # Adam sign noise on ~zero gradients
        assert np.abs(o["dec"] - ref["dec"]).max() < 2e-2 * np.abs(ref["dec"] - nsb.synthetic.make_decoders(0, bias_scale=0.05)["color"]).max()
        assert np.abs(o["cams"] - ref["cams"]).max() < 1e-4
    assert np.array_equal(got[0]["middle"], got[1]["middle"]) and np.array_equal(got[0]["dec"], got[1]["dec"])   # replicas stay bit-identical
