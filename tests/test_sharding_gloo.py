"""CPU, world_size 2, gloo: the multi-GPU path shards the (already filtered, batch-global) ray list into contiguous
slices, every rank renders its slice against replicated grids, and ONE sum all-reduce of the flat gradient arena
(loss scalar in its tail) makes every rank's Adam input identical (api.cu nsb_mapping_iter_async).
This test runs exactly that host logic with the oracle standing in for the kernels: the sliced, all-reduced
gradients must equal the single-process gradients, and the batch-global scalars of Renderer.cpp:76,93 must be
computed over the full batch (slicing first would change the z values)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

WORLD = 2


def slice_of(n, rank, world):
    per = -(-n // world)           # cdiv, as nsb_mapping_iter_async
    off = min(n, rank * per)
    return off, max(0, min(per, n - off))


def _worker(rank, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import importlib
    import nice_oracle as O
    syn = importlib.import_module("nice-slam-cpp_b200.synthetic")
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    torch.set_num_threads(2)
    small = {"coarse": (3, 2, 4), "middle": (6, 4, 8), "fine": (10, 6, 14), "color": (10, 6, 14)}
    grids = syn.make_grids(0, dims=small); decs = syn.make_decoders(0)
    depths, colors, poses = syn.make_frames(2, 0, H=60, W=80)
    cam = dict(fx=45.0, fy=45.0, cx=40.0, cy=30.0)
    tt, ts = O.t_tables()
    idx = syn.mt19937_indices(4, 120, 60 * 80)          # identical on every rank (same seed), like the CUDA path
    ro, rd, gd, gc = O.ray_sampler(0, 60, 0, 80, idx, cam["fx"], cam["fy"], cam["cx"], cam["cy"], torch.tensor(depths[0]), torch.tensor(colors[0]), torch.tensor(poses[0]))
    keep = O.inside_mask(ro, rd, gd, torch.tensor(O.BOUND))
    ro, rd, gd, gc = ro[keep], rd[keep], gd[keep], gc[keep]
    n = ro.shape[0]

    def grads(sl, gmax12, gmax):
        m = O.Model(grids, decs)
        for k in ("middle", "fine", "color"):
            m.grids[k].requires_grad_(True)
        o, c = sl
        # z values with the BATCH-GLOBAL scalars: evaluate on the full batch, then slice (what k_zvals does via stats[0])
        z = O.z_values(ro, rd, gd, m.bound, tt, ts)[o:o + c]
        pts = ro[o:o + c, None, :] + rd[o:o + c, None, :] * z[..., None]
        raw = m.eval_points(pts.reshape(-1, 3), "color").reshape(c, 48, 4)
        rgb, depth, var, _ = O.raw2outputs(raw, z, rd[o:o + c])
        loss = O.mapping_loss(gd[o:o + c], gc[o:o + c], depth, rgb, True, 0.5)
        loss.backward()
        return torch.cat([m.grids[k].grad.reshape(-1) for k in ("middle", "fine", "color")] + [loss.detach().reshape(1)])

    arena = grads(slice_of(n, rank, WORLD), None, None)
    dist.all_reduce(arena, op=dist.ReduceOp.SUM)          # the one collective of the path
    if rank == 0:
        full = grads((0, n), None, None)
        q.put((float((arena - full).abs().max()), float(full.abs().max()), int(n), [slice_of(n, r, WORLD) for r in range(WORLD)]))
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_ray_sharding_allreduce_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    err, scale, n, slices = q.get(timeout=500)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sum(c for _, c in slices) == n and slices[1][0] == slices[0][1]       # contiguous, disjoint, complete
    assert err < 1e-4 * scale, (err, scale)


def test_slice_edges():
    for n in (0, 1, 7, 5000, 5001, 39999):
        for w in (1, 2, 4, 8):
            parts = [slice_of(n, r, w) for r in range(w)]
            assert sum(c for _, c in parts) == n
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] or parts[i + 1][1] == 0 for i in range(w - 1))
