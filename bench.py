#!/usr/bin/env python
"""bench.py -- mapping rays/s (fwd + bwd + fused Adam, 48 samples/ray) of the B200-native NICE-SLAM hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): Mapper iteration of Mapper.cpp:330-465 -- sample 5 keyframes x 1000 pixels
(5000 rays/iter), inside filter, render_batch_ray("color"), L1 depth (+ colour) loss, backward to the
middle/fine/color grids and the colour decoder, fused Adam.  One "step" = one joint iteration.  The K timed steps are
spread EVENLY over the 60-iteration schedule of one optimize_map (37 geometry + 23 colour iterations, Mapper.cpp:351-358):
step k runs iteration floor(k * 60 / K) for K < 60 and iteration k mod 60 otherwise, so ANY --steps times the 37:23 mix and
`ms_per_step_by_stage` always carries both stages; Adam is re-created at every wrap as the reference does per keyframe.
Synthetic 640x480 RGB-D frames, random-init grids and decoders (datasets / pretrained decoders are offline).
N > 1: weak scaling, global batch N x 5000 rays sharded over the ranks, gradients exchanged and stepped by one fused
reduce-scatter + Adam + all-gather kernel over NVLink peer memory (--comm nccl: ncclAllReduce + Adam).
`value` = rays of all ranks / max-over-ranks device time, inputs resident in HBM, every iteration ONE cudaGraphLaunch;
`e2e` = the same loop through the host-buffer C ABI (pinned pixel indices H2D + keyframe upload every 60 steps +
loss D2H every step).  The per-kernel times behind `roofline` come from a second pass over the same K steps with the kernels
enqueued one by one between CUDA events (a graph launch cannot be timed kernel by kernel).
`configs` carries the other BASELINE.json configurations (forward-only render, tracking, dense render) with their own CPU baselines.
--impl reference times the reference's own CPU path (oracle/_ref: its Renderer.cpp + utils.h on libtorch CPU, all host threads) on
the same 5000-ray iterations.
"""
import argparse
import hashlib
import importlib
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np

RAYS_PER_GPU = 5000
N_FRAMES = 5
ITERS_PER_KEYFRAME = 60
FIRST_COLOR_ITER = 37                      # stage_of_iter: it <= 24 middle, it <= 36 middle again (Mapper.cpp:355-356), then color
# algorithmic work per ray: MAC counts x 2 x 48 samples
FLOP_FWD_RAY = 51653 * 2 * 48              # 4.96 MFLOP: three decoders forward
FLOP_BWD_COLOR_RAY = 49367 * 2 * 48        # 4.74 MFLOP: colour-stage backward (grids + colour decoder wgrad)
FLOP_BWD_GEOM_RAY = 18496 * 2 * 48         # 1.78 MFLOP: geometry-stage backward (middle + fine data grads)
FLOP_WGRAD_RAY = 15575 * 2 * 48            # 1.50 MFLOP: colour-decoder weight gradient (one MAC per forward MAC of that decoder), part of FLOP_BWD_COLOR_RAY
GATHER_BYTES_RAY = 3 * 8 * 32 * 4 * 48     # 147 456 B of voxel-corner lines per ray and direction


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        p.update({"hbm_gbs": d["hbm_gbs"], "bf16_tflops_sustained": d["bf16_tflops_sustained"], "bf16_tflops": d.get("bf16_tflops"), "source": "measured"})
    except Exception:
        pass
    return p


def load_profile_json(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except Exception:
        return {}


def schedule(K):
    """Iteration number (0..59) of timed step k: the K steps sample the 60-iteration schedule evenly."""
    if K >= ITERS_PER_KEYFRAME:
        return [k % ITERS_PER_KEYFRAME for k in range(K)]
    return [(k * ITERS_PER_KEYFRAME) // K for k in range(K)]


_CLOCK_CHILD = r"""
import sys, time, json
import pynvml as N
N.nvmlInit()
uuid = sys.argv[1]
try:
    h = N.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
except Exception:
    h = N.nvmlDeviceGetHandleByIndex(int(sys.argv[2]))
mx = int(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
bits = {"hw_slowdown": N.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": N.nvmlClocksEventReasonHwThermalSlowdown,
        "sw_thermal_slowdown": N.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": N.nvmlClocksEventReasonSwPowerCap}
import select
out = []
print("ready", flush=True)
while True:
    r, _, _ = select.select([sys.stdin], [], [], 0.003)
    if r and not sys.stdin.readline():
        break
    try:
        out.append((time.monotonic(), int(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), int(N.nvmlDeviceGetCurrentClocksEventReasons(h))))
    except Exception:
        pass
res = {"max": mx, "samples": [[t, c, [k for k, b in bits.items() if r & b]] for t, c, r in out]}
print(json.dumps(res), flush=True)
"""


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md).  NVML is polled every ~3 ms by a
    CHILD PROCESS, not by a thread of this one: a polling thread shares the GIL with the launch loop and its hiccups become
    device time on every peer at the cross-GPU barrier.  window(t0, t1) summarises the samples of one region."""

    def __init__(self, dev):
        self.p = None
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(dev).uuid)
            self.p = subprocess.Popen([sys.executable, "-c", _CLOCK_CHILD, uuid, str(dev)], stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
            if self.p.stdout.readline().strip() != "ready":
                raise RuntimeError("clock child failed")
        except Exception:
            self.p = None
        self.res = None

    def stop(self):
        if not self.p:
            return
        try:
            self.p.stdin.close()
            self.res = json.loads(self.p.stdout.readline())
            self.p.wait(timeout=5)
        except Exception:
            self.res = None
        self.p = None

    def window(self, t0, t1):
        if not self.res:
            return None
        sm = [c for t, c, r in self.res["samples"] if t0 <= t <= t1]
        rs = sorted({x for t, c, r in self.res["samples"] if t0 <= t <= t1 for x in r})
        if not sm:
            return None
        return {"sm_mhz": int(statistics.median(sm)), "sm_min_mhz": min(sm), "sm_max_mhz": self.res["max"], "reasons": rs, "samples": len(sm)}


def synthetic_inputs(nsb):
    syn = nsb.synthetic
    return syn.make_grids(0), syn.make_decoders(0), syn.make_frames(N_FRAMES + 1, 0)


# --------------------------------------------------------------------------------------------- reference arm
def _reference_backend(nsb_syn, grids, decs, cores):
    """(kind, mapping-step callable, object) -- the reference's own compiled code when oracle/_ref travelled, else the port."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import nice_oracle as O
    import refbind as R
    lr = np.array([O.DEFAULT_LR[k] for k in ("coarse", "middle", "fine", "color")], np.float32)
    if R.available():
        ref = R.Ref(grids, decs)
        ref.set_threads(cores)
        return "reference", ref, O, lr
    import torch
    torch.set_num_threads(cores)
    return "port", O.Model(grids, decs), O, lr


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path, on the host cores (rank 0 only): the SAME 5000-ray iterations."""
    if rank != 0:
        return
    nsb_syn = importlib.import_module("nice-slam-cpp_b200.synthetic")
    grids, decs = nsb_syn.make_grids(0), nsb_syn.make_decoders(0)
    depths, colors, poses = nsb_syn.make_frames(N_FRAMES, 0)
    cores = os.cpu_count() or 1
    kind, ref, O, lr = _reference_backend(nsb_syn, grids, decs, cores)
    sched = [O.STAGE_ID[s] for s in O.stage_schedule(ITERS_PER_KEYFRAME)]
    names = {v: k for k, v in O.STAGE_ID.items()}
    its = schedule(args.steps)

    def step(it, seed):
        if kind == "reference":
            ref.mapping_iters(depths, colors, poses, nsb_syn.CAM, RAYS_PER_GPU, [sched[it]], lr, seed=seed)
        else:
            O.mapping_iters(ref, depths, colors, poses, nsb_syn.CAM, RAYS_PER_GPU, [names[sched[it]]], seed=seed)
    for i in range(args.warmup):
        step(59 if i % 2 else 0, i)
    t0 = time.perf_counter()
    for k, it in enumerate(its):
        step(it, args.warmup + k)
    dt = time.perf_counter() - t0
    v = args.steps * RAYS_PER_GPU / dt
    n_color = sum(1 for it in its if it >= FIRST_COLOR_ITER)
    out = {"impl": "reference", "metric": "mapping rays/s (fwd+bwd, 48 samples/ray)", "value": v, "unit": "rays/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "mapper_iteration_5000rays_60iters_keyframe", "rays_per_gpu": RAYS_PER_GPU, "global_rays": RAYS_PER_GPU, "samples_per_ray": 48,
                      "frames": N_FRAMES, "schedule": "timed step k = iteration floor(k*60/K) (K<60) or k%60 of optimize_map", "color_steps": n_color,
                      "raydir": "utils.h:44-47 as written (the reference's own raySampler)"},
           "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": kind,
                            "sample": "%d mapping iterations x %d rays (%d geometry + %d colour), libtorch CPU autograd + torch::optim::Adam" % (args.steps, RAYS_PER_GPU, args.steps - n_color, n_color)},
           "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def cpu_baselines(nsb, want_configs=True):
    """Bounded CPU samples (rank 0, N = 1): the mapping iteration for `cpu_baseline`, and forward-only / tracking samples for the
    `configs` sub-records.  ~10-30 s of CPU work in total."""
    syn = nsb.synthetic
    grids, decs = syn.make_grids(0), syn.make_decoders(0)
    depths, colors, poses = syn.make_frames(N_FRAMES, 0)
    cores = os.cpu_count() or 1
    kind, ref, O, lr = _reference_backend(syn, grids, decs, cores)
    out = {}
    n_it = 3
    if kind == "reference":
        ref.mapping_iters(depths, colors, poses, syn.CAM, 500, [3], lr, seed=0)   # warm-up
        _, _, sec = ref.mapping_iters(depths, colors, poses, syn.CAM, RAYS_PER_GPU, [1, 3, 3], lr, seed=1)
    else:
        t0 = time.perf_counter()
        O.mapping_iters(ref, depths, colors, poses, syn.CAM, RAYS_PER_GPU, ["middle", "color", "color"], seed=1)
        sec = time.perf_counter() - t0
    out["mapping"] = {"value": n_it * RAYS_PER_GPU / sec, "unit": "rays/s", "cores": cores, "kind": kind,
                      "sample": "%d mapping iterations (1 geometry + 2 colour) x %d rays, %.1f s" % (n_it, RAYS_PER_GPU, sec)}
    if not want_configs:
        return out
    import torch
    idx = syn.mt19937_indices(3, 6000, 480 * 640)
    a, b, c_, _ = O.ray_sampler(0, 480, 0, 640, idx, syn.CAM["fx"], syn.CAM["fy"], syn.CAM["cx"], syn.CAM["cy"], torch.tensor(depths[0]), torch.tensor(colors[0]),
                                torch.tensor(poses[0]), "pinhole")
    keep = O.inside_mask(a, b, c_, torch.tensor(syn.BOUND)).numpy()
    ro, rd, gd = a.numpy()[keep][:RAYS_PER_GPU], b.numpy()[keep][:RAYS_PER_GPU], c_.numpy()[keep][:RAYS_PER_GPU]
    if kind == "reference":
        ref.render_batch_ray(rd[:500], ro[:500], "color", gd[:500])
        t0 = time.perf_counter()
        for _ in range(2):
            ref.render_batch_ray(rd, ro, "color", gd)
        sec = (time.perf_counter() - t0) / 2
    else:
        with torch.no_grad():
            t0 = time.perf_counter()
            O.render_batch_ray(ref, torch.tensor(rd), torch.tensor(ro), "color", torch.tensor(gd))
            sec = time.perf_counter() - t0
    out["forward"] = {"value": rd.shape[0] / sec, "unit": "rays/s", "cores": cores, "kind": kind, "sample": "render_batch_ray forward, %d rays, %.2f s per call" % (rd.shape[0], sec)}
    pts = np.random.RandomState(0).uniform(syn.BOUND[:, 0], syn.BOUND[:, 1], (200000, 3)).astype(np.float32)
    t0 = time.perf_counter()
    if kind == "reference":
        ref.eval_points(pts, "color")
    else:
        with torch.no_grad():
            ref.eval_points(torch.tensor(pts), "color")
    sec = time.perf_counter() - t0
    out["eval_points"] = {"value": pts.shape[0] / sec, "unit": "points/s", "cores": cores, "kind": kind, "sample": "bounded: Renderer::eval_points on 200 000 of the 256^3 lattice points, %.2f s" % sec}
    cam0 = O.get_tensor_from_camera(poses[0]); cam0 = np.asarray(cam0, np.float32).copy(); cam0[4:] += 0.01
    n_trk = 2
    if kind == "reference":
        *_, sec = ref.tracking_iters(depths[0], colors[0], cam0, syn.CAM, RAYS_PER_GPU, n_trk, 1e-2, seed=2)
    else:
        t0 = time.perf_counter()
        O.tracking_iters(ref, depths[0], colors[0], cam0, syn.CAM, RAYS_PER_GPU, n_trk, 1e-2, seed=2)
        sec = time.perf_counter() - t0
    out["tracking"] = {"value": sec / n_trk * 1e3, "unit": "ms/iter", "cores": cores, "kind": kind, "sample": "%d tracking iterations x %d rays, %.1f s" % (n_trk, RAYS_PER_GPU, sec)}
    return out


# -------------------------------------------------------------------------------------------------- our arm
def run_tracking(e, nsb, torch, ext, depths, colors, poses, frames=6, warm_frames=2):
    """ms per tracking iteration (BASELINE configs[2], Tracker.cpp:41-113): 5000 rays in the edge-cropped window, render,
    uncertainty-weighted loss with the median mask, backward to the 7-vector, Adam; 10 iterations per frame, the loss read
    back every iteration as the reference prints it (Tracker.cpp:111).  Device time from events on the library's stream,
    wall time through the host ABI (pixel indices drawn by the host mt19937 and copied H2D inside the timed region)."""
    iters = e.cfg.tracking_iters
    e.set_frame(0, depths[0], colors[0], poses[0])
    cam0 = nsb.get_tensor_from_camera(poses[0])
    cam0[4:] += 0.01     # start 1 cm off the ground-truth translation
    dev_ms, wall = 0.0, 0.0
    with torch.cuda.stream(ext):
        for f in range(warm_frames + frames):
            e.tracking_begin(0, cam0)
            e.synchronize()
            evs = []
            t0 = time.perf_counter()
            for _ in range(iters):
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record(ext); loss, _ = e.tracking_iter(None, want_grad=False); b.record(ext)
                evs.append((a, b))
            e.synchronize()
            if f >= warm_frames:
                wall += time.perf_counter() - t0
                dev_ms += sum(a.elapsed_time(b) for a, b in evs)
    n = frames * iters
    return {"workload": "tracker_5000rays_10iters_frame", "ms_per_iter": dev_ms / n, "e2e_ms_per_iter": wall / n * 1e3, "rays_per_iter": int(e.cfg.tracking_pixels), "iters_per_frame": int(iters),
            "frames_timed": frames, "last_loss": float(loss), "note": "replicas only (one GPU per frame); loss D2H + sync every iteration"}


def run_forward_only(e, nsb, torch, ext, depths, colors, poses, flush, reps=10):
    """BASELINE configs[0]: Renderer::render_batch_ray forward on 5000 rays x 48 samples (cofusion.yaml camera).  Device time with the
    rays resident (nsb_render_batch_ray_dev), and through the host ABI (rays H2D, rgb/depth/var/weights D2H inside the timed region)."""
    idx = nsb.synthetic.mt19937_indices(3, 6000, 480 * 640)
    parts = [e.get_samples(0, 0, 480, 0, 640, 3000, idx=idx[h * 3000:(h + 1) * 3000]) for h in range(2)]
    inside = np.concatenate([p[4] for p in parts])
    ro, rd, gd = (np.concatenate([p[k] for p in parts])[inside][:RAYS_PER_GPU] for k in (0, 1, 2))
    n = ro.shape[0]
    d_ro = torch.from_numpy(ro).cuda(); d_rd = torch.from_numpy(rd).cuda(); d_gd = torch.from_numpy(gd).cuda()
    o_rgb = torch.empty(n, 3, device="cuda"); o_d = torch.empty(n, device="cuda"); o_v = torch.empty(n, device="cuda"); o_w = torch.empty(n, 48, device="cuda")
    torch.cuda.synchronize()
    ms = []
    with torch.cuda.stream(ext):
        for r in range(reps + 3):
            flush.zero_()
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(ext)
            e.render_batch_ray_dev("color", n, d_rd.data_ptr(), d_ro.data_ptr(), d_gd.data_ptr(), o_rgb.data_ptr(), o_d.data_ptr(), o_v.data_ptr(), o_w.data_ptr())
            b.record(ext)
            e.synchronize()
            if r >= 3:
                ms.append(a.elapsed_time(b))
    t0 = time.perf_counter()
    for _ in range(reps):
        e.render_batch_ray(rd, ro, "color", gd)
    wall = (time.perf_counter() - t0) / reps
    return {"workload": "render_batch_ray_forward_5000rays", "rays": int(n), "ms": float(np.mean(ms)), "rays_per_s": n / (np.mean(ms) * 1e-3),
            "e2e_ms": wall * 1e3, "e2e_rays_per_s": n / wall, "h2d_bytes": int(ro.nbytes + rd.nbytes + gd.nbytes), "d2h_bytes": int(n * (3 + 1 + 1 + 48) * 4)}


def run_dense(e, nsb, torch, ext, reps=3):
    """BASELINE configs[3]: all 640 x 480 = 307 200 pixels of a frame through render_batch_ray("color"), depth-guided
    (upstream render_img; Renderer.cpp:5 keeps its ray_batch_size).  Device time without the image read-back, and end to end."""
    ms = []
    with torch.cuda.stream(ext):
        for r in range(reps + 1):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(ext); e.render_img(0, "color", True, want_outputs=False); b.record(ext)
            e.synchronize()
            if r >= 1:
                ms.append(a.elapsed_time(b))
    t0 = time.perf_counter()
    for _ in range(reps):
        e.render_img(0, "color", True)
    wall = (time.perf_counter() - t0) / reps
    n = e.cfg.H * e.cfg.W
    return {"workload": "dense_render_640x480_color", "rays": int(n), "ms": float(np.mean(ms)), "rays_per_s": n / (np.mean(ms) * 1e-3), "e2e_ms": wall * 1e3,
            "e2e_rays_per_s": n / wall, "d2h_bytes": int(n * 5 * 4), "chunk_rays": int(e.cfg.max_rays)}


def run_lattice(e, res=256):
    """SURVEY 8-f row 3 / nice_slam.yaml meshing.resolution: eval_points("color") over a res^3 lattice generated on the device, the
    stage assembly and the bound mask applied in-kernel; wall time includes the read-back of the occupancy channel."""
    e.eval_lattice("color", 32, 32, 32)
    t0 = time.perf_counter()
    occ = e.eval_lattice("color", res, res, res)
    wall = time.perf_counter() - t0
    n = res ** 3
    return {"workload": "eval_points_lattice_%d^3_color" % res, "points": int(n), "e2e_ms": wall * 1e3, "points_per_s": n / wall, "d2h_bytes": int(n * 4),
            "inside_fraction": float((occ != 100.0).mean())}


def parity_selfcheck(nsb, e, torch, dist, rank, world, local_rank, grids, decs, depths, colors, poses, cfg):
    """N > 1: four fixed-seed iterations (2 geometry + 2 colour) on the sharded engine against ONE GPU rendering the same global
    batch: loss agreement, parameter agreement (RMS relative to the RMS update) and bit-identity of the replicas across ranks."""
    n_global = cfg.mapping_pixels
    its = [0, 1, 58, 59]
    idx = nsb.synthetic.mt19937_indices(1234, len(its) * n_global, cfg.H * cfg.W).reshape(len(its), n_global)
    e.set_model(grids, decs)
    e.mapping_begin(list(range(N_FRAMES)), ITERS_PER_KEYFRAME, 1.0)
    losses = [e.mapping_iter(it, idx[k]) for k, it in enumerate(its)]
    e.mapping_end()
    got = {lv: e.get_grid(lv) for lv in ("middle", "fine", "color")}
    got["dec"] = e.get_decoder("color")
    h = hashlib.sha256()
    for k in ("middle", "fine", "color", "dec"):
        h.update(np.ascontiguousarray(got[k]).tobytes())
    digest = torch.tensor(list(h.digest()), dtype=torch.uint8, device="cuda")
    alld = [torch.empty_like(digest) for _ in range(world)]
    dist.all_gather(alld, digest)
    identical = all(bool(torch.equal(alld[0], d)) for d in alld)
    res = None
    if rank == 0:
        c1 = nsb.default_config()
        for f, _t in cfg._fields_:
            setattr(c1, f, getattr(cfg, f))
        e1 = nsb.Engine(c1, device=local_rank)       # world = 1: the whole global batch on one GPU
        e1.set_model(grids, decs)
        for f in range(N_FRAMES):
            e1.set_frame(f, depths[f], colors[f], poses[f])
        e1.mapping_begin(list(range(N_FRAMES)), ITERS_PER_KEYFRAME, 1.0)
        ref_losses = [e1.mapping_iter(it, idx[k]) for k, it in enumerate(its)]
        rms = {}
        for lv in ("middle", "fine", "color"):
            ref = e1.get_grid(lv)
            move = float(np.sqrt(((ref - grids[lv]) ** 2).mean()))
            rms[lv] = float(np.sqrt(((got[lv] - ref) ** 2).mean()) / max(move, 1e-30))
        dref = e1.get_decoder("color")
        res = {"iterations": its, "rays": int(n_global), "ranks_bit_identical": bool(identical),
               "loss_rel_vs_1gpu": float(np.max(np.abs(np.array(losses) - np.array(ref_losses)) / np.abs(ref_losses))),
               "grid_rms_vs_1gpu": rms, "dec_color_rel_vs_1gpu": float(np.abs(got["dec"] - dref).max() / max(np.abs(dref - decs["color"]).max(), 1e-30)),
               "dec_color_rms_vs_1gpu": float(np.sqrt(((got["dec"] - dref) ** 2).mean()) / max(np.sqrt(((dref - decs["color"]) ** 2).mean()), 1e-30)),
               "note": "grid_rms = RMS(param_N - param_1) / RMS(update): Adam turns gradients at the fp32 noise level into +-lr steps of random sign in any "
                       "implementation, and the sharded sum only re-associates the fp32 additions; the second colour step's gradient is taken at those parameters and the "
                       "loss is piecewise linear (L1, relu), so dec_color_rel (a max) also moves by a few % of a step between two ONE-GPU runs "
                       "(profiles/r4t_determinism_seed_sweep.log) -- dec_color_rms is the stable figure"}
        e1.close()
    return res


def run_ours(args, rank, world, local_rank):
    import torch
    nsb = importlib.import_module("nice-slam-cpp_b200")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    grids, decs, (depths, colors, poses) = synthetic_inputs(nsb)
    n_global = RAYS_PER_GPU * world
    cfg = nsb.default_config()
    cfg.mapping_pixels = n_global
    cfg.max_rays = n_global
    cfg.max_frames = N_FRAMES + 1
    cfg.tracking_pixels = RAYS_PER_GPU     # BASELINE configs[2]: 5000 rays per tracking iteration
    cfg.frustum_feature_selection = 0     # every voxel is optimised in BOTH arms (random synthetic depth gives a degenerate frustum);
                                          # the GPU frustum mask (Mapper.cpp:42-130) is covered by tests/test_gpu_parity.py
    e = nsb.Engine(cfg, device=local_rank)
    e.set_model(grids, decs)
    for f in range(N_FRAMES):
        e.set_frame(f, depths[f], colors[f], poses[f])
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(nsb.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        e.comm_init(bytes(uid.cpu().tolist()), rank, world)
        if args.comm == "p2p":   # peer-memory optimiser step: gather every rank's CUDA IPC handles, import, barrier
            mine = torch.tensor(list(e.p2p_export()), dtype=torch.uint8, device="cuda")
            allh = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allh, mine)
            ok = torch.ones(1, device="cuda")
            try:
                e.p2p_import(b"".join(bytes(h.cpu().tolist()) for h in allh), rank, world)
            except RuntimeError as ex:       # e.g. CUDA IPC not permitted in this container: every rank falls back to NCCL together
                print("rank %d: peer-memory mode unavailable (%s)" % (rank, ex), file=sys.stderr)
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok.item()) == 0.0:
                e.p2p_import(None, rank, world)
                args.comm = "nccl"
            dist.barrier()
    pix = n_global // N_FRAMES
    K, W = args.steps, args.warmup
    its = schedule(K)
    warm_its = [(0, 59)[i % 2] for i in range(W)]      # both stages (every kernel variant / captured graph) are warmed
    idx_all = nsb.synthetic.mt19937_indices(0, (K + W) * N_FRAMES * pix, cfg.H * cfg.W).reshape(K + W, N_FRAMES * pix)
    slots = list(range(N_FRAMES))
    ext = torch.cuda.ExternalStream(e.stream_ptr(), device=torch.device("cuda", local_rank))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        e.synchronize(); torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    state = {"prev": None}

    def run_step(it, idx=None, sync=False):
        if state["prev"] is None or it < state["prev"]:
            e.mapping_begin(slots, ITERS_PER_KEYFRAME, 1.0)   # fresh Adam per optimize_map (Mapper.cpp:330)
        state["prev"] = it
        return e.mapping_iter(it, idx, sync=sync)              # idx None: next row of the device-resident pool

    clocks = ClockSampler(local_rank) if rank == 0 else None

    def timed_pass(profile):
        """W warm-up + K timed steps (L2 flushed between steps, events on the library's stream)."""
        state["prev"] = None
        e.mapping_set_index_pool(idx_all)
        with torch.cuda.stream(ext):
            for it in warm_its:
                run_step(it)
            barrier()
            e.launch_count(reset=True)
            e.set_profiling(profile)
            t0 = time.monotonic()
            evs = []
            kms_stage = {"geometry": {}, "color": {}}          # per-kernel device time split by iteration kind (profile pass only)
            prev = None
            for it in its:
                flush.zero_()                                  # L2 flush between timed steps (not timed)
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record(ext); run_step(it); b.record(ext)
                evs.append((a, b))
                if profile:                                    # cumulative since set_profiling(1): difference per step (synchronises; this pass is not `value`)
                    cum = e.kernel_ms()
                    st = kms_stage["color" if it >= FIRST_COLOR_ITER else "geometry"]
                    for k_, v_ in cum.items():
                        st[k_] = st.get(k_, 0.0) + v_ - (prev[k_] if prev else 0.0)
                    prev = cum
            barrier()
            t1 = time.monotonic()
            step_ms = [a.elapsed_time(b) for a, b in evs]
            kms = e.kernel_ms() if profile else None
            if profile:
                kms = dict(kms, by_stage=kms_stage)
            launches = e.launch_count()
            e.set_profiling(False)
        return step_ms, kms, launches, (t0, t1)

    # ---- device-resident loop, one cudaGraphLaunch per iteration: `value`
    step_ms, _, launches, win = timed_pass(False)
    with torch.cuda.stream(ext):
        last_begin = max(k for k in range(K) if k == 0 or its[k] < its[k - 1])
        losses, n_inside = e.mapping_losses(0, K - last_begin)    # the steps since the last mapping_begin
    t_dev = sum(step_ms) * 1e-3
    t = torch.tensor([t_dev], device="cuda", dtype=torch.float64)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_dev_max = float(t.item())
    value = K * n_global / t_dev_max
    is_color = [it >= FIRST_COLOR_ITER for it in its]
    ms_geom = float(np.mean([m for m, c_ in zip(step_ms, is_color) if not c_])) if not all(is_color) else None
    ms_color = float(np.mean([m for m, c_ in zip(step_ms, is_color) if c_])) if any(is_color) else None
    # ---- the same K steps, kernels enqueued one by one between CUDA events: per-kernel device time for the roofline
    step_ms_eager, kms, _, _ = timed_pass(True)
    gather_ms = e.bench_gather(20)       # grid sampling alone on the last step's rays (resident grids: L1/L2 traffic)

    # ---- end-to-end loop through the host-buffer ABI: pinned indices H2D, keyframe upload per 60 steps, loss D2H per step
    e.mapping_set_index_pool(None)
    pin_idx = torch.from_numpy(idx_all).pin_memory().numpy()
    pin_depth = torch.from_numpy(depths[N_FRAMES]).pin_memory().numpy(); pin_color = torch.from_numpy(colors[N_FRAMES]).pin_memory().numpy()
    state["prev"] = None
    for i, it in enumerate(warm_its):
        run_step(it, pin_idx[i], sync=True)
    barrier()
    state["prev"] = None
    t0 = time.perf_counter()
    for k, it in enumerate(its):
        if state["prev"] is None or it < state["prev"]:
            e.set_frame(N_FRAMES - 1, pin_depth, pin_color, poses[N_FRAMES - 1])   # the new keyframe of this optimize_map
        run_step(it, pin_idx[W + k], sync=True)
    barrier()
    t_e2e = time.perf_counter() - t0
    t = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_e2e = float(t.item())
    n_begin = sum(1 for k in range(K) if k == 0 or its[k] < its[k - 1])
    frame_bytes = (depths[0].nbytes + colors[0].nbytes + 64)
    e2e = {"value": K * n_global / t_e2e, "unit": "rays/s", "h2d_bytes_per_step": int(idx_all[0].nbytes + frame_bytes * n_begin / K),
           "d2h_bytes_per_step": 16, "ms_per_step": t_e2e / K * 1e3}
    e.set_frame(N_FRAMES - 1, depths[N_FRAMES - 1], colors[N_FRAMES - 1], poses[N_FRAMES - 1])

    p2p_times = e.p2p_times() if (world > 1 and args.comm == "p2p") else None
    parity = None
    if world > 1:
        parity = parity_selfcheck(nsb, e, torch, dist, rank, world, local_rank, grids, decs, depths, colors, poses, cfg)

    # ---- the other BASELINE.json configurations (rank 0's GPU; replicas only)
    configs = None
    if rank == 0 and not args.no_configs:
        e.set_model(grids, decs)
        configs = {"forward_only": run_forward_only(e, nsb, torch, ext, depths, colors, poses, flush), "dense_render": run_dense(e, nsb, torch, ext),
                   "mesh_lattice": run_lattice(e)}
    tracking = run_tracking(e, nsb, torch, ext, depths, colors, poses)   # replicas only: every rank tracks its own frame
    if clocks:
        clocks.stop()

    if rank == 0:
        pk = peaks()
        n_color = sum(is_color)
        n_geom = K - n_color
        frac_in = float(np.mean(n_inside[n_inside > 0])) / n_global if np.any(n_inside > 0) else 1.0
        rays_rank = RAYS_PER_GPU * frac_in                 # rays that survive the inside filter, per rank and step
        traffic = load_profile_json("traffic.json")        # dram__bytes_read.sum + dram__bytes_write.sum per launch and kernel VARIANT (ncu --set full)
        by = kms.pop("by_stage")
        tc = os.environ.get("NSB_TCGEN05", "3")
        fwd_geo = {"3": "k_decode_fwd_t5", "2": "k_decode_fwd_tc16", "1": "k_decode_fwd_tc"}.get(tc, "k_decode_fwd")
        stash = os.environ.get("NSB_WGRAD_STASH", "1") != "0"
        t5_stash = tc == "3" and os.environ.get("NSB_T5_STASH", "1") != "0"
        # a colour iteration that stashes activations for k_wgrad: the tcgen05 forward writes the stash itself (default) or the colour decoder
        # runs on the warp-MMA forward beside it (NSB_T5_STASH=0)
        fwd_col = fwd_geo if (not stash or t5_stash) else "k_decode_fwd"

        def entry(name, stage, key, n_it, flop_ray, extra=None):
            ms = by[stage].get(key, 0.0)
            if n_it == 0 or ms <= 0:
                return None
            d = {"kernel": name, "iterations": stage, "launches": n_it, "avg_launch_ms": ms / n_it, "total_ms": ms,
                 "flop_per_launch": flop_ray * rays_rank, "achieved": flop_ray * rays_rank * n_it / (ms * 1e-3) * 1e-12,
                 "traffic": traffic.get("%s<%s>" % (name, stage))}
            if extra:
                d[extra] = GATHER_BYTES_RAY * (2.0 / 3.0 if stage == "geometry" and extra == "scatter_gbs" else 1.0) * rays_rank * n_it / (ms * 1e-3) * 1e-9
            return d
        ents = [entry(fwd_geo, "geometry", "decode_fwd", n_geom, FLOP_FWD_RAY, "gather_gbs"), entry(fwd_col, "color", "decode_fwd", n_color, FLOP_FWD_RAY, "gather_gbs"),
                entry("k_decode_bwd", "geometry", "decode_bwd", n_geom, FLOP_BWD_GEOM_RAY, "scatter_gbs"),
                entry("k_decode_bwd", "color", "decode_bwd", n_color, FLOP_BWD_COLOR_RAY - FLOP_WGRAD_RAY, "scatter_gbs"),
                entry("k_wgrad", "color", "wgrad", n_color, FLOP_WGRAD_RAY)]
        ents = [x for x in ents if x]
        kern = {"%s<%s>" % (x["kernel"], x["iterations"]): x for x in ents}
        dom = max(ents, key=lambda x: x["total_ms"])         # the single kernel with the largest share of the K timed steps
        ach = dom["achieved"]
        gat = load_profile_json("gather_ncu.json")        # l1tex / lts / dram bytes of k_gather_only per launch (ncu), measured L2 gather peak
        gs = {"kernel": "k_gather_only", "ms": gather_ms, "rays": int(n_global // world),
              "algorithmic_gbs": GATHER_BYTES_RAY * RAYS_PER_GPU / (gather_ms * 1e-3) * 1e-9 if gather_ms > 0 else 0.0,
              "note": "all rays of the step (no inside filter), 3 grids x 8 corners x 128 B per sample; grids (11.2 MB) are L2-resident, so the algorithmic "
                      "bytes are served by L1 (neighbouring samples share corners) and L2; the roofline fraction is L2 bytes / time against the measured L2 gather peak"}
        if gat.get("lts_t_bytes") and gather_ms > 0:
            gs.update({"l1tex_t_bytes": gat.get("l1tex_t_bytes"), "lts_t_bytes": gat["lts_t_bytes"], "dram_bytes": gat.get("dram_bytes"),
                       "l2_gbs": gat["lts_t_bytes"] / (gather_ms * 1e-3) * 1e-9, "l2_peak_gbs_measured": gat.get("l2_gather_peak_gbs"),
                       "frac_of_l2_peak": gat["lts_t_bytes"] / (gather_ms * 1e-3) * 1e-9 / gat["l2_gather_peak_gbs"] if gat.get("l2_gather_peak_gbs") else None,
                       "l1_gbs": (gat.get("l1tex_t_bytes") or 0) / (gather_ms * 1e-3) * 1e-9, "source": gat.get("_source")})
        roof = {"bound": "tensor", "kernel": "%s<%s>" % (dom["kernel"], dom["iterations"]), "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": ach / pk["bf16_tflops_sustained"], "traffic": dom["traffic"], "traffic_by_variant": {k: v for k, v in traffic.items() if not k.startswith("_")},
                "peak_source": pk["source"] + " bf16 dense (sustained, kernel timed inside the step)",
                "note": "algorithmic FLOPs (SURVEY 8-d: 2 x MACs of the three decoders x 48 samples x rays surviving the inside filter) / CUDA-event time of the kernel "
                        "(second pass over the same K steps, kernels enqueued one by one, split by iteration kind); the arithmetic is the fp32-grade fp16 two-way split "
                        "(3 tensor-core MMAs per product).  k_decode_fwd_t5 issues them as tcgen05.mma kind::f16 with operands in tensor memory (tensor pipe ~20 % busy: "
                        "the kernel is bound by the per-value ALU work -- sines, fp16 splits, relu, gather -- not by the tensor pipe); the warp-MMA kernels (backward, "
                        "stash forward, wgrad) are bound by T(HMMA) + T(ALU), legacy HMMA ceiling 556 TFLOP/s / 3 = 185 TFLOP/s algorithmic",
                "launches": dom["launches"], "avg_launch_ms": dom["avg_launch_ms"], "kernels": kern,
                "kernel_ms_total": kms, "ms_per_step_eager": float(np.mean(step_ms_eager)), "grid_sampling": gs}
        cpu = None
        try:
            if world == 1 and not args.no_cpu_baseline:
                cb = cpu_baselines(nsb, want_configs=configs is not None)
                cpu = cb["mapping"]
                if configs is not None:
                    configs["forward_only"]["cpu_baseline"] = cb.get("forward")
                    configs["mesh_lattice"]["cpu_baseline"] = cb.get("eval_points")
                    tracking["cpu_baseline"] = cb.get("tracking")
                    if cb.get("forward"):
                        configs["dense_render"]["cpu_baseline"] = dict(cb["forward"], sample="bounded: " + cb["forward"]["sample"] + " (a 5000-ray strip of the image)")
        except Exception as ex:  # the checker is optional for the bench line
            cpu = {"error": repr(ex)}
        if configs is not None:
            configs["tracking"] = tracking
        clk = clocks.window(*win) if clocks else None
        out = {"metric": "mapping rays/s (fwd+bwd, 48 samples/ray)", "value": value, "unit": "rays/s", "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": t_dev_max / K * 1e3, "ms_per_step_rank0": t_dev / K * 1e3, "ms_per_step_by_stage": {"geometry": ms_geom, "color": ms_color, "geometry_steps": n_geom, "color_steps": n_color},
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic",
               "config": {"workload": "mapper_iteration_5000rays_60iters_keyframe", "rays_per_gpu": RAYS_PER_GPU, "global_rays": n_global,
                          "samples_per_ray": 48, "frames": N_FRAMES, "schedule": "timed step k = iteration floor(k*60/K) (K<60) or k%60 of optimize_map (37 geometry + 23 colour)",
                          "launch": "one cudaGraphLaunch per iteration (device-resident iteration state)",
                          "mma": "fp16 two-way split (3 products), fp32 accumulate: fp32-grade; forward on tcgen05.mma kind::f16 (operands in tensor memory), backward / wgrad on mma.sync m16n8k16", "l2": "256 MiB flush write between timed steps", "parallelism": ("rays sharded x%d, " % world) + ("single GPU" if world == 1 else "fused reduce-scatter + Adam + all-gather kernel over NVLink peer memory" if args.comm == "p2p" else "NCCL all-reduce of grads + Adam"),
                          "raydir": "pinhole (library default)", "inside_fraction": frac_in},
               "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "tracking": tracking, "configs": configs,
               "parity": parity, "p2p_exchange": p2p_times,
               "loss_first_last": [float(losses[0]), float(losses[len(losses) - 1])]}
        print(json.dumps(out), flush=True)
    e.close()
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the forward-only / dense-render sub-records")
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"], help="multi-GPU optimiser step: fused peer-memory kernel (default) or ncclAllReduce + Adam")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 4)      # >= 3, and even: both stages get warmed
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
