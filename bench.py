#!/usr/bin/env python
"""bench.py -- mapping rays/s (fwd + bwd + fused Adam, 48 samples/ray) of the B200-native NICE-SLAM hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): Mapper iteration of Mapper.cpp:330-465 -- sample 5 keyframes x 1000 pixels
(5000 rays/iter), inside filter, render_batch_ray("color"), L1 depth (+ colour) loss, backward to the
middle/fine/color grids and the colour decoder, fused Adam.  One "step" = one joint iteration; step i runs
iteration (i mod 60) of a 60-iteration optimize_map (37 geometry + 23 colour iterations, Mapper.cpp:351-358),
Adam being re-created at every wrap as the reference does per keyframe.  Synthetic 640x480 RGB-D frames,
random-init grids and decoders (datasets / pretrained decoders are offline).
N > 1: weak scaling, global batch N x 5000 rays sharded over the ranks, one NCCL all-reduce of the flat gradient
arena per iteration.  `value` = rays of all ranks / max-over-ranks device time, inputs resident in HBM;
`e2e` = the same loop through the host-buffer C ABI (pinned pixel indices H2D + keyframe upload every 60 steps +
loss D2H every step).  --impl reference times the reference's own CPU path (oracle/_ref: its Renderer.cpp +
utils.h on libtorch CPU, all host threads) on a bounded sample (1000 rays/step, the reference's mapping.pixels).
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np

RAYS_PER_GPU = 5000
N_FRAMES = 5
ITERS_PER_KEYFRAME = 60
# algorithmic work per ray: MAC counts x 2 x 48 samples
FLOP_FWD_RAY = 51653 * 2 * 48              # 4.96 MFLOP: three decoders forward
FLOP_BWD_COLOR_RAY = 49367 * 2 * 48        # 4.74 MFLOP: colour-stage backward (grids + colour decoder wgrad)
FLOP_BWD_GEOM_RAY = 18496 * 2 * 48         # 1.78 MFLOP: geometry-stage backward (middle + fine data grads)
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


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML polled every ~2 ms from a
    thread (nvidia-smi -lms cannot start fast enough for a 60 ms region); nvidia-smi is the fallback."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, dev):
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self.th = None
        self.smi = None
        try:
            import pynvml as N
            import torch
            N.nvmlInit()
            h = None
            try:
                uuid = str(torch.cuda.get_device_properties(dev).uuid)
                h = N.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = N.nvmlDeviceGetHandleByIndex(dev)
            self.max_mhz = int(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            bits = {"hw_slowdown": N.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": N.nvmlClocksEventReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": N.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": N.nvmlClocksEventReasonSwPowerCap}

            def poll():
                while not self._stop.is_set():
                    try:
                        self.sm.append(int(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
                        r = int(N.nvmlDeviceGetCurrentClocksEventReasons(h))
                        for k, b in bits.items():
                            if r & b:
                                self.reasons.add(k)
                    except Exception:
                        pass
                    time.sleep(0.002)
            self.th = threading.Thread(target=poll, daemon=True)
            self.th.start()
        except Exception:
            self.th = None
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            try:
                self.smi = subprocess.Popen(["nvidia-smi", "-i", str(dev), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                            stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                threading.Thread(target=self._read_smi, daemon=True).start()
            except Exception:
                self.smi = None

    def _read_smi(self):
        for line in self.smi.stdout:
            r = [x.strip() for x in line.split(",")]
            if r and r[0].isdigit():
                self.sm.append(int(r[0]))
                if len(r) > 1 and r[1].isdigit():
                    self.max_mhz = int(r[1])
                for k, nm in enumerate(self.NAMES):
                    if len(r) > 2 + k and r[2 + k].lower().startswith("active"):
                        self.reasons.add(nm)

    def stop(self):
        self._stop.set()
        if self.th:
            self.th.join(timeout=1)
        if self.smi:
            self.smi.terminate()
        if not self.sm:
            return None
        return {"sm_mhz": int(statistics.median(self.sm)), "sm_min_mhz": min(self.sm), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.sm)}


def synthetic_inputs(nsb):
    syn = nsb.synthetic
    return syn.make_grids(0), syn.make_decoders(0), syn.make_frames(N_FRAMES + 1, 0)


# --------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path, on the host cores (rank 0 only)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    nsb_syn = importlib.import_module("nice-slam-cpp_b200.synthetic")
    import nice_oracle as O
    import refbind as R
    grids, decs = nsb_syn.make_grids(0), nsb_syn.make_decoders(0)
    depths, colors, poses = nsb_syn.make_frames(N_FRAMES, 0)
    sample_rays = 1000
    lr = np.array([O.DEFAULT_LR[k] for k in ("coarse", "middle", "fine", "color")], np.float32)
    sched = [O.STAGE_ID[s] for s in O.stage_schedule(ITERS_PER_KEYFRAME)]
    cores = os.cpu_count() or 1
    if R.available():
        ref = R.Ref(grids, decs)
        ref.set_threads(cores)
        kind = "reference"

        def step(i):
            ref.mapping_iters(depths, colors, poses, nsb_syn.CAM, sample_rays, [sched[i % ITERS_PER_KEYFRAME]], lr, seed=i)
    else:   # oracle/_ref needs /root/reference to build; fall back to the port of the same algorithm
        import torch
        torch.set_num_threads(cores)
        model = O.Model(grids, decs)
        names = {v: k for k, v in O.STAGE_ID.items()}
        kind = "port"

        def step(i):
            O.mapping_iters(model, depths, colors, poses, nsb_syn.CAM, sample_rays, [names[sched[i % ITERS_PER_KEYFRAME]]], seed=i)
    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    dt = time.perf_counter() - t0
    v = args.steps * sample_rays / dt
    out = {"impl": "reference", "metric": "mapping rays/s (fwd+bwd, 48 samples/ray)", "value": v, "unit": "rays/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "mapper_iteration_5000rays_60iters_keyframe", "sample": "%d rays/step" % sample_rays},
           "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": kind, "sample": "%d steps x %d rays, stage schedule of a 60-iteration optimize_map" % (args.steps, sample_rays)},
           "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def cpu_baseline_sample(nsb):
    """Bounded CPU sample for the `cpu_baseline` key of our own line: 3 colour-stage iterations x 5000 rays."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import nice_oracle as O
    import refbind as R
    syn = nsb.synthetic
    grids, decs = syn.make_grids(0), syn.make_decoders(0)
    depths, colors, poses = syn.make_frames(N_FRAMES, 0)
    cores = os.cpu_count() or 1
    lr = np.array([O.DEFAULT_LR[k] for k in ("coarse", "middle", "fine", "color")], np.float32)
    n_it = 3
    if R.available():
        ref = R.Ref(grids, decs); ref.set_threads(cores)
        ref.mapping_iters(depths, colors, poses, syn.CAM, 500, [3], lr, seed=0)   # warm-up
        _, _, sec = ref.mapping_iters(depths, colors, poses, syn.CAM, RAYS_PER_GPU, [1, 3, 3], lr, seed=1)
        kind = "reference"
    else:
        import torch
        torch.set_num_threads(cores)
        model = O.Model(grids, decs)
        t0 = time.perf_counter()
        O.mapping_iters(model, depths, colors, poses, syn.CAM, RAYS_PER_GPU, ["middle", "color", "color"], seed=1)
        sec = time.perf_counter() - t0
        kind = "port"
    return {"value": n_it * RAYS_PER_GPU / sec, "unit": "rays/s", "cores": cores, "kind": kind,
            "sample": "%d mapping iterations (1 geometry + 2 colour) x %d rays, %.1f s" % (n_it, RAYS_PER_GPU, sec)}


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
    return {"ms_per_iter": dev_ms / n, "e2e_ms_per_iter": wall / n * 1e3, "rays_per_iter": int(e.cfg.tracking_pixels), "iters_per_frame": int(iters),
            "frames_timed": frames, "last_loss": float(loss), "note": "replicas only (one GPU per frame); loss D2H + sync every iteration"}


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
    idx_all = nsb.synthetic.mt19937_indices(0, (K + W) * N_FRAMES * pix, cfg.H * cfg.W).reshape(K + W, N_FRAMES * pix)
    slots = list(range(N_FRAMES))
    ext = torch.cuda.ExternalStream(e.stream_ptr(), device=torch.device("cuda", local_rank))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        e.synchronize(); torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident loop: `value`
    e.mapping_set_index_pool(idx_all)
    def run_step(i, idx=None, sync=False):
        it = i % ITERS_PER_KEYFRAME
        if it == 0:
            e.mapping_begin(slots, ITERS_PER_KEYFRAME, 1.0)   # fresh Adam per optimize_map (Mapper.cpp:330)
        return e.mapping_iter(it, idx, sync=sync)              # idx None: next row of the device-resident pool

    with torch.cuda.stream(ext):
        for i in range(W):
            run_step(i)
        barrier()
        e.launch_count(reset=True)
        e.set_profiling(True)
        clocks = ClockSampler(local_rank) if rank == 0 else None
        evs = []
        for i in range(W, W + K):
            flush.zero_()                                  # L2 flush between timed steps (not timed)
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(ext); run_step(i); b.record(ext)
            evs.append((a, b))
        barrier()
        clk = clocks.stop() if clocks else None
        step_ms = [a.elapsed_time(b) for a, b in evs]
        t_dev = sum(step_ms) * 1e-3
        is_color = [((W + k) % ITERS_PER_KEYFRAME) > 36 for k in range(K)]
        ms_geom = float(np.mean([m for m, c_ in zip(step_ms, is_color) if not c_])) if not all(is_color) else None
        ms_color = float(np.mean([m for m, c_ in zip(step_ms, is_color) if c_])) if any(is_color) else None
        kms = e.kernel_ms()
        launches = e.launch_count()
        e.set_profiling(False)
        gather_ms = e.bench_gather(20)       # grid sampling alone on the last step's rays (resident grids: L1/L2 traffic)
        last_it = (W + K - 1) % ITERS_PER_KEYFRAME
        losses, n_inside = e.mapping_losses(0, last_it + 1)    # the iterations run since the last mapping_begin
    t = torch.tensor([t_dev], device="cuda", dtype=torch.float64)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_dev = float(t.item())
    value = K * n_global / t_dev

    # ---- end-to-end loop through the host-buffer ABI: pinned indices H2D, keyframe upload per 60 steps, loss D2H per step
    e.mapping_set_index_pool(None)
    pin_idx = torch.from_numpy(idx_all).pin_memory().numpy()
    pin_depth = torch.from_numpy(depths[N_FRAMES]).pin_memory().numpy(); pin_color = torch.from_numpy(colors[N_FRAMES]).pin_memory().numpy()
    for i in range(W):
        run_step(i, pin_idx[i], sync=True)
    barrier()
    t0 = time.perf_counter()
    for i in range(W, W + K):
        if i % ITERS_PER_KEYFRAME == 0:
            e.set_frame(N_FRAMES - 1, pin_depth, pin_color, poses[N_FRAMES - 1])   # the new keyframe of this optimize_map
        run_step(i, pin_idx[i], sync=True)
    barrier()
    t_e2e = time.perf_counter() - t0
    t = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_e2e = float(t.item())
    frame_bytes = (depths[0].nbytes + colors[0].nbytes + 64)
    e2e = {"value": K * n_global / t_e2e, "unit": "rays/s", "h2d_bytes_per_step": int(idx_all[0].nbytes + frame_bytes / ITERS_PER_KEYFRAME),
           "d2h_bytes_per_step": 8, "ms_per_step": t_e2e / K * 1e3}

    tracking = run_tracking(e, nsb, torch, ext, depths, colors, poses)   # replicas only: every rank tracks its own frame

    if rank == 0:
        pk = peaks()
        n_color = sum(1 for i in range(W, W + K) if (i % ITERS_PER_KEYFRAME) > 36)
        n_geom = K - n_color
        frac_in = float(np.mean(n_inside[n_inside > 0])) / n_global if np.any(n_inside > 0) else 1.0
        rays_rank = RAYS_PER_GPU * frac_in                 # rays that survive the inside filter, per rank and step
        bwd_flop = (n_color * FLOP_BWD_COLOR_RAY + n_geom * FLOP_BWD_GEOM_RAY) * rays_rank
        fwd_flop = K * FLOP_FWD_RAY * rays_rank
        bwd_s = kms["decode_bwd"] * 1e-3
        fwd_s = kms["decode_fwd"] * 1e-3
        traffic = {}
        try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full capture (profiles/)
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f)
        except Exception:
            pass
        fwd_name = "k_decode_fwd_tc" if os.environ.get("NSB_TCGEN05", "0") not in ("", "0") else "k_decode_fwd"
        kern = {fwd_name: {"achieved": fwd_flop / fwd_s * 1e-12 if fwd_s > 0 else 0.0, "avg_launch_ms": kms["decode_fwd"] / K, "flop_per_launch": fwd_flop / K,
                           "gather_gbs": K * GATHER_BYTES_RAY * rays_rank / fwd_s * 1e-9 if fwd_s > 0 else 0.0},
                "k_decode_bwd": {"achieved": bwd_flop / bwd_s * 1e-12 if bwd_s > 0 else 0.0, "avg_launch_ms": kms["decode_bwd"] / K, "flop_per_launch": bwd_flop / K}}
        dom = fwd_name if fwd_s >= bwd_s else "k_decode_bwd"       # the dominant kernel of the step
        ach = kern[dom]["achieved"]
        roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": ach / pk["bf16_tflops_sustained"], "traffic": traffic.get(dom), "peak_source": pk["source"] + " bf16 dense (sustained, kernel timed inside the step)",
                "note": "algorithmic FLOPs (SURVEY 8-d: 2 x MACs of the three decoders x 48 samples x rays surviving the inside filter) / CUDA-event time of the kernel; "
                        "the arithmetic is the fp32-grade fp16 two-way split (3 tensor-core MMAs per product, measured warp-MMA ceiling 556 TFLOP/s), so the design ceiling is 185 TFLOP/s algorithmic",
                "launches": K, "avg_launch_ms": kern[dom]["avg_launch_ms"], "kernels": kern,
                "kernel_ms_total": kms,
                "grid_sampling": {"kernel": "k_gather_only", "ms": gather_ms, "rays": RAYS_PER_GPU,
                                  "achieved_gbs": GATHER_BYTES_RAY * RAYS_PER_GPU / (gather_ms * 1e-3) * 1e-9 if gather_ms > 0 else 0.0,
                                  "frac_of_hbm_peak": GATHER_BYTES_RAY * RAYS_PER_GPU / (gather_ms * 1e-3) * 1e-9 / pk["hbm_gbs"] if gather_ms > 0 else 0.0,
                                  "l2_gather_peak_gbs_measured": 9100.0, "note": "all rays of the step (no inside filter), 3 grids x 8 corners x 128 B per sample; grids (11.2 MB) are L2-resident"}}
        try:
            cpu = cpu_baseline_sample(nsb) if world == 1 and not args.no_cpu_baseline else None
        except Exception as ex:  # the checker is optional for the bench line
            cpu = {"error": repr(ex)}
        out = {"metric": "mapping rays/s (fwd+bwd, 48 samples/ray)", "value": value, "unit": "rays/s", "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": t_dev / K * 1e3, "ms_per_step_by_stage": {"geometry": ms_geom, "color": ms_color}, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic",
               "config": {"workload": "mapper_iteration_5000rays_60iters_keyframe", "rays_per_gpu": RAYS_PER_GPU, "global_rays": n_global,
                          "samples_per_ray": 48, "frames": N_FRAMES, "schedule": "step i = iteration i%60 of optimize_map (37 geometry + 23 colour)",
                          "mma": "fp16 two-way split (3 products) on mma.sync m16n8k16, fp32 accumulate: fp32-grade", "l2": "256 MiB flush write between timed steps", "parallelism": ("rays sharded x%d, " % world) + ("single GPU" if world == 1 else "fused reduce-scatter + Adam + all-gather kernel over NVLink peer memory" if args.comm == "p2p" else "NCCL all-reduce of grads + Adam"),
                          "inside_fraction": frac_in},
               "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "tracking": tracking,
               "loss_first_last": [float(losses[0]), float(losses[len(losses) - 1])]}
        print(json.dumps(out), flush=True)
    e.close()
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"], help="multi-GPU optimiser step: fused peer-memory kernel (default) or ncclAllReduce + Adam")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
