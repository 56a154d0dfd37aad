"""Deterministic synthetic inputs for the ray-rendering hot path (numpy only).

Datasets and pretrained decoders are offline (BASELINE.json north_star), so every test, bench and smoke
run draws its inputs from here.  Shapes, layouts and init scales follow the reference:

* frames: 480x640, fx=fy=360, cx=320, cy=240          config/cofusion.yaml:23-29
* grids (1, c_dim, Z, Y, X) fp32, normal_(0, 0.01), fine normal_(0, 1e-4)     src/main.cpp:33-78
* grid dims xyz_len / grid_len in fp32, truncated      src/main.cpp:38,48,59,70
* decoders: xavier_uniform(gain=sqrt 2) + zero bias    src/models/MLP.cpp:65-74
  fc_c default Linear init; B = 25 * randn(3, 93)       src/models/GaussianFFT.cpp:6

The flat decoder layout is the one documented in include/nsb.h.
"""
import math
import numpy as np

BOUND = np.array([[-4.5, 3.82], [-1.5, 2.02], [-3.0, 2.76]], dtype=np.float32)  # Renderer.cpp:15
GRID_LEN = {"coarse": 2.0, "middle": 0.32, "fine": 0.16, "color": 0.16}          # config/nice_slam.yaml:8-12
LEVELS = ("coarse", "middle", "fine", "color")
CAM = dict(H=480, W=640, fx=360.0, fy=360.0, cx=320.0, cy=240.0)
EMB, HID, CDIM = 93, 32, 32


def grid_dims(level, bound=BOUND, coarse_bound_enlarge=2):
    """(Z, Y, X) exactly as main.cpp computes them: fp32 divide, then truncation."""
    xyz_len = (bound[:, 1] - bound[:, 0]).astype(np.float32)
    if level == "coarse":
        v = xyz_len * np.float32(coarse_bound_enlarge) / np.float32(GRID_LEN[level])
    else:
        v = xyz_len / np.float32(GRID_LEN[level])
    v = v.astype(np.float32)
    return int(v[2]), int(v[1]), int(v[0])


def make_grids(seed=0, c_dim=CDIM, dims=None):
    """dict level -> (1, c_dim, Z, Y, X) fp32, channel-first like the reference."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    for lv in LEVELS:
        Z, Y, X = dims[lv] if dims else grid_dims(lv)
        std = 1e-4 if lv == "fine" else 0.01
        out[lv] = (rng.standard_normal((1, c_dim, Z, Y, X)) * std).astype(np.float32)
    return out


def decoder_count(which, E=EMB, H=HID, C=CDIM):
    if which == "coarse":
        K = [C, H, H, C + H, H]
        return sum(H * k + H for k in K) + H + 1
    Cd = 2 * C if which == "fine" else C
    O = 4 if which == "color" else 1
    K = [E, H, H, E + H, H]
    return 3 * E + sum(H * k + H for k in K) + 5 * (H * Cd + H) + O * H + O


def _xavier(rng, out_f, in_f, gain):
    a = gain * math.sqrt(6.0 / (in_f + out_f))
    return rng.uniform(-a, a, size=(out_f, in_f)).astype(np.float32)


def _linear_default(rng, out_f, in_f):
    a = 1.0 / math.sqrt(in_f)
    return (rng.uniform(-a, a, size=(out_f, in_f)).astype(np.float32),
            rng.uniform(-a, a, size=(out_f,)).astype(np.float32))


def make_decoder(which, seed=0, E=EMB, H=HID, C=CDIM, bias_scale=0.0):
    """Flat fp32 parameter vector of one decoder.  bias_scale > 0 perturbs the zero biases so that parity
    tests also exercise the bias terms (the reference initialises them to zero)."""
    rng = np.random.Generator(np.random.PCG64(1000 + seed * 7 + LEVELS.index(which)))
    g = math.sqrt(2.0)
    parts = []

    def bias(n):
        return (rng.standard_normal(n) * bias_scale).astype(np.float32)

    if which == "coarse":
        K = [C, H, H, C + H, H]
        for k in K:
            parts += [_xavier(rng, H, k, g).ravel(), bias(H)]
        parts += [_xavier(rng, 1, H, g).ravel(), bias(1)]
    else:
        Cd = 2 * C if which == "fine" else C
        O = 4 if which == "color" else 1
        K = [E, H, H, E + H, H]
        parts.append((rng.standard_normal((3, E)) * 25.0).astype(np.float32).ravel())
        for k in K:
            parts += [_xavier(rng, H, k, g).ravel(), bias(H)]
        for _ in range(5):
            w, b = _linear_default(rng, H, Cd)
            parts += [w.ravel(), b]
        parts += [_xavier(rng, O, H, g).ravel(), bias(O)]
    flat = np.concatenate(parts).astype(np.float32)
    assert flat.size == decoder_count(which, E, H, C), (which, flat.size)
    return flat


def make_decoders(seed=0, bias_scale=0.0):
    return {w: make_decoder(w, seed, bias_scale=bias_scale) for w in LEVELS}


def yaw_pose(deg, t=(-0.34, 0.26, -0.12)):
    """4x4 c2w: rotation about +y by `deg`, translation at the bound centre."""
    a = math.radians(deg)
    c2w = np.eye(4, dtype=np.float32)
    c2w[0, 0], c2w[0, 2], c2w[2, 0], c2w[2, 2] = math.cos(a), math.sin(a), -math.sin(a), math.cos(a)
    c2w[:3, 3] = t
    return c2w


def make_frame(seed=0, H=CAM["H"], W=CAM["W"], zero_frac=0.02):
    """depth (H,W) ~ U[0.5,3) with `zero_frac` of the pixels forced to 0, colour (H,W,3) ~ U[0,1)."""
    rng = np.random.Generator(np.random.PCG64(5000 + seed))
    depth = rng.uniform(0.5, 3.0, size=(H, W)).astype(np.float32)
    depth[rng.uniform(size=(H, W)) < zero_frac] = 0.0
    color = rng.uniform(0.0, 1.0, size=(H, W, 3)).astype(np.float32)
    return depth, color


def make_frames(n, seed=0, H=CAM["H"], W=CAM["W"]):
    yaws = [0.0, 15.0, -15.0, 30.0, -30.0, 45.0, -45.0, 60.0]
    depths, colors, poses = [], [], []
    for f in range(n):
        d, c = make_frame(seed * 100 + f, H, W)
        depths.append(d); colors.append(c); poses.append(yaw_pose(yaws[f % len(yaws)]))
    return np.stack(depths), np.stack(colors), np.stack(poses)


def linspace_sym(n):
    """torch::linspace(0, 1, n) in fp32 with the symmetric formula ATen's scalar kernel uses
    (SURVEY.md 8-A.2 item 11).  The vectorised ATen kernels may differ in the last bit for some n, which
    is why the t-tables are inputs of the C ABI rather than recomputed on the device."""
    step = np.float32(1.0) / np.float32(n - 1)
    out = np.empty(n, dtype=np.float32)
    for i in range(n):
        out[i] = np.float32(i) * step if i < n // 2 else np.float32(1.0) - np.float32(n - 1 - i) * step
    return out


def mt19937_indices(seed_or_state, n, high):
    """`torch::randint(high, {n})` on the CPU generator == std::mt19937(seed)() % high, the stream
    continuing across calls (utils.h:32).  Pass an int seed or a RandomState to continue a stream."""
    rs = seed_or_state if isinstance(seed_or_state, np.random.RandomState) else np.random.RandomState(int(seed_or_state) & 0xFFFFFFFF)
    raw = rs.randint(0, 2 ** 32, size=n, dtype=np.uint64)
    return (raw % np.uint64(high)).astype(np.int64)
