"""nice-slam-cpp_b200 -- B200-native NICE-SLAM ray-rendering hot path.

This package is only a thin ctypes binding of the C ABI declared in include/nsb.h (libnsb.so, hand-written
sm_100a CUDA + plain C++ host code, no libtorch).  It mirrors the reference's operator interface
(Renderer::render_batch_ray, Mapper::optimize_map iterations, Tracker::optimize_cam_in_batch, get_samples,
quad2rotation ...) for the Python tests and bench.py; C++ callers use include/nsb/*.h instead.

There is NO CPU fallback: loading fails loudly if libnsb.so has not been built
(python nice-slam-cpp_b200/build.py) and `Engine(...)` raises if no sm_100 GPU is present.
The package name contains hyphens; import it with importlib.import_module("nice-slam-cpp_b200").
"""
import ctypes as C
import os
import numpy as np

from . import synthetic  # noqa: F401  (numpy-only input generator)

_HERE = os.path.dirname(os.path.abspath(__file__))
LEVELS = ("coarse", "middle", "fine", "color")
STAGE = {"coarse": 0, "middle": 1, "fine": 2, "color": 3}
F_GRID, F_WGRAD, F_RAY = 1, 2, 4
MAP_COARSE, MAP_FIX_COLOR, MAP_NO_FRUSTUM, MAP_ZERO_RATIOS, MAP_COLOR_REFINE = 1, 2, 4, 8, 14          # nsb_mapping_begin_ex flags
RAYDIR_REFERENCE, RAYDIR_PINHOLE = 0, 1
DISTNORM_PER_RAY, DISTNORM_REFERENCE = 0, 1
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)


class Config(C.Structure):
    """Mirror of struct nsb_config (include/nsb.h)."""
    _fields_ = [
        ("H", C.c_int), ("W", C.c_int), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("bound", (C.c_float * 2) * 3),
        ("grid_len", C.c_float * 4), ("coarse_bound_enlarge", C.c_int), ("grid_dim", (C.c_int * 3) * 4), ("c_dim", C.c_int),
        ("n_samples", C.c_int), ("n_surface", C.c_int), ("occupancy", C.c_int), ("dist_norm", C.c_int), ("raydir", C.c_int),
        ("mapping_pixels", C.c_int), ("mapping_iters", C.c_int), ("mapping_iters_first", C.c_int),
        ("mapping_window_size", C.c_int), ("keyframe_every", C.c_int),
        ("middle_iter_ratio", C.c_float), ("fine_iter_ratio", C.c_float), ("second_stage", C.c_int),
        ("lr_factor", C.c_float), ("lr_first_factor", C.c_float), ("stage_lr", (C.c_float * 5) * 4),
        ("mapping_w_color_loss", C.c_float),
        ("fix_fine", C.c_int), ("fix_color", C.c_int), ("frustum_feature_selection", C.c_int), ("BA", C.c_int),
        ("BA_cam_lr", C.c_float),
        ("tracking_lr", C.c_float), ("tracking_iters", C.c_int), ("tracking_pixels", C.c_int),
        ("ignore_edge_W", C.c_int), ("ignore_edge_H", C.c_int), ("handle_dynamic", C.c_int),
        ("use_color_in_tracking", C.c_int), ("w_color_loss", C.c_float),
        ("precision", C.c_int), ("max_rays", C.c_int), ("max_frames", C.c_int), ("color_refine", C.c_int),
    ]


def lib_path(variant=""):
    return os.path.join(_HERE, "libnsb%s.so" % ("_" + variant if variant else ""))


_LIBS = {}


def load_library(variant=""):
    """dlopen libnsb.so and declare the prototypes.  Raises if the library is missing: no fallback."""
    if variant == "":
        variant = os.environ.get("NSB_VARIANT", "")       # build-variant experiments (build.py VARIANTS)
    if variant in _LIBS:
        return _LIBS[variant]
    path = lib_path(variant)
    if not os.path.exists(path):
        raise RuntimeError("%s not found: build it with `python nice-slam-cpp_b200/build.py` "
                           "(this package has no CPU / PyTorch fallback)" % path)
    L = C.CDLL(path)
    L.nsb_config_default.argtypes = [C.POINTER(Config)]
    L.nsb_config_default.restype = None
    L.nsb_config_load_yaml.argtypes = [C.POINTER(Config), C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    L.nsb_grid_dims.argtypes = [C.POINTER(Config), C.c_int, _ip, _ip, _ip]
    L.nsb_grid_dims.restype = None
    L.nsb_decoder_count.argtypes = [C.c_int, C.c_int]
    L.nsb_decoder_count.restype = C.c_int64
    L.nsb_create.argtypes = [C.POINTER(Config), C.c_int, C.POINTER(C.c_void_p)]
    L.nsb_destroy.argtypes = [C.c_void_p]
    L.nsb_destroy.restype = None
    L.nsb_last_error.argtypes = [C.c_void_p]
    L.nsb_last_error.restype = C.c_char_p
    L.nsb_build_info.restype = C.c_char_p
    L.nsb_stream.argtypes = [C.c_void_p]
    L.nsb_stream.restype = C.c_void_p
    L.nsb_launch_count.argtypes = [C.c_void_p, C.c_int]
    L.nsb_launch_count.restype = C.c_int64
    for name in ("nsb_quad2rotation", "nsb_get_camera_from_tensor", "nsb_get_tensor_from_camera"):
        getattr(L, name).argtypes = [_fp, _fp]
        getattr(L, name).restype = None
    v = C.c_void_p
    L.nsb_synchronize.argtypes = [v]
    L.nsb_set_grid.argtypes = [v, C.c_int, _fp]
    L.nsb_get_grid.argtypes = [v, C.c_int, _fp]
    L.nsb_get_grid_grad.argtypes = [v, C.c_int, _fp]
    L.nsb_set_decoder.argtypes = [v, C.c_int, _fp, C.c_int64]
    L.nsb_get_decoder.argtypes = [v, C.c_int, _fp, C.c_int64]
    L.nsb_get_decoder_grad.argtypes = [v, C.c_int, _fp, C.c_int64]
    L.nsb_set_ttables.argtypes = [v, _fp, _fp]
    L.nsb_set_voxel_mask.argtypes = [v, C.c_int, _u8p]
    L.nsb_frustum_mask.argtypes = [v, C.c_int, _fp, C.c_int, _u8p, C.c_int]
    L.nsb_set_frame.argtypes = [v, C.c_int, _fp, _fp, _fp]
    L.nsb_set_frame_pose.argtypes = [v, C.c_int, _fp]
    L.nsb_seed.argtypes = [v, C.c_uint64]
    L.nsb_get_samples.argtypes = [v, C.c_int, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i64p, _fp, _fp, _fp, _fp, _u8p, _i64p]
    L.nsb_render_batch_ray.argtypes = [v, C.c_int, C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, _fp]
    L.nsb_render_batch_ray_dev.argtypes = [v, C.c_int, C.c_int, v, v, v, v, v, v, v]
    L.nsb_eval_points.argtypes = [v, C.c_int, C.c_int, _fp, _fp]
    L.nsb_get_last_zvals.argtypes = [v, C.c_int, C.c_int, _fp]
    L.nsb_render_vjp.argtypes = [v, C.c_int, C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int, _fp, _fp]
    L.nsb_mapping_begin.argtypes = [v, C.c_int, _ip, C.c_int, C.c_float]
    L.nsb_mapping_iter.argtypes = [v, C.c_int, _i64p, _fp]
    L.nsb_mapping_iter_async.argtypes = [v, C.c_int, _i64p]
    L.nsb_mapping_losses.argtypes = [v, C.c_int, C.c_int, _fp, _ip]
    L.nsb_mapping_set_index_pool.argtypes = [v, _i64p, C.c_int, C.c_int]
    L.nsb_optimize_map.argtypes = [v, C.c_int, _ip, C.c_int, C.c_float, _fp]
    L.nsb_tracking_begin.argtypes = [v, C.c_int, _fp]
    L.nsb_tracking_iter.argtypes = [v, _i64p, _fp, _fp]
    L.nsb_tracking_get_camera.argtypes = [v, _fp]
    L.nsb_comm_unique_id.argtypes = [C.c_char_p]
    L.nsb_comm_init.argtypes = [v, C.c_char_p, C.c_int, C.c_int]
    L.nsb_comm_p2p_export.argtypes = [v, C.c_char_p]
    L.nsb_comm_p2p_import.argtypes = [v, C.c_char_p, C.c_int, C.c_int]
    L.nsb_comm_rank_world.argtypes = [v, _ip, _ip]
    L.nsb_set_profiling.argtypes = [v, C.c_int]
    L.nsb_get_kernel_ms.argtypes = [v, _fp]
    L.nsb_bench_gather.argtypes = [v, C.c_int, _fp]
    L.nsb_mapping_begin_ba.argtypes = [v, C.c_int, _ip, C.c_int, C.c_float, C.c_uint32]
    L.nsb_mapping_end.argtypes = [v, _fp]
    L.nsb_mapping_cam_grads.argtypes = [v, _fp]
    L.nsb_set_frame_async.argtypes = [v, C.c_int, _fp, _fp, _fp]
    L.nsb_frames_ready.argtypes = [v]
    L.nsb_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    L.nsb_host_free.argtypes = [C.c_void_p]
    L.nsb_save_checkpoint.argtypes = [v, C.c_char_p]
    L.nsb_load_checkpoint.argtypes = [v, C.c_char_p]
    L.nsb_ray_order_source.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _ip]
    L.nsb_render_img.argtypes = [v, C.c_int, _fp, C.c_int, C.c_int, _fp, _fp, _fp]
    L.nsb_keyframe_selection_overlap.argtypes = [v, C.c_int, _fp, C.c_int, _fp, C.c_int, _i64p, C.c_int, C.c_int, _ip, _ip, _fp]
    L.nsb_get_frame_pose.argtypes = [v, C.c_int, _fp]
    L.nsb_debug_counters.argtypes = [v, C.POINTER(C.c_uint64)]
    L.nsb_eval_points_dev.argtypes = [v, C.c_int, C.c_int, v, v]
    L.nsb_eval_lattice.argtypes = [v, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp]
    L.nsb_comm_p2p_stats.argtypes = [v, C.POINTER(C.c_double), C.c_int]
    L.nsb_config_reference_literal.argtypes = [C.POINTER(Config)]
    L.nsb_config_reference_literal.restype = None
    L.nsb_mapping_begin_ex.argtypes = [v, C.c_int, _ip, C.c_int, C.c_float, C.c_uint32, C.c_int]
    L.nsb_mapping_capture_grads.argtypes = [v, C.c_int]
    L.nsb_get_captured_grid_grad.argtypes = [v, C.c_int, _fp]
    L.nsb_get_captured_decoder_grad.argtypes = [v, C.c_int, _fp, C.c_int64]
    _LIBS[variant] = L
    return L


EXPORTS = [  # every symbol include/nsb.h declares (checked by tests/test_abi.py)
    "nsb_config_default", "nsb_config_load_yaml", "nsb_grid_dims", "nsb_decoder_count", "nsb_create", "nsb_destroy",
    "nsb_last_error", "nsb_abi_version", "nsb_build_info", "nsb_synchronize", "nsb_stream", "nsb_set_grid", "nsb_get_grid",
    "nsb_get_grid_grad", "nsb_set_decoder", "nsb_get_decoder", "nsb_get_decoder_grad", "nsb_set_ttables",
    "nsb_set_voxel_mask", "nsb_frustum_mask", "nsb_set_frame", "nsb_set_frame_pose", "nsb_quad2rotation", "nsb_get_camera_from_tensor",
    "nsb_get_tensor_from_camera", "nsb_seed", "nsb_get_samples", "nsb_render_batch_ray", "nsb_render_batch_ray_dev",
    "nsb_eval_points", "nsb_get_last_zvals", "nsb_render_vjp", "nsb_mapping_begin", "nsb_mapping_iter",
    "nsb_mapping_iter_async", "nsb_mapping_losses", "nsb_mapping_set_index_pool", "nsb_optimize_map", "nsb_tracking_begin", "nsb_tracking_iter",
    "nsb_tracking_get_camera", "nsb_comm_unique_id", "nsb_comm_init", "nsb_comm_rank_world", "nsb_launch_count",
    "nsb_set_profiling", "nsb_get_kernel_ms", "nsb_debug_counters", "nsb_bench_gather",
    "nsb_mapping_begin_ba", "nsb_mapping_end", "nsb_get_frame_pose", "nsb_mapping_cam_grads", "nsb_keyframe_selection_overlap", "nsb_render_img", "nsb_ray_order_source", "nsb_set_frame_async", "nsb_frames_ready", "nsb_host_alloc", "nsb_host_free", "nsb_save_checkpoint", "nsb_load_checkpoint", "nsb_comm_p2p_export", "nsb_comm_p2p_import",
    "nsb_eval_points_dev", "nsb_eval_lattice", "nsb_comm_p2p_stats", "nsb_config_reference_literal", "nsb_mapping_begin_ex", "nsb_mapping_capture_grads", "nsb_get_captured_grid_grad", "nsb_get_captured_decoder_grad",
]


def default_config(variant=""):
    cfg = Config()
    load_library(variant).nsb_config_default(C.byref(cfg))
    return cfg


def reference_literal_config(variant=""):
    """Defaults with the two degenerate formulas of the transliteration kept as written (utils.h:44-47 ray directions, utils.h:153
    norm): the mode the bit-exact parity tests against the reference's own compiled code run in."""
    cfg = default_config(variant)
    load_library(variant).nsb_config_reference_literal(C.byref(cfg))
    return cfg


def load_yaml_config(nice_slam_yaml, dataset_yaml, variant=""):
    cfg = default_config(variant)
    err = C.create_string_buffer(512)
    rc = load_library(variant).nsb_config_load_yaml(C.byref(cfg), nice_slam_yaml.encode() if nice_slam_yaml else None,
                                                    dataset_yaml.encode() if dataset_yaml else None, err, 512)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    return cfg


def _f(a):
    return None if a is None else a.ctypes.data_as(_fp)


def _c(a, dt=np.float32):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


# host-side camera helpers (utils.h:174-231)
def quad2rotation(q4, variant=""):
    q = _c(q4); R = np.empty(9, np.float32)
    load_library(variant).nsb_quad2rotation(_f(q), _f(R))
    return R.reshape(3, 3)


def get_camera_from_tensor(cam7, variant=""):
    q = _c(cam7); RT = np.empty(12, np.float32)
    load_library(variant).nsb_get_camera_from_tensor(_f(q), _f(RT))
    return RT.reshape(3, 4)


def get_tensor_from_camera(c2w, variant=""):
    m = _c(c2w).reshape(-1)
    if m.size == 12:
        m = np.concatenate([m, np.array([0, 0, 0, 1], np.float32)])
    out = np.empty(7, np.float32)
    load_library(variant).nsb_get_tensor_from_camera(_f(m), _f(out))
    return out


class Engine:
    """One nsb_ctx (one GPU).  Method names follow the reference's API."""

    def __init__(self, cfg=None, device=0, variant=""):
        self.lib = load_library(variant)
        self.cfg = cfg if cfg is not None else default_config(variant)
        self.h = C.c_void_p()
        rc = self.lib.nsb_create(C.byref(self.cfg), device, C.byref(self.h))
        if rc != 0:
            msg = self.lib.nsb_last_error(self.h).decode() if self.h else "nsb_create failed"
            if self.h:
                self.lib.nsb_destroy(self.h)
                self.h = C.c_void_p()
            raise RuntimeError("libnsb: " + msg)
        self.grid_shape = {}
        for i, lv in enumerate(LEVELS):
            z, y, x = C.c_int(), C.c_int(), C.c_int()
            self.lib.nsb_grid_dims(C.byref(self.cfg), i, C.byref(z), C.byref(y), C.byref(x))
            self.grid_shape[lv] = (1, self.cfg.c_dim, z.value, y.value, x.value)
        self.S = self.cfg.n_samples + self.cfg.n_surface

    def close(self):
        if getattr(self, "h", None):
            self.lib.nsb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError("libnsb: " + self.lib.nsb_last_error(self.h).decode())

    # ---- state
    def set_grid(self, level, g):
        g = _c(g); assert g.shape == self.grid_shape[level], (g.shape, self.grid_shape[level])
        self._ck(self.lib.nsb_set_grid(self.h, STAGE[level], _f(g)))

    def get_grid(self, level):
        out = np.empty(self.grid_shape[level], np.float32)
        self._ck(self.lib.nsb_get_grid(self.h, STAGE[level], _f(out)))
        return out

    def get_grid_grad(self, level):
        out = np.empty(self.grid_shape[level], np.float32)
        self._ck(self.lib.nsb_get_grid_grad(self.h, STAGE[level], _f(out)))
        return out

    def decoder_count(self, which):
        return int(self.lib.nsb_decoder_count(STAGE[which], self.cfg.c_dim))

    def set_decoder(self, which, flat):
        flat = _c(flat)
        self._ck(self.lib.nsb_set_decoder(self.h, STAGE[which], _f(flat), flat.size))

    def get_decoder(self, which):
        out = np.empty(self.decoder_count(which), np.float32)
        self._ck(self.lib.nsb_get_decoder(self.h, STAGE[which], _f(out), out.size))
        return out

    def get_decoder_grad(self, which):
        out = np.empty(self.decoder_count(which), np.float32)
        self._ck(self.lib.nsb_get_decoder_grad(self.h, STAGE[which], _f(out), out.size))
        return out

    def set_model(self, grids, decoders):
        for lv in LEVELS:
            if lv in grids:
                self.set_grid(lv, grids[lv])
            if lv in decoders:
                self.set_decoder(lv, decoders[lv])

    def set_ttables(self, t32, t16):
        a, b = _c(t32), _c(t16); assert a.size == 32 and b.size == 16
        self._ck(self.lib.nsb_set_ttables(self.h, _f(a), _f(b)))

    def set_voxel_mask(self, level, mask):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self._ck(self.lib.nsb_set_voxel_mask(self.h, STAGE[level], None if m is None else m.ctypes.data_as(_u8p)))

    def frustum_mask(self, slot, level, c2w=None, install=False):
        """Mapper::get_mask_from_c2w (Mapper.cpp:42-130) on the GPU -> (Z, Y, X) bool."""
        p = _c(c2w)
        out = np.empty(self.grid_shape[level][2:], np.uint8)
        self._ck(self.lib.nsb_frustum_mask(self.h, slot, _f(p), STAGE[level], out.ctypes.data_as(_u8p), int(install)))
        return out.astype(bool)

    def set_frame(self, slot, depth, color, c2w):
        d, c, p = _c(depth), _c(color), _c(c2w)
        self._ck(self.lib.nsb_set_frame(self.h, slot, _f(d), _f(c), _f(p)))

    def set_frame_async(self, slot, depth, color, c2w):
        """Asynchronous ingest (own copy stream); the arrays must stay alive until frames_ready() or the next synchronising call."""
        d, c, p = _c(depth), _c(color), _c(c2w)
        self._pending_frames = (d, c, p)
        self._ck(self.lib.nsb_set_frame_async(self.h, slot, _f(d), _f(c), _f(p)))

    def frames_ready(self):
        self._ck(self.lib.nsb_frames_ready(self.h))
        self._pending_frames = None

    def save_checkpoint(self, path):
        self._ck(self.lib.nsb_save_checkpoint(self.h, str(path).encode()))

    def load_checkpoint(self, path):
        self._ck(self.lib.nsb_load_checkpoint(self.h, str(path).encode()))

    def seed(self, s):
        self._ck(self.lib.nsb_seed(self.h, C.c_uint64(s)))

    def synchronize(self):
        self._ck(self.lib.nsb_synchronize(self.h))

    # ---- get_samples (utils.h:141-146) + inside filter
    def get_samples(self, slot, H0, H1, W0, W1, n, idx=None, c2w=None):
        ro = np.empty((n, 3), np.float32); rd = np.empty((n, 3), np.float32); gd = np.empty(n, np.float32)
        gc = np.empty((n, 3), np.float32); ins = np.empty(n, np.uint8); io = np.empty(n, np.int64)
        ix = None if idx is None else np.ascontiguousarray(idx, np.int64)
        p = _c(c2w)
        self._ck(self.lib.nsb_get_samples(self.h, slot, _f(p), H0, H1, W0, W1, n, None if ix is None else ix.ctypes.data_as(_i64p),
                                          _f(ro), _f(rd), _f(gd), _f(gc), ins.ctypes.data_as(_u8p), io.ctypes.data_as(_i64p)))
        return ro, rd, gd, gc, ins.astype(bool), io

    # ---- Renderer::render_batch_ray (Renderer.h:13; note rays_d precedes rays_o)
    def render_batch_ray(self, rays_d, rays_o, stage, gt_depth, want_weights=True):
        rd, ro, gd = _c(rays_d), _c(rays_o), _c(gt_depth)
        n = rd.shape[0]
        S = self.S if gd is not None else self.cfg.n_samples
        rgb = np.empty((n, 3), np.float32); depth = np.empty(n, np.float32); var = np.empty(n, np.float32)
        w = np.empty((n, S), np.float32) if want_weights else None
        self._ck(self.lib.nsb_render_batch_ray(self.h, STAGE[stage], n, _f(rd), _f(ro), _f(gd), _f(rgb), _f(depth), _f(var), _f(w)))
        return rgb, depth, var, w

    def render_batch_ray_dev(self, stage, n, d_rays_d, d_rays_o, d_gt_depth, d_rgb, d_depth, d_var, d_weights):
        """Device-pointer form (ints): enqueues on the context's stream, no host synchronisation."""
        v = lambda p: None if not p else C.c_void_p(int(p))
        self._ck(self.lib.nsb_render_batch_ray_dev(self.h, STAGE[stage], int(n), v(d_rays_d), v(d_rays_o), v(d_gt_depth), v(d_rgb), v(d_depth), v(d_var), v(d_weights)))

    def render_img(self, slot, stage, use_gt_depth=True, c2w=None, want_outputs=True):
        """Dense render of every pixel of a resident frame (upstream render_img) -> rgb (H,W,3), depth (H,W), var (H,W).
        want_outputs=False renders without the image read-back (timing of the device work alone)."""
        H, W = self.cfg.H, self.cfg.W
        p = None if c2w is None else _c(np.asarray(c2w, np.float32).reshape(-1))
        if not want_outputs:
            self._ck(self.lib.nsb_render_img(self.h, slot, _f(p), STAGE[stage], int(use_gt_depth), None, None, None))
            return None
        rgb = np.empty((H, W, 3), np.float32); depth = np.empty((H, W), np.float32); var = np.empty((H, W), np.float32)
        self._ck(self.lib.nsb_render_img(self.h, slot, _f(p), STAGE[stage], int(use_gt_depth), _f(rgb), _f(depth), _f(var)))
        return rgb, depth, var

    def last_zvals(self, n, S=None):
        S = S or self.S
        z = np.empty((n, S), np.float32)
        self._ck(self.lib.nsb_get_last_zvals(self.h, n, S, _f(z)))
        return z

    # ---- Renderer::eval_points (Renderer.h:12)
    def eval_points(self, pts, stage):
        p = _c(pts); raw = np.empty((p.shape[0], 4), np.float32)
        self._ck(self.lib.nsb_eval_points(self.h, STAGE[stage], p.shape[0], _f(p), _f(raw)))
        return raw

    def eval_lattice(self, stage, nx, ny, nz, lo=None, hi=None, want_raw=False):
        """eval_points over a regular lattice generated on the device (mesh-extraction query) -> occupancy (ny, nx, nz)
        [numpy.meshgrid(x, y, z) order], and raw (ny, nx, nz, 4) when want_raw."""
        n = nx * ny * nz
        occ = np.empty(n, np.float32); raw = np.empty((n, 4), np.float32) if want_raw else None
        l = None if lo is None else _c(lo); h = None if hi is None else _c(hi)
        self._ck(self.lib.nsb_eval_lattice(self.h, STAGE[stage], nx, ny, nz, _f(l), _f(h), _f(raw), _f(occ)))
        occ = occ.reshape(ny, nx, nz)
        return (occ, raw.reshape(ny, nx, nz, 4)) if want_raw else occ

    # ---- loss.backward() through render_batch_ray
    def render_vjp(self, rays_d, rays_o, stage, gt_depth, g_rgb, g_depth, g_var, flags=F_GRID | F_WGRAD | F_RAY):
        rd, ro, gd = _c(rays_d), _c(rays_o), _c(gt_depth)
        n = rd.shape[0]
        a, b, c_ = _c(g_rgb), _c(g_depth), _c(g_var)
        drd = np.zeros((n, 3), np.float32); dro = np.zeros((n, 3), np.float32)
        self._ck(self.lib.nsb_render_vjp(self.h, STAGE[stage], n, _f(rd), _f(ro), _f(gd), _f(a), _f(b), _f(c_), flags, _f(drd), _f(dro)))
        out = {"rays_d": drd, "rays_o": dro}
        if flags & F_GRID:
            for lv in ("middle", "fine", "color"):
                out["grid_" + lv] = self.get_grid_grad(lv)
        if flags & F_WGRAD:
            out["dec_color"] = self.get_decoder_grad("color")
        return out

    # ---- Mapper::optimize_map inner loop (Mapper.cpp:330-465)
    def mapping_begin(self, slots, n_iters, lr_factor=1.0, ba_mask=0, flags=0):
        """ba_mask: bit f set -> the pose of slots[f] is optimised with the map (bundle adjustment, Mapper.cpp:305-329).
        flags: MAP_COARSE (coarse mapper), MAP_FIX_COLOR / MAP_NO_FRUSTUM (color_refine, Mapper.cpp:505-513)."""
        s = np.ascontiguousarray(slots, np.int32)
        self._n_map_frames = len(s)
        self._ck(self.lib.nsb_mapping_begin_ex(self.h, len(s), s.ctypes.data_as(_ip), n_iters, C.c_float(lr_factor), C.c_uint32(ba_mask), int(flags)))

    def mapping_capture_grads(self, on=True):
        self._ck(self.lib.nsb_mapping_capture_grads(self.h, int(on)))

    def captured_grads(self):
        """Gradient of the last iteration (before the optimiser step): grids in (1,C,Z,Y,X), the colour decoder flat."""
        out = {}
        for lv in LEVELS:
            g = np.empty(self.grid_shape[lv], np.float32)
            self._ck(self.lib.nsb_get_captured_grid_grad(self.h, STAGE[lv], _f(g)))
            out["grid_" + lv] = g
        d = np.empty(self.decoder_count("color"), np.float32)
        self._ck(self.lib.nsb_get_captured_decoder_grad(self.h, STAGE["color"], _f(d), d.size))
        out["dec_color"] = d
        return out

    def mapping_end(self):
        """BA write-back (Mapper.cpp:467-489); returns the (n_frames, 7) camera vectors."""
        cams = np.zeros((self._n_map_frames, 7), np.float32)
        self._ck(self.lib.nsb_mapping_end(self.h, _f(cams)))
        return cams

    def mapping_cam_grads(self):
        g = np.zeros((self._n_map_frames, 7), np.float32)
        self._ck(self.lib.nsb_mapping_cam_grads(self.h, _f(g)))
        return g

    def keyframe_selection_overlap(self, cur_slot, kf_c2ws, k_overlap, cur_c2w=None, idx=None, pixels=100, n_samples=16):
        """Mapper::keyframe_selection_overlap (Mapper.cpp:132-196) -> (selected keyframe indices, per-keyframe fractions)."""
        kf = _c(np.asarray(kf_c2ws, np.float32).reshape(-1, 16))
        n_kf = kf.shape[0]
        sel = np.zeros(max(n_kf, 1), np.int32); n_sel = C.c_int(0); pct = np.zeros(max(n_kf, 1), np.float32)
        ix = None if idx is None else np.ascontiguousarray(idx, np.int64)
        cur = None if cur_c2w is None else _c(np.asarray(cur_c2w, np.float32).reshape(-1))
        self._ck(self.lib.nsb_keyframe_selection_overlap(self.h, cur_slot, _f(cur), n_kf, _f(kf), k_overlap, None if ix is None else ix.ctypes.data_as(_i64p),
                                                         pixels, n_samples, sel.ctypes.data_as(_ip), C.byref(n_sel), _f(pct)))
        return sel[:n_sel.value].tolist(), pct[:n_kf]

    def get_frame_pose(self, slot):
        m = np.empty((3, 4), np.float32)
        self._ck(self.lib.nsb_get_frame_pose(self.h, slot, _f(m)))
        return m

    def mapping_iter(self, it, idx=None, sync=True):
        ix = None if idx is None else np.ascontiguousarray(idx, np.int64)
        p = None if ix is None else ix.ctypes.data_as(_i64p)
        if not sync:
            self._ck(self.lib.nsb_mapping_iter_async(self.h, it, p))
            return None
        loss = C.c_float(0)
        self._ck(self.lib.nsb_mapping_iter(self.h, it, p, C.byref(loss)))
        return loss.value

    def mapping_losses(self, first, n):
        l = np.empty(n, np.float32); k = np.empty(n, np.int32)
        self._ck(self.lib.nsb_mapping_losses(self.h, first, n, _f(l), k.ctypes.data_as(_ip)))
        return l, k

    def mapping_set_index_pool(self, idx_all):
        if idx_all is None:
            self._ck(self.lib.nsb_mapping_set_index_pool(self.h, None, 0, 0)); return
        ix = np.ascontiguousarray(idx_all, np.int64)
        self._ck(self.lib.nsb_mapping_set_index_pool(self.h, ix.ctypes.data_as(_i64p), ix.shape[0], ix.shape[1]))

    def stream_ptr(self):
        return int(self.lib.nsb_stream(self.h))

    def optimize_map(self, slots, n_iters, lr_factor=1.0):
        s = np.ascontiguousarray(slots, np.int32); l = np.empty(n_iters, np.float32)
        self._ck(self.lib.nsb_optimize_map(self.h, len(s), s.ctypes.data_as(_ip), n_iters, C.c_float(lr_factor), _f(l)))
        return l

    # ---- Tracker::optimize_cam_in_batch (Tracker.cpp:41-89)
    def tracking_begin(self, slot, cam7):
        c = _c(cam7)
        self._ck(self.lib.nsb_tracking_begin(self.h, slot, _f(c)))

    def tracking_iter(self, idx=None, want_grad=True):
        ix = None if idx is None else np.ascontiguousarray(idx, np.int64)
        loss = C.c_float(0); g = np.zeros(7, np.float32)
        self._ck(self.lib.nsb_tracking_iter(self.h, None if ix is None else ix.ctypes.data_as(_i64p), C.byref(loss), _f(g) if want_grad else None))
        return loss.value, g

    def tracking_camera(self):
        c = np.empty(7, np.float32)
        self._ck(self.lib.nsb_tracking_get_camera(self.h, _f(c)))
        return c

    # ---- multi-GPU / instrumentation
    def comm_init(self, uid, rank, world):
        self._ck(self.lib.nsb_comm_init(self.h, uid, rank, world))

    def p2p_export(self):
        """192 bytes: CUDA IPC handles of this rank's gradient arena, parameter arena and flag block."""
        buf = C.create_string_buffer(192)
        self._ck(self.lib.nsb_comm_p2p_export(self.h, buf))
        return buf.raw

    def p2p_import(self, all_handles, rank, world):
        """all_handles: the ranks' p2p_export() bytes concatenated in rank order; host-barrier afterwards."""
        if all_handles is None:          # back to the NCCL path
            self._ck(self.lib.nsb_comm_p2p_import(self.h, None, rank, world)); return
        assert len(all_handles) == 192 * world
        self._ck(self.lib.nsb_comm_p2p_import(self.h, all_handles, rank, world))

    def p2p_times(self, reset=False):
        """Device-stamped timing of the fused exchange kernel (us): wait for the slowest rank vs the exchange itself."""
        o = (C.c_double * 8)()
        self._ck(self.lib.nsb_comm_p2p_stats(self.h, o, int(reset)))
        return {"last_wait_us": o[0], "last_kernel_us": o[1], "mean_wait_us": o[2], "mean_kernel_us": o[3], "mean_exchange_us": o[3] - o[2], "exchanges": int(o[4]),
                "mean_range_bytes": o[5], "nvlink_bytes_per_direction_model": o[6], "nvlink_gbs_per_direction": o[7],
                "model": "a rank reads its 1/W slice of the exchanged range from W-1 peers and writes the updated slice to W-1 peers: (W-1)/W of the range per direction "
                         "(an upper bound for the write direction: slots whose gradient sum, m and v are all zero are not rewritten)"}

    def bench_gather(self, reps=20):
        ms = C.c_float(0)
        self._ck(self.lib.nsb_bench_gather(self.h, reps, C.byref(ms)))
        return ms.value

    def launch_count(self, reset=False):
        return int(self.lib.nsb_launch_count(self.h, int(reset)))

    def set_profiling(self, on):
        self._ck(self.lib.nsb_set_profiling(self.h, int(on)))

    def kernel_ms(self):
        ms = np.zeros(7, np.float32)
        self._ck(self.lib.nsb_get_kernel_ms(self.h, _f(ms)))
        return dict(zip(("sample", "decode_fwd", "composite", "decode_bwd", "wgrad", "adam", "comm"), ms.tolist()))


def comm_unique_id(variant=""):
    buf = C.create_string_buffer(128)
    if load_library(variant).nsb_comm_unique_id(buf) != 0:
        raise RuntimeError("libnsb: ncclGetUniqueId failed (libnccl.so.2 not found?)")
    return buf.raw
