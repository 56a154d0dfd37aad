"""TEST INFRASTRUCTURE -- ctypes binding of oracle/_ref/libnsbref*.so (the reference's own Renderer.cpp +
utils.h, see oracle/Makefile and oracle/ref_harness.cpp).  Only tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs may import this."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LEVELS = ("coarse", "middle", "fine", "color")
_f = C.POINTER(C.c_float)


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f)


def lib_path(verbatim=False):
    return os.path.join(_HERE, "_ref", "libnsbref_verbatim.so" if verbatim else "libnsbref.so")


def available(verbatim=False):
    return os.path.exists(lib_path(verbatim))


class Ref:
    """One reference context: grids + decoders held as libtorch CPU tensors."""

    def __init__(self, grids, decoders, verbatim=False, c_dim=32, E=93, H=32):
        self.lib = C.CDLL(lib_path(verbatim))
        L = self.lib
        L.ref_create.restype = C.c_void_p
        L.ref_last_error.restype = C.c_char_p
        L.ref_last_error.argtypes = [C.c_void_p]
        L.ref_decoder_count.restype = C.c_int64
        self.h = C.c_void_p(L.ref_create())
        self.c_dim, self.E, self.H = c_dim, E, H
        self.grid_shape = {}
        for i, lv in enumerate(LEVELS):
            if lv in grids:
                self.set_grid(lv, grids[lv])
            if lv in decoders:
                self.set_decoder(lv, decoders[lv])

    def __del__(self):
        try:
            self.lib.ref_destroy(self.h)
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError("oracle/_ref: " + self.lib.ref_last_error(self.h).decode(errors="replace")[:2000])

    def set_threads(self, n):
        self.lib.ref_set_threads(C.c_int(n))

    def get_threads(self):
        return self.lib.ref_get_threads()

    def _dec_dims(self, which):
        Cd = 2 * self.c_dim if which == "fine" else self.c_dim
        O = 4 if which == "color" else 1
        return Cd, O

    def set_grid(self, level, g):
        g = np.ascontiguousarray(g, dtype=np.float32)
        _, Cc, Z, Y, X = g.shape
        self.grid_shape[level] = g.shape
        self._ck(self.lib.ref_set_grid(self.h, LEVELS.index(level), _p(g), Cc, Z, Y, X))

    def get_grid(self, level):
        out = np.empty(self.grid_shape[level], dtype=np.float32)
        self._ck(self.lib.ref_get_grid(self.h, LEVELS.index(level), _p(out)))
        return out

    def set_decoder(self, which, flat):
        flat = np.ascontiguousarray(flat, dtype=np.float32)
        Cd, O = self._dec_dims(which)
        self._dec_n = getattr(self, "_dec_n", {})
        self._dec_n[which] = flat.size
        self._ck(self.lib.ref_set_decoder(self.h, LEVELS.index(which), _p(flat), C.c_int64(flat.size), self.E, self.H, Cd, O))

    def get_decoder(self, which):
        out = np.empty(self._dec_n[which], dtype=np.float32)
        self._ck(self.lib.ref_get_decoder(self.h, LEVELS.index(which), _p(out)))
        return out

    def render_batch_ray(self, rays_d, rays_o, stage, gt_depth, n_out=48):
        n = rays_d.shape[0]
        rd = np.ascontiguousarray(rays_d, np.float32); ro = np.ascontiguousarray(rays_o, np.float32)
        gd = None if gt_depth is None else np.ascontiguousarray(gt_depth, np.float32)
        rgb = np.empty((n, 3), np.float32); depth = np.empty(n, np.float32); var = np.empty(n, np.float32)
        w = np.empty((n, n_out), np.float32)
        self._ck(self.lib.ref_render_batch_ray(self.h, stage.encode(), n, _p(rd), _p(ro), _p(gd), _p(rgb), _p(depth), _p(var), _p(w)))
        return rgb, depth, var, w

    def eval_points(self, pts, stage):
        pts = np.ascontiguousarray(pts, np.float32)
        raw = np.empty((pts.shape[0], 4), np.float32)
        self._ck(self.lib.ref_eval_points(self.h, stage.encode(), pts.shape[0], _p(pts), _p(raw)))
        return raw

    def render_vjp(self, rays_d, rays_o, stage, gt_depth, g_rgb, g_depth, g_var):
        n = rays_d.shape[0]
        rd = np.ascontiguousarray(rays_d, np.float32); ro = np.ascontiguousarray(rays_o, np.float32)
        gd = None if gt_depth is None else np.ascontiguousarray(gt_depth, np.float32)
        out = {}
        gptr, dptr = [], []
        for lv in LEVELS:
            if lv in self.grid_shape:
                out["grid_" + lv] = np.zeros(self.grid_shape[lv], np.float32); gptr.append(_p(out["grid_" + lv]))
            else:
                gptr.append(None)
        for lv in LEVELS:
            if lv in getattr(self, "_dec_n", {}):
                out["dec_" + lv] = np.zeros(self._dec_n[lv], np.float32); dptr.append(_p(out["dec_" + lv]))
            else:
                dptr.append(None)
        out["rays_d"] = np.zeros((n, 3), np.float32); out["rays_o"] = np.zeros((n, 3), np.float32)
        self._ck(self.lib.ref_render_vjp(self.h, stage.encode(), n, _p(rd), _p(ro), _p(gd),
                                         _p(np.ascontiguousarray(g_rgb, np.float32)), _p(np.ascontiguousarray(g_depth, np.float32)),
                                         _p(np.ascontiguousarray(g_var, np.float32)), *gptr, *dptr, _p(out["rays_d"]), _p(out["rays_o"])))
        return out

    def capture_grads(self, it):
        """Ask the next mapping_iters call for the gradients of iteration `it` (before optimizer.step): returns the dict of host
        arrays that call will fill (grid_middle / grid_fine / grid_color / dec_color)."""
        out = {"grid_" + lv: np.zeros(self.grid_shape[lv], np.float32) for lv in ("middle", "fine", "color")}
        out["dec_color"] = np.zeros(self._dec_n["color"], np.float32)
        self._caps = getattr(self, "_caps", []) + [out]      # keep the buffers alive
        self._ck(self.lib.ref_mapping_capture_grads(self.h, int(it), _p(out["grid_middle"]), _p(out["grid_fine"]), _p(out["grid_color"]), _p(out["dec_color"])))
        return out

    def mapping_iters(self, depths, colors, c2ws, cam, mapping_pixels, stage_ids, lr_table, w_color_loss=0.5,
                      fix_fine=True, fix_color=False, seed=0, masks=None):
        depths = np.ascontiguousarray(depths, np.float32); colors = np.ascontiguousarray(colors, np.float32)
        c2ws = np.ascontiguousarray(c2ws, np.float32)
        nf, H, W = depths.shape
        st = np.ascontiguousarray(stage_ids, np.int32)
        lr = np.ascontiguousarray(lr_table, np.float32).reshape(4, 5)
        losses = np.zeros(len(st), np.float32); n_in = np.zeros(len(st), np.int32); sec = C.c_double(0)
        mp = [None, None, None]
        keep = []
        if masks is not None:
            for k, lv in enumerate(("middle", "fine", "color")):
                if masks.get(lv) is not None:
                    m = np.ascontiguousarray(masks[lv], np.uint8); keep.append(m)
                    mp[k] = m.ctypes.data_as(C.POINTER(C.c_uint8))
        self._ck(self.lib.ref_mapping_iters(self.h, nf, H, W, C.c_float(cam["fx"]), C.c_float(cam["fy"]), C.c_float(cam["cx"]), C.c_float(cam["cy"]),
                                            _p(depths), _p(colors), _p(c2ws), mapping_pixels, len(st), st.ctypes.data_as(C.POINTER(C.c_int)),
                                            _p(lr), C.c_float(w_color_loss), int(fix_fine), int(fix_color), C.c_uint64(seed),
                                            mp[0], mp[1], mp[2], _p(losses), n_in.ctypes.data_as(C.POINTER(C.c_int)), C.byref(sec)))
        return losses, n_in, sec.value

    def tracking_iters(self, depth, color, cam7, cam, pixels, n_iters, lr, edge_h=20, edge_w=20, handle_dynamic=True,
                       use_color=True, w_color_loss=0.5, seed=0):
        depth = np.ascontiguousarray(depth, np.float32); color = np.ascontiguousarray(color, np.float32)
        H, W = depth.shape
        cam7 = np.array(cam7, np.float32).copy()
        losses = np.zeros(n_iters, np.float32); g0 = np.zeros(7, np.float32); n_in = np.zeros(n_iters, np.int32); sec = C.c_double(0)
        self._ck(self.lib.ref_tracking_iters(self.h, H, W, C.c_float(cam["fx"]), C.c_float(cam["fy"]), C.c_float(cam["cx"]), C.c_float(cam["cy"]),
                                             edge_h, edge_w, _p(depth), _p(color), _p(cam7), pixels, n_iters, C.c_float(lr),
                                             int(handle_dynamic), int(use_color), C.c_float(w_color_loss), C.c_uint64(seed),
                                             _p(losses), _p(g0), n_in.ctypes.data_as(C.POINTER(C.c_int)), C.byref(sec)))
        return cam7, losses, g0, n_in, sec.value


def get_samples(H0, H1, W0, W1, n, cam, c2w, depth, color, seed):
    lib = C.CDLL(lib_path())
    H, W = depth.shape
    ro = np.empty((n, 3), np.float32); rd = np.empty((n, 3), np.float32); gd = np.empty(n, np.float32); gc = np.empty((n, 3), np.float32)
    idx = np.empty(n, np.int64)
    rc = lib.ref_get_samples(H0, H1, W0, W1, n, H, W, C.c_float(cam["fx"]), C.c_float(cam["fy"]), C.c_float(cam["cx"]), C.c_float(cam["cy"]),
                             _p(np.ascontiguousarray(c2w, np.float32)), _p(np.ascontiguousarray(depth, np.float32)),
                             _p(np.ascontiguousarray(color, np.float32)), C.c_uint64(seed), _p(ro), _p(rd), _p(gd), _p(gc),
                             idx.ctypes.data_as(C.POINTER(C.c_int64)))
    assert rc == 0
    return ro, rd, gd, gc, idx


def raw2outputs(raw, z_vals, rays_d):
    lib = C.CDLL(lib_path())
    n, S = z_vals.shape
    rgb = np.empty((n, 3), np.float32); depth = np.empty(n, np.float32); var = np.empty(n, np.float32); w = np.empty((n, S), np.float32)
    rc = lib.ref_raw2outputs(n, S, _p(np.ascontiguousarray(raw, np.float32)), _p(np.ascontiguousarray(z_vals, np.float32)),
                             _p(np.ascontiguousarray(rays_d, np.float32)), _p(rgb), _p(depth), _p(var), _p(w))
    assert rc == 0
    return rgb, depth, var, w


def quad2rotation(q4):
    lib = C.CDLL(lib_path())
    R = np.empty(9, np.float32)
    assert lib.ref_quad2rotation(_p(np.ascontiguousarray(q4, np.float32)), _p(R)) == 0
    return R.reshape(3, 3)


def get_camera_from_tensor(cam7):
    lib = C.CDLL(lib_path())
    RT = np.empty(12, np.float32)
    assert lib.ref_get_camera_from_tensor(_p(np.ascontiguousarray(cam7, np.float32)), _p(RT)) == 0
    return RT.reshape(3, 4)
