"""CPU tests (no GPU): the C-ABI library loads, exports every symbol include/nsb.h declares, its host-side
logic (config defaults, YAML subset reader, grid dims, camera utilities, stage schedule helpers) matches the
reference's values, and the product path fails loudly without a GPU instead of falling back to anything."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
import nice_oracle as O


def test_header_symbols_exported(nsb):
    hdr = open(os.path.join(ROOT, "include", "nsb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(nsb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = nsb.load_library()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(nsb.EXPORTS)
    assert L.nsb_abi_version() == 2


def test_config_defaults_are_reference_values(nsb):
    c = nsb.default_config()
    assert (c.H, c.W, c.fx, c.fy, c.cx, c.cy) == (480, 640, 360.0, 360.0, 320.0, 240.0)       # cofusion.yaml:23-29
    assert [[round(float(v), 5) for v in r] for r in c.bound] == [[-4.5, 3.82], [-1.5, 2.02], [-3.0, 2.76]]   # Renderer.cpp:15
    assert (c.n_samples, c.n_surface) == (32, 16)                                               # Renderer.cpp:9-10
    assert (c.mapping_pixels, c.mapping_iters, c.mapping_iters_first) == (1000, 60, 1500)       # cofusion.yaml:20-22
    assert abs(c.tracking_lr - 1e-2) < 1e-9 and c.tracking_iters == 10                           # Tracker.cpp:103,107
    lr = np.array([list(r) for r in c.stage_lr])
    assert np.allclose(lr, [O.DEFAULT_LR[k] for k in ("coarse", "middle", "fine", "color")])     # nice_slam.yaml:102-126


def test_grid_dims_follow_main_cpp(nsb, syn):
    """main.cpp:34-78: coarse 5x3x8, middle 18x11x26, fine/color 36x22x52."""
    c = nsb.default_config()
    L = nsb.load_library()
    want = {"coarse": (5, 3, 8), "middle": (18, 11, 26), "fine": (36, 22, 52), "color": (36, 22, 52)}
    for i, lv in enumerate(syn.LEVELS):
        z, y, x = C.c_int(), C.c_int(), C.c_int()
        L.nsb_grid_dims(C.byref(c), i, C.byref(z), C.byref(y), C.byref(x))
        assert (z.value, y.value, x.value) == want[lv] == syn.grid_dims(lv)
    assert [L.nsb_decoder_count(i, 32) for i in range(4)] == [6337, 15800, 20920, 15899] == [syn.decoder_count(w) for w in syn.LEVELS]


def test_yaml_subset_reader(nsb, tmp_path):
    ns = tmp_path / "ns.yaml"
    ns.write_text("""coarse: True
grid_len:
  coarse: 2
  middle: 0.32 # comment
  fine: 0.16
  color: 0.16
tracking:
  ignore_edge_W: 30
  w_color_loss: 0.25
  lr: 0.002
  pixels: 250
  handle_dynamic: False
mapping:
  BA: False
  fix_fine: True
  middle_iter_ratio: 0.5
  pixels: 777
  iters: 40
  keyframe_selection_method: 'overlap'
  stage:
    middle:
      decoders_lr: 0.0
      middle_lr: 0.2
    color:
      decoders_lr: 0.007
      color_lr: 0.009
cam:
  H: 680
  W: 1200
""")
    ds = tmp_path / "ds.yaml"
    ds.write_text("dataset: 'cofusion'\nmapping:\n  pixels: 1234\ncam:\n  H: 480\n  W: 640\n  fx: 361.5\n")
    c = nsb.load_yaml_config(str(ns), str(ds))
    assert (c.H, c.W) == (480, 640) and abs(c.fx - 361.5) < 1e-6          # dataset file overrides (Mapper.cpp:22-28)
    assert c.mapping_pixels == 1234 and c.mapping_iters == 40
    assert c.ignore_edge_W == 30 and c.tracking_pixels == 250 and c.handle_dynamic == 0 and c.BA == 0
    assert abs(c.w_color_loss - 0.25) < 1e-7 and abs(c.mapping_w_color_loss - 0.25) < 1e-7   # Mapper.cpp:33 reads tracking's
    assert abs(c.stage_lr[1][2] - 0.2) < 1e-7 and abs(c.stage_lr[3][0] - 0.007) < 1e-7 and abs(c.stage_lr[3][4] - 0.009) < 1e-7
    assert abs(c.middle_iter_ratio - 0.5) < 1e-7
    with pytest.raises(RuntimeError):
        nsb.load_yaml_config(str(tmp_path / "missing.yaml"), None)


def test_camera_utils_match_reference_golden(nsb):
    """nsb_quad2rotation / nsb_get_camera_from_tensor vs utils.h:174-210 outputs of oracle/_ref."""
    g = load_golden("sampling.npz")
    for q, R_ in zip(g["quats"], g["rots"]):
        assert np.abs(nsb.quad2rotation(q) - R_).max() < 1e-6
    assert np.abs(nsb.get_camera_from_tensor(g["cam7"]) - g["RT"]).max() < 1e-6
    m = np.eye(4, dtype=np.float32); m[:3, :] = g["RT"]
    c7 = nsb.get_tensor_from_camera(m)
    q = g["cam7"][:4] / np.linalg.norm(g["cam7"][:4])
    assert min(np.abs(c7[:4] - q).max(), np.abs(c7[:4] + q).max()) < 1e-6 and np.allclose(c7[4:], g["cam7"][4:])
    assert np.allclose(c7, O.get_tensor_from_camera(m), atol=1e-6)


def test_no_cpu_fallback(nsb):
    """Without a GPU nsb_create must fail with a message -- never compute on the CPU."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device|sm_"):
        nsb.Engine(nsb.default_config())


def test_product_does_not_import_oracle():
    """The product tree never references the oracle (only tests/, smoke() and bench.py's cpu legs may)."""
    pkg = os.path.join(ROOT, "nice-slam-cpp_b200")
    for dp, _, fs in os.walk(pkg):
        if os.path.basename(dp) == "build":
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert "nice_oracle" not in src and "refbind" not in src and "oracle/" not in src, os.path.join(dp, f)
