"""Debug aid: run-to-run determinism of the colour-decoder gradient and of the decoder after two colour steps (one GPU, graph path)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nsb = importlib.import_module("nice-slam-cpp_b200"); syn = nsb.synthetic
def run():
    grids = syn.make_grids(0); decs = syn.make_decoders(0, bias_scale=0.05)
    depths, colors, poses = syn.make_frames(2, 0)
    cfg = nsb.default_config(); cfg.mapping_pixels = 2000; cfg.max_rays = 2000; cfg.frustum_feature_selection = 0
    e = nsb.Engine(cfg); e.set_model(grids, decs)
    for f in range(2): e.set_frame(f, depths[f], colors[f], poses[f])
    e.mapping_capture_grads(True)
    e.seed(int(os.environ.get("DIAG_SEED", "21"))); e.mapping_begin([0, 1], 60, 1.0)
    gs = []
    mid = None
    for n, k in enumerate((0, 30, 59, 59)):
        e.mapping_iter(k); gs.append(e.captured_grads()["dec_color"].copy())
        if n == 2 and os.environ.get("DIAG_MID"):     # parameter state between the two colour iterations
            mid = {lv: e.get_grid(lv).copy() for lv in ("middle", "fine", "color")}; mid["dec"] = e.get_decoder("color").copy()
    d = e.get_decoder("color"); e.close(); MIDS.append(mid); return d, gs
MIDS = []
runs = [run() for _ in range(int(os.environ.get("DIAG_RUNS", "8")))]
d0 = syn.make_decoders(0, bias_scale=0.05)["color"]
print("T5_STASH", os.environ.get("NSB_T5_STASH", "1"))
for k, (d, gs) in enumerate(runs):
    dev = float(np.abs(d - runs[0][0]).max())
    g3 = [float(np.abs(gs[i] - runs[0][1][i]).max() / np.abs(runs[0][1][i]).max()) if np.abs(runs[0][1][i]).max() > 0 else 0.0 for i in (2, 3)]
    j = int(np.argmax(np.abs(d - runs[0][0])))
    if g3[1] > 1e-5:
        K = [93, 32, 32, 125, 32]; parts = [("B", 279)]
        for i, kk in enumerate(K): parts += [("W%d" % i, 32 * kk), ("b%d" % i, 32)]
        for i in range(5): parts += [("Fc%d" % i, 1024), ("bc%d" % i, 32)]
        parts += [("Wo", 128), ("bo", 4)]
        off = 0; txt = []
        for name, n in parts:
            dv = np.abs(gs[3][off:off + n] - runs[0][1][3][off:off + n]).max() / np.abs(runs[0][1][3]).max()
            if dv > 1e-6: txt.append("%s %.1e" % (name, dv))
            off += n
        w0 = (gs[3][279:279 + 2976] - runs[0][1][3][279:279 + 2976]).reshape(32, 93)
        cols = np.nonzero(np.abs(w0).max(0) > 1e-6 * np.abs(runs[0][1][3]).max())[0]
        print("   deviating tensors:", ", ".join(txt), "| W0 columns:", cols.tolist()[:40])
    print("run %d: decoder dev %.2e at %d | grad rel dev colour it 1: %.2e, it 2: %.2e | grad[%d] it1 %.4e it2 %.4e (gmax %.2e)" %
          (k, dev, j, g3[0], g3[1], j, gs[2][j], gs[3][j], np.abs(gs[3]).max()))
    if MIDS[k] is not None:
        print("   state after colour it 1 vs run 0:", ", ".join("%s max %.2e n(>1e-4) %d" % (kk, float(np.abs(MIDS[k][kk] - MIDS[0][kk]).max()), int((np.abs(MIDS[k][kk] - MIDS[0][kk]) > 1e-4).sum())) for kk in MIDS[k]))
