"""Debug aid: the same iteration sequence through the captured-graph path and kernel by kernel (NSB_GRAPH=0) on one GPU.  usage: diag_graph_eager.py [ba]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nsb = importlib.import_module("nice-slam-cpp_b200"); syn = nsb.synthetic
ba = len(sys.argv) > 1 and sys.argv[1] == "ba"
seq = [int(x) for x in os.environ.get("DIAG_SEQ", "0,30,59,59").split(",")]
outs = {}
for graph in ("1", "0"):
    os.environ["NSB_GRAPH"] = graph
    grids = syn.make_grids(0); decs = syn.make_decoders(0, bias_scale=0.05)
    depths, colors, poses = syn.make_frames(2, 0)
    cfg = nsb.default_config(); cfg.mapping_pixels = 2000; cfg.max_rays = 2000; cfg.frustum_feature_selection = 0; cfg.BA_cam_lr = 0.001
    e = nsb.Engine(cfg)
    e.set_model(grids, decs)
    for f in range(2):
        e.set_frame(f, depths[f], colors[f], poses[f])
    e.seed(21)
    e.mapping_begin([0, 1], 60, 1.0, ba_mask=0b10 if ba else 0)
    losses = [e.mapping_iter(k) for k in seq]
    cams = e.mapping_end()
    outs[graph] = (np.array(losses), e.get_decoder("color"), e.get_grid("fine"), cams)
    e.close()
d0 = syn.make_decoders(0, bias_scale=0.05)["color"]
print("knobs", {k: v for k, v in os.environ.items() if k.startswith("NSB_") and k != "NSB_GRAPH"}, "ba", ba, "seq", seq)
print("losses graph", outs["1"][0], "eager", outs["0"][0])
print("decoder |graph - eager| max %.3e (move %.3e); fine grid %.3e; cams %.3e" % (np.abs(outs["1"][1] - outs["0"][1]).max(), np.abs(outs["1"][1] - d0).max(),
      np.abs(outs["1"][2] - outs["0"][2]).max(), np.abs(outs["1"][3] - outs["0"][3]).max()))
