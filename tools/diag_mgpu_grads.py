"""Debug aid: colour-decoder gradient of ONE colour iteration on 2 GPUs (NCCL path, captured after the all-reduce or summed over the
ranks) against the one-GPU run, per parameter tensor.  usage: python tools/diag_mgpu_grads.py [ba]"""
import importlib, os, sys
import numpy as np
import torch.multiprocessing as mp
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parts():
    K = [93, 32, 32, 125, 32]
    out = [("B", 279)]
    for i, k in enumerate(K):
        out += [("W%d" % i, 32 * k), ("b%d" % i, 32)]
    for i in range(5):
        out += [("Fc%d" % i, 1024), ("bc%d" % i, 32)]
    out += [("Wo", 128), ("bo", 4)]
    return out


def run(rank, world, uid, q, ba, it=59):
    nsb = importlib.import_module("nice-slam-cpp_b200"); syn = nsb.synthetic
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
    e.mapping_capture_grads(not os.environ.get("DIAG_NOCAP"))
    e.mapping_begin([0, 1], 60, 1.0, ba_mask=0b10 if ba else 0)
    reps = int(os.environ.get("DIAG_REPS", "1"))
    seq = [int(x) for x in os.environ["DIAG_SEQ"].split(",")] if os.environ.get("DIAG_SEQ") else [it] * reps
    for k in seq:
        loss = e.mapping_iter(k)
    g = e.captured_grads() if not os.environ.get("DIAG_NOCAP") else {"dec_color": np.zeros(1), "grid_fine": np.zeros(1), "grid_color": np.zeros(1)}
    out = {"loss": loss, "dec": g["dec_color"], "fine": g["grid_fine"], "color": g["grid_color"], "param": e.get_decoder("color")}
    e.close()
    if q is not None:
        q.put((rank, out))
    return out


if __name__ == "__main__":
    ba = len(sys.argv) > 1 and sys.argv[1] == "ba"
    nsb = importlib.import_module("nice-slam-cpp_b200")
    if os.environ.get("DIAG_PRE"):
        run(0, 1, None, None, False)         # an earlier engine in the same process (as in the pytest module)
    ref = run(0, 1, None, None, ba)
    uid = nsb.comm_unique_id()
    ctx = mp.get_context("spawn"); q = ctx.Queue()
    procs = [ctx.Process(target=run, args=(r, 2, uid, q, ba)) for r in range(2)]
    for p in procs: p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs: p.join(timeout=60)
    print("loss", ref["loss"], got[0]["loss"], got[1]["loss"], "T5_STASH", os.environ.get("NSB_T5_STASH", "1"), "ba", ba)
    same = np.array_equal(got[0]["dec"], got[1]["dec"])
    tot = got[0]["dec"] if same else got[0]["dec"] + got[1]["dec"]
    print("ranks identical (captured after the reduction):", same)
    off = 0
    for name, n in (parts() if not os.environ.get("DIAG_NOCAP") else []):
        r = ref["dec"][off:off + n]; t = tot[off:off + n]
        den = np.abs(r).max() or 1.0
        print("%-4s rel %.2e  (ref max %.3e)" % (name, np.abs(t - r).max() / den, den))
        off += n
    d0 = importlib.import_module("nice-slam-cpp_b200").synthetic.make_decoders(0, bias_scale=0.05)["color"]
    print("decoder after the step(s): |2gpu - 1gpu| max %.3e, move %.3e" % (np.abs(got[0]["param"] - ref["param"]).max(), np.abs(ref["param"] - d0).max()))
    for lv in (("fine", "color") if not os.environ.get("DIAG_NOCAP") else ()):
        t = got[0][lv] if np.array_equal(got[0][lv], got[1][lv]) else got[0][lv] + got[1][lv]
        print("grid_%s rel %.2e" % (lv, np.abs(t - ref[lv]).max() / np.abs(ref[lv]).max()))
