"""Builds libnsb.so (the C-ABI library of include/nsb.h) in-tree with nvcc for sm_100a.

    python nice-slam-cpp_b200/build.py [--force] [--variant precise_sin]

Object files are cached under nice-slam-cpp_b200/build/ keyed by a hash of the sources and flags; the
translation units are compiled in parallel (the backward decoder instantiations dominate the build time).
nvcc cross-compiles without a GPU, so this also runs in the CPU-only container.
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
BASE = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]

UNITS = [("api.cu", []), ("decode_fwd.cu", []), ("decode_fwd_tc.cu", []), ("decode_fwd_tc16.cu", []), ("decode_fwd_t5.cu", []), ("decode_bwd_t5.cu", []), ("wgrad.cu", []), ("wgrad_fused.cu", [])] + \
        [("decode_bwd_inst.cu", ["-DNSB_BWD_COMBO=%d" % k]) for k in range(6)]
VARIANTS = {"": [], "precise_sin": ["-DNSB_PRECISE_SIN"], "tctiming": ["-DNSB_TC_TIMING"],
            # occupancy experiments: warps per CTA of the forward / backward decoder kernels
            "f20b20": ["-DNSB_FWD_WARPS=20", "-DNSB_BWD_WARPS=20"], "f16b24": ["-DNSB_BWD_WARPS=24"], "f20b24": ["-DNSB_FWD_WARPS=20", "-DNSB_BWD_WARPS=24"],
            "f24b24": ["-DNSB_FWD_WARPS=24", "-DNSB_BWD_WARPS=24"], "f16b20": ["-DNSB_BWD_WARPS=20"],
            "nohint": ["-DNSB_MBAR_HINT_NS=0"],   # mbarrier waits without the suspend-time hint (A/B)
            "scat0": ["-DNSB_SCATTER_AGG=0"],   # per-sample quad reductions instead of the warp-aggregated scatter (A/B)
            "emb1": ["-DNSB_EMB_UNROLL=1"], "emb3": ["-DNSB_EMB_UNROLL=3"], "emb6": ["-DNSB_EMB_UNROLL=6"]}


def lib_path(variant=""):
    return os.path.join(HERE, "libnsb%s.so" % ("_" + variant if variant else ""))


def _hash(flags):
    h = hashlib.sha256()
    for fn in sorted(os.listdir(CSRC)) + ["../../include/nsb.h"]:
        with open(os.path.join(CSRC, fn), "rb") as f:
            h.update(fn.encode()); h.update(f.read())
    h.update(" ".join(flags).encode())
    return h.hexdigest()[:16]


def _compile(src, flags, obj, log):
    cmd = [NVCC] + ARCH + BASE + flags + ["-Xptxas", "-v", "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout[-4000:]))
    return obj


def build(variant="", force=False, verbose=True):
    vflags = VARIANTS[variant]
    os.makedirs(OUT, exist_ok=True)
    tag = _hash(vflags)
    lib = lib_path(variant)
    stamp = os.path.join(OUT, "stamp_%s.txt" % (variant or "default"))
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read().strip() == tag:
        return lib
    jobs = []
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        for src, fl in UNITS:
            name = os.path.splitext(src)[0] + "".join(x.replace("-D", "_").replace("=", "") for x in fl) + ("_" + variant if variant else "")
            obj = os.path.join(OUT, name + ".o")
            jobs.append(ex.submit(_compile, src, vflags + fl, obj, os.path.join(OUT, name + ".log")))
        objs = [j.result() for j in jobs]
    cmd = [NVCC] + ARCH + ["-shared", "-o", lib] + objs + ["-ldl"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout[-4000:])
    with open(stamp, "w") as f:
        f.write(tag)
    if verbose:
        print("built", lib)
    return lib


if __name__ == "__main__":
    v = ""
    if "--variant" in sys.argv:
        v = sys.argv[sys.argv.index("--variant") + 1]
    build(v, force="--force" in sys.argv)
