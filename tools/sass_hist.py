"""Static SASS instruction histogram per kernel of libnsb.so (cuobjdump -sass), for profiles/.  usage: sass_hist.py [out.txt]
Lists, per kernel: instruction count, registers, and the counts of the mnemonics that identify the design (UTCHMMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UBLKCP = bulk TMA copy, SYNCS = mbarrier, HMMA = warp-level MMA, RED / ATOM, MUFU.*, F2FP, HADD2 = the fp16
split) plus the ten most frequent opcodes."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "nice-slam-cpp_b200", "libnsb.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
    if m and cur:
        regs[cur] = (int(m.group(1)), int(m.group(2)))
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "HMMA", "LDSM", "REDG", "RED", "ATOM", "ATOMG", "ATOMS", "FHADD", "MUFU", "F2FP", "HADD2", "FFMA", "LDG", "STG", "LDS", "STS", "SHFL", "BAR"]
out = []
fn = None; ops = None
def flush():
    if fn is None or not ops:
        return
    total = sum(ops.values())
    name = demangle(fn)
    if total < 150 or "k_" not in name:
        return
    r = regs.get(fn, (None, None))
    out.append("----- %s\n   instructions %d, registers %s, static shared %s" % (name[:150], total, r[0], r[1]))
    out.append("   design mnemonics: " + ", ".join("%s %d" % (k, ops[k]) for k in KEY if ops.get(k)))
    out.append("   top opcodes: " + ", ".join("%s %d" % kv for kv in ops.most_common(10)))
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        flush(); fn = m.group(1); ops = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and ops is not None:
        ops[m.group(1)] += 1
flush()
txt = "\n".join(out) + "\n"
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(txt)
print(txt)
