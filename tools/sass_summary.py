"""Opcode histogram per kernel from `cuobjdump -sass` of the built library: the evidence that the hot kernels are the Blackwell-native
ones (UTCIMMA = tcgen05.mma kind::i8, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (TMA), UTCBAR/SYNCS = tcgen05.commit / mbarrier,
DMMA = FP64 tensor pipe).   usage: python tools/sass_summary.py [out.txt]"""
import collections, re, subprocess, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from flexlmm_b200 import build
KEY = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTCBAR", "SYNCS", "DMMA", "IMMA", "HMMA", "PRMT", "DFMA", "DADD", "DMUL", "I2F", "LDS", "STS", "LDG", "STG", "ATOMG", "RED", "SHFL", "BAR"]
out = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True, check=True).stdout
kern, hist = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and kern:
        hist[kern][m.group(1).split(".")[0]] += 1
        hist[kern]["_total"] += 1
lines = ["# cuobjdump -sass flexlmm_b200/lib/libflexlmm_cuda.so (sm_100a): instruction counts per kernel", "# kernel | total | " + " ".join(KEY)]
for k, h in hist.items():
    lines.append(f"{k} | {h['_total']} | " + " ".join(f"{o}={h[o]}" for o in KEY if h[o]))
txt = "\n".join(lines) + "\n"
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(txt)
print(txt)
