"""K2q with and without L2 grouping of the SNP tile pairs (FLX_QF_GROUP_PAIRS: 0 = plain order, k = pairs per group, unset = the
library's choice): K2q ms per step on a C2-like task.   usage: qf_group_sweep.py [n] [M]"""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch, bench
import flexlmm_b200 as fx
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 303104
w = dict(bench.WORKLOADS["C2"], n=n, M=M, P=100)
dev = torch.device("cuda", 0)
task, _ = (bench.task_chol if n <= 20000 else bench.task_structured)(w, dev)
packed = bench.make_packed(n, M, 1, dev)
vi = np.arange(1, M + 1, dtype=np.int32)
for setting in [None, "0", "24", "37", "74", "148", None, "0"]:
    if setting is None:
        os.environ.pop("FLX_QF_GROUP_PAIRS", None)
    else:
        os.environ["FLX_QF_GROUP_PAIRS"] = setting
    with fx.FitModel() as fm:
        task(fm, packed, n); fm.make_resident()
        for _ in range(2):
            fm.run(vi, per_snp=False)
        ms = []
        for _ in range(4):
            fm.run(vi, per_snp=False); st = fm.stats(); ms.append((st["ms_whiten"], st["ms_total"]))
    print(f"n={n} M={M} FLX_QF_GROUP_PAIRS={setting}: K2q {np.mean([m[0] for m in ms]):.2f} ms, run {np.mean([m[1] for m in ms]):.2f} ms", flush=True)
