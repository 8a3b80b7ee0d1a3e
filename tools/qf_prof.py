"""K2q MMA-issuer cycle split (epilogue wait, TMA wait) via FitModel(verbose=2): python tools/qf_prof.py [batch_tiles] [workload] [snps]
The timing experiments behind profiles/r02_k2q_issuer_probes.txt (skipped copies, N = 32 MMAs) were compile-time edits of qf.cu and are in git history."""
import sys, numpy as np
sys.path.insert(0, ".")
import torch
import bench, flexlmm_b200 as fx
bt = int(sys.argv[1]) if len(sys.argv) > 1 else 0
name = sys.argv[2] if len(sys.argv) > 2 else "C2-small"
w = dict(bench.WORKLOADS[name]); M = int(sys.argv[3]) if len(sys.argv) > 3 else 75776
w["P"] = min(w["P"], 100)
dev = torch.device("cuda:0")
packed = bench.make_packed(w["n"], M, 7, dev)
task = (bench.task_chol if w["gen"] == "chol" else bench.task_structured)(w, dev)[0]
verbose = int(sys.argv[4]) if len(sys.argv) > 4 else 2
fm = fx.FitModel(verbose=verbose, batch_tiles=bt)
task(fm, packed, w["n"]); fm.make_resident()
vi = np.arange(1, M + 1, dtype=np.int32)
for _ in range(4):
    fm.run(vi, per_snp=False); st = fm.stats(); print({k: round(float(st[k]), 3) for k in ("ms_whiten", "ms_decode", "ms_fit", "ms_perm", "ms_total", "refit_snps")}, flush=True)
fm.close()
