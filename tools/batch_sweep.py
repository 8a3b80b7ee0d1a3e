"""Step time vs batch size (SNP tiles per batch) on a C2-like task: python tools/batch_sweep.py"""
import numpy as np
from _inputs import workload
import flexlmm_b200 as fx
task, packed, w = workload("C2-small", M=18944 * 8)
vi = np.arange(1, w["M"] + 1, dtype=np.int32)
for bt in (148, 296, 444, 592, 1184):
    fm = fx.FitModel(batch_tiles=bt, verbose=0)
    task(fm, packed, w["n"]); fm.make_resident()
    best = None
    for _ in range(4):
        fm.run(vi, per_snp=False); st = fm.stats()
        if best is None or st["ms_total"] < best["ms_total"]: best = st
    print(bt, {k: round(best[k], 3) for k in ("ms_total", "ms_decode", "ms_whiten", "ms_fit", "ms_perm")}, flush=True)
    fm.close()
