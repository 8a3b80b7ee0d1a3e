"""Refit counts and run time of the int8 route per synthetic-factor seed (bench.task_structured), n given.  usage: refit_by_seed.py n M"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch, bench
import flexlmm_b200 as fx
n = int(sys.argv[1]); M = int(sys.argv[2])
dev = torch.device("cuda", 0)
w = dict(bench.WORKLOADS["C4"], n=n, M=M, P=1000)
vi = np.arange(1, M + 1, dtype=np.int32)
for r in range(8):
    for ch in range(3):
        seed = 13 + ch + 100 * r
        task, _ = bench.task_structured(w, dev, seed=seed)
        packed = bench.make_packed(n, M, 0xF1E + 7919 * r + 104729 * ch, dev)
        with fx.FitModel() as fm:
            task(fm, packed, n); fm.prepare(); fm.make_resident()
            fm.run(vi, per_snp=False)
            t0 = time.perf_counter(); res = fm.run(vi, per_snp=True); dt = time.perf_counter() - t0
            st = fm.stats()
        print(f"rank {r} chrom {ch} seed {seed}: run {dt:.3f} s, ms_total {st['ms_total']:.0f}, refits {st['refit_snps']}, df0 {(res['lrt_df'] == 0).sum()}, min chisq {np.nanmin(res['lrt_chisq']):.3g}", flush=True)
