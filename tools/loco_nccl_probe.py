"""Why does a multi-handle (LOCO) rank spend seconds after its batches when communicators are attached?  Under torchrun (2 ranks):
three resident handles per rank, per-call stats with and without communicators.   usage: torchrun ... tools/loco_nccl_probe.py [n] [M]"""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch, bench
import flexlmm_b200 as fx
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 150000
ctx = bench.Ctx()
w = dict(bench.WORKLOADS["C4"], n=n, M=M, P=200)
fms, vi = [], np.arange(1, M + 1, dtype=np.int32)
for ch in range(3):
    task, _ = bench.task_structured(w, ctx.dev, seed=13 + ch + 100 * ctx.rank)
    packed = bench.make_packed(n, M, 7 + ctx.rank * 31 + ch, ctx.dev)
    fm = fx.FitModel(device=ctx.local_rank)
    task(fm, packed, n); fm.prepare(); fm.make_resident()
    fms.append(fm)
def rounds(tag):
    for rep in range(3):
        ctx.barrier(); t0 = time.perf_counter()
        rows = []
        for fm in fms:
            t1 = time.perf_counter(); fm.run(vi, per_snp=False); st = fm.stats()
            rows.append((time.perf_counter() - t1, st["ms_total"], st["ms_batches"], st["ms_d2h"], st["refit_snps"]))
        ctx.barrier()
        print(f"[{tag}] rank {ctx.rank} rep {rep}: step {time.perf_counter() - t0:.3f} s; per call (wall s, ms_total, ms_batches, ms_after, refits): " +
              "; ".join(f"{a:.3f} {b:.0f} {c:.0f} {d:.0f} {e}" for a, b, c, d, e in rows), flush=True)
rounds("no comm")
for fm in fms:
    bench.attach_comm(ctx, fm)
rounds("comm on every handle")
