"""Setup cost of one factor at a large n, stage by stage (FitModel(verbose=1) prints the library's stopwatch on stderr).
usage: setup_prof_big.py n [P]      the structured factor of bench.py's C4 / C5 workloads, built on the GPU"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import bench
import flexlmm_b200 as fx

n = int(sys.argv[1]); P = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
w = dict(n=n, P=P, model="additive")
dev = torch.device("cuda", 0)
task, _ = bench.task_structured(w, dev)
packed = bench.make_packed(n, 2048, 1, dev)
for rep in range(2):
    fm = fx.FitModel(verbose=1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    task(fm, packed, n)
    t1 = time.perf_counter()
    fm.prepare()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    r = fm.run(np.arange(1, 2049, dtype=np.int32), per_snp=False)
    t3 = time.perf_counter()
    print(f"rep {rep}: n={n} P={P}: inputs + flx_whitening_* + flx_whiten + null fit {t1 - t0:.2f} s, flx_prepare {t2 - t1:.2f} s, first run of 2048 SNPs {t3 - t2:.2f} s", flush=True)
    fm.close()
