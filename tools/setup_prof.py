"""Once-per-task setup cost, stage by stage (FitModel(verbose=1) prints the library's own stopwatch): python tools/setup_prof.py [workload]"""
import sys, time, numpy as np
from _inputs import workload
import flexlmm_b200 as fx
task, packed, w = workload(sys.argv[1] if len(sys.argv) > 1 else "C2-small", M=20000)
vi = np.arange(1, 101, dtype=np.int32)
for rep in range(4):
    t0 = time.perf_counter(); fm = fx.FitModel(verbose=1 if rep >= 2 else 0); t1 = time.perf_counter()
    task(fm, packed, w["n"]); t2 = time.perf_counter()
    fm.run(vi); t3 = time.perf_counter()
    fm.run(vi); t4 = time.perf_counter()
    print("create %.0f ms | whitening + null + design + perm + packed %.0f | first run (finalize + 100 SNPs) %.0f | second run %.1f" % tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3)), flush=True)
    fm.close()
