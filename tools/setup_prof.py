import sys, time, numpy as np
sys.path.insert(0, ".")
import bench, flexlmm_b200 as fx
w = dict(bench.WORKLOADS["C2-small"]); w["M"] = 20000
inp = None  # stale: see tools/qf_prof.py for the current input API
vi = np.arange(1, 101, dtype=np.int32)
for rep in range(4):
    t0 = time.perf_counter(); fm = fx.FitModel(verbose=1 if rep >= 2 else 0); t1 = time.perf_counter()
    fm.set_whitening(inp["L"]); t2 = time.perf_counter()
    fm.set_null(inp["X_null"], inp["y_mm"], inp["ll_null"], inp["df_null"]); fm.set_design(inp["M0"], inp["M1"], inp["M2"])
    fm.set_perm(inp["y_pred"], inp["U1"], inp["e_p"], inp["seeds"]); t3 = time.perf_counter()
    fm.set_packed(inp["packed"], w["n"]); t4 = time.perf_counter()
    fm.run(vi); t5 = time.perf_counter()
    fm.run(vi); t6 = time.perf_counter()
    print("create %.0f ms | set_whitening %.0f | set_null/design/perm %.0f | set_packed %.0f | first run (finalize + 100 SNPs) %.0f | second run %.1f" % tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5)), flush=True)
    fm.close()
