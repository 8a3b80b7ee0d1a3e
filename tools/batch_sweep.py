import sys, numpy as np
sys.path.insert(0, ".")
import bench, flexlmm_b200 as fx
w = dict(bench.WORKLOADS["C2-small"]); w["M"] = 18944 * 8
inp = bench.make_inputs(w, 0, 0)
vi = np.arange(1, w["M"] + 1, dtype=np.int32)
for bt in (148, 296, 444, 592, 1184):
    fm = fx.FitModel(batch_tiles=bt, verbose=0)
    fm.set_whitening(inp["L"]); fm.set_null(inp["X_null"], inp["y_mm"], inp["ll_null"], inp["df_null"]); fm.set_design(inp["M0"], inp["M1"], inp["M2"])
    fm.set_perm(inp["y_pred"], inp["U1"], inp["e_p"], inp["seeds"]); fm.set_packed(inp["packed"], w["n"]); fm.make_resident()
    best = None
    for _ in range(4):
        fm.run(vi, per_snp=False); st = fm.stats()
        if best is None or st["ms_total"] < best["ms_total"]: best = st
    print(bt, {k: round(best[k], 3) for k in ("ms_total", "ms_decode", "ms_whiten", "ms_fit", "ms_perm")}, flush=True)
    fm.close()
