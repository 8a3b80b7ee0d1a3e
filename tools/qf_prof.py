import sys, numpy as np
sys.path.insert(0, ".")
import bench, flexlmm_b200 as fx
w = dict(bench.WORKLOADS["C2-small"]); w["M"] = 18944 * 2
inp = bench.make_inputs(w, 0, 0)
bt = int(sys.argv[1]) if len(sys.argv) > 1 else 0
fm = fx.FitModel(verbose=2, batch_tiles=bt)
fm.set_whitening(inp["L"]); fm.set_null(inp["X_null"], inp["y_mm"], inp["ll_null"], inp["df_null"]); fm.set_design(inp["M0"], inp["M1"], inp["M2"])
fm.set_perm(inp["y_pred"], inp["U1"], inp["e_p"], inp["seeds"]); fm.set_packed(inp["packed"], w["n"]); fm.make_resident()
vi = np.arange(1, w["M"] + 1, dtype=np.int32)
for _ in range(3):
    fm.run(vi, per_snp=False); print(fm.stats())
