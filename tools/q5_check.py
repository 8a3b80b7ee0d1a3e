import sys, numpy as np
sys.path.insert(0, ".")
import flexlmm_b200 as fx
from flexlmm_b200 import synth
t = synth.task(5000, 3000, model="additive", P=20, seed=3, effect=0.05)
vi = np.arange(1, 3001, dtype=np.int32)
res = {}
for name, kw in (("fp64", dict(int8=False)), ("q6", dict(planes=6)), ("q5", dict(planes=5))):
    fm = fx.FitModel(**kw)
    fm.set_whitening(t["L"]); fm.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"]); fm.set_design(t["M0"], t["M1"], t["M2"])
    fm.set_perm(t["y_pred"], t["U1"], t["e_p"], t["seeds"]); fm.set_packed(t["packed"], t["n_raw"], t["pgen_order"])
    res[name] = fm.run(vi); st = fm.stats(); fm.close()
    print(name, "ms_whiten", st["ms_whiten"], "refit", st["refit_snps"])
for a in ("q6", "q5"):
    d = np.abs(res[a]["lrt_chisq"] - res["fp64"]["lrt_chisq"]) / np.maximum(res["fp64"]["lrt_chisq"], 1e-3)
    lp = np.abs(np.log10(res[a]["pval"]) - np.log10(res["fp64"]["pval"])) / (1e-8 * np.abs(np.log10(res["fp64"]["pval"])) + 1e-10)
    mp = np.abs(np.log10(res[a]["min_pval"]) - np.log10(res["fp64"]["min_pval"])) / np.abs(np.log10(res["fp64"]["min_pval"]))
    print(a, "max rel d chisq", d.max(), " -log10p err in tolerance units", lp.max(), " minp rel", mp.max())
