# Scratch: first GPU parity check of the CUDA path against the oracle on a small synthetic task.
import sys, time, numpy as np
sys.path.insert(0, ".")
import flexlmm_b200 as fx
from flexlmm_b200 import synth
from oracle import oracle as orc

def run(n, M, model, P, n_raw=None):
    t = synth.task(n, M, model=model, P=P, n_raw=n_raw)
    codes = t["codes_raw"][:, t["pgen_order"] - 1]
    fm = fx.FitModel()
    fm.set_whitening(t["L"]); fm.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"]); fm.set_design(t["M0"], t["M1"], t["M2"])
    if P: fm.set_perm(t["y_pred"], t["U1"], t["e_p"], t["seeds"])
    fm.set_packed(t["packed"], t["n_raw"], t["pgen_order"])
    vi = np.arange(1, M + 1, dtype=np.int32)
    X, cd = fm.decode_tile(vi)
    print("decode codes exact:", np.array_equal(cd, codes))
    W = fm.whiten(X[:, :min(X.shape[1], 40)])
    Wo = orc.forwardsolve(t["L"], X[:, :min(X.shape[1], 40)])
    print("whiten max rel err:", np.abs(W - Wo).max() / np.abs(Wo).max())
    t0 = time.time(); r = fm.run(vi); t1 = time.time()
    o = orc.fit_model(t["L"], t["y_mm"], t["X_null"], t["ll_null"], t["df_null"], t["M0"], t["M1"], t["M2"], codes)
    lp, lo = -np.log10(r["pval"]), -np.log10(o["pval"])
    print(f"n={n} M={M} {model} P={P}: run {t1-t0:.3f}s  df eq {np.array_equal(r['lrt_df'], o['df'])} rank eq {np.array_equal(r['rank'], o['rank'])} pivot eq {np.array_equal(r['pivot'], o['pivot'])}")
    print("  max |d chisq| rel:", np.max(np.abs(r["lrt_chisq"] - o["chisq"]) / np.maximum(np.abs(o["chisq"]), 1e-3)), " max |d -log10p|:", np.nanmax(np.abs(lp - lo) / (1e-8 * np.abs(lo) + 1e-10)), "(tolerance units)")
    print("  coef max rel:", np.max(np.abs(r["coef"] - o["coef"]) / (np.abs(o["coef"]) + 1e-6)))
    if P:
        mp = []
        for sd in t["seeds"]:
            op = orc.fit_model(t["L"], t["y_mm"], t["X_null"], t["ll_null"], t["df_null"], t["M0"], t["M1"], t["M2"], codes, perm_seed=int(sd), y_pred=t["y_pred"], U1=t["U1"], e_p=t["e_p"])
            mp.append(op["min_pval"])
        mp = np.array(mp)
        print("  minp max rel err (-log10):", np.max(np.abs(np.log10(r["min_pval"]) - np.log10(mp)) / np.abs(np.log10(mp))), r["min_pval"][:3], mp[:3])
    print("  stats", fm.stats())
    fm.close()

run(300, 50, "additive", 3)
run(300, 50, "domgxe", 3, n_raw=330)
run(1000, 200, "simple", 2)
