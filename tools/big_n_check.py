"""Large-n functional check (BASELINE configs C4/C5 shapes): structured lower-triangular L (no host Cholesky), additive model,
implicit permutations.  int8 route vs FP64 route on a slice, oracle on a few SNPs.   usage: big_n_check.py n [M] [model]"""
import sys, time, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import flexlmm_b200 as fx
from flexlmm_b200 import synth
import scipy.linalg as sla

n = int(sys.argv[1]); M = int(sys.argv[2]) if len(sys.argv) > 2 else 19000
model = sys.argv[3] if len(sys.argv) > 3 else "additive"
rng = np.random.default_rng(n)
t0 = time.time()
L = np.zeros((n, n), order="F")
U = rng.standard_normal((n, 6)) * (1.5 / np.sqrt(n)); W = rng.standard_normal((n, 6))
for j0 in range(0, n, 2048):
    j1 = min(n, j0 + 2048)
    L[:, j0:j1] = np.tril(U @ W[j0:j1].T, k=-(j0 + 1))
L[np.arange(n), np.arange(n)] = 0.8 + 0.4 * rng.random(n)
M0, M1, M2, names, Xn, nn, cov1 = synth.design(model, n, seed=3)
codes = synth.genotypes(M, n, seed=5)
y = 0.3 * cov1 + L @ rng.standard_normal(n) + 0.05 * codes[0]
y_mm = sla.solve_triangular(L, y, lower=True); X_mm = np.asfortranarray(sla.solve_triangular(L, Xn, lower=True))
nf = fx.fit_null_model(X_mm, y_mm)
packed = synth.pack_2bit(codes)
seeds = np.arange(1, 11, dtype=np.int32)
print("inputs built in %.1f s" % (time.time() - t0), flush=True)

def run(int8, m):
    t1 = time.time()
    with fx.FitModel(int8=int8, verbose=1) as fm:
        fm.set_whitening(L); fm.set_null(X_mm, y_mm, nf["ll"], nf["df"]); fm.set_design(M0, M1, M2)
        fm.set_perm_implicit(seeds); fm.set_packed(packed, n)
        r = fm.run(np.arange(1, m + 1, dtype=np.int32)); st = fm.stats()
        t2 = time.time()
        r2 = fm.run(np.arange(1, m + 1, dtype=np.int32)); st2 = fm.stats()
    print("int8=%s m=%d: first call %.1f s (setup + run), second run %.3f s device, int8_path %d refit %d" % (int8, m, t2 - t1, st2["ms_total"] / 1e3, st["int8_path"], st["refit_snps"]), flush=True)
    return r2

r8 = run(True, M)
m64 = min(M, 2000)
r64 = run(False, m64)
d = np.abs(r8["lrt_chisq"][:m64] - r64["lrt_chisq"]) / np.maximum(1e-9, np.abs(r64["lrt_chisq"]))
print("int8 vs fp64: max rel chisq diff %.2e over %d SNPs; df equal %s" % (d.max(), m64, np.array_equal(r8["lrt_df"][:m64], r64["lrt_df"])))
import oracle as orc
pick = np.array([0, 7, m64 - 1])
o = orc.fit_model(L, y_mm, X_mm, nf["ll"], nf["df"], M0, M1, M2, codes[pick])
print("oracle chisq", o["chisq"], "int8", r8["lrt_chisq"][pick], "fp64", r64["lrt_chisq"][pick])
print("rel vs oracle", np.abs(r8["lrt_chisq"][pick] - o["chisq"]) / np.abs(o["chisq"]))
print("min_pval int8", r8["min_pval"][:4])
