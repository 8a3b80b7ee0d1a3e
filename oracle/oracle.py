"""ctypes front-end to liboracle.so plus numpy restatements of the R steps that feed / consume the
hot path.  TEST INFRASTRUCTURE ONLY — PARITY UNPINNED (see flexlmm_oracle.h).

Reference call sites restated here (relative to /root/reference):
  lower_design()        modules/local/r/get_design_matrix.nf:87-121, fit_model.nf:124-127 (model.matrix)
  decorrelate()         modules/local/r/decorrelate.nf:50-77
  fit_null_model()      modules/local/r/fit_null_model.nf:33-72
  fit_model()           modules/local/r/fit_model.nf:96-181 (C: orc_fit_model)
  summarise_minp()      modules/local/r/get_null_p_dist.nf:77-79
  quantile7()           modules/local/r/make_plots.nf:31
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_sample_perm.argtypes = [C.c_int32, C.c_int64, C.c_void_p]
        L.orc_pgen_info.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
        L.orc_pgen_read_hardcalls.argtypes = [C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p]
        L.orc_pgen_read_dosage.argtypes = [C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p]
        L.orc_forwardsolve.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64]
        L.orc_lm_fit.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_double] + [C.c_void_p] * 5
        L.orc_lm_fit.restype = C.c_int
        L.orc_loglik_lm.argtypes = [C.c_void_p, C.c_int64]
        L.orc_loglik_lm.restype = C.c_double
        L.orc_pchisq_upper.argtypes = [C.c_double, C.c_int]
        L.orc_pchisq_upper.restype = C.c_double
        L.orc_fit_model.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32] + [C.c_void_p] * 6 + [C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.orc_set_seed.argtypes = [C.c_void_p, C.c_int32]
        L.orc_unif_rand.argtypes = [C.c_void_p]
        L.orc_unif_rand.restype = C.c_double
        _LIB = L
    return _LIB


class _Task(C.Structure):
    _fields_ = [("n", C.c_int64), ("k", C.c_int), ("c_null", C.c_int),
                ("L", C.c_void_p), ("y_mm", C.c_void_p), ("X_null", C.c_void_p),
                ("ll_null", C.c_double), ("df_null", C.c_int),
                ("M0", C.c_void_p), ("M1", C.c_void_p), ("M2", C.c_void_p),
                ("y_pred", C.c_void_p), ("U1", C.c_void_p), ("e_p", C.c_void_p), ("m", C.c_int64)]


def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


# ------------------------------------------------------------------ R RNG
def r_sample(seed, m):
    """set.seed(seed); sample.int(m)  -> 1-based int32 permutation (fit_model.nf:97-98)."""
    out = np.empty(m, dtype=np.int32)
    lib().orc_sample_perm(int(seed), int(m), out.ctypes.data)
    return out


def r_runif(seed, cnt):
    g = (C.c_uint32 * 626)()
    lib().orc_set_seed(g, int(seed))
    return np.array([lib().orc_unif_rand(g) for _ in range(cnt)])


# ------------------------------------------------------------------ pgen
def pgen_info(buf):
    b = np.frombuffer(buf, dtype=np.uint8)
    v, s, m = C.c_uint32(), C.c_uint32(), C.c_int()
    rc = lib().orc_pgen_info(b.ctypes.data, b.size, C.byref(v), C.byref(s), C.byref(m))
    if rc:
        raise ValueError(f"bad pgen ({rc})")
    return v.value, s.value, m.value


def pgen_read_hardcalls(buf, vidx0):
    """codes (uint8 0..3) of 0-based variant vidx0 for all raw samples (pgenlibr::ReadHardcalls, fit_model.nf:123)."""
    b = np.frombuffer(buf, dtype=np.uint8)
    _, s, _ = pgen_info(buf)
    codes = np.empty(s, dtype=np.uint8)
    rc = lib().orc_pgen_read_hardcalls(b.ctypes.data, b.size, int(vidx0), codes.ctypes.data)
    if rc:
        raise ValueError(f"pgen decode failed ({rc})")
    return codes


def pgen_read(buf, vidx0):
    """pgenlibr::Read (fit_model.nf:26,123 with use_dosage = true): doubles, dosage / 16384 where the record carries one,
    else the hard call; NaN for a missing call without a dosage."""
    b = np.frombuffer(buf, dtype=np.uint8)
    _, s, _ = pgen_info(buf)
    out = np.empty(s, dtype=np.float64)
    rc = lib().orc_pgen_read_dosage(b.ctypes.data, b.size, int(vidx0), out.ctypes.data)
    if rc:
        raise ValueError(f"pgen dosage read failed ({rc})")
    return out


def fit_model_x(L, y_mm, X_null, ll_null, df_null, design_of_x, xs, perm_seed=0, y_pred=None, U1=None, e_p=None):
    """FIT_MODEL for REAL-valued genotypes (dosages): the literal loop fit_model.nf:120-143 with the design of each SNP evaluated by
    `design_of_x(x) -> n x k matrix` (model.frame + model.matrix on the dosage vector, :124-127), built from the C primitives.
    xs: (M, n) doubles in model order.  Small cases only (Python loop)."""
    y = np.ascontiguousarray(y_mm, dtype=np.float64).copy()
    n = y.size
    if perm_seed:
        idx = r_sample(perm_seed, len(np.ravel(e_p)))
        y = np.asarray(y_pred, dtype=np.float64) + np.asarray(U1) @ np.ravel(e_p)[idx - 1]
        f0 = lm_fit(X_null, y)
        ll_null, df_null = loglik_lm(f0["residuals"]), f0["rank"] + 1
    out = dict(chisq=[], df=[], pval=[], rank=[], pivot=[], coef=[])
    min_pval = 2.0
    for x in xs:
        if np.isnan(x).any():
            raise ValueError("missing genotype in a tested variant (the reference aborts: SURVEY §8 a4)")
        X = np.asarray(design_of_x(x), dtype=np.float64)
        fit = lm_fit(forwardsolve(L, X), y)
        ll = loglik_lm(fit["residuals"])
        df = fit["rank"] + 1 - df_null
        chisq = 2.0 * (ll - ll_null)
        p = pchisq_upper(chisq, df) if df > 0 else np.nan
        if not np.isnan(p):
            min_pval = min(min_pval, p)
        for k2, v in (("chisq", chisq), ("df", df), ("pval", p), ("rank", fit["rank"]), ("pivot", fit["pivot"]), ("coef", fit["coefficients"])):
            out[k2].append(v)
    res = {k2: np.array(v) for k2, v in out.items()}
    res["min_pval"], res["nvars"] = min_pval, len(xs)
    return res


# ------------------------------------------------------------------ small numerics
def forwardsolve(L, X):
    L = _f(L); X = _f(np.array(X, dtype=np.float64, copy=True).reshape(L.shape[0], -1))
    lib().orc_forwardsolve(L.ctypes.data, L.shape[0], L.shape[0], X.ctypes.data, X.shape[1], X.shape[0])
    return X


def lm_fit(X, y, tol=1e-7):
    """.lm.fit(x, y): dict(coefficients (pivoted order), residuals, effects, rank, pivot (1-based), qr, qraux)."""
    X = _f(np.array(X, dtype=np.float64, copy=True)); n, p = X.shape
    y = np.ascontiguousarray(y, dtype=np.float64)
    coef = np.zeros(p); res = np.zeros(n); eff = np.zeros(n); piv = np.zeros(p, dtype=np.int32); qraux = np.zeros(p)
    r = lib().orc_lm_fit(X.ctypes.data, n, p, y.ctypes.data, tol, coef.ctypes.data, res.ctypes.data, eff.ctypes.data,
                         piv.ctypes.data, qraux.ctypes.data)
    return dict(coefficients=coef, residuals=res, effects=eff, rank=r, pivot=piv, qr=X, qraux=qraux)


def loglik_lm(resid):
    r = np.ascontiguousarray(resid, dtype=np.float64)
    return lib().orc_loglik_lm(r.ctypes.data, r.size)


def pchisq_upper(q, df):
    return lib().orc_pchisq_upper(float(q), int(df))


# ------------------------------------------------------------------ design lowering (model.matrix subset)
def lower_design(terms, covars, x):
    """Evaluate the model matrix of a row-wise formula for a given genotype vector x.

    terms : list of term labels in FORMULA order drawn from
            'x', 'I(x == 1)', '<name>', 'x:<name>'  (the forms the reference documents, README.md:83, docs/parameters.md:50)
    covars: dict name -> ('factor', levels, codes) | ('numeric', values)   (get_design_matrix.nf:91,99)
    Returns (matrix n*k, column names) in R terms() order: intercept, main effects in formula order,
    then interactions in formula order; factors/logicals use treatment contrasts.
    """
    x = np.asarray(x, dtype=np.float64); n = x.size

    def cols_of(name):
        if name == "x":
            return [("x", x)]
        if name == "I(x == 1)":
            return [("I(x == 1)TRUE", (x == 1).astype(np.float64))]
        kind = covars[name]
        if kind[0] == "numeric":
            return [(name, np.asarray(kind[1], dtype=np.float64))]
        levels, codes = kind[1], np.asarray(kind[2])
        return [(f"{name}{lev}", (codes == li).astype(np.float64)) for li, lev in enumerate(levels) if li > 0]

    main = [t for t in terms if ":" not in t]
    inter = [t for t in terms if ":" in t]
    names = ["(Intercept)"]; cols = [np.ones(n)]
    for t in main:
        for nm, v in cols_of(t):
            names.append(nm); cols.append(v)
    for t in inter:
        a, b = t.split(":")
        for na, va in cols_of(a):
            for nb, vb in cols_of(b):
                names.append(f"{na}:{nb}"); cols.append(va * vb)
    return np.column_stack(cols), names


def design_lookup(terms, covars, n):
    """M0, M1, M2 = model.matrix evaluated with x == 0, 1, 2 for every row."""
    out = []
    for g in (0.0, 1.0, 2.0):
        M, names = lower_design(terms, covars, np.full(n, g))
        out.append(M)
    return out[0], out[1], out[2], names


# ------------------------------------------------------------------ upstream producers (inputs of the hot path)
def decorrelate(K, s2g, s2e, y, X):
    V = s2g * np.asarray(K, dtype=np.float64) + s2e * np.eye(len(y))
    L = np.linalg.cholesky(V)                      # t(chol(V)), decorrelate.nf:57
    return dict(y_mm=forwardsolve(L, y)[:, 0], X_mm=forwardsolve(L, X), L=L)


def fit_null_model(y_mm, X_mm):
    fit = lm_fit(X_mm, y_mm)                        # lm.fit and .lm.fit share dqrls
    ll = loglik_lm(fit["residuals"])
    e = fit["residuals"]
    n = len(y_mm)
    Q = np.linalg.qr(np.asarray(X_mm, dtype=np.float64))[0]
    V = np.eye(n) - Q @ Q.T
    w, vec = np.linalg.eigh(V)
    order = np.argsort(-w, kind="stable")           # eigen() returns decreasing eigenvalues
    w, vec = w[order], vec[:, order]
    U1 = vec[:, w > 0.9]
    e_p = U1.T @ e
    return dict(e_p=e_p, y_pred=y_mm - e, U1=np.asfortranarray(U1), ll=ll, df=fit["rank"] + 1, rank=fit["rank"])


# ------------------------------------------------------------------ the task
def fit_model(L, y_mm, X_null, ll_null, df_null, M0, M1, M2, geno, perm_seed=0, y_pred=None, U1=None, e_p=None):
    """One FIT_MODEL task. geno: (M, n) uint8 codes in model order. Returns dict of per-SNP arrays + min_pval, nvars."""
    L = _f(L); y_mm = np.ascontiguousarray(y_mm, dtype=np.float64); X_null = _f(X_null)
    M0 = _f(M0); M1 = _f(M1); M2 = _f(M2)
    n, k = M0.shape
    geno = np.ascontiguousarray(geno, dtype=np.uint8); Mv = geno.shape[0]
    assert geno.shape[1] == n
    t = _Task()
    t.n, t.k, t.c_null = n, k, X_null.shape[1]
    t.L, t.y_mm, t.X_null = L.ctypes.data, y_mm.ctypes.data, X_null.ctypes.data
    t.ll_null, t.df_null = float(ll_null), int(df_null)
    t.M0, t.M1, t.M2 = M0.ctypes.data, M1.ctypes.data, M2.ctypes.data
    keep = []
    if perm_seed:
        y_pred = np.ascontiguousarray(y_pred, dtype=np.float64); U1 = _f(U1); e_p = np.ascontiguousarray(np.ravel(e_p), dtype=np.float64)
        keep = [y_pred, U1, e_p]
        t.y_pred, t.U1, t.e_p, t.m = y_pred.ctypes.data, U1.ctypes.data, e_p.ctypes.data, e_p.size
    chisq = np.zeros(Mv); df = np.zeros(Mv, dtype=np.int32); pval = np.zeros(Mv); rank = np.zeros(Mv, dtype=np.int32)
    pivot = np.zeros((Mv, k), dtype=np.int32); coef = np.zeros((Mv, k))
    minp = C.c_double(); nv = C.c_int64()
    rc = lib().orc_fit_model(C.byref(t), geno.ctypes.data, Mv, int(perm_seed), chisq.ctypes.data, df.ctypes.data,
                             pval.ctypes.data, rank.ctypes.data, pivot.ctypes.data, coef.ctypes.data, C.byref(minp), C.byref(nv))
    del keep
    if rc == -2:
        raise ValueError("missing genotype in a tested variant (the reference aborts: SURVEY §8 a4)")
    if rc:
        raise RuntimeError(f"oracle failed ({rc})")
    return dict(chisq=chisq, df=df, pval=pval, rank=rank, pivot=pivot, coef=coef, min_pval=minp.value, nvars=nv.value)


# ------------------------------------------------------------------ downstream consumers
def summarise_minp(rows):
    """rows: iterable of (permutation, chr, min_pval, nvars) -> {permutation: (min over chr, sum nvars)}."""
    out = {}
    for perm, _chr, mp, nv in rows:
        a = out.get(perm)
        out[perm] = (mp, nv) if a is None else (min(a[0], mp), a[1] + nv)
    return out


def quantile7(x, p):
    """R quantile(x, p) default type 7 (make_plots.nf:31)."""
    xs = np.sort(np.asarray(x, dtype=np.float64)); n = xs.size
    h = (n - 1) * p
    lo = int(np.floor(h)); hi = min(lo + 1, n - 1)
    return xs[lo] + (h - lo) * (xs[hi] - xs[lo])
