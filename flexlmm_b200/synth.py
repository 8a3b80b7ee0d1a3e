"""Deterministic synthetic inputs of the shapes BASELINE.json names.  Host-side data generation
only (numpy); used by tests, bench.py and __graft_entry__.smoke().  The upstream steps that would produce these
inputs in the pipeline (DECORRELATE, FIT_NULL_MODEL, GET_DESIGN_MATRIX) are restated with plain numpy here because
they are inputs of the hot path, not part of it.
"""
import numpy as np


def pack_2bit(codes):
    """codes: (M, N_raw) uint8 in 0..3 -> (M, ceil(N_raw/4)) uint8, pgen mode-0x02 record layout."""
    codes = np.asarray(codes, dtype=np.uint8)
    M, N = codes.shape
    pad = (-N) % 4
    if pad:
        codes = np.concatenate([codes, np.zeros((M, pad), dtype=np.uint8)], axis=1)
    c = codes.reshape(M, -1, 4)
    return (c[:, :, 0] | (c[:, :, 1] << 2) | (c[:, :, 2] << 4) | (c[:, :, 3] << 6)).astype(np.uint8)


def genotypes(M, n_raw, seed=0, missing_rate=0.0):
    rng = np.random.default_rng(seed)
    maf = rng.uniform(0.05, 0.5, size=M)
    g = rng.binomial(2, maf[:, None], size=(M, n_raw)).astype(np.uint8)
    if missing_rate > 0:
        g[rng.random((M, n_raw)) < missing_rate] = 3
    return g


def covariance_factor(n, rank=256, seed=13):
    """V = 0.5 B B'/r + 0.5 I, L = chol(V) (lower)."""
    rng = np.random.default_rng(seed)
    r = min(rank, n)
    B = rng.standard_normal((n, r))
    V = 0.5 * (B @ B.T) / r + 0.5 * np.eye(n)
    return np.linalg.cholesky(V)


def design(model, n, seed=11):
    """Returns (M0, M1, M2, colnames, X_null_raw, null_names) for the named BASELINE model.

    'additive'  : y ~ x + cov1                       vs y ~ cov1     (k=3, c=2, s=1)
    'domgxe'    : y ~ x + I(x == 1) + x:cov1 + cov1  vs y ~ cov1     (k=5, c=2, s=3)
    'simple'    : y ~ x                              vs y ~ 1        (k=2, c=1, s=1)
    Column order is R's terms() order (intercept, main effects in formula order, interactions).
    """
    rng = np.random.default_rng(seed)
    cov1 = rng.standard_normal(n)
    one = np.ones(n)

    def mm(x):
        if model == "additive":
            return np.column_stack([one, x, cov1]), ["(Intercept)", "x", "cov1"]
        if model == "domgxe":
            return np.column_stack([one, x, (x == 1).astype(float), cov1, x * cov1]), ["(Intercept)", "x", "I(x == 1)TRUE", "cov1", "x:cov1"]
        if model == "simple":
            return np.column_stack([one, x]), ["(Intercept)", "x"]
        raise ValueError(model)

    M0, names = mm(np.zeros(n)); M1, _ = mm(np.ones(n)); M2, _ = mm(np.full(n, 2.0))
    if model == "simple":
        Xn, nn = one[:, None], ["(Intercept)"]
    else:
        Xn, nn = np.column_stack([one, cov1]), ["(Intercept)", "cov1"]
    return M0, M1, M2, names, Xn, nn, cov1


def design_at(model, x, cov1):
    """model.matrix of the named BASELINE model for a REAL-valued genotype vector x (dosages, pgenlibr::Read): what
    model.frame + model.matrix (fit_model.nf:124-127) evaluate per SNP.  design() is this at x = 0, 1, 2."""
    x = np.asarray(x, dtype=np.float64)
    one = np.ones(x.size)
    if model == "additive":
        return np.column_stack([one, x, cov1])
    if model == "domgxe":
        return np.column_stack([one, x, (x == 1).astype(float), cov1, x * cov1])
    if model == "simple":
        return np.column_stack([one, x])
    raise ValueError(model)


def solve_lower(L, B):
    """forwardsolve(L, B) with LAPACK (input generation only)."""
    try:
        from scipy.linalg import solve_triangular
        return solve_triangular(L, B, lower=True)
    except Exception:
        return np.linalg.solve(L, B)


def null_fit(y_mm, X_mm, with_U1=True):
    """FIT_NULL_MODEL restated with an implicit-complement U1 (last n-c columns of the Householder Q)."""
    n, c = X_mm.shape
    if with_U1:
        Q, _ = np.linalg.qr(X_mm, mode="complete")
        Qc, U1 = Q[:, :c], np.asfortranarray(Q[:, c:])
    else:
        Qc, _ = np.linalg.qr(X_mm)
        U1 = None
    fitted = Qc @ (Qc.T @ y_mm)
    e = y_mm - fitted
    rss = float(e @ e)
    ll = 0.5 * (-n * (np.log(2 * np.pi) + 1 - np.log(n) + np.log(rss)))
    return dict(e_p=None if U1 is None else U1.T @ e, y_pred=fitted, U1=U1, ll=ll, df=c + 1)


def task(n, M, model="additive", P=0, seed=0, n_raw=None, subset=False, effect=0.0):
    """A complete synthetic FIT_MODEL task."""
    n_raw = n if n_raw is None else n_raw
    rng = np.random.default_rng(1000 + seed)
    codes_raw = genotypes(M, n_raw, seed=seed)
    if subset or n_raw != n:
        order = np.sort(rng.choice(n_raw, size=n, replace=False)).astype(np.int32) + 1
    else:
        order = np.arange(1, n + 1, dtype=np.int32)
    L = covariance_factor(n, seed=13 + seed)
    M0, M1, M2, names, Xn, nn, cov1 = design(model, n, seed=11 + seed)
    z = np.random.default_rng(17 + seed).standard_normal(n)
    y = 0.3 * cov1 + L @ z
    if effect:
        y = y + effect * codes_raw[0, order - 1]
    y_mm = solve_lower(L, y)
    X_mm = solve_lower(L, Xn)
    nf = null_fit(y_mm, X_mm, with_U1=P > 0)
    return dict(n=n, M=M, model=model, P=P, L=L, y_mm=y_mm, X_null=np.asfortranarray(X_mm), ll_null=nf["ll"], df_null=nf["df"],
                M0=M0, M1=M1, M2=M2, names=names, codes_raw=codes_raw, packed=pack_2bit(codes_raw), n_raw=n_raw,
                pgen_order=order, y_pred=nf["y_pred"], U1=nf["U1"], e_p=nf["e_p"], seeds=np.arange(1, P + 1, dtype=np.int32))
