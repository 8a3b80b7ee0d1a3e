"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the committed golden fixtures.

Bar: genotype decode and permutation index order bit-exact; lrt_df / rank / pivot exact;
-log10 p and the minP threshold within |d| <= 1e-8 * |value| + 1e-10 in FP64; coefficients within 1e-8 relative.
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL, ATOL = 1e-8, 1e-10        # the stated tolerance on -log10 p


def assert_logp_close(p_gpu, p_ref, chisq_ref=None, df_ref=None, N=None):
    """|d(-log10 p)| <= 1e-8 |-log10 p| + 1e-10, plus — when chisq_ref / df_ref / N are given — the reference's OWN rounding floor:
    R forms lrt_chisq = 2 (ll - ll.null) from two numbers of size N/2 log(RSS) (fit_model.nf:134-136), so one or two ulps in
    each RSS move chi-square by ~N eps in absolute terms; for chi-square -> 0 that is a visible relative change of p
    (SURVEY App. B.5).  The bar therefore also accepts what a chi-square within 8 N eps of the reference's maps to."""
    p_gpu, p_ref = np.asarray(p_gpu, dtype=float), np.asarray(p_ref, dtype=float)
    assert np.array_equal(np.isnan(p_gpu), np.isnan(p_ref))
    ok = ~np.isnan(p_ref)
    with np.errstate(divide="ignore"):
        a, b = -np.log10(p_gpu[ok]), -np.log10(p_ref[ok])
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)                       # p -> 0 underflow must give exactly Inf like R
    tol = RTOL * np.abs(b[fin]) + ATOL
    if chisq_ref is not None:
        from scipy.stats import chi2
        q, d = np.asarray(chisq_ref, dtype=float)[ok][fin], np.asarray(df_ref)[ok][fin]
        dq = 8.0 * N * np.finfo(float).eps
        tol = tol + np.abs(chi2.logsf(np.maximum(q + dq, 0.0), d) - chi2.logsf(np.maximum(q - dq, 0.0), d)) / np.log(10.0)
    err = np.abs(a[fin] - b[fin]) - tol
    assert err.size == 0 or err.max() <= 0, f"-log10 p off by {err.max():.3e} beyond tolerance"


def make(fx, t, perm=True, pgen_path=None, batch_tiles=0):
    fm = fx.FitModel(batch_tiles=batch_tiles)
    fm.set_whitening(t["L"])
    fm.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"])
    fm.set_design(t["M0"], t["M1"], t["M2"])
    if perm and t.get("P", 0):
        fm.set_perm(t["y_pred"], t["U1"], t["e_p"], t["seeds"])
    if pgen_path is not None:
        fm.open_pgen(pgen_path, t["pgen_order"])
    else:
        fm.set_packed(t["packed"], t["n_raw"], t["pgen_order"])
    return fm


def oracle_run(orc, t, codes, seed=0):
    kw = dict(perm_seed=seed, y_pred=t["y_pred"], U1=t["U1"], e_p=t["e_p"]) if seed else {}
    return orc.fit_model(t["L"], t["y_mm"], t["X_null"], t["ll_null"], t["df_null"], t["M0"], t["M1"], t["M2"], codes, **kw)


def oracle_parallel(orc, t, codes, seeds=(), threads=None):
    """The oracle over many SNPs at the benched shapes: SNP chunks on host threads (the C call releases the GIL; L and U1 are
    shared, not copied).  Returns (per-SNP dict of the ORIG task, {seed: min_pval})."""
    from concurrent.futures import ThreadPoolExecutor
    M = codes.shape[0]
    threads = threads or max(1, min(os.cpu_count() or 1, 32, M))
    cuts = np.linspace(0, M, threads + 1).astype(int)
    chunks = [(cuts[i], cuts[i + 1]) for i in range(threads) if cuts[i + 1] > cuts[i]]
    jobs = [(sd, lo, hi) for sd in (0, *seeds) for lo, hi in chunks]
    with ThreadPoolExecutor(threads) as ex:
        outs = list(ex.map(lambda j: oracle_run(orc, t, codes[j[1]:j[2]], seed=int(j[0])), jobs))
    orig = [o for (sd, _, _), o in zip(jobs, outs) if sd == 0]
    per_snp = {k2: np.concatenate([o[k2] for o in orig]) for k2 in ("chisq", "df", "pval", "rank", "pivot", "coef")}
    minp = {int(sd): min(o["min_pval"] for (s2, _, _), o in zip(jobs, outs) if s2 == sd) for sd in seeds}
    return per_snp, minp


def pick_fields(r, idx):
    return {k2: r[k2][idx] for k2 in ("lrt_df", "rank", "pivot", "pval", "lrt_chisq", "coef")}


def compare(r, o, N=None):
    """N: number of samples, for the reference's own chi-square rounding floor (assert_logp_close); matters from n ~ 5,000 on."""
    assert np.array_equal(r["lrt_df"], o["df"])
    assert np.array_equal(r["rank"], o["rank"])
    assert np.array_equal(r["pivot"], o["pivot"])
    if N is None:
        assert_logp_close(r["pval"], o["pval"])
    else:
        assert_logp_close(r["pval"], o["pval"], o["chisq"], o["df"], N)
    np.testing.assert_allclose(r["lrt_chisq"], o["chisq"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(r["coef"], o["coef"], rtol=1e-8, atol=1e-10)


# ------------------------------------------------------------------------------------------------ K1 decode
@pytest.mark.parametrize("n,n_raw,model", [(4, 4, "simple"), (5, 9, "additive"), (6, 6, "domgxe"), (127, 200, "additive"), (128, 128, "domgxe"),
                                           (129, 400, "simple"), (300, 300, "domgxe"), (1000, 1777, "additive")])
def test_decode_bit_exact(fx, orc, n, n_raw, model):
    from flexlmm_b200 import synth
    t = synth.task(n, 77, model=model, n_raw=n_raw, seed=n)
    rng = np.random.default_rng(n)
    t["pgen_order"] = (rng.permutation(t["n_raw"])[: t["n"]] + 1).astype(np.int32)       # non-monotone subset + reorder
    codes = t["codes_raw"][:, t["pgen_order"] - 1]
    with make(fx, t, perm=False) as fm:
        vi = rng.permutation(77).astype(np.int32) + 1                                    # arbitrary variant order
        X, cd = fm.decode_tile(vi)
    assert np.array_equal(cd, codes[vi - 1])
    s = X.shape[1] // 77
    snp_cols = [j for j in range(t["M0"].shape[1]) if not (np.array_equal(t["M0"][:, j], t["M1"][:, j]) and np.array_equal(t["M0"][:, j], t["M2"][:, j]))]
    assert len(snp_cols) == s
    for q, v in enumerate(vi):
        g = codes[v - 1]
        for ti, j in enumerate(snp_cols):
            want = np.where(g == 0, t["M0"][:, j], np.where(g == 1, t["M1"][:, j], t["M2"][:, j]))
            assert np.array_equal(X[:, q * s + ti], want)                               # bit-exact design columns


@pytest.mark.parametrize("mode", [0x02, 0x10])
def test_decode_from_pgen_file_all_vrtypes(fx, orc, tmp_path, mode):
    from flexlmm_b200 import synth
    from pgen_writer import write_pgen
    t = synth.task(211, 60, model="additive", n_raw=300, seed=2)
    c = t["codes_raw"].copy()
    c[4] = 0; c[5] = 2; c[5, 3] = 1
    vt = ([0, 1, 2, 3, 4, 6, 7, 2, 3, 1] * 6) if mode == 0x10 else None
    path = tmp_path / "t.pgen"
    path.write_bytes(write_pgen(c, vt, mode=mode))
    with make(fx, t, perm=False, pgen_path=path) as fm:
        vi = np.arange(1, 61, dtype=np.int32)
        _, cd = fm.decode_tile(vi, want_X=False)
    assert np.array_equal(cd, c[:, t["pgen_order"] - 1])
    buf = path.read_bytes()
    for v in range(60):
        assert np.array_equal(orc.pgen_read_hardcalls(buf, v)[t["pgen_order"] - 1], cd[v])


def test_missing_genotype_is_fatal_like_the_reference(fx):
    from flexlmm_b200 import synth
    t = synth.task(100, 40, model="additive", seed=3)
    c = t["codes_raw"].copy(); c[17, 5] = 3
    t["packed"] = synth.pack_2bit(c)
    with make(fx, t, perm=False) as fm:
        _, cd = fm.decode_tile(np.arange(1, 41, dtype=np.int32), want_X=False)
        assert cd[17, 5] == 3
        with pytest.raises(fx.FlxError) as e:
            fm.run(np.arange(1, 41, dtype=np.int32))
        assert e.value.code == -2 and "var_idx[17]" in str(e.value)
        fm.run(np.arange(1, 17, dtype=np.int32))            # variants without missing calls still run


# ------------------------------------------------------------------------------------------------ K2 whiten
@pytest.mark.parametrize("n", [12, 129, 500, 1000])
def test_whiten_matches_forwardsolve(fx, orc, n):
    from flexlmm_b200 import synth
    L = synth.covariance_factor(n, seed=n)
    rng = np.random.default_rng(n)
    X = np.asfortranarray(rng.integers(0, 3, size=(n, 150)).astype(float) * rng.standard_normal((n, 1)))
    with fx.FitModel() as fm:
        fm.set_whitening(L)
        W = fm.whiten(X)
    Wo = orc.forwardsolve(L, X)
    assert np.abs(W - Wo).max() <= 1e-12 * np.abs(Wo).max()


def test_whitening_accepts_explicit_inverse(fx):
    from flexlmm_b200 import synth
    L = synth.covariance_factor(200, seed=1)
    X = np.asfortranarray(np.random.default_rng(0).standard_normal((200, 5)))
    with fx.FitModel() as a, fx.FitModel() as b:
        a.set_whitening(L); b.set_whitening(np.linalg.inv(L), is_inverse=True)
        np.testing.assert_allclose(a.whiten(X), b.whiten(X), rtol=1e-12, atol=1e-13)


# ------------------------------------------------------------------------------------------------ K3 fit + LRT
@pytest.mark.parametrize("n,M,model,n_raw", [(300, 200, "additive", None), (300, 130, "domgxe", 333), (1000, 300, "simple", None),
                                             (64, 500, "domgxe", None), (2000, 64, "additive", 2100)])
def test_fit_matches_oracle(fx, orc, n, M, model, n_raw):
    from flexlmm_b200 import synth
    t = synth.task(n, M, model=model, n_raw=n_raw, seed=n + M, effect=0.2)
    codes = t["codes_raw"][:, t["pgen_order"] - 1]
    with make(fx, t, perm=False) as fm:
        r = fm.run(np.arange(1, M + 1, dtype=np.int32))
    compare(r, oracle_run(orc, t, codes))
    assert r["nvars_tested"] == M


def test_rank_deficient_snps(fx, orc):
    """Monomorphic SNPs, SNPs without heterozygotes (all-zero I(x==1) column), a SNP equal to a covariate: the dqrdc2 rank rule."""
    from flexlmm_b200 import synth
    n = 240
    t = synth.task(n, 12, model="domgxe", seed=7)
    c = t["codes_raw"].copy()
    c[0] = 0; c[1] = 2; c[2] = 1                                  # monomorphic
    c[3] = np.where(c[3] == 1, 0, c[3]); c[4] = np.where(c[4] == 1, 2, c[4])   # no hets
    c[5] = c[6]                                                   # duplicated SNP
    c[7] = 0; c[7, :3] = 2                                        # the test-asset pattern: few carriers, no hets
    t["packed"] = synth.pack_2bit(c)
    with make(fx, t, perm=False) as fm:
        r = fm.run(np.arange(1, 13, dtype=np.int32))
    o = oracle_run(orc, t, c)
    compare(r, o)
    assert r["lrt_df"][0] == 0 and np.isnan(r["pval"][0])           # lrt_df == 0 -> NA (fit_model.nf:138)
    assert r["lrt_df"][3] == 2 and r["rank"][3] == 4
    assert r["lrt_chisq"][5] == r["lrt_chisq"][6]


def test_covariate_collinear_with_snp(fx, orc):
    """A later NULL-model column collinear with (1, x): dqrdc2 rejects the covariate, not x; df drops to 0."""
    from flexlmm_b200 import synth
    n = 150
    t = synth.task(n, 6, model="additive", seed=11)
    g0 = t["codes_raw"][0].astype(float)
    cov = 2.0 * g0 - 1.0
    one = np.ones(n)
    t["M0"], t["M1"], t["M2"] = (np.column_stack([one, np.full(n, v), cov]) for v in (0.0, 1.0, 2.0))
    Xn = np.column_stack([one, cov])
    y = np.random.default_rng(1).standard_normal(n)
    t["y_mm"] = orc.forwardsolve(t["L"], y)[:, 0]; t["X_null"] = orc.forwardsolve(t["L"], Xn)
    nm = orc.fit_null_model(t["y_mm"], t["X_null"])
    t["ll_null"], t["df_null"] = nm["ll"], nm["df"]
    with make(fx, t, perm=False) as fm:
        r = fm.run(np.arange(1, 7, dtype=np.int32))
    o = oracle_run(orc, t, t["codes_raw"])
    compare(r, o)
    assert r["rank"][0] == 2 and r["pivot"][0].tolist() == [1, 2, 3] and r["lrt_df"][0] == 0


def test_extra_covariate_in_real_model_only(fx, orc):
    """Nested models where the real model adds an SNP-independent column too (y ~ x + cov1 vs y ~ 1): lrt_df = 2."""
    from flexlmm_b200 import synth
    n = 180
    t = synth.task(n, 40, model="additive", seed=13)
    Xn = np.ones((n, 1))
    t["X_null"] = orc.forwardsolve(t["L"], Xn)
    nm = orc.fit_null_model(t["y_mm"], t["X_null"])
    t["ll_null"], t["df_null"] = nm["ll"], nm["df"]
    with make(fx, t, perm=False) as fm:
        r = fm.run(np.arange(1, 41, dtype=np.int32))
    o = oracle_run(orc, t, t["codes_raw"])
    compare(r, o)
    assert set(r["lrt_df"].tolist()) == {2}


def test_huge_effect_underflows_to_zero_like_r(fx, orc):
    from flexlmm_b200 import synth
    n = 400
    t = synth.task(n, 5, model="simple", seed=17)
    g = t["codes_raw"][0].astype(float)
    y = 50.0 * g + 1e-3 * np.random.default_rng(0).standard_normal(n)
    t["y_mm"] = orc.forwardsolve(t["L"], y)[:, 0]
    nm = orc.fit_null_model(t["y_mm"], t["X_null"])
    t["ll_null"], t["df_null"] = nm["ll"], nm["df"]
    with make(fx, t, perm=False) as fm:
        r = fm.run(np.arange(1, 6, dtype=np.int32))
    o = oracle_run(orc, t, t["codes_raw"])
    assert o["pval"][0] == 0.0 and r["pval"][0] == 0.0
    np.testing.assert_allclose(r["lrt_chisq"], o["chisq"], rtol=1e-6)


# ------------------------------------------------------------------------------------------------ C1: the reference's own test assets
def test_c1_test_assets_golden(fx, tmp_path):
    C1 = json.load(open(os.path.join(ROOT, "tests", "golden", "c1_test_asset.json")))
    path = tmp_path / "in.pgen"
    path.write_bytes(bytes.fromhex(C1["pgen_hex"]))
    for case in C1["cases"]:
        n = len(case["y_mm"])
        with fx.FitModel() as fm:
            fm.set_whitening(np.array(case["Lmat"]))
            fm.set_null(np.array(case["X_null"]).reshape(n, -1), case["y_mm"], case["ll_null"], case["df_null"])
            fm.set_design(np.array(case["M0"]), np.array(case["M1"]), np.array(case["M2"]))
            fm.set_perm(case["y_pred"], np.array(case["U1"]), case["e_p"], [1, 2])
            fm.open_pgen(path, np.arange(1, n + 1))
            for chrom, g in case["chr"].items():
                r = fm.run(np.array(g["var_idx"], dtype=np.int32))
                assert r["lrt_df"].tolist() == g["df"] and r["rank"].tolist() == g["rank"] and r["pivot"].tolist() == g["pivot"]
                assert_logp_close(r["pval"], [np.nan if v is None else v for v in g["pval"]])
                np.testing.assert_allclose(r["lrt_chisq"], g["chisq"], rtol=1e-8, atol=1e-9)
                np.testing.assert_allclose(r["coef"], g["coef"], rtol=1e-8, atol=1e-10)
                for si, sd in enumerate(("1", "2")):
                    assert_logp_close([r["min_pval"][si]], [g["perm"][sd]["min_pval"]])
                    assert fx.perm_indices(int(sd), len(case["e_p"])).tolist() == g["perm"][sd]["idx"]
                assert r["nvars_tested"] == g["perm"]["1"]["nvars"]


def test_golden_synth_small(fx):
    from flexlmm_b200 import synth
    G = np.load(os.path.join(ROOT, "tests", "golden", "synth_small.npz"))
    for name in ("additive", "domgxe", "simple"):
        n, M, P, n_raw, seed = G[f"{name}.args"].tolist()
        t = synth.task(n, M, model=name, P=P, n_raw=None if n_raw == n else n_raw, seed=seed)
        with make(fx, t) as fm:
            r = fm.run(np.arange(1, M + 1, dtype=np.int32))
        compare(r, {k2: G[f"{name}.{k2}"] for k2 in ("chisq", "df", "pval", "rank", "pivot", "coef")})
        assert_logp_close(r["min_pval"], G[f"{name}.min_pval"])


# ------------------------------------------------------------------------------------------------ K4 permutations
@pytest.mark.parametrize("n,M,model,P", [(200, 150, "additive", 5), (160, 90, "domgxe", 4), (500, 40, "simple", 131)])
def test_permutation_minp_matches_oracle(fx, orc, n, M, model, P):
    from flexlmm_b200 import synth
    t = synth.task(n, M, model=model, P=P, seed=n)
    codes = t["codes_raw"]
    with make(fx, t) as fm:
        r = fm.run(np.arange(1, M + 1, dtype=np.int32))
    nseed = min(P, 6)
    want = np.array([oracle_run(orc, t, codes, seed=int(s))["min_pval"] for s in t["seeds"][:nseed]])
    assert_logp_close(r["min_pval"][:nseed], want)
    assert r["nvars_tested"] == M
    thr_gpu = fx.minp_threshold(r["min_pval"][:nseed], 0.05)
    assert abs(np.log10(thr_gpu) - np.log10(orc.quantile7(want, 0.05))) <= RTOL * abs(np.log10(thr_gpu)) + ATOL


def test_permutation_with_rank_deficient_snps(fx, orc):
    from flexlmm_b200 import synth
    t = synth.task(120, 30, model="domgxe", P=3, seed=23)
    c = t["codes_raw"].copy()
    for v in range(0, 30, 3):
        c[v] = np.where(c[v] == 1, 0, c[v])            # df = 2 SNPs mixed with df = 3 SNPs
    c[1] = 0                                            # df = 0 (NA) is skipped by min(..., na.rm = TRUE)
    t["packed"] = synth.pack_2bit(c)
    with make(fx, t) as fm:
        r = fm.run(np.arange(1, 31, dtype=np.int32))
    want = np.array([oracle_run(orc, t, c, seed=int(s))["min_pval"] for s in t["seeds"]])
    assert_logp_close(r["min_pval"], want)


# ------------------------------------------------------------------------------------------------ BASELINE-size properties (n = 5,000)
@pytest.fixture(scope="module")
def big(fx):
    from flexlmm_b200 import synth
    t = synth.task(5000, 6000, model="additive", P=100, seed=1)
    fm = make(fx, t)
    r = fm.run(np.arange(1, 6001, dtype=np.int32))
    yield t, fm, r
    fm.close()


def test_big_batching_and_order_invariance(fx, big):
    """Per-SNP results do not depend on which other SNPs share the batch / tile, nor on their order (bit-exact)."""
    t, fm, r = big
    rng = np.random.default_rng(0)
    sub = np.sort(rng.choice(6000, size=777, replace=False)).astype(np.int32) + 1
    r2 = fm.run(sub, minp=False)
    assert np.array_equal(r2["lrt_chisq"], r["lrt_chisq"][sub - 1])
    assert np.array_equal(r2["coef"], r["coef"][sub - 1])
    perm = rng.permutation(6000).astype(np.int32) + 1
    r3 = fm.run(perm, minp=False)
    assert np.array_equal(r3["pval"], r["pval"][perm - 1])


def test_big_sample_against_oracle(fx, orc, big):
    t, fm, r = big
    pick = np.array([0, 1, 2, 3000, 5999])
    o = oracle_run(orc, t, t["codes_raw"][pick])
    compare({k2: r[k2][pick] for k2 in ("lrt_df", "rank", "pivot", "pval", "lrt_chisq", "coef")}, o)


def test_c2_shape_256_snps_and_two_seeds_against_oracle(fx, orc, big):
    """The headline shape (BASELINE configs[1]: n = 5,000, y ~ x + cov1 vs y ~ cov1, int8 route): 256 SNPs spread over the
    run against the oracle for every per-SNP output, and min_pval of three seeds against the oracle's PERM tasks over the
    same 256 SNPs (fit_model.nf:96-143 literally: the whole design re-whitened per SNP and per task)."""
    t, fm, r = big
    assert fm.stats()["int8_path"] == 1
    pick = np.unique(np.linspace(0, 5999, 256).astype(int))
    o, minp = oracle_parallel(orc, t, t["codes_raw"][pick], seeds=(1, 2, 100))
    compare(pick_fields(r, pick), o, N=t["n"])
    r_sub = fm.run((pick + 1).astype(np.int32), per_snp=False)
    assert r_sub["nvars_tested"] == pick.size
    assert_logp_close(r_sub["min_pval"][[0, 1, 99]], [minp[1], minp[2], minp[100]])
    # same SNPs on the FP64 route
    with fx.FitModel(int8=False) as f2:
        f2.set_whitening(t["L"]); f2.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"]); f2.set_design(t["M0"], t["M1"], t["M2"])
        f2.set_perm(t["y_pred"], t["U1"], t["e_p"], [1, 2, 100]); f2.set_packed(t["packed"], t["n_raw"], t["pgen_order"])
        r64 = f2.run((pick + 1).astype(np.int32))
        assert f2.stats()["int8_path"] == 0 and f2.stats()["fallback_reason"] == 1
    compare(r64, o, N=t["n"])
    assert_logp_close(r64["min_pval"], [minp[1], minp[2], minp[100]])


def test_big_perm_path_agrees_with_orig_path(fx, big):
    """K4 (one contraction for all permutations) must equal running the ORIG path on the permuted phenotype itself:
    min over SNPs of the per-SNP p-values == min_pval[seed].  Size-independent check of the permutation identity."""
    t, fm, r = big
    n = t["n"]
    for si in (0, 57, 99):
        seed = int(t["seeds"][si])
        idx = fx.perm_indices(seed, len(t["e_p"]))
        y_star = t["y_pred"] + t["U1"] @ t["e_p"][idx - 1]
        Q, _ = np.linalg.qr(t["X_null"])
        e = y_star - Q @ (Q.T @ y_star)
        ll = 0.5 * (-n * (np.log(2 * np.pi) + 1 - np.log(n) + np.log(e @ e)))
        with fx.FitModel() as f2:
            f2.set_whitening(t["L"]); f2.set_null(t["X_null"], y_star, ll, t["df_null"]); f2.set_design(t["M0"], t["M1"], t["M2"])
            f2.set_packed(t["packed"], t["n_raw"], t["pgen_order"])
            r2 = f2.run(np.arange(1, 6001, dtype=np.int32))
        assert_logp_close([r["min_pval"][si]], [np.nanmin(r2["pval"])])


def test_big_scale_invariance(fx, big):
    """lrt_chisq is invariant to rescaling the phenotype (linearity of the fit)."""
    t, fm, r = big
    n = t["n"]
    with fx.FitModel() as f2:
        f2.set_whitening(t["L"])
        f2.set_null(t["X_null"], 3.0 * t["y_mm"], t["ll_null"] - n * np.log(3.0), t["df_null"])
        f2.set_design(t["M0"], t["M1"], t["M2"]); f2.set_packed(t["packed"], t["n_raw"], t["pgen_order"])
        r2 = f2.run(np.arange(1, 1001, dtype=np.int32))
    np.testing.assert_allclose(r2["lrt_chisq"], r["lrt_chisq"][:1000], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(r2["coef"], 3.0 * r["coef"][:1000], rtol=1e-9, atol=1e-12)


def test_empty_and_single_variant(fx, big):
    t, fm, r = big
    r0 = fm.run(np.zeros(0, dtype=np.int32))
    assert r0["nvars_tested"] == 0 and r0["lrt_chisq"].size == 0 and (r0["min_pval"] == 2.0).all()
    r1 = fm.run(np.array([4321], dtype=np.int32))
    assert r1["lrt_chisq"][0] == r["lrt_chisq"][4320]
    with pytest.raises(fx.FlxError):
        fm.run(np.array([0], dtype=np.int32))             # var_idx is 1-based
    with pytest.raises(fx.FlxError):
        fm.run(np.array([6001], dtype=np.int32))


def test_default_batching_ramp_refit_and_staged_results(fx):
    """80,000 SNPs through the default int8 batching (ramp 18,944 / 37,888 / 75,776-SNP batches from host memory, per-batch
    result copies into pinned staging, FP64 refit of the monomorphic SNPs scattered over the batches) must reproduce, bit for
    bit, the run with one tile per SM and the run from HBM-resident genotypes."""
    from flexlmm_b200 import synth
    n, M = 200, 80000
    t = synth.task(n, M, model="additive", P=4, seed=41)
    c = t["codes_raw"]
    c[::997] = 0                                   # monomorphic: refit in every batch
    c[5::1499] = 2
    t["packed"] = synth.pack_2bit(c)
    vi = np.arange(1, M + 1, dtype=np.int32)
    res = []
    for bt, resident in ((0, False), (148, False), (0, True)):
        with make(fx, t, perm=True, batch_tiles=bt) as fm:
            if resident:
                fm.make_resident()
            r = fm.run(vi)
            st = fm.stats()
            assert st["int8_path"] == 1 and st["refit_snps"] >= M // 997
            res.append(r)
    for r in res[1:]:
        for k2 in ("lrt_chisq", "pval", "lrt_df", "rank", "pivot", "coef", "min_pval"):
            assert np.array_equal(res[0][k2], r[k2], equal_nan=True), k2
    assert (res[0]["lrt_df"][::997] == 0).all() and np.isnan(res[0]["pval"][::997]).all()


def test_run_reuses_result_arrays(fx, big):
    """run(out=previous result) overwrites the caller's arrays in place (no new allocations on a streaming caller's path)."""
    t, fm, r = big
    vi = np.arange(1, 501, dtype=np.int32)
    a = fm.run(vi)
    keep = {k2: a[k2].copy() for k2 in ("lrt_chisq", "pval", "lrt_df", "rank", "pivot", "coef")}
    ids = {k2: a[k2].ctypes.data for k2 in keep}
    a["lrt_chisq"][:] = -1.0
    b = fm.run(vi, out=a)
    for k2 in keep:
        assert b[k2].ctypes.data == ids[k2]
        assert np.array_equal(b[k2], keep[k2], equal_nan=True)
    c = fm.run(vi[:100], out=b)                     # other size: fresh arrays
    assert c["lrt_chisq"].shape == (100,) and np.array_equal(c["lrt_chisq"], keep["lrt_chisq"][:100])


# ------------------------------------------------------------------------------------------------ K2q: exact int8 tensor-core path
def _run(fx, t, int8, M, perm=True):
    fm = fx.FitModel(int8=int8)
    fm.set_whitening(t["L"]); fm.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"]); fm.set_design(t["M0"], t["M1"], t["M2"])
    if perm and t["P"]:
        fm.set_perm(t["y_pred"], t["U1"], t["e_p"], t["seeds"])
    fm.set_packed(t["packed"], t["n_raw"], t["pgen_order"])
    r = fm.run(np.arange(1, M + 1, dtype=np.int32))
    st = fm.stats()
    fm.close()
    return r, st


@pytest.mark.parametrize("n,M,model", [(700, 300, "additive"), (5000, 2000, "additive"), (1300, 500, "simple")])
def test_int8_path_equals_fp64_path(fx, n, M, model):
    """The int8 digit-plane path (K2q) and the FP64 DMMA path (K2) are two arithmetics for the same numbers."""
    from flexlmm_b200 import synth
    t = synth.task(n, M, model=model, P=7, seed=n, effect=0.1)
    r8, st8 = _run(fx, t, True, M)
    r64, st64 = _run(fx, t, False, M)
    assert st8["int8_path"] == 1 and st64["int8_path"] == 0
    assert np.array_equal(r8["lrt_df"], r64["lrt_df"]) and np.array_equal(r8["pivot"], r64["pivot"])
    np.testing.assert_allclose(r8["lrt_chisq"], r64["lrt_chisq"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(r8["coef"], r64["coef"], rtol=1e-9, atol=1e-12)
    assert_logp_close(r8["pval"], r64["pval"])
    assert_logp_close(r8["min_pval"], r64["min_pval"])


def test_int8_path_refits_collinear_snps_in_fp64(fx, orc):
    """Monomorphic / nearly-constant SNPs lose digits in x'V^-1x - |u|^2: they must be flagged and re-fitted on the FP64 path."""
    from flexlmm_b200 import synth
    n, M = 400, 60
    t = synth.task(n, M, model="additive", P=3, seed=5)
    c = t["codes_raw"].copy()
    c[0] = 0; c[1] = 2; c[2] = 1          # monomorphic: x collinear with the intercept
    c[3] = 0; c[3, 0] = 1                 # a single carrier: legitimate, tiny variance
    t["packed"] = synth.pack_2bit(c)
    r8, st8 = _run(fx, t, True, M)
    assert st8["int8_path"] == 1 and st8["refit_snps"] >= 3
    o = oracle_run(orc, t, c)
    compare(r8, o)
    want = np.array([oracle_run(orc, t, c, seed=int(s))["min_pval"] for s in t["seeds"]])
    assert_logp_close(r8["min_pval"], want)
    assert r8["lrt_df"][0] == 0 and np.isnan(r8["pval"][0])


def test_int8_path_for_an_indicator_term(fx, orc):
    """A single SNP term with integer values other than x itself: y ~ I(x == 1) + cov1."""
    from flexlmm_b200 import synth
    n, M = 350, 90
    t = synth.task(n, M, model="additive", P=2, seed=8)
    cov = t["M0"][:, 2]
    one = np.ones(n)
    t["M0"], t["M1"], t["M2"] = (np.column_stack([one, np.full(n, v), cov]) for v in (0.0, 1.0, 0.0))
    r8, st8 = _run(fx, t, True, M)
    assert st8["int8_path"] == 1
    compare(r8, oracle_run(orc, t, t["codes_raw"]))


def test_int8_path_not_used_for_non_separable_terms(fx, orc):
    """A term that is not g(x) * c[sample] with an integer triple g (here exp(x * cov)) stays on the FP64 route."""
    from flexlmm_b200 import synth
    n, M = 300, 50
    t = synth.task(n, M, model="additive", P=0, seed=1)
    cov = t["M0"][:, 2]
    one = np.ones(n)
    t["M0"], t["M1"], t["M2"] = (np.column_stack([one, np.exp(0.3 * v * cov), cov]) for v in (0.0, 1.0, 2.0))
    r, st = _run(fx, t, True, M, perm=False)
    assert st["int8_path"] == 0
    compare(r, oracle_run(orc, t, t["codes_raw"]))


@pytest.mark.parametrize("n,M", [(300, 200), (1500, 700)])
def test_int8_path_multi_term_model(fx, orc, n, M):
    """y ~ x + I(x==1) + x:cov1 + cov1: three SNP terms, two integer triples, two per-sample scales -> six int8 passes.
    Compared with the FP64 route (same numbers, other arithmetic) and with the oracle (per-SNP results and min-p)."""
    from flexlmm_b200 import synth
    t = synth.task(n, M, model="domgxe", P=5, seed=n + 1, effect=0.1)
    r8, st8 = _run(fx, t, True, M)
    r64, st64 = _run(fx, t, False, M)
    assert st8["int8_path"] == 1 and st64["int8_path"] == 0
    assert st8["refit_snps"] < M // 4
    assert np.array_equal(r8["lrt_df"], r64["lrt_df"]) and np.array_equal(r8["pivot"], r64["pivot"])
    np.testing.assert_allclose(r8["lrt_chisq"], r64["lrt_chisq"], rtol=1e-8, atol=1e-9)
    assert_logp_close(r8["pval"], r64["pval"])
    assert_logp_close(r8["min_pval"], r64["min_pval"])
    if n <= 300:
        compare(r8, oracle_run(orc, t, t["codes_raw"]))
        want = np.array([oracle_run(orc, t, t["codes_raw"], seed=int(s))["min_pval"] for s in t["seeds"]])
        assert_logp_close(r8["min_pval"], want)


def test_int8_path_four_terms_two_scales_two_codes(fx, orc):
    """y ~ x + I(x==1) + x:cov + I(x==1):cov + cov: every (scale, genotype code) combination in use — 12 int8 forms."""
    from flexlmm_b200 import synth
    n, M = 380, 150
    t = synth.task(n, M, model="additive", P=3, seed=23, effect=0.1)
    cov = t["M0"][:, 2]
    one = np.ones(n)
    t["M0"], t["M1"], t["M2"] = (np.column_stack([one, np.full(n, float(v)), np.full(n, float(v == 1)), cov, v * cov, (v == 1) * cov]) for v in (0, 1, 2))
    r8, st8 = _run(fx, t, True, M)
    r64, st64 = _run(fx, t, False, M)
    assert st8["int8_path"] == 1 and st64["int8_path"] == 0 and st8["int8_passes"] == 8
    assert np.array_equal(r8["lrt_df"], r64["lrt_df"]) and np.array_equal(r8["pivot"], r64["pivot"])
    np.testing.assert_allclose(r8["lrt_chisq"], r64["lrt_chisq"], rtol=1e-8, atol=1e-9)
    assert_logp_close(r8["min_pval"], r64["min_pval"])
    compare(r8, oracle_run(orc, t, t["codes_raw"]))


@pytest.mark.parametrize("n,n_raw,prefix", [(301, 301, True), (130, 257, True), (203, 333, False), (64, 67, True)])
def test_code_tiles_any_record_alignment(fx, orc, n, n_raw, prefix):
    """K1q builds its 2-bit code tiles from the records' own 32-bit words: record lengths of any residue mod 4 (every start alignment),
    records longer than the model's samples (identity-prefix order: the staged path masks the tail), and a gathering order."""
    from flexlmm_b200 import synth
    M = 150
    t = synth.task(n, M, model="domgxe", P=2, n_raw=n_raw, seed=n + n_raw, effect=0.1)
    if prefix:
        t["pgen_order"] = np.arange(1, n + 1, dtype=np.int32)
    r8, st8 = _run(fx, t, True, M)
    assert st8["int8_path"] == 1
    codes = t["codes_raw"][:, t["pgen_order"] - 1]
    compare(r8, oracle_run(orc, t, codes))
    want = np.array([oracle_run(orc, t, codes, seed=int(sd))["min_pval"] for sd in t["seeds"]])
    assert_logp_close(r8["min_pval"], want)


def test_code_tiles_negative_triple(fx, orc):
    """A centred genotype coding (x - 1: values -1, 0, 1) is a tile set whose codes are not its values: the expanders' and the
    epilogue's table path."""
    from flexlmm_b200 import synth
    n, M = 333, 140
    t = synth.task(n, M, model="additive", P=2, seed=77, effect=0.1)
    cov = t["M0"][:, 2]
    one = np.ones(n)
    t["M0"], t["M1"], t["M2"] = (np.column_stack([one, np.full(n, float(v - 1)), cov]) for v in (0, 1, 2))
    r8, st8 = _run(fx, t, True, M)
    r64, st64 = _run(fx, t, False, M)
    assert st8["int8_path"] == 1 and st64["int8_path"] == 0
    assert np.array_equal(r8["lrt_df"], r64["lrt_df"]) and np.array_equal(r8["pivot"], r64["pivot"])
    compare(r8, oracle_run(orc, t, t["codes_raw"]))
    assert_logp_close(r8["min_pval"], r64["min_pval"])


def test_int8_path_factor_interaction_and_rare_variants(fx, orc):
    """The reference's test formula shape, y ~ x + I(x==1) + x:group + group with a two-level factor (indicator scale), on
    SNPs many of which have no hom-alt call: there I(x==1) duplicates x exactly and is dropped by the rank rule WITHOUT a
    trip through the FP64 refit."""
    from flexlmm_b200 import synth
    n, M = 420, 160
    t = synth.task(n, M, model="additive", P=3, seed=12, effect=0.1)
    rng = np.random.default_rng(9)
    grp = (rng.random(n) < 0.4).astype(float)
    c = t["codes_raw"].copy()
    c[:80][c[:80] == 2] = 1                       # the first 80 SNPs carry no hom-alt genotype
    c[5] = np.where(grp > 0, c[5], 0)             # x:group == x for this SNP: a second exact duplicate
    t["packed"] = synth.pack_2bit(c)
    one = np.ones(n)
    t["M0"], t["M1"], t["M2"] = (np.column_stack([one, np.full(n, float(v)), np.full(n, float(v == 1)), v * grp, grp]) for v in (0, 1, 2))
    Xn = np.column_stack([one, grp])
    t["X_null"] = orc.forwardsolve(t["L"], Xn)
    nm = orc.fit_null_model(t["y_mm"], t["X_null"])
    t["ll_null"], t["df_null"] = nm["ll"], nm["df"]
    t["y_pred"], t["U1"], t["e_p"] = nm["y_pred"], nm["U1"], nm["e_p"]
    r8, st8 = _run(fx, t, True, M)
    assert st8["int8_path"] == 1
    assert st8["refit_snps"] <= 8, st8["refit_snps"]
    o = oracle_run(orc, t, c)
    compare(r8, o)
    assert (r8["lrt_df"][:80] <= 2).all()
    want = np.array([oracle_run(orc, t, c, seed=int(s))["min_pval"] for s in t["seeds"]])
    assert_logp_close(r8["min_pval"], want)


def _factor_task(orc, n, M, levels, seed, dominance=True, P=3):
    """y ~ x [+ I(x == 1)] + x:group + group vs y ~ group with a `levels`-level factor (treatment contrasts): the reference's own
    test formula shape (conf/test.config:35-36: covar1 is a factor)."""
    from flexlmm_b200 import synth
    t = synth.task(n, M, model="additive", P=P, seed=seed, effect=0.1)
    rng = np.random.default_rng(seed)
    g = rng.integers(0, levels, size=n)
    ind = [(g == lv).astype(float) for lv in range(1, levels)]
    one = np.ones(n)

    def mm(v):
        cols = [one, np.full(n, float(v))] + ([np.full(n, float(v == 1))] if dominance else []) + ind + [v * i for i in ind]
        return np.column_stack(cols)
    t["M0"], t["M1"], t["M2"] = mm(0), mm(1), mm(2)
    t["X_null"] = orc.forwardsolve(t["L"], np.column_stack([one] + ind))
    nm = orc.fit_null_model(t["y_mm"], t["X_null"])
    t["ll_null"], t["df_null"] = nm["ll"], nm["df"]
    t["y_pred"], t["U1"], t["e_p"] = nm["y_pred"], nm["U1"], nm["e_p"]
    return t


@pytest.mark.parametrize("levels,dominance,passes", [(3, False, 6), (3, True, 8), (4, False, 8), (2, True, 5)])
def test_int8_path_multi_level_factor_interaction(fx, orc, levels, dominance, passes):
    """x:factor with more than two levels used to drop to the FP64 route without a word (three scales).  The indicator columns are now
    folded into the int8 genotype tiles as sample masks (up to four tile sets sharing the unscaled digit planes); the two-level case
    keeps its cheaper plain plan (5 pass units)."""
    t = _factor_task(orc, 380, 140, levels, seed=20 + levels, dominance=dominance)
    r8, st8 = _run(fx, t, True, 140)
    assert st8["int8_path"] == 1 and st8["fallback_reason"] == 0 and st8["int8_passes"] == passes
    r64, st64 = _run(fx, t, False, 140)
    assert st64["int8_path"] == 0
    o = oracle_run(orc, t, t["codes_raw"])
    compare(r8, o)
    compare(r64, o)
    want = np.array([oracle_run(orc, t, t["codes_raw"], seed=int(sd))["min_pval"] for sd in t["seeds"]])
    assert_logp_close(r8["min_pval"], want)
    assert_logp_close(r64["min_pval"], want)


def test_fp64_fallback_is_reported(fx, orc, capfd):
    """A design the int8 plan cannot hold (x:factor with five levels plus dominance: five tile sets) runs on the FP64 route and SAYS so:
    flx_stats.fallback_reason and one line on stderr."""
    t = _factor_task(orc, 300, 60, 5, seed=31, dominance=True, P=0)
    with fx.FitModel() as fm:
        fm.set_whitening(t["L"]); fm.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"]); fm.set_design(t["M0"], t["M1"], t["M2"])
        fm.set_packed(t["packed"], t["n_raw"], t["pgen_order"])
        r = fm.run(np.arange(1, 61, dtype=np.int32))
        st = fm.stats()
    assert st["int8_path"] == 0 and st["fallback_reason"] == 3
    assert "FP64 route" in capfd.readouterr().err
    compare(r, oracle_run(orc, t, t["codes_raw"]))


def test_int8_digit_planes_when_the_leading_digit_would_be_128(fx, orc):
    """Regression: a row (or column) whose largest entry has a mantissa above 127/128 needs one more bit of headroom in
    its power-of-two scale, or its leading balanced digit is 128 and wraps.  V^-1 = diag(1 / sigma^2) is built so that
    half of the rows sit exactly there."""
    from flexlmm_b200 import synth
    n, M = 330, 120
    t = synth.task(n, M, model="additive", P=3, seed=17, effect=0.1)
    y_raw = t["L"] @ t["y_mm"]; Xn_raw = t["L"] @ t["X_null"]
    rng = np.random.default_rng(2)
    target = np.where(rng.random(n) < 0.5, 0.9995, 0.7) * 2.0 ** rng.integers(-2, 3, n)     # wanted 0.5 * V^-1_ii
    sigma = np.sqrt(0.5 / target)
    t["L"] = np.diag(sigma)
    t["y_mm"] = y_raw / sigma; t["X_null"] = Xn_raw / sigma[:, None]
    nm = orc.fit_null_model(t["y_mm"], t["X_null"])
    t["ll_null"], t["df_null"] = nm["ll"], nm["df"]
    t["y_pred"], t["U1"], t["e_p"] = nm["y_pred"], nm["U1"], nm["e_p"]
    r8, st8 = _run(fx, t, True, M)
    assert st8["int8_path"] == 1 and st8["refit_snps"] == 0
    compare(r8, oracle_run(orc, t, t["codes_raw"]))
    want = np.array([oracle_run(orc, t, t["codes_raw"], seed=int(s))["min_pval"] for s in t["seeds"]])
    assert_logp_close(r8["min_pval"], want)


def test_five_plane_option_stays_within_the_parity_bar(fx):
    """FLX_FLAG_PLANES_5 (40-bit digit planes, one sixth less K2q work): documented as ~1e-9 relative in lrt_chisq; the
    parity bar of the path is 1e-8 in -log10 p."""
    from flexlmm_b200 import synth
    n, M = 1500, 600
    for model in ("additive", "domgxe"):
        t = synth.task(n, M, model=model, P=5, seed=77, effect=0.1)
        r6, _ = _run(fx, t, True, M)
        fm = fx.FitModel(planes=5)
        fm.set_whitening(t["L"]); fm.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"]); fm.set_design(t["M0"], t["M1"], t["M2"])
        fm.set_perm(t["y_pred"], t["U1"], t["e_p"], t["seeds"]); fm.set_packed(t["packed"], t["n_raw"], t["pgen_order"])
        r5 = fm.run(np.arange(1, M + 1, dtype=np.int32)); st5 = fm.stats(); fm.close()
        assert st5["int8_path"] == 1
        assert np.array_equal(r5["lrt_df"], r6["lrt_df"]) and np.array_equal(r5["pivot"], r6["pivot"])
        np.testing.assert_allclose(r5["lrt_chisq"], r6["lrt_chisq"], rtol=2e-8, atol=1e-9)
        with np.errstate(divide="ignore"):
            a, b = -np.log10(r5["pval"]), -np.log10(r6["pval"])
        ok = np.isfinite(a) & np.isfinite(b)
        assert np.all(np.abs(a[ok] - b[ok]) <= 1e-8 * np.abs(b[ok]) + 1e-9)


def test_int8_path_at_larger_n(fx, orc):
    """n = 12,000 (47 row tiles of 256, 188 k-blocks): int8 route vs FP64 route vs the oracle on a handful of SNPs."""
    from flexlmm_b200 import synth
    n, M = 12000, 300
    t = synth.task(n, M, model="additive", P=0, seed=4, effect=0.05)
    r8, st8 = _run(fx, t, True, M, perm=False)
    r64, st64 = _run(fx, t, False, M, perm=False)
    assert st8["int8_path"] == 1 and st64["int8_path"] == 0
    np.testing.assert_allclose(r8["lrt_chisq"], r64["lrt_chisq"], rtol=1e-9, atol=1e-10)
    assert_logp_close(r8["pval"], r64["pval"])
    pick = np.array([0, 150, 299])
    o = oracle_run(orc, t, t["codes_raw"][pick])
    compare({k2: r8[k2][pick] for k2 in ("lrt_df", "rank", "pivot", "pval", "lrt_chisq", "coef")}, o)


@pytest.fixture(scope="module")
def c3_task():
    from flexlmm_b200 import synth
    return synth.task(20000, 64, model="domgxe", P=1, seed=6, effect=0.03)


@pytest.mark.parametrize("int8", [True, False])
def test_c3_shape_slice_against_oracle(fx, orc, c3_task, int8):
    """BASELINE configs[2] shape: n = 20,000, y ~ x + I(x == 1) + x:cov1 + cov1 vs y ~ cov1 (k = 5, s = 3): 64 SNPs, every
    per-SNP output and the min_pval of seed 1 against the oracle, on the int8 route (5 pass units) and on the FP64 route."""
    t = c3_task
    o, minp = oracle_parallel(orc, t, t["codes_raw"], seeds=(1,))
    with fx.FitModel(int8=int8) as fm:
        fm.set_whitening(t["L"]); fm.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"]); fm.set_design(t["M0"], t["M1"], t["M2"])
        fm.set_perm(t["y_pred"], t["U1"], t["e_p"], t["seeds"]); fm.set_packed(t["packed"], t["n_raw"], t["pgen_order"])
        r = fm.run(np.arange(1, 65, dtype=np.int32))
        st = fm.stats()
    assert st["int8_path"] == (1 if int8 else 0) and (not int8 or st["int8_passes"] == 5)
    compare(r, o, N=t["n"])
    assert_logp_close(r["min_pval"], [minp[1]])


def _structured_factor(n, seed):
    """A non-singular lower-triangular factor without an n^3 host Cholesky: strict lower triangle of a rank-6 product plus a
    positive diagonal (column blocks, so that the 8 n^2-byte matrix is written once)."""
    rng = np.random.default_rng(seed)
    L = np.zeros((n, n), order="F")
    U = rng.standard_normal((n, 6)) * (1.5 / np.sqrt(n)); W = rng.standard_normal((n, 6))
    for j0 in range(0, n, 2048):
        j1 = min(n, j0 + 2048)
        L[:, j0:j1] = np.tril(U @ W[j0:j1].T, k=-(j0 + 1))
    L[np.arange(n), np.arange(n)] = 0.8 + 0.4 * rng.random(n)
    return L


def test_c4_shape_n50000_against_oracle(fx, orc):
    """BASELINE configs[3] shape (n = 50,000, additive, one LOCO factor): a 20 GB factor inverted in place on the device, int8 route
    against the FP64 route on 600 SNPs and against the oracle on 4 of them, permutations from the implicit complement."""
    import scipy.linalg as sla
    from flexlmm_b200 import synth
    try:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        avail = 0
    if avail < 60 * 2 ** 30:
        pytest.skip("needs 60 GB of free host memory for the 20 GB factor and the oracle's copies")
    n, M = 50000, 600
    L = _structured_factor(n, 50)
    M0, M1, M2, names, Xn, nn, cov1 = synth.design("additive", n, seed=3)
    codes = synth.genotypes(M, n, seed=5)
    rng = np.random.default_rng(8)
    y = 0.3 * cov1 + L @ rng.standard_normal(n) + 0.05 * codes[0]
    y_mm = sla.solve_triangular(L, y, lower=True, check_finite=False)
    X_mm = np.asfortranarray(sla.solve_triangular(L, Xn, lower=True, check_finite=False))
    nf = fx.fit_null_model(X_mm, y_mm)
    packed = synth.pack_2bit(codes)
    vi = np.arange(1, M + 1, dtype=np.int32)
    res = {}
    for int8 in (True, False):
        with fx.FitModel(int8=int8) as fm:
            fm.set_whitening(L); fm.set_null(X_mm, y_mm, nf["ll"], nf["df"]); fm.set_design(M0, M1, M2)
            fm.set_perm_implicit([1, 2, 3]); fm.set_packed(packed, n)
            res[int8] = fm.run(vi)
            assert fm.stats()["int8_path"] == (1 if int8 else 0)
    np.testing.assert_allclose(res[True]["lrt_chisq"], res[False]["lrt_chisq"], rtol=2e-9, atol=1e-9)
    assert_logp_close(res[True]["pval"], res[False]["pval"])
    assert_logp_close(res[True]["min_pval"], res[False]["min_pval"])
    pick = np.array([0, 7, 311, 599])
    t = dict(L=L, y_mm=y_mm, X_null=X_mm, ll_null=nf["ll"], df_null=nf["df"], M0=M0, M1=M1, M2=M2)
    o, _ = oracle_parallel(orc, t, codes[pick], threads=4)
    compare(pick_fields(res[True], pick), o, N=n)
    compare(pick_fields(res[False], pick), o, N=n)


def test_two_factors_on_one_handle(fx, orc):
    """LOCO gives every chromosome its own Cholesky factor (lmm.nf:27-46): re-setting the whitening matrix, null model and
    permutations on a LIVE handle must give bit for bit what two fresh handles give, and both must match the oracle."""
    from flexlmm_b200 import synth
    ta = synth.task(700, 150, model="additive", P=3, seed=41)
    tb = synth.task(700, 150, model="additive", P=3, seed=42)
    tc = synth.task(450, 90, model="additive", P=3, seed=43)          # a different n on the same handle
    vi = np.arange(1, 151, dtype=np.int32)
    fresh = []
    for t in (ta, tb):
        with make(fx, t) as fm:
            fresh.append(fm.run(vi))
    fm = make(fx, ta)
    live = [fm.run(vi)]
    for t in (tb, tc):
        fm.set_whitening(t["L"]); fm.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"]); fm.set_design(t["M0"], t["M1"], t["M2"])
        fm.set_perm(t["y_pred"], t["U1"], t["e_p"], t["seeds"]); fm.set_packed(t["packed"], t["n_raw"], t["pgen_order"])
        live.append(fm.run(np.arange(1, t["M"] + 1, dtype=np.int32)))
    fm.close()
    for a, b in zip(fresh, live[:2]):
        for k2 in ("lrt_chisq", "pval", "coef", "pivot", "min_pval"):
            assert np.array_equal(a[k2], b[k2], equal_nan=True), k2
    for t, r in zip((ta, tb, tc), live):
        compare(r, oracle_run(orc, t, t["codes_raw"]))
        want = np.array([oracle_run(orc, t, t["codes_raw"], seed=int(sd))["min_pval"] for sd in t["seeds"]])
        assert_logp_close(r["min_pval"], want)
    # a stale null model (vectors of the old n) must be refused, not read past its end
    with make(fx, tc) as f2:
        f2.set_whitening(ta["L"])
        f2.set_packed(ta["packed"], ta["n_raw"], ta["pgen_order"])
        with pytest.raises(fx.FlxError):
            f2.run(vi)


@pytest.mark.parametrize("int8", [True, False])
def test_perm_minp_when_the_null_lacks_a_constant_column_of_the_real_model(fx, orc, int8):
    """y ~ x + cov1 against y ~ 1: lrt_df = 2 > s = 1 for every SNP.  The permutation maxima are bucketed by the rank the SNP
    columns add, not by lrt_df itself, so min_pval must follow the oracle (it used to come back as 2.0)."""
    from flexlmm_b200 import synth
    n = 220
    t = synth.task(n, 60, model="additive", P=4, seed=15)
    t["codes_raw"][7] = 1                                             # a monomorphic SNP: lrt_df = 1 from cov1 alone
    t["packed"] = synth.pack_2bit(t["codes_raw"])
    t["X_null"] = orc.forwardsolve(t["L"], np.ones((n, 1)))
    nm = orc.fit_null_model(t["y_mm"], t["X_null"])
    for k2 in ("ll", "df", "y_pred", "U1", "e_p"):
        t[{"ll": "ll_null", "df": "df_null"}.get(k2, k2)] = nm[k2]
    with fx.FitModel(int8=int8) as fm:
        fm.set_whitening(t["L"]); fm.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"]); fm.set_design(t["M0"], t["M1"], t["M2"])
        fm.set_perm(t["y_pred"], t["U1"], t["e_p"], t["seeds"]); fm.set_packed(t["packed"], t["n_raw"], t["pgen_order"])
        r = fm.run(np.arange(1, 61, dtype=np.int32))
    o = oracle_run(orc, t, t["codes_raw"])
    compare(r, o)
    assert sorted(set(r["lrt_df"].tolist())) == [1, 2]
    want = np.array([oracle_run(orc, t, t["codes_raw"], seed=int(sd))["min_pval"] for sd in t["seeds"]])
    assert np.all(want <= 1.0)
    assert_logp_close(r["min_pval"], want)
    # only SNPs without any SNP contribution: the bucket that is counted but never reduced
    r1 = None
    with fx.FitModel(int8=int8) as fm:
        fm.set_whitening(t["L"]); fm.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"]); fm.set_design(t["M0"], t["M1"], t["M2"])
        fm.set_perm(t["y_pred"], t["U1"], t["e_p"], t["seeds"]); fm.set_packed(t["packed"], t["n_raw"], t["pgen_order"])
        r1 = fm.run(np.array([8], dtype=np.int32))
    want1 = np.array([oracle_run(orc, t, t["codes_raw"][7:8], seed=int(sd))["min_pval"] for sd in t["seeds"]])
    assert_logp_close(r1["min_pval"], want1)


@pytest.mark.parametrize("int8", [True, False])
def test_many_covariates(fx, orc, int8):
    """k = 20 design columns (intercept, x, 18 covariates): wide constant block in K3b, 32-column u/b block in K4q pass A."""
    from flexlmm_b200 import synth
    n, M, nc = 600, 150, 18
    t = synth.task(n, M, model="additive", P=4, seed=21, effect=0.1)
    rng = np.random.default_rng(3)
    C = rng.standard_normal((n, nc))
    C[:, 5] = C[:, 4] * 2.0 - C[:, 3]          # an exactly collinear covariate: rejected by the rank rule in both models
    one = np.ones(n)
    t["M0"], t["M1"], t["M2"] = (np.column_stack([one, np.full(n, v), C]) for v in (0.0, 1.0, 2.0))
    Xn = np.column_stack([one, C])
    y = orc.forwardsolve(np.eye(n), t["L"] @ t["y_mm"])[:, 0] + C[:, 0] * 0.3
    t["y_mm"] = orc.forwardsolve(t["L"], y)[:, 0]; t["X_null"] = orc.forwardsolve(t["L"], Xn)
    nm = orc.fit_null_model(t["y_mm"], t["X_null"])
    t["ll_null"], t["df_null"] = nm["ll"], nm["df"]
    t["y_pred"], t["U1"], t["e_p"] = nm["y_pred"], nm["U1"], nm["e_p"]
    r, st = _run(fx, t, int8, M)
    assert st["int8_path"] == (1 if int8 else 0)
    o = oracle_run(orc, t, t["codes_raw"])
    compare(r, o)
    assert set(r["rank"].tolist()) == {19} and r["pivot"][0].tolist()[-1] == 8     # column 8 = the collinear covariate
    want = np.array([oracle_run(orc, t, t["codes_raw"], seed=int(s))["min_pval"] for s in t["seeds"]])
    assert_logp_close(r["min_pval"], want)


def test_reset_permutations_between_runs(fx, orc):
    """flx_set_perm after a run re-derives the setup; results follow the new seeds (and stay on the int8 route for small n)."""
    from flexlmm_b200 import synth
    t = synth.task(300, 80, model="additive", P=3, seed=31)
    fm = make(fx, t)
    vi = np.arange(1, 81, dtype=np.int32)
    r1 = fm.run(vi)
    fm.set_perm(t["y_pred"], t["U1"], t["e_p"], [7, 8])
    r2 = fm.run(vi)
    st = fm.stats()
    fm.close()
    assert st["int8_path"] == 1 and r2["min_pval"].size == 2
    want = np.array([oracle_run(orc, t, t["codes_raw"], seed=s)["min_pval"] for s in (7, 8)])
    assert_logp_close(r2["min_pval"], want)
    assert np.array_equal(r1["lrt_chisq"], r2["lrt_chisq"])


# ------------------------------------------------------------------------------------------------ DECORRELATE on the GPU (§8f rank 1)
def test_decorrelate_matches_reference_steps(fx, orc):
    """decorrelate.nf:50-77: V = s2g K + s2e I, L = t(chol(V)), forwardsolve of y and X; then the hot path runs on that factor."""
    from flexlmm_b200 import synth
    n, M = 500, 60
    rng = np.random.default_rng(2)
    Z = rng.standard_normal((n, 700)); K = Z @ Z.T / 700
    t = synth.task(n, M, model="additive", P=2, seed=6)
    y = rng.standard_normal(n); Xn = np.column_stack([np.ones(n), t["M0"][:, 2]])
    want = orc.decorrelate(K, 0.6, 0.4, y, Xn)
    with fx.FitModel() as fm:
        got = fm.decorrelate(K, 0.6, 0.4, y, Xn)
        assert got["clipped_eigs"] == 0
        np.testing.assert_allclose(got["L"], want["L"], rtol=1e-11, atol=1e-13)
        assert np.all(np.triu(got["L"], 1) == 0)
        np.testing.assert_allclose(got["y_mm"], want["y_mm"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(got["X_mm"], want["X_mm"], rtol=1e-10, atol=1e-12)
        # the handle is ready: same results as feeding L back through flx_set_whitening
        nm = orc.fit_null_model(want["y_mm"], want["X_mm"])
        fm.set_null(got["X_mm"], got["y_mm"], nm["ll"], nm["df"]); fm.set_design(t["M0"], t["M1"], t["M2"])
        fm.set_packed(t["packed"], t["n_raw"], t["pgen_order"])
        r = fm.run(np.arange(1, M + 1, dtype=np.int32))
    o = orc.fit_model(want["L"], want["y_mm"], want["X_mm"], nm["ll"], nm["df"], t["M0"], t["M1"], t["M2"], t["codes_raw"])
    compare(r, o)


def test_decorrelate_eigenvalue_clipping_fallback(fx):
    """decorrelate.nf:58-69: a GRM with slightly negative eigenvalues is forced positive definite; too negative is an error."""
    n = 120
    rng = np.random.default_rng(5)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.abs(rng.standard_normal(n)) + 0.1
    lam[:3] = [-2e-7, -5e-8, -3e-9]           # clearly below the rounding of syevd (an exact 0 comes back as +-1e-16: either side of the rule)
    K = (Q * lam) @ Q.T
    K = 0.5 * (K + K.T)
    y = rng.standard_normal(n); X = np.ones((n, 1))
    with fx.FitModel() as fm:
        got = fm.decorrelate(K, 1.0, 0.0, y, X)                 # V = K: not positive definite
        assert got["clipped_eigs"] == 3
        w, U = np.linalg.eigh(K)
        w2 = np.where(w <= 0, 1e-10, w)
        V2 = (U * np.abs(w2)) @ U.T
        np.testing.assert_allclose(got["L"] @ got["L"].T, V2, rtol=1e-7, atol=1e-12)
        K_bad = K - 1e-3 * np.outer(Q[:, 0], Q[:, 0])
        with pytest.raises(fx.FlxError) as e:
            fm.decorrelate(K_bad, 1.0, 0.0, y, X)
        assert "min_allowed_eig" in str(e.value)


# ------------------------------------------------------------------------------------------------ implicit U1 (§8f rank 2)
@pytest.mark.parametrize("model,int8", [("additive", True), ("domgxe", False)])
def test_implicit_complement_permutations(fx, orc, model, int8):
    """flx_set_perm_implicit == flx_set_perm with the (y_pred, U1, e_p) of flx_fit_null_model == the oracle's PERM tasks on them."""
    from flexlmm_b200 import synth
    n, M, P = 260, 70, 4
    t = synth.task(n, M, model=model, P=P, seed=12)
    nf = fx.fit_null_model(t["X_null"], t["y_mm"], want_U1=True)
    assert nf["df"] == t["df_null"] and nf["ll"] == pytest.approx(t["ll_null"], rel=1e-12)
    vi = np.arange(1, M + 1, dtype=np.int32)
    with fx.FitModel(int8=int8) as fm:
        fm.set_whitening(t["L"]); fm.set_null(t["X_null"], t["y_mm"], nf["ll"], nf["df"]); fm.set_design(t["M0"], t["M1"], t["M2"])
        fm.set_packed(t["packed"], t["n_raw"], t["pgen_order"])
        fm.set_perm_implicit(t["seeds"])
        r_imp = fm.run(vi)
        fm.set_perm(nf["y_pred"], nf["U1"], nf["e_p"], t["seeds"])
        r_exp = fm.run(vi)
    assert_logp_close(r_imp["min_pval"], r_exp["min_pval"])
    t2 = dict(t); t2.update(y_pred=nf["y_pred"], U1=nf["U1"], e_p=nf["e_p"], ll_null=nf["ll"], df_null=nf["df"])
    want = np.array([oracle_run(orc, t2, t["codes_raw"], seed=int(s))["min_pval"] for s in t["seeds"]])
    assert_logp_close(r_imp["min_pval"], want)


def test_pgen_device_expansion_long_difflists_and_foreign_ld_bases(fx, tmp_path):
    """K0 on records whose difflists span more than 32 groups of 64 entries (two rounds of the warp scan over the group byte
    counts), requested as a shuffled subset so that most LD records refer to a base variant that is not part of the batch."""
    from flexlmm_b200 import synth
    from pgen_writer import write_pgen
    rng = np.random.default_rng(31)
    n_raw, M = 4100, 240
    c = rng.integers(0, 3, size=(M, n_raw)).astype(np.uint8)
    for v in range(1, M):
        if v % 5:
            c[v] = c[v - 1]
            flip = rng.choice(n_raw, size=int(rng.integers(1, 200)), replace=False)
            c[v, flip] = rng.integers(0, 4, size=flip.size)         # includes missing calls (code 3)
    vt = [0] + [int(x) for x in rng.choice([1, 2, 2, 3, 3, 4, 6, 7, 0], size=M - 1)]
    path = tmp_path / "long.pgen"
    path.write_bytes(write_pgen(c, vt))
    n = 1500
    t = synth.task(n, 8, model="simple", seed=3)
    order = (rng.permutation(n_raw)[:n] + 1).astype(np.int32)
    with fx.FitModel() as fm:
        fm.set_whitening(t["L"]); fm.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"]); fm.set_design(t["M0"], t["M1"], t["M2"])
        fm.open_pgen(path, order)
        vi = (rng.permutation(M)[:150] + 1).astype(np.int32)
        _, cd = fm.decode_tile(vi, want_X=False)
        assert np.array_equal(cd, c[vi - 1][:, order - 1])
        _, cd_all = fm.decode_tile(np.arange(1, M + 1, dtype=np.int32), want_X=False)
        assert np.array_equal(cd_all, c[:, order - 1])


def test_pgen_malformed_record_is_reported(fx, tmp_path):
    """A difflist whose sample ids run past the sample count must surface as FLX_ERR_PGEN, not as silent garbage."""
    from flexlmm_b200 import synth
    from pgen_writer import write_pgen
    rng = np.random.default_rng(4)
    n, M = 300, 12
    c = rng.integers(0, 3, size=(M, n)).astype(np.uint8)
    buf = bytearray(write_pgen(c, [0] + [4] * (M - 1)))
    t = synth.task(n, 8, model="simple", seed=1)
    # the last record is a vrtype-4 difflist: blow up one of its first-sample-id bytes (2 bytes per group at n = 300)
    body = bytes(buf)
    bad = bytearray(body)
    bad[-(len(body) // (4 * M)) - 40:] = b"\xff" * len(bad[-(len(body) // (4 * M)) - 40:])
    path = tmp_path / "bad.pgen"
    path.write_bytes(bytes(bad))
    with fx.FitModel() as fm:
        fm.set_whitening(t["L"]); fm.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"]); fm.set_design(t["M0"], t["M1"], t["M2"])
        fm.open_pgen(path, None)
        with pytest.raises(fx.FlxError, match="pgen"):
            fm.decode_tile(np.arange(1, M + 1, dtype=np.int32), want_X=False)


def test_pgen_mode10_threaded_expansion(fx, tmp_path):
    """Enough variants that the compressed records of a batch are gathered by several host threads; LD chains throughout."""
    from flexlmm_b200 import synth
    from pgen_writer import write_pgen
    rng = np.random.default_rng(9)
    n, M = 97, 5000
    c = rng.integers(0, 3, size=(M, n)).astype(np.uint8)
    for v in range(1, M):                       # LD-friendly: most variants differ from the previous one in a few samples
        if v % 7:
            c[v] = c[v - 1]
            flip = rng.choice(n, size=3, replace=False)
            c[v, flip] = rng.integers(0, 3, size=3)
    vt = [0] + [int(x) for x in rng.choice([0, 1, 2, 2, 2, 3, 4, 6], size=M - 1)]
    path = tmp_path / "big.pgen"
    path.write_bytes(write_pgen(c, vt))
    t = synth.task(n, 8, model="simple", seed=1)
    with fx.FitModel() as fm:
        fm.set_whitening(t["L"]); fm.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"]); fm.set_design(t["M0"], t["M1"], t["M2"])
        fm.open_pgen(path, None)
        info = fm.pgen_info()
        assert info == dict(variant_ct=M, sample_ct=n, has_dosage=False)
        vi = np.arange(1, M + 1, dtype=np.int32)
        _, cd = fm.decode_tile(vi, want_X=False)
        assert np.array_equal(cd, c)
        r = fm.run(vi)
        assert r["nvars_tested"] == M and np.isfinite(r["lrt_chisq"]).all()


def test_fit_model_from_process_files(fx, orc, tmp_path):
    """The process interface on files: .pgen (compressed records) + .pvar + .psam in a shuffled sample order, a var_idx
    subset, ORIG TSV and PERM TSV out; numbers against the oracle, text against the R formatting rules."""
    import gzip
    from flexlmm_b200 import synth
    from pgen_writer import write_pgen
    n, n_raw, Mtot = 150, 220, 90
    t = synth.task(n, Mtot, model="domgxe", n_raw=n_raw, P=3, seed=19, effect=0.2)
    rng = np.random.default_rng(6)
    raw_ids = [f"id{q:04d}" for q in rng.permutation(n_raw)]
    samples = [raw_ids[i - 1] for i in t["pgen_order"]]
    (tmp_path / "g.pgen").write_bytes(write_pgen(t["codes_raw"], [0] + [int(x) for x in rng.choice([0, 1, 2, 4, 6], size=Mtot - 1)]))
    (tmp_path / "g.psam").write_text("#FID\tIID\tSEX\n" + "".join(f"fam\t{s}\tNA\n" for s in raw_ids))
    (tmp_path / "g.pvar").write_text("##source=test\n#CHROM\tPOS\tID\tREF\tALT\n" + "".join(f"7\t{1000 + 10 * v}\tsnp{v}\tA\tC\n" for v in range(Mtot)))
    vi = np.arange(5, 65, dtype=np.int32)
    res = fx.fit_model_files(str(tmp_path / "out"), tmp_path / "g.pgen", tmp_path / "g.pvar", tmp_path / "g.psam", samples, vi, t["names"],
                             t["L"], t["y_mm"], t["X_null"], t["ll_null"], t["df_null"], t["M0"], t["M1"], t["M2"],
                             y_pred=t["y_pred"], U1=t["U1"], e_p=t["e_p"], seeds=t["seeds"])
    codes = t["codes_raw"][vi - 1][:, t["pgen_order"] - 1]
    compare(res, oracle_run(orc, t, codes))
    lines = gzip.open(tmp_path / "out.tsv.gwas.gz", "rt").read().splitlines()
    assert lines[0] == "chr\tpos\tid\tref\talt\tlrt_chisq\tlrt_df\tpval\tbeta" and len(lines) == 1 + vi.size
    f = lines[1].split("\t")
    assert f[:5] == ["7", "1040", "snp4", "A", "C"] and float(f[5]) == pytest.approx(res["lrt_chisq"][0], rel=1e-14)
    assert f[8].startswith("(Intercept)~") and f[8].count(",") == len(t["names"]) - 1
    perm = gzip.open(tmp_path / "out.perm.tsv.gz", "rt").read().splitlines()
    assert len(perm) == 3 and perm[0].split("\t")[:2] == ["1", "7"] and perm[0].split("\t")[3] == str(vi.size)
    want = np.array([oracle_run(orc, t, codes, seed=int(sd))["min_pval"] for sd in t["seeds"]])
    assert_logp_close(res["min_pval"], want)
