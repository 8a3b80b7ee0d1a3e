"""The R pin: compare the CPU oracle (CPU test) and the CUDA path (gpu test) with the outputs of the reference's own R
statements (tests/golden/make_r_fixture.R runs fit_model.nf:96-143 literally on the committed input bundles).

R is not installed in the build image, so the *.out.txt files may be absent: the comparisons then SKIP with that reason
(parity stays "unpinned"); the bundle round trip and the oracle-vs-GPU agreement on the same bundles are always checked.
A maintainer pins parity with:   Rscript tests/golden/make_r_fixture.R && python -m pytest tests/test_r_pin.py
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import r_pin  # noqa: E402

CASES = r_pin.case_paths()
NEED_R = "no R output for this bundle: run `Rscript tests/golden/make_r_fixture.R` (base R only) to pin parity"
RTOL, ATOL = 1e-8, 1e-10


def logp_close(a, b):
    a, b = np.ravel(np.asarray(a, dtype=float)), np.ravel(np.asarray(b, dtype=float))
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(b)
    with np.errstate(divide="ignore"):
        la, lb = -np.log10(a[ok]), -np.log10(b[ok])
    fin = np.isfinite(lb)
    assert np.array_equal(np.isfinite(la), fin)
    assert np.all(np.abs(la[fin] - lb[fin]) <= RTOL * np.abs(lb[fin]) + ATOL)


def oracle_outputs(orc, b):
    geno = np.ascontiguousarray(b["geno"].T.astype(np.uint8))
    args = (b["L"], b["y_mm"][:, 0], b["X_null"], float(b["ll_null"][0, 0]), int(b["df_null"][0, 0]), b["M0"], b["M1"], b["M2"], geno)
    o = orc.fit_model(*args)
    perm = {int(sd): orc.fit_model(*args, perm_seed=int(sd), y_pred=b["y_pred"][:, 0], U1=b["U1"], e_p=b["e_p"][:, 0]) for sd in b["seeds"][:, 0]}
    return o, perm


def check_against_r(r, chisq, df, pval, rank, pivot, coef, minp):
    """r: parsed R output.  pivot / coef given SNP-major (M x k); R wrote k x M."""
    assert np.array_equal(np.asarray(df), r["lrt_df"][:, 0].astype(int))
    assert np.array_equal(np.asarray(rank), r["rank"][:, 0].astype(int))
    assert np.array_equal(np.asarray(pivot), r["pivot"].T.astype(int))
    np.testing.assert_allclose(chisq, r["lrt_chisq"][:, 0], rtol=1e-8, atol=1e-9)
    logp_close(pval, r["pval"][:, 0])
    np.testing.assert_allclose(coef, r["coef"].T, rtol=1e-8, atol=1e-10)
    for sd, v in minp.items():
        logp_close([v], r[f"perm{sd}.min_pval"])


def test_bundles_exist_and_round_trip(tmp_path):
    assert len(CASES) >= 7, "tests/golden/r_pin/*.in.txt are committed fixtures (tests/golden/make_golden.py writes them)"
    b = r_pin.read_bundle(CASES[0][1])
    r_pin.write_bundle(tmp_path / "x.in.txt", b)
    b2 = r_pin.read_bundle(tmp_path / "x.in.txt")
    assert b.keys() == b2.keys()
    for k in b:
        assert (b[k] == b2[k]) if isinstance(b[k], str) else np.array_equal(b[k], b2[k], equal_nan=True)
    src = open(os.path.join(r_pin.HERE, "golden", "make_r_fixture.R")).read()
    for stmt in ("set.seed(perm_seed)", "sample(e.p)", ".lm.fit(x = X.mm, y = y.mm)", "stats:::logLik.lm(fit)", "forwardsolve(L, X)",
                 "model.matrix(model[-2], data = model_frame)", "pchisq(lrt_chisq, df = lrt_df, lower.tail = FALSE)", "min(pval, min_pval, na.rm = TRUE)"):
        assert stmt in src, f"make_r_fixture.R must run the reference's own statement `{stmt}` (fit_model.nf:96-143)"


@pytest.mark.parametrize("name,inp,out", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_r(orc, name, inp, out):
    if out is None:
        pytest.skip(NEED_R)
    b, r = r_pin.read_bundle(inp), r_pin.read_bundle(out)
    o, perm = oracle_outputs(orc, b)
    check_against_r(r, o["chisq"], o["df"], o["pval"], o["rank"], o["pivot"], o["coef"], {sd: p["min_pval"] for sd, p in perm.items()})
    np.testing.assert_allclose(float(b["ll_null"][0, 0]), r["ll_null_r"][0, 0], rtol=1e-12)
    for sd in perm:
        assert np.array_equal(orc.r_sample(sd, b["e_p"].shape[0]), r[f"perm{sd}.sample_idx"][:, 0].astype(np.int32))      # bit-exact index order
        logp_close(perm[sd]["pval"], r[f"perm{sd}.pval"][:, 0])


def test_oracle_kat_matches_r(orc):
    path = os.path.join(r_pin.PIN_DIR, "kat.out.txt")
    if not os.path.exists(path):
        pytest.skip(NEED_R)
    r = r_pin.read_bundle(path)
    for sd in (1, 42, 123):
        assert np.array_equal(orc.r_sample(sd, 10), r[f"sample10.seed{sd}"][:, 0].astype(np.int32))
    for m in (9, 298, 4998, 70001):
        assert np.array_equal(orc.r_sample(7, m), r[f"sample_int{m}.seed7"][:, 0].astype(np.int32))
    np.testing.assert_allclose(orc.r_runif(1, 3), r["runif3.seed1"][:, 0], rtol=0, atol=1e-15)
    q = r["pchisq_q"][:, 0]
    for d in range(1, 7):
        logp_close([orc.pchisq_upper(x, d) for x in q], r[f"pchisq_upper.df{d}"][:, 0])
    np.testing.assert_allclose(orc.quantile7([0.3, 0.01, 0.2, 0.05, 0.5, 0.04, 0.09], 0.05), r["quantile7"][0, 0], rtol=1e-14)


def gpu_outputs(fx, b, int8):
    geno = b["geno"].T.astype(np.uint8)
    from flexlmm_b200 import synth
    seeds = b["seeds"][:, 0].astype(np.int32)
    with fx.FitModel(int8=int8) as fm:
        fm.set_whitening(b["L"]); fm.set_null(b["X_null"], b["y_mm"][:, 0], float(b["ll_null"][0, 0]), int(b["df_null"][0, 0]))
        fm.set_design(b["M0"], b["M1"], b["M2"]); fm.set_perm(b["y_pred"][:, 0], b["U1"], b["e_p"][:, 0], seeds)
        fm.set_packed(synth.pack_2bit(geno), geno.shape[1])
        return fm.run(np.arange(1, geno.shape[0] + 1, dtype=np.int32)), seeds


@pytest.mark.gpu
@pytest.mark.parametrize("int8", [True, False])
@pytest.mark.parametrize("name,inp,out", CASES, ids=[c[0] for c in CASES])
def test_gpu_matches_r_or_oracle(fx, orc, name, inp, out, int8):
    """Through the C ABI on the bundles R runs on.  With R outputs present: GPU vs R.  Always: GPU vs oracle on the same bundle."""
    b = r_pin.read_bundle(inp)
    g, seeds = gpu_outputs(fx, b, int8)
    o, perm = oracle_outputs(orc, b)
    ref = dict(lrt_chisq=o["chisq"].reshape(-1, 1), lrt_df=o["df"].reshape(-1, 1), pval=o["pval"].reshape(-1, 1), rank=o["rank"].reshape(-1, 1),
               pivot=o["pivot"].T, coef=o["coef"].T, **{f"perm{sd}.min_pval": np.array([[p["min_pval"]]]) for sd, p in perm.items()})
    for r in ([ref] if out is None else [ref, r_pin.read_bundle(out)]):
        check_against_r(r, g["lrt_chisq"], g["lrt_df"], g["pval"], g["rank"], g["pivot"], g["coef"], {int(sd): g["min_pval"][i] for i, sd in enumerate(seeds)})
    for sd in seeds:
        assert np.array_equal(fx.perm_indices(int(sd), b["e_p"].shape[0]), orc.r_sample(int(sd), b["e_p"].shape[0]))
