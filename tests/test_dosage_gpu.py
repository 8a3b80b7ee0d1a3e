"""use_dosage = true (the reference's default, nextflow.config:59): pgenlibr::Read semantics through the C ABI — K0's dosage-track
decoder, K1's evaluation of the design on real-valued x, the FP64 route — against the oracle's restatement of Read and of the
literal loop on dosage vectors (oracle.pgen_read, oracle.fit_model_x)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_parity_gpu import assert_logp_close  # noqa: E402


def dosage_file(tmp_path, t, rng, frac=0.3, missing_under_dosage=True):
    """A mode-0x10 .pgen with every main-track vrtype and every dosage-track mode (none / list / unconditional / bitarray)."""
    import pgen_writer as pw
    codes = t["codes_raw"].copy()
    M, N = codes.shape
    dos = np.full((M, N), 65535, dtype=np.uint16)
    m = rng.random((M, N)) < frac
    dos[m] = rng.integers(0, 32769, size=int(m.sum()))
    exact = rng.random((M, N)) < 0.05                    # some dosages exactly on a hard-call value: I(x == 1) must fire for 16384
    dos[m & exact] = (16384 * rng.integers(0, 3, size=int((m & exact).sum()))).astype(np.uint16)
    if missing_under_dosage:
        hide = m & (rng.random((M, N)) < 0.2)
        codes[hide] = 3                                   # a missing hard call UNDER a dosage is fine: Read returns the dosage
    dm = [(v % 4) for v in range(M)]
    for v in range(M):
        if dm[v] == 0:
            dos[v] = 65535
            codes[v] = t["codes_raw"][v]
    vr = [[0, 1, 2, 3, 4, 6, 7][v % 7] for v in range(M)]
    path = tmp_path / "dos.pgen"
    data = pw.write_pgen(codes, vr, dosages=dos, dosage_modes=dm)
    path.write_bytes(data)
    return str(path), data, codes, dos, dm


def setup(fx, t, path, probes=True, use=True, int8=True):
    from flexlmm_b200 import synth
    fm = fx.FitModel(int8=int8)
    fm.set_whitening(t["L"]); fm.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"]); fm.set_design(t["M0"], t["M1"], t["M2"])
    if probes:
        n = t["n"]
        fm.set_design_probes(synth.design_at(t["model"], np.full(n, 0.5), t["cov1"]), synth.design_at(t["model"], np.full(n, 1.5), t["cov1"]))
    if t.get("P", 0):
        fm.set_perm(t["y_pred"], t["U1"], t["e_p"], t["seeds"])
    fm.set_use_dosage(use)
    fm.open_pgen(path, t["pgen_order"])
    return fm


def task(n, M, model, P, seed, n_raw=None):
    from flexlmm_b200 import synth
    t = synth.task(n, M, model=model, P=P, seed=seed, n_raw=n_raw)
    t["cov1"] = synth.design(model, n, seed=11 + seed)[6]
    return t


@pytest.mark.parametrize("n,n_raw", [(60, 60), (255, 300), (256, 256), (700, 1000)])
def test_read_dosage_bit_exact(fx, orc, tmp_path, n, n_raw):
    """K0 dosage tracks + K1: the x column equals pgenlibr::Read's buffer gathered by pgen_order, bit for bit."""
    t = task(n, 56, "additive", 0, 5, n_raw=n_raw)
    path, data, codes, dos, dm = dosage_file(tmp_path, t, np.random.default_rng(n))
    assert fx.FitModel  # noqa
    with setup(fx, t, path) as fm:
        assert fm.pgen_info()["has_dosage"]
        X, _ = fm.decode_tile(np.arange(1, 57, dtype=np.int32), want_codes=False)
    for v in range(56):
        want = orc.pgen_read(data, v)[t["pgen_order"] - 1]
        got = X[:, v]
        ok = ~np.isnan(want)
        assert np.array_equal(got[ok], want[ok]), v
    with setup(fx, t, path, use=False) as fm:             # use_dosage = false: ReadHardcalls on the same file
        X, c = fm.decode_tile(np.arange(1, 57, dtype=np.int32))
    assert np.array_equal(c, codes[:, t["pgen_order"] - 1])


@pytest.mark.parametrize("model", ["additive", "domgxe"])
def test_fit_on_dosages_matches_oracle(fx, orc, tmp_path, model):
    from flexlmm_b200 import synth
    t = task(150, 48, model, 3, 8, n_raw=171)
    path, data, codes, dos, dm = dosage_file(tmp_path, t, np.random.default_rng(3))
    xs = np.stack([orc.pgen_read(data, v)[t["pgen_order"] - 1] for v in range(48)])
    # variants whose Read buffer still holds an NA (missing call, no dosage) abort the reference: test the others
    good = np.where(~np.isnan(xs).any(axis=1))[0]
    assert good.size >= 30
    vi = (good + 1).astype(np.int32)
    with setup(fx, t, path) as fm:
        r = fm.run(vi)
        st = fm.stats()
    assert st["int8_path"] == 0 and st["fallback_reason"] == 6
    des = lambda x: synth.design_at(model, x, t["cov1"])      # noqa: E731
    o = orc.fit_model_x(t["L"], t["y_mm"], t["X_null"], t["ll_null"], t["df_null"], des, xs[good])
    assert np.array_equal(r["lrt_df"], o["df"]) and np.array_equal(r["rank"], o["rank"]) and np.array_equal(r["pivot"], o["pivot"])
    np.testing.assert_allclose(r["lrt_chisq"], o["chisq"], rtol=1e-8, atol=1e-9)
    assert_logp_close(r["pval"], o["pval"])
    np.testing.assert_allclose(r["coef"], o["coef"], rtol=1e-8, atol=1e-10)
    want = [orc.fit_model_x(t["L"], t["y_mm"], t["X_null"], t["ll_null"], t["df_null"], des, xs[good], perm_seed=int(sd), y_pred=t["y_pred"],
                            U1=t["U1"], e_p=t["e_p"])["min_pval"] for sd in t["seeds"]]
    assert_logp_close(r["min_pval"], want)
    if model == "domgxe":
        assert (xs[good] == 1.0).any() and ((xs[good] > 0.9) & (xs[good] < 1.1) & (xs[good] != 1.0)).any()    # I(x == 1) had both cases to tell apart


def test_missing_call_without_dosage_is_fatal_with_dosage_is_not(fx, orc, tmp_path):
    t = task(80, 8, "additive", 0, 9)
    import pgen_writer as pw
    codes = t["codes_raw"].copy()
    dos = np.full(codes.shape, 65535, dtype=np.uint16)
    codes[2, 5] = 3; dos[2, 5] = 20000          # missing call under a dosage: fine
    codes[6, 7] = 3                              # missing call, no dosage: NA -> the reference aborts
    dos[:, 0] = 100
    path = tmp_path / "m.pgen"
    path.write_bytes(pw.write_pgen(codes, [0] * 8, dosages=dos, dosage_modes=[1] * 8))
    with setup(fx, t, str(path)) as fm:
        r = fm.run(np.array([1, 2, 3, 4, 5, 6], dtype=np.int32))
        assert np.isfinite(r["lrt_chisq"]).all()
        with pytest.raises(fx.FlxError) as e:
            fm.run(np.arange(1, 9, dtype=np.int32))
        assert e.value.code == -2


def test_use_dosage_needs_probes_and_refuses_what_it_cannot_evaluate(fx, tmp_path):
    from flexlmm_b200 import synth
    t = task(90, 12, "additive", 0, 10)
    path, *_ = dosage_file(tmp_path, t, np.random.default_rng(1), missing_under_dosage=False)
    with setup(fx, t, path, probes=False) as fm:
        with pytest.raises(fx.FlxError) as e:
            fm.run(np.arange(1, 13, dtype=np.int32))
        assert e.value.code == -7 and "flx_set_design_probes" in str(e.value)
    # y ~ I(x^2) + cov1: neither affine in x nor an indicator
    n = t["n"]
    sq = lambda x: np.column_stack([np.ones(n), x * x, t["cov1"]])      # noqa: E731
    with fx.FitModel() as fm:
        fm.set_whitening(t["L"]); fm.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"])
        fm.set_design(sq(np.zeros(n)), sq(np.ones(n)), sq(np.full(n, 2.0)))
        fm.set_design_probes(sq(np.full(n, 0.5)), sq(np.full(n, 1.5)))
        fm.set_use_dosage(True); fm.open_pgen(path, t["pgen_order"])
        with pytest.raises(fx.FlxError) as e:
            fm.run(np.arange(1, 13, dtype=np.int32))
        assert e.value.code == -1 and "neither affine" in str(e.value)
        fm.set_use_dosage(False)                                        # hard calls: any function of x is a lookup
        assert np.isfinite(fm.run(np.arange(1, 13, dtype=np.int32))["lrt_chisq"]).all()


def test_use_dosage_on_a_file_without_dosages_is_the_hard_call_path(fx, tmp_path):
    import pgen_writer as pw
    t = task(300, 64, "additive", 2, 11)
    path = tmp_path / "h.pgen"
    path.write_bytes(pw.write_pgen(t["codes_raw"], [[0, 1, 4][v % 3] for v in range(64)]))
    vi = np.arange(1, 65, dtype=np.int32)
    with setup(fx, t, str(path), use=True) as fm:
        a = fm.run(vi); sa = fm.stats()
    with setup(fx, t, str(path), use=False, probes=False) as fm:
        b = fm.run(vi)
    assert sa["int8_path"] == 1 and sa["fallback_reason"] == 0
    for f in ("lrt_chisq", "pval", "coef", "min_pval"):
        assert np.array_equal(a[f], b[f], equal_nan=True)
