"""Host-side logic and the C-ABI boundary, without a GPU: the library loads, exports every symbol the header declares,
its host-only entry points (R-exact permutation order, type-7 quantile) are bit-exact against the oracle, and the product
path fails loudly when no B200 is present (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "flexlmm_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(flx_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(fx):
    L = ctypes.CDLL(fx.lib_path())
    names = header_functions()
    assert len(names) >= 18
    for nm in names:
        assert hasattr(L, nm), f"{nm} declared in include/flexlmm_cuda.h but not exported"
    assert sorted(fx._capi.EXPORTS) == names


def test_ctypes_structs_mirror_the_header():
    """The Python mirror of the C ABI structs must list the header's fields in the header's order (a drifted flx_stats would
    silently shift every counter bench.py reads)."""
    import re
    from flexlmm_b200 import _capi
    hdr = open(os.path.join(ROOT, "include", "flexlmm_cuda.h")).read()

    def fields(struct_name):
        end = re.search(r"\}\s*" + struct_name + r"\s*;", hdr).start()
        body = hdr[hdr.rindex("typedef struct", 0, end):end].split("{", 1)[1]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                out.append(re.sub(r"[\*\s]", " ", decl).split()[-1])
        return out

    assert fields("flx_config") == [f for f, _ in _capi.FlxConfig._fields_]
    assert fields("flx_snp_out") == [f for f, _ in _capi.FlxSnpOut._fields_]
    assert fields("flx_stats") == [f for f, _ in _capi.FlxStats._fields_]


def test_version_and_error_string(fx):
    L = fx.load()
    assert b"sm_100a" in L.flx_version()
    assert isinstance(L.flx_last_error(), bytes)


def test_no_gpu_means_loud_failure(fx):
    if fx.load().flx_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(fx.FlxError) as e:
        fx.FitModel()
    assert e.value.code == -8 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "flexlmm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".c")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f == "synth.py", f"{f} mentions the oracle"
    assert "oracle" not in open(os.path.join(ROOT, "include", "flexlmm_cuda.h")).read().lower()


@pytest.mark.parametrize("m", [1, 2, 9, 10, 298, 4998, 70001])
def test_perm_indices_bit_exact_vs_oracle(fx, orc, m):
    for seed in (1, 2, 42, 100, 2**31 - 1):
        assert np.array_equal(fx.perm_indices(seed, m), orc.r_sample(seed, m))


def test_perm_indices_known_answers(fx):
    import json
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "kat.json")))
    for seed, want in kat["sample10"].items():
        assert fx.perm_indices(int(seed), 10).tolist() == want


def test_minp_threshold_type7(fx, orc):
    rng = np.random.default_rng(0)
    for P in (1, 2, 10, 100, 1000):
        x = rng.random(P)
        for p in (0.05, 0.5, 0.0, 1.0, 0.013):
            assert fx.minp_threshold(x, p) == pytest.approx(orc.quantile7(x, p), abs=1e-16)
            assert fx.minp_threshold(x, p) == pytest.approx(np.quantile(x, p), abs=1e-15)


def test_r_number_formatting(fx):
    from flexlmm_b200.fit_model import _r_num
    assert _r_num(9.98853558118183) == "9.98853558118183"
    assert _r_num(0.006776681232578928) == "0.00677668123257893"
    assert _r_num(2.0) == "2"
    assert _r_num(float("nan")) == "NA"
    assert _r_num(1e-20) == "1e-20"
    assert _r_num(1.5e-7) == "1.5e-07"


def test_tsv_writers(fx, tmp_path):
    import gzip
    res = dict(coef=np.array([[1.0, -0.5]]), lrt_chisq=np.array([3.25]), lrt_df=np.array([1]), pval=np.array([0.0714]))
    p = tmp_path / "a.tsv.gwas.gz"
    fx.write_gwas_tsv(p, [("14", 100, "rs1", "A", "G")], ["(Intercept)", "x"], res)
    lines = gzip.open(p, "rt").read().splitlines()
    assert lines[0].split("\t") == ["chr", "pos", "id", "ref", "alt", "lrt_chisq", "lrt_df", "pval", "beta"]
    assert lines[1] == "14\t100\trs1\tA\tG\t3.25\t1\t0.0714\t(Intercept)~1,x~-0.5"
    q = tmp_path / "a.perm.tsv.gz"
    fx.write_perm_tsv(q, [1, 2], "14", [0.5, 0.25], 11)
    assert gzip.open(q, "rt").read() == "1\t14\t0.5\t11\n2\t14\t0.25\t11\n"
    with pytest.raises(ValueError):
        fx.write_perm_tsv(q, [1], "14", [2.0], 11)       # fit_model.nf:177 stopifnot


def test_synth_packing_roundtrip():
    from flexlmm_b200 import synth
    g = synth.genotypes(7, 13, seed=1, missing_rate=0.1)
    p = synth.pack_2bit(g)
    assert p.shape == (7, 4)
    un = np.stack([(p[:, i >> 2] >> (2 * (i & 3))) & 3 for i in range(13)], axis=1)
    assert np.array_equal(un, g)


def test_shard_bounds():
    from flexlmm_b200.sharding import shard_bounds, shard_var_idx
    for M in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(M, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == M
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_var_idx(np.arange(1, 11), 1, 3).tolist() == [5, 6, 7]


def test_loco_chromosome_sharding():
    from flexlmm_b200.loco import chromosomes_of_rank
    for world in (1, 2, 4, 8):
        parts = [chromosomes_of_rank(22, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(22))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_golden_synth_matches_oracle_now(orc):
    """The committed golden vectors are reproducible from the committed generator."""
    from flexlmm_b200 import synth
    G = np.load(os.path.join(ROOT, "tests", "golden", "synth_small.npz"))
    n, M, P, n_raw, seed = G["additive.args"].tolist()
    t = synth.task(n, M, model="additive", P=P, n_raw=None if n_raw == n else n_raw, seed=seed)
    codes = t["codes_raw"][:, t["pgen_order"] - 1]
    o = orc.fit_model(t["L"], t["y_mm"], t["X_null"], t["ll_null"], t["df_null"], t["M0"], t["M1"], t["M2"], codes)
    np.testing.assert_allclose(o["chisq"], G["additive.chisq"], rtol=1e-12)
    assert np.array_equal(o["pivot"], G["additive.pivot"])


def test_r_shim_compiles_against_stub():
    """r_shim.c (the .Call binding R uses) must at least parse and type-check against the C ABI header; R itself is not
    installed in this image, so R's API is declared by a stub header."""
    import subprocess
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Wno-cast-function-type", "-Werror",
                        "-I" + os.path.join(ROOT, "flexlmm_b200", "r", "stub"), "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "flexlmm_b200", "r", "r_shim.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_shim_and_module_cover_every_abi_step():
    shim = open(os.path.join(ROOT, "flexlmm_b200", "r", "r_shim.c")).read()
    # one entry point per reference step, in the several-GPUs-from-one-caller form (ndev = 1 is the plain handle)
    for fn in ("flx_multi_create", "flx_multi_set_whitening", "flx_multi_set_null", "flx_multi_set_design", "flx_multi_set_design_probes",
               "flx_multi_set_use_dosage", "flx_multi_set_perm", "flx_multi_open_pgen", "flx_multi_run", "flx_multi_destroy", "flx_last_error",
               "flx_perm_indices"):
        assert fn in shim
    nf = open(os.path.join(ROOT, "flexlmm_b200", "r", "fit_model.nf")).read()
    # same process interface as the reference (fit_model.nf:7-16)
    for line in ("tuple val(meta), path(mm), path(var_idx), path(null_model_fit), val(perm_seed)",
                 "tuple val(meta2), path(pgen), path(pvar), path(psam)", "path model", "tuple val(meta3), path(model_frame)", "val use_dosage",
                 "path \"versions.yml\" , emit: versions", "design_at(0.5)", "design_at(1.5)", "FLEXLMM_DEVICES", "chr\\\\tpos\\\\tid\\\\tref\\\\talt\\\\tlrt_chisq\\\\tlrt_df\\\\tpval\\\\tbeta"):
        assert line in nf, line


def test_cat_perm_and_summarise_by_chr(fx, orc, tmp_path):
    """get_null_p_dist.nf: concatenate per-(seed, chr) outputs, min over chromosomes, type-7 threshold (make_plots.nf:31-32)."""
    import gzip
    rng = np.random.default_rng(0)
    seeds = [1, 2, 3, 4]
    paths, rows = [], []
    for chrom in ("14", "15"):
        mp = rng.random(4) * 0.2
        p = tmp_path / f"{chrom}.perm.tsv.gz"
        fx.write_perm_tsv(p, seeds, chrom, mp, 11)
        paths.append(p)
        rows += [(s, chrom, float(f"{m:.15g}"), 11) for s, m in zip(seeds, mp)]
    cat = tmp_path / "all.tsv.gz"
    fx.cat_perm(paths, cat)
    got = fx.read_perm_table(cat)
    assert [(a, b, d) for a, b, _, d in got] == [(a, b, d) for a, b, _, d in rows]
    table = fx.summarise_by_chr(got, 4)
    want = orc.summarise_minp(got)
    assert [(t[0], t[1], t[2]) for t in table] == [(k, want[k][0], want[k][1]) for k in sorted(want)]
    assert all(t[2] == 22 for t in table)
    with pytest.raises(ValueError):
        fx.summarise_by_chr(got, 5)
    out = tmp_path / "gw.tsv.gz"
    fx.write_genome_wide_min_p(out, table)
    assert gzip.open(out, "rt").readline() == '"permutation"\t"min_pval"\t"nvars"\n'
    th = fx.thresholds(table, n_tests=22, p_thr=0.05)
    assert th["bonferroni"] == 0.05 / 22
    assert th["perm_thr"] == pytest.approx(orc.quantile7([t[1] for t in table], 0.05), abs=1e-16)


def test_fit_null_model_implicit_basis(fx, orc):
    """FIT_NULL_MODEL without the n x n eigen: lm.fit quantities equal the oracle's; the implicit complement basis is an
    orthonormal basis of the orthogonal complement of the null design (what fit_null_model.nf:52-65 needs of U1)."""
    rng = np.random.default_rng(0)
    n = 150
    X = np.column_stack([np.ones(n), rng.standard_normal(n), np.zeros(n), rng.standard_normal(n)])
    X[:, 3] = 2 * X[:, 1] - 1                                   # rank 2 of 4
    X = np.column_stack([X, rng.standard_normal(n)])           # rank 3 of 5
    y = rng.standard_normal(n)
    r = fx.fit_null_model(X, y, want_U1=True)
    f = orc.lm_fit(X, y)
    assert r["df"] == f["rank"] + 1 == 4 and r["m"] == n - 3
    assert r["ll"] == pytest.approx(orc.loglik_lm(f["residuals"]), rel=1e-13)
    np.testing.assert_allclose(r["resid"], f["residuals"], atol=1e-13)
    np.testing.assert_allclose(r["y_pred"], y - f["residuals"], atol=1e-13)
    U1 = r["U1"]
    np.testing.assert_allclose(U1.T @ U1, np.eye(r["m"]), atol=1e-13)
    np.testing.assert_allclose(U1.T @ X, 0, atol=1e-13)
    np.testing.assert_allclose(U1.T @ f["residuals"], r["e_p"], atol=1e-13)
    np.testing.assert_allclose(U1 @ r["e_p"], f["residuals"], atol=1e-13)     # e = U1 e.p: the decorrelation is lossless


def test_pvar_psam_readers(fx, tmp_path):
    """The text inputs next to the .pgen (fit_model.nf:34-52,70): header conventions, var_idx selection, match()."""
    import gzip
    pv = tmp_path / "a.pvar"
    pv.write_text("##fileformat=VCFv4.2\n##contig=<ID=1>\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\n"
                  "1\t100\trs1\tA\tG\t.\n1\t250\trs2\tC\tT\t.\nX\t7\trs3\tG\tGA\t.\n")
    assert fx.read_pvar(pv) == [("1", 100, "rs1", "A", "G"), ("1", 250, "rs2", "C", "T"), ("X", 7, "rs3", "G", "GA")]
    assert fx.read_pvar(pv, var_idx=[3, 1]) == [("X", 7, "rs3", "G", "GA"), ("1", 100, "rs1", "A", "G")]
    bim = tmp_path / "b.pvar.gz"
    with gzip.open(bim, "wt") as f:
        f.write("2\tv1\t0\t55\tT\tC\n")
    assert fx.read_pvar(bim) == [("2", 55, "v1", "C", "T")]
    ps = tmp_path / "a.psam"
    ps.write_text("#FID\tIID\tSEX\nf1\ts1\t1\nf1\ts2\t2\nf2\ts3\tNA\n")
    assert fx.read_psam(ps) == ["s1", "s2", "s3"]
    ps2 = tmp_path / "b.psam"
    ps2.write_text("#IID\tSEX\ns9\t1\ns8\t2\n")
    assert fx.read_psam(ps2) == ["s9", "s8"]
    fam = tmp_path / "c.psam"
    fam.write_text("f1 s1 0 0 1 -9\nf1 s2 0 0 2 -9\n")
    assert fx.read_psam(fam) == ["s1", "s2"]
    assert fx.pgen_order(["s3", "s1"], ["s1", "s2", "s3", "s1"]).tolist() == [3, 1]
    with pytest.raises(ValueError, match="not in the .psam"):
        fx.pgen_order(["zz"], ["s1"])
