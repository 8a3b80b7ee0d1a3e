"""Generates the committed golden fixtures.  Run HERE (needs /root/reference for the test assets):

    python tests/golden/make_golden.py

c1_test_asset.json : BASELINE config C1 — the reference's own -profile test data (assets/test_data: 12 samples x 22 SNPs),
                     both formula pairs (conf/test.config:35-36 and BASELINE's y ~ x vs y ~ 1), L = I and a GRM-based L,
                     inputs + oracle outputs (ORIG and PERM seeds 1,2).  The reference has no golden outputs of its own
                     (parity unpinned); the values are the oracle's, which is pinned on SURVEY.md Appendix B.
synth_small.npz    : small synthetic tasks (seeded) with oracle outputs, so the GPU tests can also be checked against
                     numbers computed in this container.
kat.json           : SURVEY.md Appendix B known answers.
r_pin/*.in.txt     : input bundles for tests/golden/make_r_fixture.R, the script a maintainer with R runs to PIN the oracle and the
                     CUDA path on R's own outputs (tests/test_r_pin.py picks the resulting r_pin/*.out.txt up when present).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from flexlmm_b200 import synth  # noqa: E402

REF = "/root/reference/assets/test_data"


def read_tsv(path):
    rows = [l.split() for l in open(path).read().strip().splitlines()]
    return rows[0], rows[1:]


def c1():
    pgen = open(f"{REF}/pgen/in.pgen", "rb").read()
    psam = [l.split("\t")[0] for l in open(f"{REF}/pgen/in.psam").read().strip().splitlines()[1:]]
    pvar = [l.split("\t") for l in open(f"{REF}/pgen/in.pvar").read().strip().splitlines() if not l.startswith("##")]
    hdr, pvar = pvar[0], pvar[1:]
    _, ph = read_tsv(f"{REF}/tsv/pheno.tsv"); _, cv = read_tsv(f"{REF}/tsv/covar.tsv"); _, qc = read_tsv(f"{REF}/tsv/qcovar.tsv")
    samples = [r[0] for r in ph]
    assert samples == psam == [r[0] for r in cv] == [r[0] for r in qc]
    pheno = np.array([float(r[1]) for r in ph])
    y = (pheno - pheno.mean()) / pheno.std(ddof=1)       # standardise = true (conf/test.config:32; plink2 --variance-standardize)
    levels = sorted(set(r[1] for r in cv))
    cov_codes = np.array([levels.index(r[1]) for r in cv])
    qcov = np.array([float(r[1]) for r in qc])
    covars = {"covar1": ("factor", levels, cov_codes), "qcovar1": ("numeric", qcov)}
    n = len(samples)
    nvar, nsamp, mode = orc.pgen_info(pgen)
    codes = np.stack([orc.pgen_read_hardcalls(pgen, v) for v in range(nvar)])
    chroms = [r[0] for r in pvar]
    out = dict(pgen_hex=pgen.hex(), samples=samples, pvar_header=hdr, pvar=pvar, codes=codes.tolist(), y=y.tolist(),
               covar1=cov_codes.tolist(), covar1_levels=levels, qcovar1=qcov.tolist(), cases=[])
    # a GRM-like PSD matrix from the genotypes themselves plus noise (AIREML is out of scope; variance components fixed at 0.5/0.5)
    rng = np.random.default_rng(5)
    Z = rng.standard_normal((n, 6)); K = Z @ Z.T / 6
    for lname, L in (("identity", np.eye(n)), ("grm", np.linalg.cholesky(0.5 * K + 0.5 * np.eye(n)))):
        for fname, terms, null_terms in (("test_config", ["x", "I(x == 1)", "x:covar1", "covar1", "qcovar1"], ["covar1", "qcovar1"]),
                                         ("baseline_simple", ["x"], [])):
            M0, M1, M2, names = orc.design_lookup(terms, covars, n)
            Xn, nn = orc.lower_design(null_terms, covars, np.zeros(n))
            y_mm = orc.forwardsolve(L, y)[:, 0]; X_mm = orc.forwardsolve(L, Xn)
            nm = orc.fit_null_model(y_mm, X_mm)
            case = dict(L=lname, formula=fname, colnames=names, null_colnames=nn, Lmat=L.tolist(), y_mm=y_mm.tolist(), X_null=X_mm.tolist(),
                        ll_null=nm["ll"], df_null=int(nm["df"]), M0=M0.tolist(), M1=M1.tolist(), M2=M2.tolist(),
                        y_pred=nm["y_pred"].tolist(), U1=nm["U1"].tolist(), e_p=nm["e_p"].tolist(), chr={})
            for chrom in sorted(set(chroms)):
                vidx = [i + 1 for i, c in enumerate(chroms) if c == chrom]
                g = codes[[v - 1 for v in vidx]]
                o = orc.fit_model(L, y_mm, X_mm, nm["ll"], nm["df"], M0, M1, M2, g)
                perm = {}
                for sd in (1, 2):
                    p = orc.fit_model(L, y_mm, X_mm, nm["ll"], nm["df"], M0, M1, M2, g, perm_seed=sd, y_pred=nm["y_pred"], U1=nm["U1"], e_p=nm["e_p"])
                    perm[str(sd)] = dict(min_pval=p["min_pval"], nvars=int(p["nvars"]), idx=orc.r_sample(sd, len(nm["e_p"])).tolist())
                case["chr"][chrom] = dict(var_idx=vidx, chisq=o["chisq"].tolist(), df=o["df"].tolist(), pval=[None if np.isnan(v) else v for v in o["pval"]],
                                          rank=o["rank"].tolist(), pivot=o["pivot"].tolist(), coef=o["coef"].tolist(), perm=perm)
            out["cases"].append(case)
    json.dump(out, open(f"{HERE}/c1_test_asset.json", "w"))


def synth_small():
    out = {}
    for name, (n, M, model, P, n_raw) in {"additive": (200, 64, "additive", 3, None), "domgxe": (150, 48, "domgxe", 2, 170),
                                           "simple": (257, 33, "simple", 2, None)}.items():
        t = synth.task(n, M, model=model, P=P, n_raw=n_raw, seed=3)
        codes = t["codes_raw"][:, t["pgen_order"] - 1]
        o = orc.fit_model(t["L"], t["y_mm"], t["X_null"], t["ll_null"], t["df_null"], t["M0"], t["M1"], t["M2"], codes)
        mp = [orc.fit_model(t["L"], t["y_mm"], t["X_null"], t["ll_null"], t["df_null"], t["M0"], t["M1"], t["M2"], codes, perm_seed=int(s),
                            y_pred=t["y_pred"], U1=t["U1"], e_p=t["e_p"])["min_pval"] for s in t["seeds"]]
        for k2 in ("chisq", "df", "pval", "rank", "pivot", "coef"):
            out[f"{name}.{k2}"] = o[k2]
        out[f"{name}.min_pval"] = np.array(mp)
        out[f"{name}.args"] = np.array([n, M, P, n_raw or n, 3])
    np.savez_compressed(f"{HERE}/synth_small.npz", **out)


def r_pin_inputs():
    """Input bundles of the R pin (tests/golden/make_r_fixture.R): the C1 test assets (both formula pairs, L = I and a GRM-based L,
    all 22 variants as one task) and three small synthetic tasks.  Everything R needs to run fit_model.nf:96-143 literally, plus
    the design lookups M0/M1/M2 the C ABI takes (R ignores them)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import r_pin
    os.makedirs(r_pin.PIN_DIR, exist_ok=True)
    c = json.load(open(f"{HERE}/c1_test_asset.json"))
    formulas = {"test_config": "y ~ x + I(x == 1) + x:covar1 + covar1 + qcovar1", "baseline_simple": "y ~ x"}
    codes = np.array(c["codes"], dtype=np.float64)            # (22, 12), already in model order (samples == psam)
    for case in c["cases"]:
        name = f"c1_{case['L']}_{case['formula']}"
        items = {"case": name, "formula": formulas[case["formula"]], "L": np.array(case["Lmat"]), "y_mm": case["y_mm"],
                 "X_null": np.array(case["X_null"]).reshape(len(case["y_mm"]), -1), "ll_null": [case["ll_null"]], "df_null": [case["df_null"]],
                 "y_pred": case["y_pred"], "U1": np.array(case["U1"]), "e_p": np.array(case["e_p"]).reshape(-1, 1), "geno": codes.T,
                 "seeds": [1, 2], "M0": np.array(case["M0"]), "M1": np.array(case["M1"]), "M2": np.array(case["M2"]),
                 "factorlevels.covar1": " ".join(c["covar1_levels"]), "factorcodes.covar1": c["covar1"], "numeric.qcovar1": c["qcovar1"]}
        assert all(" " not in lev for lev in c["covar1_levels"])
        r_pin.write_bundle(os.path.join(r_pin.PIN_DIR, name + ".in.txt"), items)
    formulas = {"additive": "y ~ x + cov1", "domgxe": "y ~ x + I(x == 1) + x:cov1 + cov1", "simple": "y ~ x"}
    for model, (n, M, n_raw) in {"additive": (60, 24, None), "domgxe": (50, 24, 57), "simple": (65, 16, None)}.items():
        t = synth.task(n, M, model=model, P=2, n_raw=n_raw, seed=21, effect=0.4)
        # the reference's own null fit (eigen basis U1) instead of the synthetic Householder one, so U1 is what R would hold
        nm = orc.fit_null_model(t["y_mm"], t["X_null"])
        cov1 = synth.design(model, n, seed=11 + 21)[6]
        geno = t["codes_raw"][:, t["pgen_order"] - 1].astype(np.float64)
        if model == "domgxe":
            j = int(np.argmax(geno.mean(axis=1)))
            geno[j] = np.where(geno[j] == 1, 2, geno[j])      # polymorphic without heterozygotes: I(x == 1) is all zero (the rank-deficient path)
            geno[5] = 2.0                                      # monomorphic
        items = {"case": f"synth_{model}", "formula": formulas[model], "L": t["L"], "y_mm": t["y_mm"], "X_null": t["X_null"], "ll_null": [nm["ll"]],
                 "df_null": [nm["df"]], "y_pred": nm["y_pred"], "U1": nm["U1"], "e_p": np.asarray(nm["e_p"]).reshape(-1, 1), "geno": geno.T,
                 "seeds": [1, 2], "M0": t["M0"], "M1": t["M1"], "M2": t["M2"]}
        if model != "simple":
            items["numeric.cov1"] = cov1
        r_pin.write_bundle(os.path.join(r_pin.PIN_DIR, f"synth_{model}.in.txt"), items)


def kat():
    json.dump({
        "sample10": {"1": [9, 4, 7, 1, 2, 5, 3, 10, 6, 8], "42": [1, 5, 10, 8, 2, 4, 6, 9, 7, 3], "123": [3, 10, 2, 8, 6, 9, 1, 7, 5, 4]},
        "runif_seed1": [0.2655087, 0.3721239, 0.5728534],
        "test_asset_identity": {"rank": 5, "pivot": [1, 2, 4, 5, 6, 3], "lrt_df": 2, "lrt_chisq": 9.98853558118183, "pval": 0.006776681232578928,
                                "rss0": 4.561540527982139, "coef": [-0.744845, -0.04904, 1.803825, 0.078684, -0.879886, 0]},
        "pchisq": [[3.841458820694124, 1, 0.05]],
    }, open(f"{HERE}/kat.json", "w"), indent=1)


if __name__ == "__main__":
    orc.build()
    c1(); synth_small(); kat(); r_pin_inputs()
    print("golden fixtures written to", HERE)
