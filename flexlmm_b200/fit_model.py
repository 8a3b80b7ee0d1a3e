"""Host-side mirror of the reference's FIT_MODEL process (/root/reference/modules/local/r/fit_model.nf).

Same inputs (whitened y / null design / Cholesky factor from DECORRELATE, null fit from FIT_NULL_MODEL, variant
indices, the real model as a genotype lookup, pgen + sample order, permutation seeds), same outputs (per-SNP
`chr pos id ref alt lrt_chisq lrt_df pval beta` table, per-permutation `seed chr min_pval nvars` lines).  All
numerical work happens in libflexlmm_cuda.so through the C ABI; nothing here computes on the CPU.
"""
import ctypes as C
import gzip

import numpy as np

from . import _capi
from ._capi import FlxConfig, FlxSnpOut, FlxStats, check, f64, ptr


def perm_indices(seed, m):
    """R's `set.seed(seed); sample.int(m)` (fit_model.nf:97-98), 1-based int32."""
    out = np.empty(int(m), dtype=np.int32)
    check(_capi.load().flx_perm_indices(int(seed), int(m), out.ctypes.data))
    return out


def minp_threshold(min_pval, p_thr=0.05):
    """quantile(min_pval, p_thr), R type 7 — the Westfall-Young threshold (make_plots.nf:31)."""
    x = f64(min_pval)
    return _capi.load().flx_minp_threshold(x.ctypes.data, x.size, float(p_thr))


def fit_null_model(X_mm, y_mm, want_U1=False):
    """FIT_NULL_MODEL without the n x n eigen (fit_null_model.nf:33-72): ll, df, y_pred, resid, e_p and, on request, an explicit
    copy of the implicit Householder complement basis U1.  Host-only."""
    X = f64(X_mm, fortran=True); y = f64(y_mm)
    n = y.size
    X = X.reshape(n, -1, order="F")
    ll, df, m = C.c_double(), C.c_int32(), C.c_int64()
    y_pred = np.empty(n); resid = np.empty(n); e_p = np.empty(n)
    check(_capi.load().flx_fit_null_model(ptr(X), X.shape[1], ptr(y), n, C.byref(ll), C.byref(df), ptr(y_pred), ptr(resid), ptr(e_p), None, C.byref(m)))
    out = dict(ll=ll.value, df=df.value, y_pred=y_pred, resid=resid, e_p=e_p[: m.value].copy(), m=m.value)
    if want_U1:
        U1 = np.empty((n, m.value), order="F")
        check(_capi.load().flx_fit_null_model(ptr(X), X.shape[1], ptr(y), n, None, None, None, None, None, ptr(U1), None))
        out["U1"] = U1
    return out


class FitModel:
    """One FIT_MODEL task bound to one B200.  Method order mirrors the R script: inputs (:30-66), then run (:120-171)."""

    def __init__(self, device=0, batch_tiles=0, verbose=0, int8=True, planes=6):
        self._lib = _capi.load()
        self._h = C.c_void_p()
        cfg = FlxConfig(int(device), int(batch_tiles), int(verbose), (0 if int8 else 1) | (4 if planes == 5 else 0))
        check(self._lib.flx_create(C.byref(cfg), C.byref(self._h)))
        self._keep = {}
        self.n = None
        self.k = None
        self.P = 0

    def close(self):
        if self._h:
            self._lib.flx_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- inputs
    def set_whitening(self, L, is_inverse=False):
        """L = l.mm[["L"]] (fit_model.nf:61), lower Cholesky factor of the phenotype covariance."""
        L = f64(L, fortran=True)
        assert L.ndim == 2 and L.shape[0] == L.shape[1]
        self.n = L.shape[0]
        check(self._lib.flx_set_whitening(self._h, ptr(L), self.n, self.n, int(bool(is_inverse))))

    def whitening_begin(self, n):
        """Streaming form of set_whitening (n = 100,000: the 80 GB factor exists once, on the device)."""
        self.n = int(n)
        check(self._lib.flx_whitening_begin(self._h, self.n))

    def whitening_cols(self, j0, ncols, cols_ptr, ld):
        """Columns [j0, j0 + ncols) of L from a raw host or device address (column-major, leading dimension ld)."""
        check(self._lib.flx_whitening_cols(self._h, int(j0), int(ncols), int(cols_ptr), int(ld)))

    def whitening_end(self, is_inverse=False):
        check(self._lib.flx_whitening_end(self._h, int(bool(is_inverse))))

    def decorrelate(self, K, s2g, s2e, y, X, min_allowed_eig=-1e-6, eig_replacement=1e-10, want_L=True):
        """DECORRELATE on the GPU (decorrelate.nf:50-77).  Returns dict(L, y_mm, X_mm, clipped_eigs); the factor becomes this
        handle's whitening matrix without a host round trip."""
        K = f64(K, fortran=True); n = K.shape[0]
        y = f64(y); X = f64(X, fortran=True).reshape(n, -1, order="F")
        L = np.empty((n, n), order="F") if want_L else None
        y_mm = np.empty(n); X_mm = np.empty_like(X, order="F")
        nclip = C.c_int32(0)
        check(self._lib.flx_decorrelate(self._h, ptr(K), n, float(s2g), float(s2e), ptr(y), ptr(X), X.shape[1], float(min_allowed_eig),
                                        float(eig_replacement), ptr(L), ptr(y_mm), ptr(X_mm), C.byref(nclip)))
        self.n = n
        return dict(L=L, y_mm=y_mm, X_mm=X_mm, clipped_eigs=nclip.value)

    def set_null(self, X_null, y_mm, ll_null, df_null):
        """X.mm, y.mm (fit_model.nf:59-60) and ll.null with attr df (fit_model.nf:63)."""
        X = f64(X_null, fortran=True).reshape(self.n, -1, order="F")
        y = f64(y_mm)
        check(self._lib.flx_set_null(self._h, ptr(X), X.shape[1], ptr(y), float(ll_null), int(df_null)))

    def set_design(self, M0, M1, M2):
        """model.matrix(model[-2], .) evaluated with x = 0, 1, 2 (fit_model.nf:124-127)."""
        M0, M1, M2 = (f64(m, fortran=True) for m in (M0, M1, M2))
        self.k = M0.shape[1]
        self.s = int(sum(not (np.array_equal(M0[:, j], M1[:, j]) and np.array_equal(M0[:, j], M2[:, j])) for j in range(self.k)))
        check(self._lib.flx_set_design(self._h, ptr(M0), ptr(M1), ptr(M2), self.k))

    def set_design_probes(self, M05, M15):
        """model.matrix evaluated with x = 0.5 and x = 1.5: how the SNP columns continue between hard calls (needed by use_dosage)."""
        M05, M15 = (f64(m, fortran=True) for m in (M05, M15))
        check(self._lib.flx_set_design_probes(self._h, ptr(M05), ptr(M15)))

    def set_use_dosage(self, use_dosage=True):
        """fit_model.nf:26: pgenlibr::Read (dosages) instead of ReadHardcalls."""
        check(self._lib.flx_set_use_dosage(self._h, int(bool(use_dosage))))

    def set_perm(self, y_pred, U1, e_p, seeds):
        """y.mm.pred, U1, e.p (fit_model.nf:64-66) and the seeds (workflows/flexlmm.nf:67)."""
        seeds = np.ascontiguousarray(seeds, dtype=np.int32)
        self.P = int(seeds.size)
        if self.P == 0:
            check(self._lib.flx_set_perm(self._h, None, None, None, 0, None, 0))
            return
        y_pred = f64(y_pred); U1 = f64(U1, fortran=True); e_p = f64(np.ravel(e_p))
        check(self._lib.flx_set_perm(self._h, ptr(y_pred), ptr(U1), ptr(e_p), e_p.size, ptr(seeds), self.P))

    def set_perm_implicit(self, seeds):
        """Permutations from the implicit Householder complement of the null design given to set_null (no U1 matrix)."""
        seeds = np.ascontiguousarray(seeds, dtype=np.int32)
        self.P = int(seeds.size)
        check(self._lib.flx_set_perm_implicit(self._h, ptr(seeds), self.P))

    def open_pgen(self, path, pgen_order=None, n=None):
        """pgenlibr::NewPgen + match(samples, psam$IID) (fit_model.nf:40,70). pgen_order is 1-based."""
        n = self.n if n is None else n
        po = None if pgen_order is None else np.ascontiguousarray(pgen_order, dtype=np.int32)
        check(self._lib.flx_open_pgen(self._h, str(path).encode(), ptr(po), int(n)))

    def pgen_info(self):
        v, sc, d = C.c_int64(), C.c_int64(), C.c_int32()
        check(self._lib.flx_pgen_info(self._h, C.byref(v), C.byref(sc), C.byref(d)))
        return dict(variant_ct=v.value, sample_ct=sc.value, has_dosage=bool(d.value))

    def set_packed(self, packed, raw_sample_ct, pgen_order=None, n=None):
        """Packed 2-bit records (M_total x bytes_per_variant uint8) already in host memory."""
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        assert packed.ndim == 2
        n = self.n if n is None else n
        po = None if pgen_order is None else np.ascontiguousarray(pgen_order, dtype=np.int32)
        self._keep["packed"] = packed  # the library reads it during run
        check(self._lib.flx_set_packed(self._h, ptr(packed), packed.shape[0], packed.shape[1], int(raw_sample_ct), ptr(po), int(n)))

    def make_resident(self):
        check(self._lib.flx_make_resident(self._h))

    def set_whitening_bcast(self, L, n, root=0, is_inverse=False):
        """Collective over the attached communicator: only `root` passes L (others None); L^-1 is broadcast over NVLink."""
        self.n = int(n)
        Lh = None if L is None else f64(L, fortran=True)
        check(self._lib.flx_set_whitening_bcast(self._h, ptr(Lh), self.n, self.n, int(bool(is_inverse)), int(root)))

    def set_perm_bcast(self, y_pred, U1, e_p, seeds, root=0):
        """Collective: only `root` passes U1 (others None); it is uploaded once and broadcast over NVLink."""
        seeds = np.ascontiguousarray(seeds, dtype=np.int32)
        self.P = int(seeds.size)
        y_pred = f64(y_pred); e_p = f64(np.ravel(e_p))
        U1h = None if U1 is None else f64(U1, fortran=True)
        check(self._lib.flx_set_perm_bcast(self._h, ptr(y_pred), ptr(U1h), ptr(e_p), e_p.size, ptr(seeds), self.P, int(root)))

    def comm_attach(self, unique_id, rank, nranks):
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        assert uid.size == 128
        check(self._lib.flx_comm_attach(self._h, ptr(uid), int(rank), int(nranks)))

    # ---- the hot loop
    def run(self, var_idx, per_snp=True, minp=None, out=None):
        """var_idx: 1-based variant numbers (get_var_idx.nf:34-35).  Returns a dict.  `out`: the dict of an earlier call with
        the same number of variants, whose per-SNP arrays are overwritten instead of allocating (and page-faulting) new ones."""
        var_idx = np.ascontiguousarray(var_idx, dtype=np.int32)
        M = var_idx.size
        res = {}
        prev, out = out, None
        if per_snp:
            shapes = dict(lrt_chisq=((M,), np.float64), pval=((M,), np.float64), lrt_df=((M,), np.int32), rank=((M,), np.int32),
                          pivot=((M, self.k), np.int32), coef=((M, self.k), np.float64))
            reuse = prev is not None and all(f in prev and prev[f].shape == sh and prev[f].dtype == dt and prev[f].flags.c_contiguous
                                             for f, (sh, dt) in shapes.items())
            res = {f: (prev[f] if reuse else np.empty(sh, dtype=dt)) for f, (sh, dt) in shapes.items()}
            out = FlxSnpOut(*(res[f].ctypes.data for f in ("lrt_chisq", "pval", "lrt_df", "rank", "pivot", "coef")))
        want_minp = (self.P > 0) if minp is None else bool(minp)
        mp = np.empty(self.P) if want_minp else None
        nv = C.c_int64(0)
        check(self._lib.flx_run(self._h, ptr(var_idx), M, C.byref(out) if out is not None else None, ptr(mp), C.byref(nv)))
        res["nvars_tested"] = nv.value
        if want_minp:
            res["min_pval"] = mp
        return res

    def prepare(self):
        """Run the once-per-task setup now (otherwise the first run() does it)."""
        check(self._lib.flx_prepare(self._h))

    def stats(self):
        st = FlxStats()
        check(self._lib.flx_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    # ---- test hooks
    def decode_tile(self, var_idx, want_X=True, want_codes=True):
        var_idx = np.ascontiguousarray(var_idx, dtype=np.int32)
        M = var_idx.size
        X = np.empty((self.n, M * self.s), order="F") if want_X else None
        codes = np.empty((M, self.n), dtype=np.uint8) if want_codes else None
        check(self._lib.flx_decode_tile(self._h, ptr(var_idx), M, ptr(X), ptr(codes)))
        return X, codes

    def whiten(self, X):
        X = f64(X, fortran=True).reshape(self.n, -1, order="F")
        W = np.empty_like(X, order="F")
        check(self._lib.flx_whiten(self._h, ptr(X), X.shape[1], ptr(W)))
        return W


class FitModelMulti:
    """One FIT_MODEL task on several GPUs of one box from ONE caller thread (flx_multi_*): the factor and U1 are uploaded once
    and broadcast over NVLink, var_idx is split into one contiguous SNP block per device, results come back in var_idx order and
    are bit-identical to a single device's."""

    def __init__(self, devices, batch_tiles=0, verbose=0, int8=True, planes=6):
        self._lib = _capi.load()
        self._h = C.c_void_p()
        devs = np.ascontiguousarray(devices, dtype=np.int32)
        cfg = FlxConfig(0, int(batch_tiles), int(verbose), (0 if int8 else 1) | (4 if planes == 5 else 0))
        check(self._lib.flx_multi_create(C.byref(cfg), ptr(devs), devs.size, C.byref(self._h)))
        self.ndev = int(devs.size)
        self._keep = {}
        self.n = self.k = None
        self.P = 0

    def close(self):
        if self._h:
            self._lib.flx_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_whitening(self, L, is_inverse=False):
        L = f64(L, fortran=True)
        self.n = L.shape[0]
        check(self._lib.flx_multi_set_whitening(self._h, ptr(L), self.n, self.n, int(bool(is_inverse))))

    def set_null(self, X_null, y_mm, ll_null, df_null):
        X = f64(X_null, fortran=True).reshape(self.n, -1, order="F"); y = f64(y_mm)
        check(self._lib.flx_multi_set_null(self._h, ptr(X), X.shape[1], ptr(y), float(ll_null), int(df_null)))

    def set_design(self, M0, M1, M2):
        M0, M1, M2 = (f64(m, fortran=True) for m in (M0, M1, M2))
        self.k = M0.shape[1]
        check(self._lib.flx_multi_set_design(self._h, ptr(M0), ptr(M1), ptr(M2), self.k))

    def set_perm(self, y_pred, U1, e_p, seeds):
        seeds = np.ascontiguousarray(seeds, dtype=np.int32)
        self.P = int(seeds.size)
        y_pred = f64(y_pred); U1 = f64(U1, fortran=True); e_p = f64(np.ravel(e_p))
        check(self._lib.flx_multi_set_perm(self._h, ptr(y_pred), ptr(U1), ptr(e_p), e_p.size, ptr(seeds), self.P))

    def set_perm_implicit(self, seeds):
        seeds = np.ascontiguousarray(seeds, dtype=np.int32)
        self.P = int(seeds.size)
        check(self._lib.flx_multi_set_perm_implicit(self._h, ptr(seeds), self.P))

    def open_pgen(self, path, pgen_order=None, n=None):
        po = None if pgen_order is None else np.ascontiguousarray(pgen_order, dtype=np.int32)
        check(self._lib.flx_multi_open_pgen(self._h, str(path).encode(), ptr(po), int(self.n if n is None else n)))

    def set_packed(self, packed, raw_sample_ct, pgen_order=None, n=None):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        po = None if pgen_order is None else np.ascontiguousarray(pgen_order, dtype=np.int32)
        self._keep["packed"] = packed
        check(self._lib.flx_multi_set_packed(self._h, ptr(packed), packed.shape[0], packed.shape[1], int(raw_sample_ct), ptr(po), int(self.n if n is None else n)))

    def make_resident(self):
        check(self._lib.flx_multi_make_resident(self._h))

    def prepare(self):
        check(self._lib.flx_multi_prepare(self._h))

    def run(self, var_idx, per_snp=True):
        var_idx = np.ascontiguousarray(var_idx, dtype=np.int32)
        M = var_idx.size
        res, out = {}, None
        if per_snp:
            res = dict(lrt_chisq=np.empty(M), pval=np.empty(M), lrt_df=np.empty(M, dtype=np.int32), rank=np.empty(M, dtype=np.int32),
                       pivot=np.empty((M, self.k), dtype=np.int32), coef=np.empty((M, self.k)))
            out = FlxSnpOut(*(res[f].ctypes.data for f in ("lrt_chisq", "pval", "lrt_df", "rank", "pivot", "coef")))
        mp = np.empty(self.P) if self.P > 0 else None
        nv = C.c_int64(0)
        check(self._lib.flx_multi_run(self._h, ptr(var_idx), M, C.byref(out) if out is not None else None, ptr(mp), C.byref(nv)))
        res["nvars_tested"] = nv.value
        if mp is not None:
            res["min_pval"] = mp
        return res

    def stats(self, device_index=0):
        st = FlxStats()
        check(self._lib.flx_multi_get_stats(self._h, int(device_index), C.byref(st)))
        return st.as_dict()


def _r_num(x):
    """as.character(<double>) as used by sprintf("%s", .) in fit_model.nf:157-168: 15 significant digits."""
    if np.isnan(x):
        return "NA"
    if np.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    if x == int(x) and abs(x) < 1e15:
        return str(int(x))
    s = f"{x:.15g}"
    if "e" in s:
        mant, exp = s.split("e")
        sign = exp[0] if exp[0] in "+-" else "+"
        digits = exp.lstrip("+-").lstrip("0") or "0"
        s = f"{mant}e{sign}{int(digits):02d}"
    return s


def write_gwas_tsv(path, pvar_rows, colnames, res):
    """ORIG-mode output (fit_model.nf:112-116,145-169): header + one line per SNP, gz."""
    with gzip.open(path, "wt") as f:
        f.write("chr\tpos\tid\tref\talt\tlrt_chisq\tlrt_df\tpval\tbeta\n")
        for i, (chrom, pos, vid, ref, alt) in enumerate(pvar_rows):
            beta = ",".join(f"{nm}~{_r_num(c)}" for nm, c in zip(colnames, res["coef"][i]))
            f.write("\t".join([str(chrom), str(pos), str(vid), str(ref), str(alt), _r_num(res["lrt_chisq"][i]),
                               str(int(res["lrt_df"][i])), _r_num(res["pval"][i]), beta]) + "\n")


def write_perm_tsv(path, seeds, chrom, min_pval, nvars):
    """PERM-mode output (fit_model.nf:176-181): one headerless line per seed; CAT_PERM concatenates gz members."""
    with gzip.open(path, "wt") as f:
        for sd, mp in zip(seeds, min_pval):
            if not (0 <= mp <= 1):
                raise ValueError("min_pval outside [0, 1] (fit_model.nf:177 stopifnot)")
            f.write(f"{int(sd)}\t{chrom}\t{_r_num(mp)}\t{int(nvars)}\n")


def fit_model_task(L, y_mm, X_null, ll_null, df_null, M0, M1, M2, var_idx, pgen=None, packed=None, raw_sample_ct=None,
                   pgen_order=None, y_pred=None, U1=None, e_p=None, seeds=(), device=0, per_snp=True):
    """One call = FIT_MODEL_ORIG plus every FIT_MODEL_PERM task of a (chromosome, phenotype) pair (lmm.nf:47-71)."""
    with FitModel(device=device) as fm:
        fm.set_whitening(L)
        fm.set_null(X_null, y_mm, ll_null, df_null)
        fm.set_design(M0, M1, M2)
        if len(seeds):
            fm.set_perm(y_pred, U1, e_p, seeds)
        if pgen is not None:
            fm.open_pgen(pgen, pgen_order)
        else:
            fm.set_packed(packed, raw_sample_ct, pgen_order)
        return fm.run(var_idx, per_snp=per_snp)


def fit_model_files(prefix, pgen, pvar, psam, samples, var_idx, colnames, L, y_mm, X_null, ll_null, df_null, M0, M1, M2,
                    y_pred=None, U1=None, e_p=None, seeds=(), chrom=None, device=0):
    """The FIT_MODEL process on its own files (fit_model.nf:7-16): reads `pvar` / `psam`, matches the model's `samples`
    to the .psam (:70), tests the variants `var_idx` of `pgen`, and writes `<prefix>.tsv.gwas.gz` (ORIG, :112-116,145-169)
    and, when seeds are given, `<prefix>.perm.tsv.gz` (one line per seed, :176-181).  Returns the result dict."""
    from . import pfiles
    rows = pfiles.read_pvar(pvar, var_idx)
    order = pfiles.pgen_order(samples, pfiles.read_psam(psam))
    res = fit_model_task(L, y_mm, X_null, ll_null, df_null, M0, M1, M2, var_idx, pgen=pgen, pgen_order=order,
                         y_pred=y_pred, U1=U1, e_p=e_p, seeds=seeds, device=device)
    write_gwas_tsv(f"{prefix}.tsv.gwas.gz", rows, colnames, res)
    if len(seeds):
        chroms = {r[0] for r in rows}
        write_perm_tsv(f"{prefix}.perm.tsv.gz", seeds, chrom if chrom is not None else (chroms.pop() if len(chroms) == 1 else "NA"),
                       res["min_pval"], res["nvars_tested"])
    return res
