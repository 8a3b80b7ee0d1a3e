"""Downstream consumers of the hot path's outputs, mirrored on the host: CAT_PERM + SUMMARISE_BY_CHR
(/root/reference/modules/local/r/get_null_p_dist.nf:24-25,73-83) and the two significance thresholds of the Manhattan
plot (/root/reference/modules/local/r/make_plots.nf:31-32).  Pure bookkeeping on P numbers per chromosome; the type-7
quantile itself is the C ABI's flx_minp_threshold.
"""
import gzip

from . import _capi
from .fit_model import _r_num, minp_threshold


def cat_perm(paths, out_path):
    """CAT_PERM: header line + the headerless per-(seed, chromosome) lines of every FIT_MODEL_PERM output (gz members
    are simply concatenated by the reference, get_null_p_dist.nf:24-25)."""
    with gzip.open(out_path, "wt") as out:
        out.write("permutation\tchr\tmin_pval\tnvars\n")
        for p in paths:
            with gzip.open(p, "rt") as f:
                for line in f:
                    out.write(line)


def read_perm_table(path):
    rows = []
    with gzip.open(path, "rt") as f:
        header = f.readline().rstrip("\n").split("\t")
        assert header == ["permutation", "chr", "min_pval", "nvars"], header
        for line in f:
            a = line.rstrip("\n").split("\t")
            rows.append((int(a[0]), a[1], float(a[2]), int(a[3])))
    return rows


def summarise_by_chr(rows, nperms):
    """SUMMARISE_BY_CHR: per permutation, min over chromosomes of min_pval and sum of nvars (get_null_p_dist.nf:77-79);
    the reference asserts one row per permutation (:80).  Returns [(permutation, min_pval, nvars)] sorted by permutation
    (merge() sorts by the key)."""
    acc = {}
    for perm, _chr, mp, nv in rows:
        if perm in acc:
            acc[perm] = (min(acc[perm][0], mp), acc[perm][1] + nv)
        else:
            acc[perm] = (mp, nv)
    if len(acc) != int(nperms):
        raise ValueError(f"nrow(out) == {len(acc)} but nperms == {nperms} (get_null_p_dist.nf:80 stopifnot)")
    return [(p, acc[p][0], acc[p][1]) for p in sorted(acc)]


def write_genome_wide_min_p(path, table):
    """write.table(out, sep = "\\t", row.names = FALSE, col.names = TRUE) (get_null_p_dist.nf:81-83): R quotes the header."""
    with gzip.open(path, "wt") as f:
        f.write('"permutation"\t"min_pval"\t"nvars"\n')
        for perm, mp, nv in table:
            f.write(f"{perm}\t{_r_num(mp)}\t{nv}\n")


def thresholds(table, n_tests, p_thr=0.05):
    """perm_thr = quantile(min_pval, p_thr) and bonferroni = p_thr / nrow(gwas) (make_plots.nf:31-32)."""
    mins = [t[1] for t in table]
    return dict(perm_thr=minp_threshold(mins, p_thr), bonferroni=p_thr / n_tests)
