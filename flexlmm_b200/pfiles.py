"""Host-side text inputs of FIT_MODEL that sit next to the .pgen: the .pvar variant table and the .psam sample table
(fit_model.nf:34-52,69-71).  In the pipeline these stay in R; this is the Python mirror used with `FitModel.open_pgen`
and `write_gwas_tsv`.  Plain text or gzip."""
import gzip

import numpy as np


def _open(path):
    path = str(path)
    return gzip.open(path, "rt") if path.endswith(".gz") else open(path, "rt")


def read_pvar(path, var_idx=None):
    """Rows (chrom, pos, id, ref, alt) of a .pvar: '##' lines are comments, the '#CHROM ...' line is the header
    (fit_model.nf:34-38 strips the first '#' to make it visible); a headerless file is read as .bim
    (chrom, id, cm, pos, alt, ref).  var_idx: optional 1-based row selection, like `[var_idx, ]`."""
    cols, rows = None, []
    with _open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if not line or line.startswith("##"):
                continue
            if line.startswith("#"):
                cols = line[1:].split()
                continue
            parts = line.split()
            if cols is None:
                if len(parts) < 6:
                    raise ValueError(f"{path}: headerless .pvar needs the six .bim columns")
                rows.append((parts[0], int(parts[3]), parts[1], parts[5], parts[4]))
            else:
                rec = dict(zip(cols, parts))
                rows.append((rec["CHROM"], int(rec["POS"]), rec["ID"], rec["REF"], rec["ALT"]))
    if var_idx is not None:
        rows = [rows[int(i) - 1] for i in np.asarray(var_idx)]
    return rows


def read_psam(path):
    """IID column of a .psam (fit_model.nf:41-52: tab separated, header '#IID ...' or '#FID IID ...'); a headerless file is
    read as .fam (FID IID ...)."""
    iids, cols = [], None
    with _open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            if line.startswith("#") and cols is None and not iids:
                cols = [c.lstrip("#") for c in line.split("\t")]
                if "IID" not in cols:
                    raise ValueError(f"{path}: no IID column in the .psam header")
                continue
            if cols is None:
                iids.append(line.split()[1])
            else:
                iids.append(line.split("\t")[cols.index("IID")])
    return iids


def pgen_order(samples, psam_iids):
    """match(samples, psam$IID) (fit_model.nf:70): 1-based position of every model sample in the .psam; a sample that is
    not there is an error (the reference would index with NA)."""
    pos = {}
    for i, s in enumerate(psam_iids):
        pos.setdefault(s, i + 1)          # match() returns the first occurrence
    try:
        return np.array([pos[s] for s in samples], dtype=np.int32)
    except KeyError as e:
        raise ValueError(f"sample {e.args[0]!r} of the model frame is not in the .psam") from None
