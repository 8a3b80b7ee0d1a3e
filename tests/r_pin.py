"""Plain-text bundles shared with tests/golden/make_r_fixture.R (the R side needs base R only): records of two lines,
"@name nrow ncol" + nrow*ncol numbers in column-major order, and "$name text" for strings."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PIN_DIR = os.path.join(HERE, "golden", "r_pin")


def write_bundle(path, items):
    with open(path, "w") as f:
        for name, v in items.items():
            if isinstance(v, str):
                f.write(f"${name} {v}\n")
                continue
            a = np.asarray(v, dtype=np.float64)
            a = a.reshape(-1, 1) if a.ndim < 2 else a
            f.write(f"@{name} {a.shape[0]} {a.shape[1]}\n")
            f.write(" ".join("NA" if np.isnan(x) else repr(float(x)) for x in a.ravel(order="F")) + "\n")


def read_bundle(path):
    out = {}
    with open(path) as f:
        lines = f.read().split("\n")
    i = 0
    while i < len(lines):
        ln = lines[i]
        if ln.startswith("$"):
            name, _, text = ln[1:].partition(" ")
            out[name] = text
            i += 1
        elif ln.startswith("@"):
            name, nr, nc = ln[1:].split(" ")
            nr, nc = int(nr), int(nc)
            vals = np.array([np.nan if t == "NA" else float(t) for t in lines[i + 1].split()], dtype=np.float64)
            assert vals.size == nr * nc, (path, name)
            out[name] = vals.reshape((nr, nc), order="F")
            i += 2
        else:
            i += 1
    return out


def case_paths():
    """[(case name, input bundle, R output or None)] for every committed input bundle."""
    if not os.path.isdir(PIN_DIR):
        return []
    out = []
    for fn in sorted(os.listdir(PIN_DIR)):
        if fn.endswith(".in.txt"):
            o = os.path.join(PIN_DIR, fn[:-7] + ".out.txt")
            out.append((fn[:-7], os.path.join(PIN_DIR, fn), o if os.path.exists(o) else None))
    return out
