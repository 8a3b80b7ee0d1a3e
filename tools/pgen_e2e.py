"""End-to-end from a compressed .pgen (mode 0x10, LD chains): flx_open_pgen + flx_run vs the same variants as packed host records."""
import sys, time, numpy as np, tempfile, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import flexlmm_b200 as fx
from flexlmm_b200 import synth
from pgen_writer import write_pgen
n, M = 5000, int(sys.argv[1]) if len(sys.argv) > 1 else 60000
rng = np.random.default_rng(1)
c = rng.integers(0, 3, size=(M, n)).astype(np.uint8)
for v in range(1, M):
    if v % 16:
        c[v] = c[v - v % 16]; f = rng.choice(n, size=40, replace=False); c[v, f] = rng.integers(0, 3, size=40)   # LD: few differences from the base variant
vt = [0 if v % 16 == 0 else 2 for v in range(M)]
t0 = time.time(); blob = write_pgen(c, vt); print("pgen written: %.1f MB (raw 2-bit would be %.1f MB) in %.0f s" % (len(blob) / 1e6, M * n / 4e6, time.time() - t0), flush=True)
path = os.path.join(tempfile.mkdtemp(), "ld.pgen"); open(path, "wb").write(blob)
t = synth.task(n, 8, model="additive", P=100, seed=2)
vi = np.arange(1, M + 1, dtype=np.int32)
def run(src):
    with fx.FitModel(batch_tiles=148) as fm:
        fm.set_whitening(t["L"]); fm.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"]); fm.set_design(t["M0"], t["M1"], t["M2"])
        fm.set_perm(t["y_pred"], t["U1"], t["e_p"], t["seeds"])
        if src == "pgen": fm.open_pgen(path, None)
        else: fm.set_packed(synth.pack_2bit(c), n)
        fm.run(vi)
        best = 1e9
        for _ in range(3):
            t1 = time.perf_counter(); r = fm.run(vi); best = min(best, time.perf_counter() - t1); st = fm.stats()
        print("%-6s wall %.1f ms  device %.1f ms  h2d %.1f MB  -> %.3g SNP-tests/s end to end" % (src, best * 1e3, st["ms_total"], st["h2d_bytes"] / 1e6, M * 101 / best), flush=True)
        return r
ra = run("pgen"); rb = run("packed")
assert np.array_equal(ra["lrt_chisq"], rb["lrt_chisq"]) and np.array_equal(ra["min_pval"], rb["min_pval"])
print("identical results")
