"""LOCO driver: one FIT_MODEL task per chromosome, each with its OWN Cholesky factor, on one GPU.

The reference's leave-one-chromosome-out mode (subworkflows/local/lmm.nf:27-46) scatters AIREML / DECORRELATE / FIT_NULL_MODEL /
FIT_MODEL per chromosome: chromosome c is tested against the factor of the GRM built WITHOUT c, so a genome-wide run walks
22 factors (BASELINE.json configs[3]: n = 50,000, 20 GB each).  A factor's setup (upload, in-place inverse, V^-1 row tiles ->
digit planes: ~2 n^3 / 3 FP64 flops) costs more than the chromosome's tests, so the driver keeps TWO handles per GPU and sets
chromosome k+1 up on the idle one, from a host thread, while chromosome k runs (flx_prepare; the handles share no state).
Across GPUs chromosomes — not SNP blocks — are the unit of sharding: a factor is then set up once in the box
instead of once per GPU, and the only exchange stays the min over chromosomes of the length-P min-p vector
(get_null_p_dist.nf:77).
"""
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .fit_model import FitModel


def chromosomes_of_rank(n_chrom, rank, world):
    """Chromosome k goes to rank k % world (sizes shrink with the chromosome number, so round-robin balances better than blocks)."""
    return [c for c in range(n_chrom) if c % world == rank]


class LocoRunner:
    def __init__(self, device=0, handles=2, **fit_model_kwargs):
        """handles = 2: ping-pong (production: a genome of 22 factors never fits at once).  handles >= len(tasks): every factor stays
        resident on its own handle after the first pass, so rerun() repeats the tests without any setup (benchmarks)."""
        self.fms = [FitModel(device=device, **fit_model_kwargs) for _ in range(max(1, handles))]
        self.setup_s = []     # host seconds each chromosome's setup took (on the side thread, overlapped with the previous run)
        self.run_s = []
        self.var_idx = []

    def close(self):
        for fm in self.fms:
            fm.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @staticmethod
    def _prepare(fm, task):
        t0 = time.perf_counter()
        var_idx = task(fm)               # set_whitening / set_null / set_design / set_perm* / set_packed | open_pgen -> var_idx
        fm.prepare()
        return np.ascontiguousarray(var_idx, dtype=np.int32), time.perf_counter() - t0

    def run(self, tasks, per_snp=True):
        """tasks: one callable per chromosome, `task(fm) -> var_idx`, that hands the chromosome's inputs to the FitModel it is
        given.  Chromosome k+1 is set up on the next handle while chromosome k runs.  Returns the per-chromosome result dicts;
        `min_pval` of the genome is their elementwise minimum (genome_min_pval)."""
        self.setup_s, self.run_s, self.var_idx = [], [], []
        results = []
        nf = len(self.fms)
        with ThreadPoolExecutor(1) as side:
            fut = side.submit(self._prepare, self.fms[0], tasks[0])
            for k in range(len(tasks)):
                var_idx, ts = fut.result()
                self.setup_s.append(ts)
                self.var_idx.append(var_idx)
                if k + 1 < len(tasks) and nf > 1:
                    fut = side.submit(self._prepare, self.fms[(k + 1) % nf], tasks[k + 1])
                t0 = time.perf_counter()
                results.append(self.fms[k % nf].run(var_idx, per_snp=per_snp))
                self.run_s.append(time.perf_counter() - t0)
                if k + 1 < len(tasks) and nf == 1:
                    fut = side.submit(self._prepare, self.fms[0], tasks[k + 1])
        return results

    def rerun(self, per_snp=True, out=None):
        """The tests of every chromosome again on the resident factors (needs handles >= chromosomes of the last run())."""
        if len(self.var_idx) > len(self.fms):
            raise RuntimeError("rerun() needs one handle per chromosome: the factors of a ping-pong run are gone")
        out = out or [None] * len(self.var_idx)
        return [self.fms[k].run(v, per_snp=per_snp, out=out[k]) for k, v in enumerate(self.var_idx)]

    @staticmethod
    def genome_min_pval(results):
        """min over chromosomes per permutation (get_null_p_dist.nf:77)."""
        mp = [r["min_pval"] for r in results if "min_pval" in r]
        return np.minimum.reduce(mp) if mp else None
