#!/usr/bin/env python
"""bench.py — SNP-tests/s (including permutations) of the per-SNP LRT engine on B200.

  python bench.py --gpus N --steps K --warmup W             # this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W      # the reference's R loop restated on the host cores (oracle port)
  python bench.py --workload C3|C4|C5 [--full]               # the other BASELINE shapes as stand-alone lines

A "step" is one full pass of the hot path (decode -> whiten -> fit/LRT -> permutation contraction -> min-p) over the
workload's SNP set.  N=1 workload = BASELINE.json configs[1] (C2): synthetic n=5,000 x 1M SNPs, additive y ~ x + cov1 vs
y ~ cov1, 100 permutations.  For N>1 every rank runs the same per-GPU workload on its own SNP shard (weak scaling) and only the
length-P min-p vector is min-allreduced with NCCL inside the library.

`value`   : SNP-tests/s with the packed genotypes already resident in HBM, from the WALL clock around the K timed flx_run calls
            (barrier + synchronize on both sides, max over ranks); `device_ms_per_step` is the same loop on CUDA events.
`e2e`     : the same metric through the C ABI with HOST buffers (packed records H2D and per-SNP results D2H inside the timed
            region), wall clock.
`roofline`: the dominant kernel (K2q: the exact int8 quadratic forms; K2 on the FP64 route), algorithmic ops per launch over
            its CUDA-event time, against the tensor-pipe figure measured on this pool.
`cpu_baseline`: the oracle port of the reference loop timed on this box's host cores on a bounded SNP sample.
`shapes`  : the BASELINE shapes C3 / C4 (LOCO, two factors with overlapped setup) / C5 as per-GPU slices, measured in the same
            run after the headline numbers (n and P as BASELINE.json states them, M a stated slice).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

# name: n, per-GPU SNPs (slice / full = BASELINE M over 8 GPUs), model, P.  C4: `chroms` LOCO factors per GPU, M per chromosome.
WORKLOADS = {
    "C2": dict(n=5000, M=1_000_000, M_full=1_000_000, model="additive", P=100, s=1, c=2, k=3, chroms=1, gen="chol",
               desc="synthetic n=5,000 x 1M SNPs, additive y ~ x + cov1 vs y ~ cov1, 100 permutations"),
    "C3": dict(n=20000, M=100_000, M_full=625_000, model="domgxe", P=1000, s=3, c=2, k=5, chroms=1, gen="chol",
               desc="synthetic n=20,000, dominance + GxE y ~ x + I(x==1) + x:cov1 + cov1 vs y ~ cov1, 1,000 permutations (configs[2]: 5M SNPs over 8 GPUs = 625,000 per GPU)"),
    "C4": dict(n=50000, M=37_888, M_full=454_545, model="additive", P=1000, s=1, c=2, k=3, chroms=2, chroms_full=3, gen="structured",
               desc="synthetic n=50,000, LOCO: one 20 GB Cholesky factor per chromosome, additive, 1,000 permutations (configs[3]: 10M SNPs / 22 chromosomes = 454,545 per chromosome, chromosomes sharded over 8 GPUs = 3 per GPU)"),
    "C5": dict(n=100000, M=37_888, M_full=250_000, model="additive", P=10000, s=1, c=2, k=3, chroms=1, gen="structured",
               desc="synthetic n=100,000 (FP64 L^-1 80 GB), additive, 10,000 permutations (configs[4]: 2M SNPs over 8 GPUs = 250,000 per GPU)"),
    "C2-small": dict(n=5000, M=100_000, M_full=100_000, model="additive", P=100, s=1, c=2, k=3, chroms=1, gen="chol",
                     desc="synthetic n=5,000 x 100k SNPs (reduced M, NOT the headline config)"),
    "tiny": dict(n=1000, M=20_000, M_full=20_000, model="additive", P=20, s=1, c=2, k=3, chroms=1, gen="chol", desc="debug"),
    "tiny-loco": dict(n=1500, M=9_000, M_full=9_000, model="additive", P=20, s=1, c=2, k=3, chroms=3, gen="structured", desc="debug (LOCO driver)"),
}


def measured_peaks():
    d = {"hbm_gbs": 6546.2, "bf16_tflops": 1657.4}
    try:
        with open(os.path.join(HERE, "MEASURED_PEAKS.json")) as f:
            d.update(json.load(f))
    except Exception:
        pass
    return d


def i8_peak():
    try:
        with open(os.path.join(HERE, "profiles", "i8_peak.json")) as f:
            d = json.load(f)
        return d["umma_i8_tops"], "measured tcgen05.mma kind::i8 M128 N256 on this pool (tools/qf_probe.cu, profiles/i8_peak.json); 2 x MEASURED_PEAKS bf16 burst = %.0f" % (2 * measured_peaks()["bf16_tflops"])
    except Exception:
        return 2 * measured_peaks()["bf16_tflops"], "2 x MEASURED_PEAKS.json bf16_tflops (int8 tensor rate is twice bf16 on B200)"


def fp64_peak():
    try:
        with open(os.path.join(HERE, "profiles", "fp64_peak.json")) as f:
            d = json.load(f)
        return d["cublas_dgemm_8192_tflops_sustained"], "measured cuBLAS DGEMM 8192^3 on this pool (profiles/fp64_peak.json)"
    except Exception:
        return 37.0, "fallback: nominal B200 FP64 37 TFLOP/s"


# ------------------------------------------------------------------------------------------------ synthetic inputs
def make_packed(n, M, seed, dev):
    """maf ~ U(0.05, 0.5), g ~ Binomial(2, maf), no missing; packed 2-bit records (pgen mode 0x02 body), N_raw = n.  Generated
    on the GPU as input plumbing, returned as a HOST array like a memory-mapped .pgen would be."""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    bpv = (n + 3) // 4
    npad = bpv * 4
    host = torch.empty((M, bpv), dtype=torch.uint8)
    chunk = max(1, min(M, (1 << 27) // max(n, 1)))
    for s0 in range(0, M, chunk):
        m = min(chunk, M - s0)
        maf = 0.05 + 0.45 * torch.rand((m, 1), generator=g, device=dev)
        a = (torch.rand((m, npad), generator=g, device=dev) < maf).to(torch.uint8)
        a += (torch.rand((m, npad), generator=g, device=dev) < maf).to(torch.uint8)
        if npad != n:
            a[:, n:] = 0
        a = a.view(m, bpv, 4)
        host[s0:s0 + m].copy_(a[:, :, 0] | (a[:, :, 1] << 2) | (a[:, :, 2] << 4) | (a[:, :, 3] << 6))
        del a, maf
    return host.numpy()


def task_chol(w, dev, seed=13):
    """n <= 20,000: V = 0.5 B B'/r + 0.5 I, L = chol(V), explicit U1 from the complete QR of the whitened null design
    (what FIT_NULL_MODEL hands over, in the Householder basis).  Returns a callable task(fm) for FitModel / LocoRunner."""
    import torch
    from flexlmm_b200 import synth
    n, P = w["n"], w["P"]
    g2 = torch.Generator(device=dev); g2.manual_seed(seed)
    r = min(256, n)
    B = torch.randn((n, r), generator=g2, device=dev, dtype=torch.float64)
    V = 0.5 * (B @ B.T) / r + 0.5 * torch.eye(n, device=dev, dtype=torch.float64)
    L = torch.linalg.cholesky(V)
    del V, B
    M0, M1, M2, names, Xn, nn, cov1 = synth.design(w["model"], n, seed=11)
    z = torch.randn((n,), generator=g2, device=dev, dtype=torch.float64)
    y = 0.3 * torch.as_tensor(cov1, device=dev) + L @ z
    rhs = torch.cat([y[:, None], torch.as_tensor(Xn, device=dev)], dim=1)
    sol = torch.linalg.solve_triangular(L, rhs, upper=False)
    y_mm, X_mm = sol[:, 0].contiguous(), sol[:, 1:].contiguous()
    c = X_mm.shape[1]
    Q, _ = torch.linalg.qr(X_mm, mode="complete")
    Qc, U1 = Q[:, :c], Q[:, c:]
    fitted = Qc @ (Qc.T @ y_mm)
    e = y_mm - fitted
    rss = float(e @ e)
    ll = 0.5 * (-n * (np.log(2 * np.pi) + 1 - np.log(n) + np.log(rss)))
    inp = dict(L=L.cpu().numpy(), y_mm=y_mm.cpu().numpy(), X_null=np.asfortranarray(X_mm.cpu().numpy()), ll_null=ll, df_null=c + 1,
               M0=M0, M1=M1, M2=M2, y_pred=fitted.cpu().numpy(), U1=np.asfortranarray(U1.cpu().numpy()), e_p=(U1.T @ e).cpu().numpy(),
               seeds=np.arange(1, P + 1, dtype=np.int32))
    del Q, U1, L
    torch.cuda.empty_cache()

    def task(fm, packed, n_raw, bcast_rank=None):
        if bcast_rank is None:
            fm.set_whitening(inp["L"])
        else:   # N > 1: rank 0 alone uploads and inverts the factor; L^-1 and U1 reach the other GPUs by ncclBroadcast over NVLink
            fm.set_whitening_bcast(inp["L"] if bcast_rank == 0 else None, n, root=0)
        fm.set_null(inp["X_null"], inp["y_mm"], inp["ll_null"], inp["df_null"])
        fm.set_design(inp["M0"], inp["M1"], inp["M2"])
        if bcast_rank is None:
            fm.set_perm(inp["y_pred"], inp["U1"], inp["e_p"], inp["seeds"])
        else:
            fm.set_perm_bcast(inp["y_pred"], inp["U1"] if bcast_rank == 0 else None, inp["e_p"], inp["seeds"], root=0)
        fm.set_packed(packed, n_raw)
    return task, inp


def task_structured(w, dev, seed=13):
    """n >= 50,000: no n^3 Cholesky and no second n x n matrix anywhere: a non-singular lower-triangular factor (strict lower
    triangle of a rank-6 product + positive diagonal) is produced on the GPU in column blocks and handed to the library by
    DEVICE pointer (flx_whitening_cols); X.mm = L^-1 [1, cov1] comes from flx_whiten; y.mm = 0.3 X.mm[, cov1] + z is L^-1 y for
    y = 0.3 cov1 + L z; null fit by flx_fit_null_model, permutations from the implicit Householder complement."""
    import torch
    import flexlmm_b200 as fx
    from flexlmm_b200 import synth
    n, P = w["n"], w["P"]
    M0, M1, M2, names, Xn, nn, cov1 = synth.design(w["model"], n, seed=11)
    seeds = np.arange(1, P + 1, dtype=np.int32)

    def task(fm, packed, n_raw, bcast_rank=None):
        g2 = torch.Generator(device=dev); g2.manual_seed(seed)
        # the strictly lower part stays well below the diagonal in norm (row norms <= ~0.7): like chol(s2g K + s2e I), the factor is
        # well conditioned.  (A scale of 1.5 / sqrt(n) gave, for some seeds, a V^-1 under which thousands of SNPs were numerically in the
        # span of the covariates: they all went through the FP64 re-fit and that rank ran 8x longer than its peers.)
        U = torch.randn((n, 6), generator=g2, device=dev, dtype=torch.float64) * (0.5 / np.sqrt(n))
        W = torch.randn((n, 6), generator=g2, device=dev, dtype=torch.float64)
        d = 0.8 + 0.4 * torch.rand((n,), generator=g2, device=dev, dtype=torch.float64)
        z = torch.randn((n,), generator=g2, device=dev, dtype=torch.float64).cpu().numpy()
        rows = torch.arange(n, device=dev)
        fm.whitening_begin(n)
        blk = 2048
        for j0 in range(0, n, blk):
            j1 = min(n, j0 + blk)
            cols = torch.arange(j0, j1, device=dev)
            bt = W[j0:j1] @ U.T                                   # bt[jj, i] = L[i, j0 + jj]: row-major (ncols, n) = column-major n x ncols
            bt.mul_((rows[None, :] > cols[:, None]).to(bt.dtype))
            bt[torch.arange(j1 - j0, device=dev), cols] = d[j0:j1]
            torch.cuda.current_stream().synchronize()
            fm.whitening_cols(j0, j1 - j0, bt.data_ptr(), n)
            del bt
        fm.whitening_end()
        X_mm = fm.whiten(Xn)
        y_mm = 0.3 * X_mm[:, 1] + z
        nf = fx.fit_null_model(X_mm, y_mm)
        fm.set_null(X_mm, y_mm, nf["ll"], nf["df"])
        fm.set_design(M0, M1, M2)
        fm.set_perm_implicit(seeds)
        fm.set_packed(packed, n_raw)
    return task, None


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """One long-running `nvidia-smi -lms` process writing to a file (started before the warm-up steps, so it is sampling by the
    time the timed region begins): nothing forks from this process during the timed region.  Samples are kept by timestamp."""

    Q = "timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None
        self.t_begin = self.t_end = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="flx_clocks_", suffix=".csv")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.device), "-lms", "20"],
                                         stdout=fd, stderr=subprocess.DEVNULL)
            os.close(fd)
        except Exception:
            self.proc = None

    def begin(self):
        import datetime
        self.t_begin = datetime.datetime.now()

    def end(self):
        import datetime
        self.t_end = datetime.datetime.now()

    def stop(self):
        import datetime
        rows, mx = [], None
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            try:
                for ln in open(self.path).read().strip().splitlines():
                    o = [x.strip() for x in ln.split(",")]
                    if len(o) < 7:
                        continue
                    try:
                        ts = datetime.datetime.strptime(o[0], "%Y/%m/%d %H:%M:%S.%f")
                        rows.append((ts, float(o[1]), [nm for nm, v in zip(self.NAMES, o[3:]) if v.lower().startswith("active")]))
                        mx = float(o[2])
                    except ValueError:
                        continue
                os.unlink(self.path)
            except Exception:
                pass
        inside = [r for r in rows if self.t_begin and self.t_end and self.t_begin <= r[0] <= self.t_end]
        window = "timed region"
        if not inside:
            inside, window = rows, "whole sampler life (no sample fell inside the timed region)"
        med = float(np.median([r[1] for r in inside])) if inside else None
        reasons = sorted({nm for r in inside for nm in r[2]})
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": reasons, "samples": len(inside), "window": window, "sampler": "nvidia-smi -lms 20 (one separate process)"}


# ------------------------------------------------------------------------------------------------ CPU baseline (oracle port)
class CpuArm:
    """The reference loop (fit_model.nf:96-171) restated in C (oracle), one single-threaded task per host core like Nextflow's
    scatter of single-threaded R processes.  The task (factor, null fit, genotypes) is built ONCE; every step runs `snps` SNPs x
    (1 ORIG + `tasks`-1 PERM) per core on a persistent thread pool (the C call releases the GIL, the threads share L and U1
    read-only) and only the C loop is timed."""

    def __init__(self, w, cores, snps, tasks=3):
        for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):      # BLAS pinned to 1 thread per task
            os.environ[v] = "1"
        from concurrent.futures import ThreadPoolExecutor
        from oracle import oracle as orc
        from flexlmm_b200 import synth
        orc.build()
        self.orc, self.cores, self.snps, self.tasks, self.w = orc, cores, snps, tasks, w
        self.t = synth.task(w["n"], snps * cores, model=w["model"], P=tasks - 1, seed=0)
        self.pool = ThreadPoolExecutor(cores)
        self.sample = (f"{snps} SNPs x (1 ORIG + {tasks - 1} PERM tasks) per core on {cores} cores, n={w['n']}, reference loop order "
                       "(whole design re-whitened per SNP and per task), task built once, C loop timed alone")

    def _work(self, ci):
        t, orc = self.t, self.orc
        codes = t["codes_raw"][ci * self.snps:(ci + 1) * self.snps]
        t0 = time.perf_counter()
        orc.fit_model(t["L"], t["y_mm"], t["X_null"], t["ll_null"], t["df_null"], t["M0"], t["M1"], t["M2"], codes)
        for sd in range(1, self.tasks):
            orc.fit_model(t["L"], t["y_mm"], t["X_null"], t["ll_null"], t["df_null"], t["M0"], t["M1"], t["M2"], codes, perm_seed=sd,
                          y_pred=t["y_pred"], U1=t["U1"], e_p=t["e_p"])
        return time.perf_counter() - t0

    def step(self):
        """-> (SNP-tests/s over all cores, seconds of the slowest core)"""
        times = list(self.pool.map(self._work, range(self.cores)))
        return self.cores * self.snps * self.tasks / max(times), max(times)

    def close(self):
        self.pool.shutdown()


def cpu_sample_snps(n, seconds):
    """SNPs per core so that a step costs ~`seconds` of CPU: the port does ~0.1 s per SNP-test at n = 5,000, k = 3 (scales with n^2 k)."""
    return max(1, int(round(seconds / 3 / (0.1 * (n / 5000.0) ** 2))))


# ------------------------------------------------------------------------------------------------ rooflines
def stage_roofline(w, acc, steps, int8_path, M, handles=1):
    """Every stage against the roofline that bounds it: algorithmic bytes or ops per step / CUDA-event time."""
    n, P, s, c = w["n"], w["P"], w["s"], w["c"]
    n128 = (n + 127) // 128 * 128
    n256 = (n + 255) // 256 * 256
    hbm = measured_peaks()["hbm_gbs"]
    out = {}

    def put(name, work, unit, ms, peak, bound):
        if ms <= 0:
            return
        ach = work / (ms * 1e-3 / steps) / (1e9 if unit == "GB/s" else 1e12)
        out[name] = {"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak, "ms_per_step": ms / steps}

    raw = (n + 3) // 4
    if int8_path:
        passes = max(1, int(round(acc.get("int8_passes", steps) / (steps * handles))))
        na = 2 if w["model"] == "domgxe" else 1
        put("K1q decode (2-bit records -> 2-bit code tiles)", M * (raw + n256 / 4.0) * na, "GB/s", acc["ms_decode"], hbm, "hbm")
        put("K2q qf_i8 (g'C V^-1 C g, 6 planes x %d pass units)" % passes, 6.0 * n * n * M * passes, "TFLOP/s", acc["ms_whiten"], i8_peak()[0], "tensor(i8)")
        put("K4q (u, b, b* for all seeds, 6 planes) + K3b", 6.0 * 2 * n * (c + 1 + P) * M * s, "TFLOP/s", acc["ms_fit"], i8_peak()[0], "tensor(i8)")
        put("perm_reduce (b*' Sm b* / RSS_f*, max per seed)", 8.0 * M * s * P, "GB/s", acc["ms_perm"], hbm, "hbm")
    else:
        put("K1 decode_2bit_tile", M * (raw + 8.0 * n128 * s), "GB/s", acc["ms_decode"], hbm, "hbm")
        put("K2 whiten_tile (DMMA)", float(s) * n * n * M, "TFLOP/s", acc["ms_whiten"], fp64_peak()[0], "tensor(f64)")
        put("K3a snp_gram (2 sweeps) + K3b", M * (2 * 8.0 * n128 * s), "GB/s", acc["ms_fit"], hbm, "hbm")
        put("K4 perm_contract (DMMA)", 2.0 * n * s * P * M, "TFLOP/s", acc["ms_perm"], fp64_peak()[0], "tensor(f64)")
    return out


def dominant_roofline(w, acc, steps, M, handles=1):
    n = w["n"]
    int8_path = bool(acc.get("int8_path", 0))
    ms_whiten = acc["ms_whiten"]
    if int8_path:
        peak, peak_src = i8_peak()
        passes = max(1, int(round(acc.get("int8_passes", steps) / (steps * handles))))
        work = 6.0 * n * n * M * steps * passes
        kname = "qf_i8_kernel (K2q: g'C V^-1 C g, tcgen05.mma kind::i8, 6 exact base-256 digit planes, %d triangular-pass unit%s per batch)" % (passes, "" if passes == 1 else "s")
        tfile = "qf_traffic.json"
    else:
        peak, peak_src = fp64_peak()
        work = float(w["s"]) * n * n * M * steps
        kname = "panel_gemm_kernel<0,1> (K2 whiten_tile, FP64 DMMA)"
        tfile = "whiten_traffic.json"
    achieved = work / (ms_whiten * 1e-3) / 1e12
    traffic = None
    if w["n"] == 5000:
        try:
            with open(os.path.join(HERE, "profiles", tfile)) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
    nl = max(1, acc["launches_whiten"])
    return {"kernel": kname, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "ops_counted": "int8 ops executed: 6 planes x n^2 per SNP per pass unit (the FP64-equivalent s n^2 figure is fp64_equivalent_tflops)" if int8_path else "FP64 flops: s n^2 per SNP",
            "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "launches": int(acc["launches_whiten"]), "avg_launch_ms": ms_whiten / nl,
            "algorithmic_ops_per_launch": work / nl,
            "fp64_equivalent_tflops": float(w["s"]) * n * n * M * steps / (ms_whiten * 1e-3) / 1e12, "fp64_dgemm_peak_tflops": fp64_peak()[0]}


# ------------------------------------------------------------------------------------------------ one workload on this rank's GPU
class Ctx:
    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]


def attach_comm(ctx, fm):
    import flexlmm_b200 as fx
    torch = ctx.torch
    uid = torch.zeros(128, dtype=torch.uint8)
    if ctx.rank == 0:
        buf = np.zeros(128, dtype=np.uint8)
        fx._capi.check(fx.load().flx_comm_unique_id(buf.ctypes.data))
        uid = torch.from_numpy(buf)
    uid = uid.cuda(); ctx.dist.broadcast(uid, 0)
    fm.comm_attach(uid.cpu().numpy(), ctx.rank, ctx.world)


def run_workload(ctx, name, w, steps, warmup, route="auto", full=False, clocks=False):
    """-> dict of measurements (rank 0 uses it).  Single factor: one FitModel; LOCO (chroms > 1): LocoRunner, factors resident."""
    import flexlmm_b200 as fx
    from flexlmm_b200.loco import LocoRunner
    torch = ctx.torch
    n, P = w["n"], w["P"]
    M = w["M_full"] if full else w["M"]
    chroms = w.get("chroms_full", w["chroms"]) if full else w["chroms"]
    int8 = (route == "auto")
    gen = task_chol if w["gen"] == "chol" else task_structured
    # the packed genotypes of all local ranks live in host memory (like a page-cached .pgen): never ask for more than a third of it
    try:
        import psutil
        avail = psutil.virtual_memory().available
        need = M * chroms * ((n + 3) // 4) * max(1, ctx.world)
        if need > avail / 3:
            M = max(1024, int(M * (avail / 3) / need))
    except Exception:
        pass
    out = {"workload": name, "n": n, "permutations": P, "snps_per_gpu": M * chroms, "chromosomes_per_gpu": chroms}

    t_gen0 = time.perf_counter()
    packs = [make_packed(n, M, 0xF1E + 7919 * ctx.rank + 104729 * ch, ctx.dev) for ch in range(chroms)]
    tasks = [gen(w, ctx.dev, seed=13 + ch + (100 * ctx.rank if chroms > 1 else 0))[0] for ch in range(chroms)]
    out["input_generation_s"] = time.perf_counter() - t_gen0
    var_idx = np.arange(1, M + 1, dtype=np.int32)
    tests_step_rank = M * chroms * (1 + P)

    if chroms == 1:
        fm = fx.FitModel(device=ctx.local_rank, int8=int8)
        if ctx.world > 1:
            attach_comm(ctx, fm)
        ctx.barrier()
        t0 = time.perf_counter()
        tasks[0](fm, packs[0], n, bcast_rank=(ctx.rank if ctx.world > 1 else None))
        fm.prepare()
        out["setup_s_once"] = time.perf_counter() - t0
        fms = [fm]
        runner = None
    else:
        # LOCO: every chromosome its own factor.  First pass through the LOCO driver (setup of chromosome k+1 overlapped with the
        # tests of chromosome k; measured); one handle per chromosome, so the factors then stay resident for the timed steps
        # (the step = the tests, as for C2; the factor setup is reported beside it)
        runner = LocoRunner(device=ctx.local_rank, handles=chroms, int8=int8)
        loco_tasks = [(lambda f, ch=ch: (tasks[ch](f, packs[ch], n), var_idx)[1]) for ch in range(chroms)]
        t0 = time.perf_counter()
        runner.run(loco_tasks, per_snp=False)
        out["loco_first_pass_s"] = time.perf_counter() - t0
        out["loco_setup_s_per_factor"] = list(runner.setup_s)
        out["loco_first_run_s_per_factor"] = list(runner.run_s)
        out["setup_s_once"] = float(np.sum(runner.setup_s))
        fms = runner.fms
        if ctx.world > 1:
            for fm in fms:        # every handle reduces its own min-p vector over the ranks; min over chromosomes afterwards
                attach_comm(ctx, fm)

    def one_step(per_snp, res):
        acc = {}
        mp = None
        for ch, fm in enumerate(fms):
            r = fm.run(var_idx, per_snp=per_snp, out=res[ch])
            res[ch] = r
            st = fm.stats()
            for kk, v in st.items():
                acc[kk] = acc.get(kk, 0) + v
            if "min_pval" in r:
                mp = r["min_pval"] if mp is None else np.minimum(mp, r["min_pval"])
        return acc, mp

    def timed(nsteps, per_snp):
        tot, res, mp = {}, [None] * len(fms), None
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(nsteps):
            acc, mp = one_step(per_snp, res)
            for kk, v in acc.items():
                tot[kk] = tot.get(kk, 0) + v
        ctx.barrier()
        wall = time.perf_counter() - t0
        dev_s, wall = ctx.max_over_ranks([tot.get("ms_total", 0.0) * 1e-3, wall])
        return dev_s, wall, tot, mp

    # ---- e2e: host buffers, H2D + D2H inside the timed region
    timed(max(1, min(warmup, 3)), True)
    e2e_dev, e2e_wall, e2e_acc, _ = timed(steps, True)
    # ---- value: genotypes resident in HBM
    for fm in fms:
        fm.make_resident()
    clk = ClockSampler(ctx.local_rank) if (clocks and ctx.rank == 0) else None
    if clk:
        clk.start()
    timed(max(3, warmup) if chroms == 1 and n <= 20000 else max(1, min(warmup, 2)), False)
    if clk:
        clk.begin()
    dev_s, wall_s, acc, mp = timed(steps, False)
    if clk:
        clk.end()
        out["clocks"] = clk.stop()
    total_tests = tests_step_rank * ctx.world * steps
    Mtot = M * chroms
    out.update({"value": total_tests / wall_s, "ms_per_step": wall_s / steps * 1e3, "device_ms_per_step": dev_s / steps * 1e3,
                "lib_wall_ms_per_step": acc.get("ms_wall", 0.0) / steps,
                "e2e": {"value": total_tests / e2e_wall, "unit": "SNP-tests/s", "h2d_bytes_per_step": e2e_acc["h2d_bytes"] / steps,
                        "d2h_bytes_per_step": e2e_acc["d2h_bytes"] / steps, "ms_per_step_wall": e2e_wall / steps * 1e3, "device_ms_per_step": e2e_dev / steps * 1e3},
                "gpu_launches": int(acc["launches"]), "int8_path": bool(acc.get("int8_path", 0)), "fallback_reason": int(acc.get("fallback_reason", 0) / max(1, steps * chroms)),
                "refit_snps_per_step": acc.get("refit_snps", 0) / steps,
                "stage_ms_per_step": {kk: acc[kk] / steps for kk in ("ms_decode", "ms_whiten", "ms_fit", "ms_perm", "ms_h2d", "ms_d2h")},
                "roofline": dominant_roofline(w, acc, steps, Mtot, chroms), "stage_roofline": stage_roofline(w, acc, steps, bool(acc.get("int8_path", 0)), Mtot, chroms),
                "min_pval_head": [float(x) for x in (mp[:3] if mp is not None else [])]})
    if chroms > 1:
        # the whole per-GPU LOCO job as the pipeline would run it once: factor setups (overlapped with the previous chromosome's
        # tests) + tests; extrapolated to the full chromosome size from the measured per-SNP step time
        t_tests_full = (wall_s / steps) / (M * chroms) * w["M_full"] * w.get("chroms_full", chroms)
        t_setup_full = float(np.mean(out["loco_setup_s_per_factor"])) * w.get("chroms_full", chroms)
        out["loco_job"] = {"first_pass_measured_s": out["loco_first_pass_s"], "tests_s_full_job_per_gpu": t_tests_full, "setup_s_full_job_per_gpu": t_setup_full,
                           "snp_tests_per_s_with_setup_per_gpu": w["M_full"] * w.get("chroms_full", chroms) * (1 + P) / (t_tests_full + t_setup_full),
                           "note": "chromosomes (not SNP blocks) are sharded over the GPUs: each factor is set up once in the box; SNP-block sharding would repeat all 22 setups on every GPU"}
    for fm in fms:
        fm.close()
    del packs, tasks
    torch.cuda.empty_cache()
    return out


def config_of(name, w, world, full, M, chroms):
    return {"workload": f"{name}: {w['desc']}", "n": w["n"], "snps_per_gpu": M * chroms, "permutations": w["P"], "chromosomes_per_gpu": chroms,
            "design": {"k": w["k"], "c": w["c"], "s": w["s"]},
            "arithmetic": "exact int8 route: genotypes travel as 2-bit codes and are expanded to int8 in the SM, tcgen05.mma kind::i8 on 6 base-256 digit planes of V^-1 (48 bits below a power-of-two scale per 16 rows), s32 accumulation, FP64 everywhere else; FP64 DMMA re-fit of near-collinear SNPs",
            "sharding": ("chromosomes per GPU (LOCO), " if chroms > 1 else "SNP blocks per GPU, ") + "NCCL min-allreduce of min-p" if world > 1 else "single GPU",
            "l2_policy": "no flush needed: each step streams the packed genotypes (C2: 1.25 GB) and, per batch of 75,776 SNPs, 2-bit code tiles (C2: 97 MB, re-read for each of the 20 row-tile rounds) + digit planes (C2: 79 MB): above the 126 MB L2"}


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--full", action="store_true", help="per-GPU SNP count of the BASELINE config at 8 GPUs instead of the default slice")
    ap.add_argument("--shapes", default="auto", help="comma list of extra BASELINE shapes measured after the headline (auto: C3,C4,C5 with the default C2 workload on one GPU; none)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--route", default="auto", choices=["auto", "fp64"], help="fp64: force the FP64 DMMA route (FLX_FLAG_NO_INT8)")
    ap.add_argument("--cpu-snps", type=int, default=0, help="SNPs per core for the CPU baseline sample (0 = auto)")
    a = ap.parse_args()
    w = WORKLOADS[a.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    M = w["M_full"] if a.full else w["M"]
    chroms = w.get("chroms_full", w["chroms"]) if a.full else w["chroms"]
    config = config_of(a.workload, w, world, a.full, M, chroms)

    if a.impl == "reference":
        if rank != 0:
            return 0
        # each "step" = a bounded sample of the workload on all host cores, ~3 s of CPU work per core
        snps = a.cpu_snps or cpu_sample_snps(w["n"], 3.0)
        t0 = time.perf_counter()
        arm = CpuArm(w, cores, snps)
        t_build = time.perf_counter() - t0
        vals = []
        for i in range(a.warmup + a.steps):
            rate, tmax = arm.step()
            if i >= a.warmup:
                vals.append((rate, tmax))
        arm.close()
        tests = cores * snps * arm.tasks
        tmean = float(np.mean([v[1] for v in vals]))
        rate = tests / tmean
        print(json.dumps({"impl": "reference", "metric": "SNP-tests/sec incl. permutations", "value": rate, "unit": "SNP-tests/s",
                          "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": tmean * 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": rate, "unit": "SNP-tests/s", "cores": cores, "kind": "port", "sample": arm.sample},
                          "task_build_s": t_build,
                          "e2e": {"value": rate, "unit": "SNP-tests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    ctx = Ctx()
    res = run_workload(ctx, a.workload, w, a.steps, a.warmup, route=a.route, full=a.full, clocks=True)

    shapes = {}
    # auto: the extra shapes ride on the single-GPU headline run only — under torchrun a failure of ONE rank inside an extra shape would leave the
    # others in a collective and take the headline line with it (every shape has its own multi-GPU line: --workload C3|C4|C5 --full)
    want = [] if a.shapes == "none" else (["C3", "C4", "C5"] if (a.shapes == "auto" and a.workload == "C2" and a.route == "auto" and world == 1) else
                                          [x for x in a.shapes.split(",") if x in WORKLOADS and a.shapes != "auto"])
    for nm in want:
        t0 = time.perf_counter()
        try:
            r = run_workload(ctx, nm, WORKLOADS[nm], steps=2, warmup=1, route=a.route, full=False, clocks=False)
            r["config"] = config_of(nm, WORKLOADS[nm], world, False, WORKLOADS[nm]["M"], WORKLOADS[nm]["chroms"])
            r["steps"], r["warmup"], r["unit"] = 2, 1, "SNP-tests/s"
            shapes[nm] = r
        except Exception as e:   # a failure in an extra shape must not take the headline line with it
            shapes[nm] = {"error": f"{type(e).__name__}: {e}"[:500]}
            torch.cuda.empty_cache()
        shapes[nm]["bench_s"] = time.perf_counter() - t0

    if rank == 0:
        line = {"metric": "SNP-tests/sec incl. permutations", "value": res["value"], "unit": "SNP-tests/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64 (dominant contraction: int8 x int8 -> s32 on 6 exact digit planes = 48-bit fixed point)" if res["int8_path"] else "f64",
                "data": "synthetic", "config": config}
        for kk in ("e2e", "gpu_launches", "roofline", "int8_path", "fallback_reason", "refit_snps_per_step", "stage_ms_per_step", "stage_roofline", "device_ms_per_step",
                   "lib_wall_ms_per_step", "setup_s_once", "input_generation_s", "clocks", "min_pval_head", "loco_job", "loco_setup_s_per_factor", "loco_first_run_s_per_factor"):
            if kk in res:
                line[kk] = res[kk]
        line["timing"] = "value and e2e: wall clock around the K timed flx_run calls, barrier + cudaDeviceSynchronize on both sides, max over ranks; device_ms_per_step: CUDA events on the library's stream"
        if shapes:
            line["shapes"] = shapes
        if not a.no_cpu_baseline and world == 1:
            snps = a.cpu_snps or cpu_sample_snps(w["n"], 12.0)
            arm = CpuArm(w, cores, snps)
            arm.step()                                   # warm the caches / page in
            rate, tmax = arm.step()
            arm.close()
            line["cpu_baseline"] = {"value": rate, "unit": "SNP-tests/s", "cores": cores, "kind": "port", "sample": arm.sample + f" ({tmax:.1f} s)"}
        print(json.dumps(line))
    if ctx.dist is not None:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
