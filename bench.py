#!/usr/bin/env python
"""bench.py — SNP-tests/s (including permutations) of the per-SNP LRT engine on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W     # the reference's R loop restated on the host cores (oracle port)

A "step" is one full pass of the hot path (decode -> whiten -> fit/LRT -> permutation contraction -> min-p) over the
workload's whole SNP set.  N=1 workload = BASELINE.json configs[1]: synthetic n=5,000 x 1M SNPs, additive
y ~ x + cov1 vs y ~ cov1, 100 permutations.  For N>1 every rank runs the same per-GPU workload on its own SNP shard
(weak scaling) and only the length-P min-p vector is min-allreduced with NCCL inside the library.

`value`  : SNP-tests/s with the packed genotypes already resident in HBM.
`e2e`    : the same metric through the C ABI with HOST buffers (packed records H2D and per-SNP results D2H inside the
           timed region).
`roofline`: the dominant kernel (K2 whiten_tile), algorithmic FLOPs (s*n^2 per SNP) over its CUDA-event time, against the
           cuBLAS DGEMM figure measured on this pool (profiles/fp64_peak.json).
`cpu_baseline`: the oracle port of the reference loop timed on this box's host cores on a bounded SNP sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

WORKLOADS = {
    # name: (n, M, model, P, s, c, k)
    "C2": dict(n=5000, M=1_000_000, model="additive", P=100, s=1, c=2, k=3,
               desc="synthetic n=5,000 x 1M SNPs, additive y ~ x + cov1 vs y ~ cov1, 100 permutations"),
    "C2-small": dict(n=5000, M=100_000, model="additive", P=100, s=1, c=2, k=3,
                     desc="synthetic n=5,000 x 100k SNPs (reduced M, NOT the headline config)"),
    "tiny": dict(n=1000, M=20_000, model="additive", P=20, s=1, c=2, k=3, desc="debug"),
    "A20k": dict(n=20000, M=100_000, model="additive", P=1000, s=1, c=2, k=3,
                 desc="synthetic n=20,000, additive, 1,000 permutations, 100k SNPs per GPU (int8 route at a larger n)"),
    "C3-1gpu": dict(n=20000, M=100_000, model="domgxe", P=1000, s=3, c=2, k=5,
                    desc="synthetic n=20,000, dominance + GxE, 1,000 permutations, 100k SNPs per GPU (per-GPU slice of configs[2])"),
}


def i8_peak():
    try:
        with open(os.path.join(HERE, "profiles", "i8_peak.json")) as f:
            d = json.load(f)
        return d["umma_i8_tops"], "measured tcgen05.mma kind::i8 M128 N256 on this pool (tools/qf_probe.cu, profiles/i8_peak.json)"
    except Exception:
        return 2 * 1657.4, "2 x MEASURED_PEAKS.json bf16_tflops (int8 tensor rate is twice bf16 on B200)"


def fp64_peak():
    try:
        with open(os.path.join(HERE, "profiles", "fp64_peak.json")) as f:
            d = json.load(f)
        return d["cublas_dgemm_8192_tflops_sustained"], "measured cuBLAS DGEMM 8192^3 on this pool (profiles/fp64_peak.json)"
    except Exception:
        return 37.0, "fallback: nominal B200 FP64 37 TFLOP/s"


# ------------------------------------------------------------------------------------------------ synthetic inputs
def make_inputs(w, rank, device):
    """Deterministic synthetic task.  Heavy pieces (genotypes, Cholesky, complement basis) are generated
    with torch on the GPU purely as input plumbing, then handed to the library as HOST arrays like R would."""
    import torch
    from flexlmm_b200 import synth

    n, M, P = w["n"], w["M"], w["P"]
    dev = torch.device("cuda", device)
    g = torch.Generator(device=dev)
    g.manual_seed(0xF1E + 7919 * rank)
    # genotypes: maf ~ U(0.05, 0.5), g ~ Binomial(2, maf), no missing; packed 2-bit, N_raw = n
    bpv = (n + 3) // 4
    packed = torch.empty((M, bpv), dtype=torch.uint8, device=dev)
    chunk = max(1, min(M, (1 << 28) // max(n, 1)))
    npad = bpv * 4
    for s0 in range(0, M, chunk):
        m = min(chunk, M - s0)
        maf = 0.05 + 0.45 * torch.rand((m, 1), generator=g, device=dev)
        a = (torch.rand((m, npad), generator=g, device=dev) < maf).to(torch.uint8)
        a += (torch.rand((m, npad), generator=g, device=dev) < maf).to(torch.uint8)
        if npad != n:
            a[:, n:] = 0
        a = a.view(m, bpv, 4)
        packed[s0:s0 + m] = a[:, :, 0] | (a[:, :, 1] << 2) | (a[:, :, 2] << 4) | (a[:, :, 3] << 6)
        del a, maf
    packed_host = torch.empty((M, bpv), dtype=torch.uint8, pin_memory=False)
    packed_host.copy_(packed)
    del packed
    # covariance factor, design, phenotype, null fit (same for every rank)
    g2 = torch.Generator(device=dev); g2.manual_seed(13)
    r = min(256, n)
    B = torch.randn((n, r), generator=g2, device=dev, dtype=torch.float64)
    V = 0.5 * (B @ B.T) / r + 0.5 * torch.eye(n, device=dev, dtype=torch.float64)
    L = torch.linalg.cholesky(V)
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
    e_p = (U1.T @ e)
    out = dict(n=n, M=M, P=P, packed=packed_host.numpy(), L=L.cpu().numpy(), y_mm=y_mm.cpu().numpy(),
               X_null=np.asfortranarray(X_mm.cpu().numpy()), ll_null=ll, df_null=c + 1, M0=M0, M1=M1, M2=M2, names=names,
               y_pred=fitted.cpu().numpy(), U1=np.asfortranarray(U1.cpu().numpy()), e_p=e_p.cpu().numpy(),
               seeds=np.arange(1, P + 1, dtype=np.int32))
    del Q, U1, V, B, L
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, device):
        self.device = device
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None

    def _loop(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.device)],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(o[0])); self.max_mhz = float(o[1])
                for nm, v in zip(names, o[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        self._t = threading.Thread(target=self._loop, daemon=True); self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU baseline (oracle port)
def _cpu_worker(args):
    (n, k, snps, tasks, seed) = args
    from oracle import oracle as orc
    from flexlmm_b200 import synth
    t = synth.task(n, snps, model="additive" if k == 3 else "domgxe", P=tasks - 1, seed=seed)
    codes = t["codes_raw"]
    t0 = time.perf_counter()
    orc.fit_model(t["L"], t["y_mm"], t["X_null"], t["ll_null"], t["df_null"], t["M0"], t["M1"], t["M2"], codes)
    for sd in range(1, tasks):
        orc.fit_model(t["L"], t["y_mm"], t["X_null"], t["ll_null"], t["df_null"], t["M0"], t["M1"], t["M2"], codes, perm_seed=sd,
                      y_pred=t["y_pred"], U1=t["U1"], e_p=t["e_p"])
    return time.perf_counter() - t0


def cpu_reference_rate(w, snps_per_core, tasks, cores):
    """The reference loop (fit_model.nf:96-171) restated in C (oracle), one independent single-threaded task per core like
    Nextflow's scatter of single-threaded R processes.  Returns SNP-tests/s over all cores and the wall time."""
    import multiprocessing as mp
    from oracle import oracle as orc
    orc.build()
    ctx = mp.get_context("spawn")
    args = [(w["n"], w["k"], snps_per_core, tasks, i) for i in range(cores)]
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        times = pool.map(_cpu_worker, args)
        wall = time.perf_counter() - t0
    tests = cores * snps_per_core * tasks
    return tests / max(times), max(times), wall


def stage_roofline(w, acc, steps, int8_path):
    """Every stage against the roofline that bounds it: algorithmic bytes or ops per step / CUDA-event time."""
    n, M, P, s, c = w["n"], w["M"], w["P"], w["s"], w["c"]
    n128 = (n + 127) // 128 * 128
    n256 = (n + 255) // 256 * 256
    hbm = 6546.2
    try:
        with open(os.path.join(HERE, "MEASURED_PEAKS.json")) as f:
            hbm = json.load(f)["hbm_gbs"]
    except Exception:
        pass
    out = {}

    def put(name, work, unit, ms, peak, bound):
        if ms <= 0:
            return
        ach = work / (ms * 1e-3 / steps) / (1e9 if unit == "GB/s" else 1e12)
        out[name] = {"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak, "ms_per_step": ms / steps}

    raw = (n + 3) // 4
    if int8_path:
        put("K1q decode_i8", M * (raw + n256), "GB/s", acc["ms_decode"], hbm, "hbm")
        passes = max(1, int(round(acc.get("int8_passes", steps) / steps)))
        put("K2q qf_i8 (g'C V^-1 C g, 6 planes x %d pass units)" % passes, 6.0 * n * n * M * passes, "TFLOP/s", acc["ms_whiten"], i8_peak()[0], "tensor(i8)")
        put("K4q (u, b, b* for all seeds, 6 planes) + K3b", 6.0 * 2 * n * (c + 1 + P) * M * s, "TFLOP/s", acc["ms_fit"], i8_peak()[0], "tensor(i8)")
        put("perm_reduce (b*' Sm b* / RSS_f*, max per seed)", 8.0 * M * s * P, "GB/s", acc["ms_perm"], hbm, "hbm")
    else:
        put("K1 decode_2bit_tile", M * (raw + 8.0 * n128 * s), "GB/s", acc["ms_decode"], hbm, "hbm")
        put("K2 whiten_tile (DMMA)", float(s) * n * n * M, "TFLOP/s", acc["ms_whiten"], fp64_peak()[0], "tensor(f64)")
        put("K3a snp_gram (2 sweeps) + K3b", M * (2 * 8.0 * n128 * s), "GB/s", acc["ms_fit"], hbm, "hbm")
        put("K4 perm_contract (DMMA)", 2.0 * n * s * P * M, "TFLOP/s", acc["ms_perm"], fp64_peak()[0], "tensor(f64)")
    return out


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--route", default="auto", choices=["auto", "fp64"], help="fp64: force the FP64 DMMA route (FLX_FLAG_NO_INT8)")
    ap.add_argument("--cpu-snps", type=int, default=0, help="SNPs per core for the CPU baseline sample (0 = auto)")
    a = ap.parse_args()
    w = WORKLOADS[a.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    tests_per_step_rank = w["M"] * (1 + w["P"])
    config = {"workload": f"{a.workload}: {w['desc']}", "n": w["n"], "snps_per_gpu": w["M"], "permutations": w["P"],
              "design": {"k": w["k"], "c": w["c"], "s": w["s"]}, "sharding": "SNP blocks per GPU, NCCL min-allreduce of min-p" if world > 1 else "single GPU",
              "l2_policy": "no flush needed: each step streams 1.25 GB of packed genotypes and, per batch of 75,776 SNPs, 388 MB of int8 tiles + 79 MB of digit planes (FP64 route: 2 x 775 MB tiles + 107 MB L^-1 per 18,944 SNPs), all above the 126 MB L2"}

    if a.impl == "reference":
        if rank != 0:
            return 0
        # each "step" = a bounded sample of the workload on all host cores
        auto = max(4, int(round(16 * (5000.0 / w["n"]) ** 2)))       # ~3 s of CPU work per step on every core
        snps = a.cpu_snps or auto
        tasks = 3  # 1 ORIG + 2 PERM tasks per SNP sample
        vals = []
        for i in range(a.warmup + a.steps):
            rate, tmax, wall = cpu_reference_rate(w, snps, tasks, cores)
            if i >= a.warmup:
                vals.append((rate, tmax))
        rate = float(np.mean([v[0] for v in vals])); tmax = float(np.mean([v[1] for v in vals]))
        sample = f"{snps} SNPs x (1 ORIG + {tasks - 1} PERM tasks) per core on {cores} cores, n={w['n']}, whole design re-whitened per SNP and per task"
        print(json.dumps({"impl": "reference", "metric": "SNP-tests/sec incl. permutations", "value": rate, "unit": "SNP-tests/s",
                          "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": tmax * 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": rate, "unit": "SNP-tests/s", "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": rate, "unit": "SNP-tests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import torch
    import flexlmm_b200 as fx
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    inp = make_inputs(w, rank, local_rank)
    fm = fx.FitModel(device=local_rank, int8=(a.route == "auto"))
    t_setup0 = time.perf_counter()
    fm.set_whitening(inp["L"])
    fm.set_null(inp["X_null"], inp["y_mm"], inp["ll_null"], inp["df_null"])
    fm.set_design(inp["M0"], inp["M1"], inp["M2"])
    fm.set_perm(inp["y_pred"], inp["U1"], inp["e_p"], inp["seeds"])
    fm.set_packed(inp["packed"], w["n"])
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = np.zeros(128, dtype=np.uint8)
            fx._capi.check(fx.load().flx_comm_unique_id(buf.ctypes.data))
            uid = torch.from_numpy(buf)
        uid = uid.cuda(); dist.broadcast(uid, 0); uid = uid.cpu().numpy()
        fm.comm_attach(uid, rank, world)
    var_idx = np.arange(1, w["M"] + 1, dtype=np.int32)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, per_snp):
        """returns (max-over-ranks seconds for nsteps, accumulated stats)"""
        acc = {}
        barrier()
        t0 = time.perf_counter()
        dev_ms = 0.0
        res = None
        for _ in range(nsteps):
            res = fm.run(var_idx, per_snp=per_snp, out=res)     # a streaming caller reuses its result buffers
            st = fm.stats()
            dev_ms += st["ms_total"]
            for kk, v in st.items():
                acc[kk] = acc.get(kk, 0) + v
        barrier()
        wall = time.perf_counter() - t0
        tt = torch.tensor([dev_ms * 1e-3, wall], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt[0]), float(tt[1]), acc, res

    # ---- e2e: host buffers, H2D + D2H inside the timed region
    timed(max(1, a.warmup), True)  # first call also does the once-per-task setup (L^-1, permutation prologue)
    t_setup = time.perf_counter() - t_setup0
    e2e_dev, e2e_wall, e2e_acc, _ = timed(a.steps, True)
    # ---- value: genotypes resident in HBM
    fm.make_resident()
    timed(max(3, a.warmup), False)
    clk = ClockSampler(local_rank)
    if rank == 0:
        clk.start()
    dev_s, wall_s, acc, res = timed(a.steps, False)
    clocks = clk.stop() if rank == 0 else None
    total_tests = tests_per_step_rank * world * a.steps
    value = total_tests / dev_s
    e2e_value = total_tests / e2e_wall

    if rank == 0:
        n = w["n"]
        ms_whiten = acc["ms_whiten"]
        int8_path = bool(acc.get("int8_path", 0))
        if int8_path:
            # K2q: Q = 6 digit planes x n(n+1)/2 int8 MACs per SNP (2 ops per MAC), on the i8 tensor pipe
            peak, peak_src = i8_peak()
            passes = max(1, int(round(acc.get("int8_passes", a.steps) / a.steps)))
            work = 6.0 * n * n * w["M"] * a.steps * passes
            kname = "qf_i8_kernel (K2q: g'C V^-1 C g, tcgen05.mma kind::i8, 6 exact digit planes, %d triangular-pass unit%s per batch)" % (passes, "" if passes == 1 else "s")
            tfile = "qf_traffic.json"
        else:
            peak, peak_src = fp64_peak()
            work = float(w["s"]) * n * n * w["M"] * a.steps          # algorithmic: s*n^2 per SNP (triangular solve)
            kname = "panel_gemm_kernel<0,1> (K2 whiten_tile, FP64 DMMA)"
            tfile = "whiten_traffic.json"
        achieved = work / (ms_whiten * 1e-3) / 1e12
        fp64_equiv = float(w["s"]) * n * n * w["M"] * a.steps / (ms_whiten * 1e-3) / 1e12
        traffic = None
        try:
            with open(os.path.join(HERE, "profiles", tfile)) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {"metric": "SNP-tests/sec incl. permutations", "value": value, "unit": "SNP-tests/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": dev_s / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "e2e": {"value": e2e_value, "unit": "SNP-tests/s", "h2d_bytes_per_step": e2e_acc["h2d_bytes"] / a.steps,
                        "d2h_bytes_per_step": e2e_acc["d2h_bytes"] / a.steps, "ms_per_step_wall": e2e_wall / a.steps * 1e3},
                "gpu_launches": int(acc["launches"]),
                "roofline": {"kernel": kname, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                             "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                             "launches": int(acc["launches_whiten"]), "avg_launch_ms": ms_whiten / max(1, acc["launches_whiten"]),
                             "algorithmic_ops_per_launch": work / max(1, acc["launches_whiten"]),
                             "fp64_equivalent_tflops": fp64_equiv, "fp64_dgemm_peak_tflops": fp64_peak()[0]},
                "int8_path": int8_path, "refit_snps_per_step": acc.get("refit_snps", 0) / a.steps,
                "stage_ms_per_step": {kk: acc[kk] / a.steps for kk in ("ms_decode", "ms_whiten", "ms_fit", "ms_perm", "ms_h2d", "ms_d2h")},
                "stage_roofline": stage_roofline(w, acc, a.steps, int8_path),
                "wall_ms_per_step": wall_s / a.steps * 1e3, "setup_s_once": t_setup, "clocks": clocks,
                "min_pval_head": [float(x) for x in res["min_pval"][:3]]}
        if not a.no_cpu_baseline and world == 1:
            auto = max(4, int(round(64 * (5000.0 / n) ** 2)))
            snps = a.cpu_snps or auto
            rate, tmax, wall = cpu_reference_rate(w, snps, 3, cores)
            line["cpu_baseline"] = {"value": rate, "unit": "SNP-tests/s", "cores": cores, "kind": "port",
                                    "sample": f"{snps} SNPs x (1 ORIG + 2 PERM tasks) per core on {cores} cores ({tmax:.1f} s), n={n}, reference loop order (whole design re-whitened per SNP and per task)"}
        print(json.dumps(line))
    fm.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
