"""The N>1 path on CPU: world_size-2 gloo run of the sharding + min-p reduction logic.  Each rank computes its
SNP shard with the oracle (no GPU here); the min-allreduce over ranks must equal the single-process minimum, per-SNP
outputs must concatenate in var_idx order, and nvars must sum."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from flexlmm_b200 import synth
    from flexlmm_b200.sharding import shard_bounds, allreduce_minp, gather_per_snp
    from oracle import oracle as orc
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    t = synth.task(60, 21, model="additive", P=3, seed=9)
    codes = t["codes_raw"]
    lo, hi = shard_bounds(21, rank, world)
    args = (t["L"], t["y_mm"], t["X_null"], t["ll_null"], t["df_null"], t["M0"], t["M1"], t["M2"], codes[lo:hi])
    o = orc.fit_model(*args)
    local = np.array([orc.fit_model(*args, perm_seed=int(s), y_pred=t["y_pred"], U1=t["U1"], e_p=t["e_p"])["min_pval"] for s in t["seeds"]])
    mp_all, nv = allreduce_minp(local, hi - lo, dist)
    chisq = gather_per_snp(o["chisq"], 21, rank, world, dist)
    if rank == 0:
        q.put((mp_all, nv, chisq))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_minp_allreduce_matches_single_process():
    sys.path.insert(0, ROOT)
    from flexlmm_b200 import synth
    from oracle import oracle as orc
    orc.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    mp_all, nv, chisq = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    t = synth.task(60, 21, model="additive", P=3, seed=9)
    args = (t["L"], t["y_mm"], t["X_null"], t["ll_null"], t["df_null"], t["M0"], t["M1"], t["M2"], t["codes_raw"])
    full = orc.fit_model(*args)
    want = np.array([orc.fit_model(*args, perm_seed=int(s), y_pred=t["y_pred"], U1=t["U1"], e_p=t["e_p"])["min_pval"] for s in t["seeds"]])
    assert nv == 21
    np.testing.assert_array_equal(mp_all, want)          # min is exact
    np.testing.assert_array_equal(chisq, full["chisq"])  # per-SNP outputs in var_idx order
