"""SNP-block sharding over N GPUs (torchrun, one rank per GPU): the NCCL min-allreduce inside flx_run must give every rank the
min-p vector and the tested-variant count of the unsharded run, bit for bit.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/nccl_parity.py"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
import flexlmm_b200 as fx
from flexlmm_b200 import synth, sharding
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for model in ("additive", "domgxe"):
    n, M, P = 600, 40000, 8
    t = synth.task(n, M, model=model, P=P, seed=7, effect=0.05)      # same task on every rank
    vi = np.arange(1, M + 1, dtype=np.int32)
    def handle(comm):
        fm = fx.FitModel(device=local)
        fm.set_whitening(t["L"]); fm.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"]); fm.set_design(t["M0"], t["M1"], t["M2"])
        fm.set_perm(t["y_pred"], t["U1"], t["e_p"], t["seeds"]); fm.set_packed(t["packed"], t["n_raw"], t["pgen_order"])
        if comm:
            uid = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                buf = np.zeros(128, dtype=np.uint8)
                fx._capi.check(fx.load().flx_comm_unique_id(buf.ctypes.data))
                uid = torch.from_numpy(buf)
            uid = uid.cuda(); dist.broadcast(uid, 0)
            fm.comm_attach(uid.cpu().numpy(), rank, world)
        return fm
    fm = handle(True)
    mine = sharding.shard_var_idx(vi, rank, world)
    r = fm.run(mine)
    fm.close()
    ref = handle(False).run(vi)
    lo, hi = sharding.shard_bounds(M, rank, world)
    ok = (np.array_equal(r["min_pval"], ref["min_pval"]) and r["nvars_tested"] == M and np.array_equal(r["lrt_chisq"], ref["lrt_chisq"][lo:hi]))
    flag = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{model}: {world} ranks, shard sizes {[sharding.shard_bounds(M, q, world)[1] - sharding.shard_bounds(M, q, world)[0] for q in range(world)]}, "
              f"min_pval / nvars / per-SNP slices identical to the unsharded run on every rank: {bool(flag.item())}", flush=True)
    assert ok
dist.destroy_process_group()
