"""Several GPUs from one caller (flx_multi_*), the two-process NCCL path (flx_comm_attach + flx_set_whitening_bcast), and the LOCO
driver.  The >= 2-GPU tests skip on a one-GPU box (run them with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`);
the one-device forms of the same entry points always run."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = ("lrt_chisq", "pval", "lrt_df", "rank", "pivot", "coef", "min_pval")


def _single(fx, t, vi, int8=True):
    with fx.FitModel(int8=int8) as fm:
        fm.set_whitening(t["L"]); fm.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"]); fm.set_design(t["M0"], t["M1"], t["M2"])
        fm.set_perm(t["y_pred"], t["U1"], t["e_p"], t["seeds"]); fm.set_packed(t["packed"], t["n_raw"], t["pgen_order"])
        return fm.run(vi)


def _multi(fx, t, vi, devices, implicit=False):
    with fx.FitModelMulti(devices) as fm:
        fm.set_whitening(t["L"]); fm.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"]); fm.set_design(t["M0"], t["M1"], t["M2"])
        if implicit:
            fm.set_perm_implicit(t["seeds"])
        else:
            fm.set_perm(t["y_pred"], t["U1"], t["e_p"], t["seeds"])
        fm.set_packed(t["packed"], t["n_raw"], t["pgen_order"])
        r = fm.run(vi)
        return r, [fm.stats(i) for i in range(len(devices))]


def test_multi_with_one_device_equals_the_plain_handle(fx):
    from flexlmm_b200 import synth
    t = synth.task(400, 333, model="domgxe", P=5, seed=61, n_raw=431)
    vi = np.arange(1, 334, dtype=np.int32)
    a = _single(fx, t, vi)
    b, st = _multi(fx, t, vi, [0])
    for f in FIELDS:
        assert np.array_equal(a[f], b[f], equal_nan=True), f
    assert b["nvars_tested"] == 333 and st[0]["snps"] == 333


def _ngpu(fx):
    return fx.load().flx_device_count()


def test_two_devices_one_caller_bit_identical(fx, orc):
    """flx_multi_run on 2 GPUs: SNP blocks per device, L^-1 and U1 broadcast over NVLink, min-p allreduced: every output bit for bit what
    one GPU returns, in var_idx order; and the oracle agrees."""
    if _ngpu(fx) < 2:
        pytest.skip("needs 2 GPUs")
    from flexlmm_b200 import synth
    t = synth.task(900, 2001, model="additive", P=7, seed=62)
    vi = np.arange(1, 2002, dtype=np.int32)
    a = _single(fx, t, vi)
    b, st = _multi(fx, t, vi, [0, 1])
    for f in FIELDS:
        assert np.array_equal(a[f], b[f], equal_nan=True), f
    assert b["nvars_tested"] == 2001 and st[0]["snps"] == 1001 and st[1]["snps"] == 1000
    c, _ = _multi(fx, t, vi[:1], [1, 0])                  # fewer SNPs than devices: one device gets an empty block
    assert np.array_equal(c["lrt_chisq"], a["lrt_chisq"][:1]) and c["nvars_tested"] == 1
    o = orc.fit_model(t["L"], t["y_mm"], t["X_null"], t["ll_null"], t["df_null"], t["M0"], t["M1"], t["M2"], t["codes_raw"][:50])
    np.testing.assert_allclose(b["lrt_chisq"][:50], o["chisq"], rtol=1e-8, atol=1e-9)


def test_two_devices_error_on_one_device_reaches_the_caller(fx):
    """A missing genotype in ONE device's SNP block: the call returns FLX_ERR_MISSING_GENOTYPE (nobody hangs in the allreduce)."""
    if _ngpu(fx) < 2:
        pytest.skip("needs 2 GPUs")
    from flexlmm_b200 import synth
    t = synth.task(300, 400, model="additive", P=3, seed=63)
    t["codes_raw"][350, 5] = 3
    t["packed"] = synth.pack_2bit(t["codes_raw"])
    with pytest.raises(fx.FlxError) as e:
        _multi(fx, t, np.arange(1, 401, dtype=np.int32), [0, 1])
    assert e.value.code == -2


def _nccl_worker(rank, world, q_uid, q_out, bad_rank):
    sys.path.insert(0, ROOT)
    import flexlmm_b200 as fx
    from flexlmm_b200 import synth
    from flexlmm_b200.sharding import shard_bounds
    t = synth.task(500, 1000, model="additive", P=4, seed=64)
    if bad_rank >= 0:
        lo, hi = shard_bounds(1000, bad_rank, world)
        t["codes_raw"][lo + 3, 7] = 3
        t["packed"] = synth.pack_2bit(t["codes_raw"])
    if rank == 0:
        uid = np.zeros(128, dtype=np.uint8)
        fx._capi.check(fx.load().flx_comm_unique_id(uid.ctypes.data))
        for _ in range(world - 1):
            q_uid.put(uid)
    else:
        uid = q_uid.get(timeout=120)
    fm = fx.FitModel(device=rank)
    fm.comm_attach(uid, rank, world)
    fm.set_whitening_bcast(t["L"] if rank == 0 else None, t["n"], root=0)          # only rank 0 uploads and inverts
    fm.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"]); fm.set_design(t["M0"], t["M1"], t["M2"])
    fm.set_perm_bcast(t["y_pred"], t["U1"] if rank == 0 else None, t["e_p"], t["seeds"], root=0)
    fm.set_packed(t["packed"], t["n_raw"], t["pgen_order"])
    lo, hi = shard_bounds(1000, rank, world)
    try:
        r = fm.run(np.arange(lo + 1, hi + 1, dtype=np.int32))
        q_out.put((rank, "ok", r["min_pval"], r["nvars_tested"], r["lrt_chisq"]))
    except fx.FlxError as e:
        q_out.put((rank, "err", e.code, str(e), None))
    fm.close()


@pytest.mark.parametrize("bad_rank", [-1, 1])
def test_two_processes_nccl_sharding(fx, bad_rank):
    """One process per GPU (what bench.py does under torchrun): sharded run == unsharded run bit for bit; with a bad record in rank 1's
    shard BOTH ranks come back with an error instead of rank 0 hanging in ncclAllReduce."""
    if _ngpu(fx) < 2:
        pytest.skip("needs 2 GPUs")
    import multiprocessing as mp
    from flexlmm_b200 import synth
    ctx = mp.get_context("spawn")
    q_uid, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, q_uid, q_out, bad_rank)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q_out.get(timeout=300) for _ in range(2)], key=lambda o: o[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if bad_rank < 0:
        t = synth.task(500, 1000, model="additive", P=4, seed=64)
        full = _single(fx, t, np.arange(1, 1001, dtype=np.int32))
        for o in outs:
            assert o[1] == "ok" and o[3] == 1000
            assert np.array_equal(o[2], full["min_pval"])
        assert np.array_equal(np.concatenate([outs[0][4], outs[1][4]]), full["lrt_chisq"])
    else:
        assert outs[0][1] == "err" and outs[1][1] == "err"
        assert outs[1][2] == -2                                        # the rank that saw the missing genotype says so
        assert outs[0][2] == -2 and "another rank" in outs[0][3]      # the other one learns it from the status exchange


def test_loco_driver_ping_pong_equals_fresh_handles(fx, orc):
    """lmm.nf:27-46: every chromosome has its own factor.  Three chromosomes through the LOCO driver (two handles, chromosome k+1 set
    up on a side thread while k runs) give bit for bit what three fresh handles give; genome-wide min over chromosomes."""
    from flexlmm_b200 import synth
    from flexlmm_b200.loco import LocoRunner, chromosomes_of_rank
    ts = [synth.task(600 + 40 * c, 700 + 100 * c, model="additive", P=4, seed=70 + c) for c in range(3)]

    def mk(t):
        def task(fm):
            fm.set_whitening(t["L"]); fm.set_null(t["X_null"], t["y_mm"], t["ll_null"], t["df_null"]); fm.set_design(t["M0"], t["M1"], t["M2"])
            fm.set_perm(t["y_pred"], t["U1"], t["e_p"], t["seeds"]); fm.set_packed(t["packed"], t["n_raw"], t["pgen_order"])
            return np.arange(1, t["M"] + 1, dtype=np.int32)
        return task

    fresh = [_single(fx, t, np.arange(1, t["M"] + 1, dtype=np.int32)) for t in ts]
    with LocoRunner(handles=2) as lr:
        res = lr.run([mk(t) for t in ts])
        assert len(lr.setup_s) == 3 and len(lr.run_s) == 3
        with pytest.raises(RuntimeError):
            lr.rerun()
    for a, b in zip(fresh, res):
        for f in FIELDS:
            assert np.array_equal(a[f], b[f], equal_nan=True), f
    assert np.array_equal(LocoRunner.genome_min_pval(res), np.minimum.reduce([r["min_pval"] for r in fresh]))
    with LocoRunner(handles=3) as lr:
        lr.run([mk(t) for t in ts], per_snp=False)
        again = lr.rerun()
    for a, b in zip(fresh, again):
        assert np.array_equal(a["lrt_chisq"], b["lrt_chisq"]) and np.array_equal(a["min_pval"], b["min_pval"])
    o = orc.fit_model(ts[2]["L"], ts[2]["y_mm"], ts[2]["X_null"], ts[2]["ll_null"], ts[2]["df_null"], ts[2]["M0"], ts[2]["M1"], ts[2]["M2"], ts[2]["codes_raw"][:40])
    np.testing.assert_allclose(res[2]["lrt_chisq"][:40], o["chisq"], rtol=1e-8, atol=1e-9)
    assert chromosomes_of_rank(22, 3, 8) == [3, 11, 19] and sorted(sum((chromosomes_of_rank(22, r, 8) for r in range(8)), [])) == list(range(22))
