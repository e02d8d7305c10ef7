"""Multi-GPU parity (needs >= 2 GPUs; skipped on a single-GPU box): one process per GPU, NCCL, time-range sharding.
Every rank must reproduce the oracle's LM trajectory and end with the full optimised trajectory."""
import os

import numpy as np
import pytest

from calico_b200 import _capi, synthetic

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, lib, cfg, chunk, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if chunk:
        os.environ["CB2_CHUNK_CPS"] = chunk
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oracle import oracle_py
    truth, prob = synthetic.generate(cfg, oracle_py.oracle_api, noise=True)
    a = _capi.CApi(lib)
    a.set_device(rank)
    pa = prob.clone()
    ids = pa.push(a)
    a.comm_init_torch(world, rank)
    # two calls with different iteration counts on the same handle: the ranks must stay in lockstep across calls
    summ0, log0 = a.optimize(_capi.Options(minimizer_progress_to_stdout=0, max_num_iterations=2))
    summ, log = a.optimize(_capi.Options(minimizer_progress_to_stdout=0))
    pa.pull(a, ids)
    out = {"rank": rank, "term": summ.termination_type, "costs": [it.cost for it in log], "ok": [it.step_is_successful for it in log],
           "costs0": [it.cost for it in log0],
           "ctrl": pa.spline.ctrl.copy(), "intr": [s.intr.copy() for s in pa.sensors], "launches": a.stats().kernel_launches,
           "res": [a.get_residuals(sid) for sid in ids]}
    if rank == 0:
        o = oracle_py.oracle_api()
        po = prob.clone()
        ido = po.push(o)
        so0, lo0 = o.optimize(oracle_py.OracleOptions(linear_solver=1, num_threads=os.cpu_count() or 1, max_num_iterations=2))
        so, lo = o.optimize(oracle_py.OracleOptions(linear_solver=1, num_threads=os.cpu_count() or 1))
        po.pull(o, ido)
        out["oracle"] = {"term": so.termination_type, "costs": [it.cost for it in lo], "ok": [it.step_is_successful for it in lo],
                         "costs0": [it.cost for it in lo0],
                         "ctrl": po.spline.ctrl.copy(), "intr": [s.intr.copy() for s in po.sensors], "res": [o.get_residuals(sid) for sid in ido]}
    q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("world,cfg,chunk", [(2, "small", None), (2, "small", "9"), (2, "tiny", "6"), (4, "small", "7"), (8, "C2", None)])
def test_multi_gpu_lm_matches_oracle(world, cfg, chunk, product_lib, oracle):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, product_lib, cfg, chunk, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = next(o for o in outs if "oracle" in o)["oracle"]
    for o in outs:
        assert o["launches"] > 0
        assert o["term"] == ref["term"]
        assert o["ok"] == ref["ok"]
        np.testing.assert_allclose(o["costs0"], ref["costs0"], rtol=1e-6)
        np.testing.assert_allclose(o["costs"], ref["costs"], rtol=1e-6)
        # every rank holds EVERY measurement's residual (Sensor::UpdateResiduals, camera.cpp:70-80), not just its shard's
        for (r1, v1), (r2, v2) in zip(o["res"], ref["res"]):
            assert (v1 == v2).all() and v1.all()
            np.testing.assert_allclose(r1, r2, rtol=1e-6, atol=1e-6 * max(1.0, np.abs(r2).max()))
        np.testing.assert_allclose(o["ctrl"], ref["ctrl"], rtol=1e-6, atol=1e-6)
        for a, b in zip(o["intr"], ref["intr"]):
            np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-9)


def _timeout_worker(rank, world, port, lib, q):
    """Rank 0 runs Optimize; rank 1 creates the communicator and then never joins a collective."""
    import time
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["CB2_COLLECTIVE_TIMEOUT_S"] = "4"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oracle import oracle_py
    truth, prob = synthetic.generate("tiny", oracle_py.oracle_api, noise=True)
    a = _capi.CApi(lib)
    a.set_device(rank)
    prob.clone().push(a)
    a.comm_init_torch(world, rank)
    if rank == 0:
        t0 = time.time()
        try:
            a.optimize(_capi.Options(minimizer_progress_to_stdout=0))
            q.put(("no error", time.time() - t0, ""))
        except _capi.CalicoError as e:
            q.put((e.code, time.time() - t0, str(e)))
    else:
        time.sleep(12)       # never enters the solve: rank 0's first collective has no partner
        q.put(("idle", 0.0, ""))
    q.close()
    q.join_thread()          # the queue's feeder thread must flush before the hard exit below
    os._exit(0)              # the aborted communicator cannot take part in a clean process-group shutdown


@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs")
def test_missing_rank_times_out_instead_of_hanging(product_lib, oracle):
    """SURVEY §5 (NCCL failure detection): a collective a peer never joins ends with status 13 after CB2_COLLECTIVE_TIMEOUT_S, the
    communicator aborted — not with a hang until an external watchdog kills the job (what happened to the round-1 N = 2 scaling run)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + 7) % 2000
    procs = [ctx.Process(target=_timeout_worker, args=(r, 2, port, product_lib, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=90) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    code, seconds, msg = next(o for o in outs if o[0] != "idle")
    print("rank 0:", code, "%.1f s" % seconds, msg)
    assert code == _capi.INTERNAL, (code, msg)
    assert "collective" in msg.lower() or "nccl" in msg.lower()
    assert 3.0 <= seconds <= 30.0
