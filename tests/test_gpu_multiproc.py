"""One process per rank on the GPU: the record-sharded scan with its real exchange.

Two (and three) processes map each other's exchange buffers with CUDA IPC and the EM kernel of iteration k+1 reads the
peers' partial counts directly.  On a box with one GPU the ranks share cuda:0 (gloo carries the handles and the
barriers; NCCL refuses two ranks on one device); with two or more GPUs every rank takes its own device, the process
group is NCCL and the all-gather variant of the exchange is exercised as well."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, exchange, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    try:
        import torch
        import torch.distributed as dist
        import golden_lib as G
        import oracle_lib as O
        from gapprober_b200 import FigbirdContext, multigpu
        ndev = torch.cuda.device_count()
        dev = rank if ndev >= world else 0
        torch.cuda.set_device(dev)
        backend = "nccl" if ndev >= world else "gloo"
        dist.init_process_group(backend, init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        case = G.Case("multi_gap_err1pct")
        with FigbirdContext(case.figbird_model(), device=dev) as fb:
            gfs = [fb.GapFiller().initGapFiller(0, case.D, g["orig"], g["recs"], g["window"]) for g in case.gaps]
            merged = multigpu.fill_partitioned(fb, gfs, rank, world, by="record", exchange=exchange,
                                               device=torch.device("cuda", dev))
            liks = [fb.scanLiks(i) for i in range(len(case.gaps))]
        ok = sorted(merged) == list(range(len(case.gaps)))
        why = ""
        for gi, g in enumerate(case.gaps):
            if merged[gi]["filled"] != g["filled"] or merged[gi]["best_estimate"] != g["best_est"]:
                ok, why = False, f"gap {gi}: filled sequence differs"
            if not np.allclose(liks[gi], g["liks"], rtol=1e-9, atol=0):
                ok, why = False, f"gap {gi}: likelihoods differ"
            if rank == 0:     # the rank-ordered grouping, bit for bit (the oracle restates it)
                sh = O.fill_gap(case.oracle_model(), g["window"], case.D, g["orig"], g["recs"], shards=world)
                if not np.array_equal(liks[gi], sh["liks"]) or merged[gi]["filled"] != sh["filled"]:
                    ok, why = False, f"gap {gi}: differs from the sharded oracle"
        # every rank must hold the same likelihood bits (the rank-ordered sum is the same computation everywhere)
        allv = [None] * world
        dist.all_gather_object(allv, [x.tobytes() for x in liks])
        if any(v != allv[0] for v in allv):
            ok, why = False, "ranks disagree on likelihood bits"
        q.put((rank, ok, why, backend))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:   # report instead of hanging the parent on q.get
        q.put((rank, False, repr(e), "?"))


def _run(world, exchange):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, exchange, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.kill()
    assert all(ok for _, ok, _, _ in res), res
    return res


@pytest.mark.parametrize("world", [2, 3])
def test_record_sharded_ipc_exchange(built, world):
    """auto: device-side barrier with one GPU per rank, host barrier when the ranks share cuda:0."""
    _run(world, "auto")


def test_record_sharded_ipc_host_barrier(built):
    _run(2, "p2p-host")


def test_record_sharded_nccl_exchange(built):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("the NCCL all-gather variant needs one GPU per rank")
    res = _run(2, "nccl")
    assert all(b == "nccl" for _, _, _, b in res)
