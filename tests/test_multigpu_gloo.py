"""The N>1 path on CPU: two gloo ranks split one gap set with the library's dealing rule (fb_partition_plan, the
rule fb_set_partition applies), fill their shares, and merge with gapprober_b200.multigpu.gather_results.  There is
no GPU here, so the ORACLE stands in for the fill inside this test only; what is under test is the partition and
the merge."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import golden_lib as G
    import oracle_lib as O
    from gapprober_b200 import multigpu
    from gapprober_b200.figbird import partition_plan
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    case = G.Case("multi_gap_err1pct")
    m = case.oracle_model()
    cost = [g["orig"] * max(1, len(g["recs"])) for g in case.gaps]
    owner = partition_plan(cost, world)
    local = {}
    for gi, g in enumerate(case.gaps):
        if owner[gi] == rank:
            r = O.fill_gap(m, g["window"], case.D, g["orig"], g["recs"])
            local[gi] = dict(filled=r["filled"], best_estimate=r["best_estimate"], liks=r["liks"])
    merged = multigpu.gather_results(local)
    ok = sorted(merged) == list(range(len(case.gaps)))
    for gi, g in enumerate(case.gaps):
        ok = ok and merged[gi]["filled"] == g["filled"] and merged[gi]["best_estimate"] == g["best_est"]
        ok = ok and np.array_equal(merged[gi]["liks"], g["liks"])
    q.put((rank, ok, len(local)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_and_gather(built):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert sum(n for _, _, n in res) == 5 and all(n > 0 for _, _, n in res)     # every gap filled once, both ranks used


def test_merge_liks_rules():
    import pytest
    from gapprober_b200.multigpu import merge_liks
    nan = float("nan")
    assert list(merge_liks([[1.0, nan, nan], [nan, 2.0, nan], [nan, nan, -0.5]])) == [1.0, 2.0, -0.5]
    assert list(merge_liks([[1.0, 2.0], [1.0, nan]])) == [1.0, 2.0]
    with pytest.raises(RuntimeError):
        merge_liks([[1.0, nan], [nan, nan]])            # an estimate nobody scanned
    with pytest.raises(RuntimeError):
        merge_liks([[1.0, 2.0], [1.5, nan]])            # two ranks disagree


class _FakeCtx:
    """Stands in for FigbirdContext in the exchange test: holds per-gap likelihood arrays with NaN for foreign estimates."""

    def __init__(self, liks):
        self.liks = [l.copy() for l in liks]

    def scanLiks(self, i):
        return self.liks[i]

    def mergeLiks(self, i, merged):
        import numpy as np
        own = ~np.isnan(self.liks[i])
        assert np.array_equal(self.liks[i][own], merged[own])
        self.liks[i] = np.array(merged, copy=True)


def _exchange_worker(rank, world, port, q):
    sys.path.insert(0, os.path.dirname(HERE))
    from gapprober_b200 import multigpu
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    full = [rng.normal(-3000, 50, size=n) for n in (51, 66, 7)]
    mine = []
    for g in full:
        m = g.copy()
        m[np.arange(g.size) % world != rank] = np.nan        # each estimate scanned by exactly one rank
        mine.append(m)
    ctx = _FakeCtx(mine)
    multigpu.exchange_liks(ctx, len(full))
    q.put((rank, all(np.array_equal(a, b) for a, b in zip(ctx.liks, full))))
    dist.barrier()
    dist.destroy_process_group()


def test_likelihood_exchange_two_ranks():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res
