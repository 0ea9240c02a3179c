"""One process per GPU: split a gap set, fill, and bring the filled sequences together.

Gaps are independent (figbird.cpp:2819-2855 fills them one after another), so the data path needs no collective:
every rank registers the whole gap set, FigbirdContext.setPartition deals the gaps (longest-processing-time first on
the library's cost estimate) and each rank fills its share on its own GPU.  The only exchange is the gather of the
small per-gap results, done here with torch.distributed (NCCL between GPUs, gloo in the CPU tests).
"""
import torch.distributed as dist


def gather_results(local_results, group=None):
    """local_results: {gap_index: result dict} of the gaps this rank owns -> the merged dict of all ranks, on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(local_results)
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, dict(local_results), group=group)
    merged = {}
    for rank, part in enumerate(parts):
        for k, v in part.items():
            if k in merged:
                raise RuntimeError(f"gap {k} was filled by two ranks (second: {rank})")
            merged[k] = v
    return merged


def merge_liks(parts):
    """Per-estimate likelihood arrays of all ranks (NaN = not scanned there) -> one array; ranks must agree where they overlap."""
    import numpy as np
    out = np.array(parts[0], dtype=np.float64, copy=True)
    for p in parts[1:]:
        p = np.asarray(p, dtype=np.float64)
        take = np.isnan(out) & ~np.isnan(p)
        both = ~np.isnan(out) & ~np.isnan(p)
        if np.any(out[both] != p[both]):
            raise RuntimeError("two ranks computed different likelihoods for one estimate")
        out[take] = p[take]
    if np.any(np.isnan(out)):
        raise RuntimeError("an estimate was scanned by no rank")
    return out


def fill_partitioned(ctx, gap_fillers, rank, world, group=None, by="gap", device=None, **shard_kw):
    """Partition + fill + gather: returns {gap_index: result dict} for ALL gaps on every rank.

    by="gap": gaps are dealt to ranks (fb_set_partition); no exchange before the final gather.
    by="estimate": (gap, estimate) units are dealt (fb_set_partition_units) — for gap sets dominated by few gaps; the
    per-estimate likelihoods (a few doubles per gap) are all-gathered between the scan and the finalize.
    by="record": every rank scans every unit over its block of each gap's records; the partial counts are exchanged
    every EM iteration (RecordShardSession; results agree with the other two to rounding, not to the bit).
    device: the rank's torch CUDA device when the process group is NCCL (control scalars and the likelihood exchange
    then live on the GPU); None under gloo."""
    if world == 1:
        ctx.fillGaps()
        return {i: gf.result() for i, gf in enumerate(gap_fillers)}
    if by == "gap":
        ctx.setPartition(rank, world)
        ctx.fillGaps()
        local = {i: gf.result() for i, gf in enumerate(gap_fillers) if ctx.ownsGap(i)}
        return gather_results(local, group)
    if by == "record":
        with RecordShardSession(ctx, rank, world, group=group, device=device, **shard_kw) as sess:
            sess.scan()
            ctx._last_exchange = sess.exchange
        ctx.runFinalize()
        local = {i: gf.result() for i, gf in enumerate(gap_fillers) if ctx.hasResult(i)}
        return gather_results(local, group)
    ctx.setPartitionUnits(rank, world)
    ctx.upload()
    ctx.runScan()
    exchange_liks(ctx, len(gap_fillers), group, device=device)
    ctx.runFinalize()
    local = {i: gf.result() for i, gf in enumerate(gap_fillers) if ctx.hasResult(i)}
    return gather_results(local, group)


def exchange_liks(ctx, n_gaps, group=None, device=None):
    """The one collective of the estimate-level partition: every rank holds the likelihoods of the estimates it scanned
    (NaN elsewhere); one all-reduce(MAX) over the concatenated vector (NaN -> -inf) gives every rank all of them.  Each
    estimate is scanned by exactly one rank, so MAX only ever picks between a value and -inf: the result is exact.
    NCCL over NVLink when the ranks sit on GPUs (pass device), gloo on CPU."""
    import numpy as np
    import torch
    mine = [ctx.scanLiks(i) for i in range(n_gaps)]
    flat = np.concatenate(mine) if mine else np.zeros(0)
    t = torch.from_numpy(np.where(np.isnan(flat), -np.inf, flat))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    merged = t.cpu().numpy()
    if np.any(np.isneginf(merged)):
        raise RuntimeError("an estimate was scanned by no rank")
    off = 0
    for i, m in enumerate(mine):
        ctx.mergeLiks(i, merged[off:off + m.size])
        off += m.size


class _DeviceDoubles:
    """A device buffer of the library seen through the CUDA array interface (so torch can wrap it without a copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class RecordShardSession:
    """The exchange of a record-sharded scan (fb_set_partition_records, BASELINE config 3).

    Every rank runs every (gap, estimate) unit over its block of the gap's records, one kernel launch per EM
    iteration; the launch of iteration k+1 starts each unit by adding up, in rank order, the partial counts all ranks
    wrote in iteration k.
    exchange="p2p": the ranks map each other's buffers with CUDA IPC; the kernel loads the peers' partial counts
      straight over NVLink and the iterations are separated by a device-side barrier (flags the ranks store into each
      other's memory) — the whole scan is enqueued at once, no collective and no host round trip inside it.  Needs one
      GPU per rank.
    exchange="p2p-host": the same mappings, but the host synchronises and runs a barrier after every iteration (works
      with several ranks on one GPU, which is how the single-GPU tests run it).
    exchange="nccl": the partial counts are all-gathered with NCCL into a local buffer the kernel reads.
    exchange="auto": p2p under an NCCL process group (one GPU per rank), p2p-host under gloo; nccl when the IPC mapping
      is refused.
    The session keeps the mappings for repeated scans of the same uploaded gap set."""

    def __init__(self, ctx, rank, world, group=None, device=None, exchange="auto"):
        import torch
        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.num_iter = ctx.model.num_iter
        ctx.setPartitionRecords(rank, world)
        ctx.upload()
        self.nc, self.nm = ctx.shardSizes()
        self.own = ctx.shardBuffers()                       # counts slot 0, counts slot 1, maxima, flags
        self._opened = []
        if world == 1:
            self.exchange = "p2p"
            self.peers = [[p] for p in self.own]
            return
        nccl_group = dist.get_backend(group) == "nccl"
        if exchange == "auto":
            exchange = "p2p" if nccl_group else "p2p-host"
            fallback = nccl_group
        else:
            fallback = False
        if exchange == "p2p" and not nccl_group:
            raise RuntimeError("exchange='p2p' needs one GPU per rank (an NCCL process group); use 'p2p-host'")
        self.exchange = exchange
        if exchange in ("p2p", "p2p-host") and not self._map_peers():
            if not fallback:
                raise RuntimeError("CUDA IPC mapping of the peers' exchange buffers failed")
            self.exchange = "nccl"
        if self.exchange == "nccl":
            if not nccl_group:
                raise RuntimeError("exchange='nccl' needs an NCCL process group")
            self._own_t = [torch.as_tensor(_DeviceDoubles(p, n), device=self.device)
                           for p, n in zip(self.own[:3], (self.nc, self.nc, self.nm))]
            self._gath_c = torch.empty(world * self.nc, dtype=torch.float64, device=self.device)
            self._gath_m = torch.empty(world * self.nm, dtype=torch.float64, device=self.device)

    def _map_peers(self):
        import torch
        handles = [self.ctx.ipcExport(p) for p in self.own]
        allh = [None] * self.world
        dist.all_gather_object(allh, handles, group=self.group)
        peers, ok = [[None] * self.world for _ in range(4)], 1
        try:
            for q in range(self.world):
                for k in range(4):
                    if q == self.rank:
                        peers[k][q] = self.own[k]
                    else:
                        peers[k][q] = self.ctx.ipcOpen(allh[q][k])
                        self._opened.append(peers[k][q])
        except Exception:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self._ctl_device())
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)      # all ranks take the same path
        if int(flag.item()) == 0:
            self._close_mapped()
            return False
        self.peers = peers
        return True

    def _barrier(self):
        """Every rank's launches of this iteration have finished (shardScanIter returns after the device has)."""
        if self.world == 1:
            return
        import torch
        t = torch.zeros(1, dtype=torch.int32, device=self._ctl_device())
        dist.all_reduce(t, group=self.group)
        t.item()

    def _ctl_device(self):
        """Where the control-plane scalars live: on the GPU under NCCL, on the host under gloo (CPU tests, or several
        ranks sharing one GPU, which NCCL refuses)."""
        return self.device if dist.get_backend(self.group) == "nccl" else "cpu"

    def scan(self):
        """All EM iterations + the closing reduction; leaves ctx ready for runFinalize() with identical likelihoods
        and consensus strings on every rank."""
        import torch
        ctx = self.ctx
        if self.exchange == "p2p" and self.world > 1:
            # A rank whose barrier wait gave up returns an error; a rank that had already finished its pass may not
            # know yet.  The status is reduced over the ranks before anybody uses results, and a failed pass is void
            # on every rank: flags are reset everywhere and all ranks raise.
            failed, err = 0, None
            try:
                ctx.shardScanAll(*self.peers)
            except Exception as e:       # noqa: BLE001 - the status goes through the all-reduce below
                failed, err = 1, e
            t = torch.tensor([failed], dtype=torch.int32, device=self._ctl_device())
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            if int(t.item()):
                torch.cuda.synchronize(self.device)
                self._barrier()
                ctx.shardReset()
                self._barrier()
                raise RuntimeError(f"record-sharded scan failed on at least one rank (this rank: {err or 'ok'}); the pass is void")
            return
        if self.exchange in ("p2p", "p2p-host"):
            for it in range(self.num_iter):
                ctx.shardScanIter(it, self.peers[(it - 1) & 1] if it else None, self.own[it & 1])
                self._barrier()
            ctx.shardScanFinish(self.peers[(self.num_iter - 1) & 1], self.peers[2])
            self._barrier()       # nobody rewrites a buffer a slower rank is still reading
            return
        rows_c = [self._gath_c.data_ptr() + 8 * q * self.nc for q in range(self.world)]
        rows_m = [self._gath_m.data_ptr() + 8 * q * self.nm for q in range(self.world)]
        for it in range(self.num_iter):
            ctx.shardScanIter(it, rows_c if it else None, self.own[it & 1])
            if self.nc:
                dist.all_gather_into_tensor(self._gath_c, self._own_t[it & 1], group=self.group)
            torch.cuda.synchronize(self.device)
        if self.nm:
            dist.all_gather_into_tensor(self._gath_m, self._own_t[2], group=self.group)
        torch.cuda.synchronize(self.device)
        ctx.shardScanFinish(rows_c, rows_m)

    def _close_mapped(self):
        for p in self._opened:
            try:
                self.ctx.ipcClose(p)
            except Exception:
                pass
        self._opened = []

    def close(self):
        if self._opened:
            self._barrier()
            self._close_mapped()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def scan_record_sharded_local(ctxs):
    """Record-sharded scan with all shards in ONE process (contexts on the same or on peer-accessible devices): the
    contexts hand each other their buffers directly.  ctxs[r] must hold the same gaps, partitioned as rank r of len(ctxs)."""
    world = len(ctxs)
    for r, c in enumerate(ctxs):
        c.setPartitionRecords(r, world)
        c.upload()
    bufs = [c.shardBuffers() for c in ctxs]
    peers = [[bufs[q][k] for q in range(world)] for k in range(3)]      # counts slot 0, counts slot 1, maxima
    num_iter = ctxs[0].model.num_iter
    for it in range(num_iter):
        for r, c in enumerate(ctxs):
            c.shardScanIter(it, peers[(it - 1) & 1] if it else None, bufs[r][it & 1])
    for c in ctxs:
        c.shardScanFinish(peers[(num_iter - 1) & 1], peers[2])
