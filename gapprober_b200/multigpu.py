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


def fill_partitioned(ctx, gap_fillers, rank, world, group=None, by="gap"):
    """Partition + fill + gather: returns {gap_index: result dict} for ALL gaps on every rank.

    by="gap": gaps are dealt to ranks (fb_set_partition); no exchange before the final gather.
    by="estimate": (gap, estimate) units are dealt (fb_set_partition_units) — for gap sets dominated by few gaps; the
    per-estimate likelihoods (a few doubles per gap) are all-gathered between the scan and the finalize."""
    if world == 1:
        ctx.fillGaps()
        return {i: gf.result() for i, gf in enumerate(gap_fillers)}
    if by == "gap":
        ctx.setPartition(rank, world)
        ctx.fillGaps()
        local = {i: gf.result() for i, gf in enumerate(gap_fillers) if ctx.ownsGap(i)}
        return gather_results(local, group)
    ctx.setPartitionUnits(rank, world)
    ctx.upload()
    ctx.runScan()
    exchange_liks(ctx, len(gap_fillers), group)
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
