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


def fill_partitioned(ctx, gap_fillers, rank, world, group=None):
    """setPartition + fillGaps + gather: returns {gap_index: result dict} for ALL gaps on every rank."""
    if world > 1:
        ctx.setPartition(rank, world)
    ctx.fillGaps()
    local = {i: gf.result() for i, gf in enumerate(gap_fillers) if world == 1 or ctx.ownsGap(i)}
    return gather_results(local, group)
