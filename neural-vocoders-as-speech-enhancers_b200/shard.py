"""Per-utterance sharding of an inference job across the GPUs of one box.

Utterances are independent: weights are replicated, each rank vocodes
its own share, and there is no collective on the data path.  These helpers are pure host
logic (tested with a world_size-2 gloo group on CPU)."""
from __future__ import annotations

from typing import List, Sequence


def shard_range(n_items: int, world_size: int, rank: int) -> range:
    """Contiguous, balanced partition: the first ``n_items % world_size`` ranks get one extra."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    base, extra = divmod(max(0, n_items), world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def shard_by_cost(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Ragged utterances: longest-first greedy binning by cost (mel frames).  Deterministic:
    ties go to the lowest rank; every index appears exactly once."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world_size
    bins: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda q: (loads[q], q))
        bins[r].append(i)
        loads[r] += costs[i]
    for b in bins:
        b.sort()
    return bins


def bucket_by_length(lengths: Sequence[int], max_batch: int) -> List[List[int]]:
    """Ragged utterances on ONE GPU: groups of utterances of exactly equal length, at most
    ``max_batch`` each, longest first.  Equal-length groups need no padding, so every utterance
    comes out bit-identical to vocoding it alone (the reference's loop, infers/inference_hifigan.py:67),
    whereas padding would change the last receptive field of the shorter ones."""
    if max_batch < 1:
        raise ValueError("max_batch must be >= 1")
    by_len = {}
    for i, n in enumerate(lengths):
        by_len.setdefault(int(n), []).append(i)
    out: List[List[int]] = []
    for n in sorted(by_len, reverse=True):
        idx = by_len[n]
        out += [idx[s:s + max_batch] for s in range(0, len(idx), max_batch)]
    return out


def bucket_padded(lengths: Sequence[int], max_batch: int, max_pad: float = 0.25) -> List[List[int]]:
    """Ragged utterances on ONE GPU with padding: utterances sorted by length (longest first) are cut into groups of at most
    ``max_batch`` whose shortest member is at least ``(1 - max_pad)`` of the longest, so that at most ``max_pad`` of a
    group's rows are padding.  The kernels mask every utterance at its own length (``nvse_generator_forward_ragged``), so
    each result is still bit-identical to vocoding the utterance alone.  Deterministic; every index appears exactly once."""
    if max_batch < 1 or not 0.0 <= max_pad < 1.0:
        raise ValueError("max_batch must be >= 1 and 0 <= max_pad < 1")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    out: List[List[int]] = []
    cur: List[int] = []
    for i in order:
        if cur and (len(cur) >= max_batch or int(lengths[i]) < (1.0 - max_pad) * int(lengths[cur[0]])):
            out.append(cur)
            cur = []
        cur.append(i)
    if cur:
        out.append(cur)
    return out


def allreduce_gradients(parameters, group=None, average=True):
    """Data-parallel training (SURVEY.md §8e "collective, training only"): ONE all-reduce over a single flat buffer
    of every gradient (13.9 M fp32 = 55.7 MB for HiFi-GAN V1; on NVSwitch the cost is launch latency, not links, so
    there is no bucketing), then the mean is scattered back into each ``p.grad``.  Call between ``backward()`` and
    ``optimizer.step()``; parameters without a gradient are skipped (every rank must skip the same ones).
    NCCL on GPUs, gloo in the CPU tests.  Returns the number of elements reduced."""
    import torch
    import torch.distributed as dist
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads or not dist.is_available() or not dist.is_initialized():
        return 0
    world = dist.get_world_size(group)
    if world == 1:
        return 0
    flat = torch.cat([g.reshape(-1).to(torch.float32) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.mul_(1.0 / world)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return off
