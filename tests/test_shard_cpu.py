"""Per-utterance sharding: pure host logic + a world_size-2 gloo run on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import pkg


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 128, 1024, 1025):
        for ws in (1, 2, 3, 4, 8):
            seen = []
            for r in range(ws):
                rg = pkg.shard_range(n, ws, r)
                seen += list(rg)
                assert abs(len(rg) - n / ws) < 1
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        pkg.shard_range(4, 2, 2)


def test_shard_by_cost_balances_ragged_lengths():
    costs = [862, 100, 400, 862, 50, 300, 700, 20, 500, 90]
    bins = pkg.shard_by_cost(costs, 4)
    assert sorted(i for b in bins for i in b) == list(range(len(costs)))
    loads = [sum(costs[i] for i in b) for b in bins]
    assert max(loads) - min(loads) <= max(costs)
    assert bins == pkg.shard_by_cost(costs, 4)  # deterministic


def test_bucket_by_length_groups_equal_lengths_only():
    lengths = [22050, 4000, 22050, 4000, 4000, 513, 22050, 4000]
    groups = pkg.bucket_by_length(lengths, 2)
    assert sorted(i for g in groups for i in g) == list(range(len(lengths)))
    for g in groups:
        assert len(g) <= 2 and len({lengths[i] for i in g}) == 1
    assert [lengths[g[0]] for g in groups] == sorted((lengths[g[0]] for g in groups), reverse=True)
    assert pkg.bucket_by_length([], 4) == []
    with pytest.raises(ValueError):
        pkg.bucket_by_length([1], 0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_utts):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = pkg.shard_range(n_utts, world, rank)
        # stand-in for "vocode my utterances": a per-utterance checksum that only depends on the index
        local = torch.tensor([float(i * i + 1) for i in mine], dtype=torch.float64)
        count = torch.tensor([len(mine)], dtype=torch.int64)
        dist.all_reduce(count)                      # only bookkeeping crosses ranks, never audio
        assert int(count) == n_utts
        gathered = [None] * world
        dist.all_gather_object(gathered, (list(mine), local.tolist()))
        order = [i for idx, _ in gathered for i in idx]
        vals = [v for _, vs in gathered for v in vs]
        assert order == list(range(n_utts))
        assert vals == [float(i * i + 1) for i in range(n_utts)]   # identical to the 1-rank result
        t = torch.tensor([1.0 + rank])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)   # bench.py's max-over-ranks timing reduction
        assert float(t) == float(world)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    mp.spawn(_worker, args=(2, _free_port(), 37), nprocs=2, join=True)


def _grad_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Conv1d(4, 8, 3), torch.nn.Conv1d(8, 1, 3))
        frozen = net[1].bias
        frozen.requires_grad_(False)
        for i, p in enumerate(net.parameters()):
            if p.requires_grad:
                p.grad = torch.full_like(p, float(rank + 1) * (i + 1))   # rank-dependent "local" gradient
        n = pkg.allreduce_gradients(list(net.parameters()))
        assert n == sum(p.numel() for p in net.parameters() if p.requires_grad)
        for i, p in enumerate(net.parameters()):
            if p.requires_grad:
                assert torch.equal(p.grad, torch.full_like(p, (i + 1) * (1 + world) / 2.0))   # the mean over ranks
        assert frozen.grad is None
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce():
    """Data-parallel training: one flat all-reduce of the generator gradients, mean written back per parameter."""
    mp.spawn(_grad_worker, args=(2, _free_port()), nprocs=2, join=True)


def test_allreduce_gradients_is_a_no_op_without_a_process_group():
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    assert pkg.allreduce_gradients([p]) == 0 and torch.equal(p.grad, torch.full((3,), 2.0))
