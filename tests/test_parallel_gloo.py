"""Host-side logic of the multi-GPU path, exercised with world_size = 2 over gloo on CPU: shard
bookkeeping and the reduction of per-rank first-error words to the error the reference reports.
(The device kernels themselves run per rank with no collective; they are covered by -m gpu.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ndarray_interp_b200 import parallel as P


def test_shard_bounds_cover_everything_once():
    for total in (0, 1, 7, 1 << 20, (1 << 28) + 3):
        for world in (1, 2, 3, 8):
            spans = [P.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert P.column_shards(1024, 8)[3] == (384, 512)


def test_error_word_rebase_and_combine():
    assert P.global_error_word(P.ERR_NONE, 100) == P.ERR_NONE
    assert P.global_error_word(5, 100) == 105
    assert P.global_error_word(2 * 5 + 1, 100, two_d=True) == 2 * 105 + 1
    assert P.combine_error_words([P.ERR_NONE, 700, 512]) == 512
    assert P.combine_error_words([P.ERR_NONE, P.ERR_NONE]) == P.ERR_NONE


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q_all, lo_hi_expected, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g0, gl = 0.0, 1.0
        lo, hi = P.shard_bounds(len(q_all), world, rank)
        assert (lo, hi) == lo_hi_expected[rank]
        mine = q_all[lo:hi]
        # what the fused K7 reduction computes on a rank: smallest local index out of range
        bad = np.flatnonzero(~((mine >= g0) & (mine <= gl)))
        local = int(bad[0]) if len(bad) else P.ERR_NONE
        got = P.first_error(local, lo)
        # tables "replicated by broadcast": every rank must end with rank 0's tensor
        t = torch.arange(8, dtype=torch.float64) * (1 if rank == 0 else 0)
        dist.broadcast(t, 0)
        # column-sharded coefficients all-gathered back to full width
        w = 6
        c_lo, c_hi = P.column_shards(w, world)[rank]
        part = torch.arange(c_lo, c_hi, dtype=torch.float64).repeat(3, 1)
        parts = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(parts, part)
        full = torch.cat(parts, dim=1)
        results[rank] = (got, t.tolist(), full[0].tolist())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bad_positions", [[], [100, 900], [511, 512]])
def test_first_error_across_two_ranks(bad_positions):
    world = 2
    rng = np.random.default_rng(0)
    q = rng.uniform(0.0, 1.0, 1024)
    for p in bad_positions:
        q[p] = 2.0
    expected = [P.shard_bounds(len(q), world, r) for r in range(world)]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), q, expected, results), nprocs=world, join=True)
    want = (min(bad_positions), 0) if bad_positions else None
    for r in range(world):
        got, bcast, gathered = results[r]
        assert got == want                      # the reference's answer: first failing query overall
        assert bcast == list(map(float, range(8)))
        assert gathered == list(map(float, range(6)))
