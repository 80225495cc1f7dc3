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


def _data_worker(rank, world, port, results, info):
    """the multi-GPU data flow of bench.py on CPU, the oracle standing in for the kernels: spline built on a column
    shard (info: the build the kernels would report -- 0 reference order, -m partition blocks), coefficients
    all-gathered, queries evaluated per contiguous block, blocks gathered"""
    from oracle import oracle_py as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)                       # same tables on every rank ("replicated")
        n, w, nq = 200, 8, 1001
        x = np.cumsum(rng.uniform(0.5, 1.5, n))
        y = rng.normal(size=(n, w))
        q = np.sort(rng.uniform(x[0], x[-1], nq))
        q[700] = x[-1] + 1.0                                    # a failing query in the second block
        c_lo, c_hi = P.column_shards(w, world)[rank]
        st, a_sh, b_sh = O.spline_build_as(x, np.ascontiguousarray(y[:, c_lo:c_hi]), {"kind": "Natural"}, info)
        assert st == 0
        ga = [torch.empty((n - 1, c_hi - c_lo), dtype=torch.float64) for _ in range(world)]
        gb = [torch.empty((n - 1, c_hi - c_lo), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(ga, torch.from_numpy(a_sh))
        dist.all_gather(gb, torch.from_numpy(b_sh))
        a_full, b_full = torch.cat(ga, dim=1).contiguous().numpy(), torch.cat(gb, dim=1).contiguous().numpy()
        lo, hi = P.shard_bounds(nq, world, rank)
        st, out, bad = O.interp1d_cubic(x, y, a_full, b_full, q[lo:hi], 0, out=np.full((hi - lo, w), -1.0))
        local = bad if st != 0 else P.ERR_NONE
        first = P.first_error(local, lo)
        outs = [torch.empty((P.shard_bounds(nq, world, r)[1] - P.shard_bounds(nq, world, r)[0], w), dtype=torch.float64)
                for r in range(world)]
        dist.all_gather(outs, torch.from_numpy(out)) if len({o.shape for o in outs}) == 1 else None
        results[rank] = (a_full, b_full, first, out, (lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("info", [0, -8], ids=["reference-order", "partition"])
def test_sharded_build_allgather_and_sharded_evaluation_equal_the_single_process_result(info):
    from oracle import oracle_py as O
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_data_worker, args=(world, _free_port(), results, info), nprocs=world, join=True)
    rng = np.random.default_rng(11)
    n, w, nq = 200, 8, 1001
    x = np.cumsum(rng.uniform(0.5, 1.5, n))
    y = rng.normal(size=(n, w))
    q = np.sort(rng.uniform(x[0], x[-1], nq))
    q[700] = x[-1] + 1.0
    st, a, b = O.spline_build_as(x, y, {"kind": "Natural"}, info)
    st, ref, bad = O.interp1d_cubic(x, y, a, b, q, 0, out=np.full((nq, w), -1.0))
    assert bad == 700
    for r in range(world):
        a_full, b_full, first, out, (lo, hi) = results[r]
        assert np.array_equal(a_full, a) and np.array_equal(b_full, b)      # columns are independent: shards == full build
        assert first == (700, 0)                                            # the batch-wide first failure on every rank
        upto = min(hi, 700) - lo                                            # rows before it equal the single-process rows
        assert np.array_equal(out[:max(upto, 0)], ref[lo:lo + max(upto, 0)])


def test_bench_workload_sharding_strong_and_weak():
    import bench
    assert bench.queries_per_gpu(bench.WORKLOADS["c5a"], 8) == 1 << 25 and bench.queries_per_gpu(bench.WORKLOADS["c5a"], 1) == 1 << 28
    assert bench.queries_per_gpu(bench.WORKLOADS["c2"], 8) == 1 << 20
    c2 = bench.WORKLOADS["c2"]
    assert bench.algorithmic_bytes(c2, 1 << 20) == 8 * (1 << 20) + 8 * 1024 * (1 << 20) + 8 * (4096 + 4096 * 1024 + 2 * 4095 * 1024)
    assert abs(bench.algorithmic_bytes(c2, 1 << 20) / 1e9 - 8.699) < 0.001
    assert abs(bench.algorithmic_bytes(bench.WORKLOADS["c3"], 1 << 24) / 1e9 - 1.145) < 0.001
    assert abs(bench.algorithmic_bytes(bench.WORKLOADS["c4"], 1 << 24) / 1e9 - 0.805) < 0.001
