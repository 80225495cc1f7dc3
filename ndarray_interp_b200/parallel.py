"""Query / column sharding across the GPUs of one box.

The evaluation path needs no collective: every rank evaluates a contiguous block of the flattened
query index against replicated tables.  What the host has to get right is the bookkeeping --
which block a rank owns, and how the per-rank first-error words combine into the error the
reference would have reported (the first failing query in row-major order, interp1d/mod.rs:321).
These helpers are backend-agnostic (`torch.distributed` with nccl on GPUs, gloo in the CPU tests).
"""
ERR_NONE = 2 ** 64 - 1


def shard_bounds(total, world, rank):
    """contiguous block [lo, hi) of `total` items owned by `rank` (sizes differ by at most one)"""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_error_word(local_word, shard_lo, two_d=False):
    """a rank's local error word rebased to the global query index (ERR_NONE stays ERR_NONE).
    2-D words are 2*index + axis (x before y, bilinear.rs:71-80)."""
    if local_word == ERR_NONE:
        return ERR_NONE
    return local_word + (2 * shard_lo if two_d else shard_lo)


def combine_error_words(words):
    """the error the reference would report for the whole batch: the smallest global word"""
    return min(words) if words else ERR_NONE


def first_error(local_word, shard_lo, two_d=False, group=None):
    """all-reduce(MIN) of the rebased error words; returns (first_bad_index, axis) or None"""
    import torch
    import torch.distributed as dist
    word = global_error_word(local_word, shard_lo, two_d)
    # int64 cannot hold 2**64-1: MIN over (flag, value) with the sentinel mapped to int64 max
    t = torch.tensor([word if word != ERR_NONE else 2 ** 63 - 1], dtype=torch.int64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        backend = dist.get_backend(group)
        if backend == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    w = int(t.item())
    if w == 2 ** 63 - 1:
        return None
    return (w >> 1, w & 1) if two_d else (w, 0)


def column_shards(w, world):
    """[lo, hi) column blocks for sharded spline construction (the tridiagonal factors depend only
    on x, cubic_spline.rs:440-451, so every rank recomputes them instead of communicating)"""
    return [shard_bounds(w, world, r) for r in range(world)]
