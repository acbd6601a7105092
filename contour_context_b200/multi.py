"""Multi-GPU plumbing of the query path (torch.distributed is plumbing only).

The path shards by QUERY SCAN (independent objects): every rank holds a replica of the database descriptors (a 5 000-scan
DB is ~0.2 GB, a 20 000-scan DB < 6 GB of the 180 GB HBM), ingests, searches, scores and refines its own slice of the query
batch.  The ONE exchange step publishes every rank's per-query outcome (c2g_query_result, include/c2g_types.h: candidate
poses with their scores, correlation and SE(2) - what ContourDB::queryRangedKNN returns plus the runner-up candidates) to all
ranks, so that each of them holds the loop closures of the whole batch: one all-gather of 3.1 KB per query scan.
(Round 1 gathered the fixed-stride hint / pair-score tables - 130 KB per query scan, ~85 % empty slots, never consumed.)
"""
import numpy as np


def shard_range(n_items: int, world: int, rank: int):
    """Contiguous, balanced [begin, end) slice of `n_items` for `rank` (first `n_items % world` ranks get one more)."""
    base, rem = divmod(n_items, world)
    beg = rank * base + min(rank, rem)
    return beg, beg + base + (1 if rank < rem else 0)


def all_gather_records(local, world: int = None, out=None, async_op: bool = False):
    """All-gather a rank-local uint8 record buffer (torch tensor, any device the process group supports) into one tensor of
    world * len(local) bytes, rank-major.  Equal sizes on every rank (fixed-stride records make that true by construction).
    Returns the gathered tensor, or (tensor, work handle) when async_op."""
    import torch
    import torch.distributed as dist

    world = world or dist.get_world_size()
    if out is None:
        out = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
    if dist.get_backend() == "gloo":  # CPU tests
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local)
        torch.cat(parts, out=out)
        return (out, None) if async_op else out
    work = dist.all_gather_into_tensor(out, local, async_op=async_op)
    return (out, work) if async_op else out


def split_gathered(buf, world: int, dtype: np.dtype):
    """View the gathered byte buffer as [world, n_records] structured numpy records (host side)."""
    a = buf.cpu().numpy().view(dtype)
    return a.reshape(world, -1)


def loop_closures(results: np.ndarray, first_query_id: int = 0):
    """What a consumer of the exchange reads: (query id, candidate gidx, refined correlation, T_fine) of every query of a
    gathered result block that returned a candidate (ContourDB::queryRangedKNN's ret_size = 1, contour_db.h:639)."""
    out = []
    for j, r in enumerate(results):
        if r["n_cand"] > 0 and r["best"] >= 0:
            c = r["cand"][r["best"]]
            out.append((first_query_id + j, int(c["cand_gidx"]), float(c["corr_fine"]), np.array(c["T_fine"])))
    return out


def verify_foreign_block(gathered: np.ndarray, owner: int, recomputed: np.ndarray) -> int:
    """Number of records of `owner`'s block in the gathered table that differ (bytes) from `recomputed`, the same queries
    run on THIS rank's replica.  0 proves both that the exchange delivered the records intact and that sharding by query
    leaves every per-query result unchanged (DYNAMIC_THRES=0, CMakeLists.txt:21)."""
    blk = gathered[owner]
    assert blk.shape == recomputed.shape
    return int(sum(1 for a, b in zip(blk, recomputed) if a.tobytes() != b.tobytes()))
