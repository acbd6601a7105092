"""Multi-GPU plumbing of the query path (torch.distributed is plumbing only).

The path shards by QUERY SCAN: every rank holds a replica of the database descriptors (a 5 000-scan DB is ~0.2 GB, a
20 000-scan DB < 1 GB of the 180 GB HBM), ingests and scores its own slice of the query batch, and the only exchange
step is one all-gather of the fixed-stride per-pair score records (include/c2g_types.h: c2g_hint 16 B + c2g_pair_score
128 B per hint slot), after which every rank holds the complete score table of the global batch.
"""
import numpy as np


def shard_range(n_items: int, world: int, rank: int):
    """Contiguous, balanced [begin, end) slice of `n_items` for `rank` (first `n_items % world` ranks get one more)."""
    base, rem = divmod(n_items, world)
    beg = rank * base + min(rank, rem)
    return beg, beg + base + (1 if rank < rem else 0)


def all_gather_records(local, world: int = None):
    """All-gather a rank-local uint8 record buffer (torch tensor, any device the process group supports) into one tensor of
    world * len(local) bytes, rank-major. Equal sizes on every rank (fixed-stride records make that true by construction)."""
    import torch
    import torch.distributed as dist

    world = world or dist.get_world_size()
    out = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
    if dist.get_backend() == "gloo":  # CPU tests
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local)
        torch.cat(parts, out=out)
    else:
        dist.all_gather_into_tensor(out, local)
    return out


def split_gathered(buf, world: int, dtype: np.dtype):
    """View the gathered byte buffer as [world, n_records] structured numpy records (host side)."""
    a = buf.cpu().numpy().view(dtype)
    return a.reshape(world, -1)
