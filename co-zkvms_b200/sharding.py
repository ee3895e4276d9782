"""Point-range sharding of one MSM across processes (one process per GPU) and the host-side combine.

MSM(s, P) = sum over shards of MSM(s[I_g], P[I_g]) for any partition of the index set - the reference's own worker-subnet
scheme: `split_ck` hands each worker a contiguous slice of the SRS (co-noir-spartan/co-spartan/src/utils.rs:38-83) and
`combine_comm` sums the chunk commitments (snarks-core/src/poly/commitment.rs:56-63).  No data-path collective: each
rank produces one 72-byte partial sum; rank 0 adds them on the host.
"""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous [lo, hi) of rank `rank` when n points are split over `world` ranks (sizes differ by at most one)."""
    lo = n * rank // world
    hi = n * (rank + 1) // world
    return lo, hi


def gather_partials(partial72, group=None):
    """All ranks contribute a (k, 72) uint8 array of partial sums; every rank gets the (world, k, 72) stack.
    Uses the process group's CPU (gloo) side: the payload is a few bytes."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(partial72, dtype=np.uint8).copy())
    world = dist.get_world_size(group)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    return np.stack([o.numpy() for o in outs])


def combine(partials, g1_sum):
    """partials: (world, k, 72) -> (k, 72) sums, by the host-side group law of the engine library (cozk_g1_sum)."""
    partials = np.asarray(partials, dtype=np.uint8)
    world, k = partials.shape[0], partials.shape[1]
    return np.stack([g1_sum(partials[:, j, :]) for j in range(k)])
