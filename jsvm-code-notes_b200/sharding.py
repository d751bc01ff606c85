"""Multi-GPU layout of the path: independent (current, reference) pairs, one final MV-field gather.

In hierarchical-B coding a picture of temporal level T references only pictures of levels < T
(ENC/GOPEncoder.cpp:1322-1329, coding order :530-640), so all (current, reference) pairs of one level are
independent.  They are dealt round-robin to the ranks (one process per GPU); each rank holds its own copy
of the planes it needs, there is no halo and no data-path collective.  The only exchange is the gather of
the per-job result records to the rank that owns mode decision -- torch.distributed (NCCL over NVLink on
GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_pairs(pairs, rank, world):
    """Round-robin whole (current, reference) pairs; returns [(global_index, pair), ...] of this rank."""
    return [(i, p) for i, p in enumerate(pairs) if i % world == rank]


def pair_owner(global_index, world):
    return global_index % world


def pairs_of_rank(n_pairs, rank, world):
    """How many of n_pairs round-robin pairs rank `rank` owns."""
    return (n_pairs - rank + world - 1) // world if rank < n_pairs else 0


def gather_results(local, dst=0, group=None, n_pairs=None, jobs_per_pair=None):
    """Gather per-rank result tensors on `dst`; returns the list (rank order) there, else None.

    Round-robin dealing gives ranks different numbers of pairs whenever n_pairs % world != 0 while the gather needs equal
    sizes: with n_pairs / jobs_per_pair given every rank pads its records to ceil(n_pairs / world) pairs and the returned
    tensors are trimmed back to what each rank owns (assemble_level expects exactly that).  Without them the sizes must
    already agree."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return [local]
    rank = dist.get_rank(group)
    send = local
    if n_pairs is not None:
        rows = -(-n_pairs // world) * jobs_per_pair
        if local.shape[0] != pairs_of_rank(n_pairs, rank, world) * jobs_per_pair:
            raise ValueError(f"rank {rank} holds {local.shape[0]} records, expected {pairs_of_rank(n_pairs, rank, world) * jobs_per_pair}")
        if local.shape[0] < rows:
            send = torch.cat([local, local.new_zeros((rows - local.shape[0],) + tuple(local.shape[1:]))])
    out = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, out, dst=dst, group=group)
    if out is not None and n_pairs is not None:
        out = [t[: pairs_of_rank(n_pairs, r, world) * jobs_per_pair] for r, t in enumerate(out)]
    return out


def assemble_level(gathered, pairs_per_rank, jobs_per_pair, world, n_pairs):
    """Reorder gathered per-rank records (rank-major) into the global pair order of the level."""
    rec = gathered[0].shape[-1] if gathered[0].dim() > 1 else None
    out = [None] * n_pairs
    for r in range(world):
        mine = [i for i in range(n_pairs) if i % world == r]
        for k, gi in enumerate(mine):
            out[gi] = gathered[r][k * jobs_per_pair:(k + 1) * jobs_per_pair]
    return torch.cat(out) if rec is None else torch.cat(out, 0)
