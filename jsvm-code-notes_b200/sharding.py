"""Multi-GPU layout of the path: independent (current, reference) pairs, one final MV-field gather.

In hierarchical-B coding a picture of temporal level T references only pictures of levels < T
(ENC/GOPEncoder.cpp:1322-1329, coding order :530-640), so all (current, reference) pairs of one level are
independent.  They are dealt to the ranks (one process per GPU) in CONTIGUOUS blocks of the level's pair list: the two pairs
of a current picture and the pairs of neighbouring pictures then land on the same rank and share their planes (a rank that
owns pairs 2k, 2k + 1 uploads 3 planes, not 4; a whole GOP's top level needs 17 planes, not 32).  Each rank holds its own copy
of the planes it needs, there is no halo and no data-path collective.  The only exchange is the gather of the per-job result
records to the rank that owns mode decision: jsvm_me_gather_results_device (NCCL inside the library) on GPUs; the
torch.distributed form below (gloo) covers the same layout logic in the CPU tests.
"""
import torch
import torch.distributed as dist


def rank_range(n_pairs, rank, world):
    """[begin, end) of the contiguous block of pairs rank `rank` owns (the first n_pairs % world ranks hold one pair more)."""
    base, extra = divmod(n_pairs, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_pairs(pairs, rank, world):
    """Whole (current, reference) pairs in contiguous blocks; returns [(global_index, pair), ...] of this rank."""
    begin, end = rank_range(len(pairs), rank, world)
    return [(i, pairs[i]) for i in range(begin, end)]


def pair_owner(global_index, world, n_pairs):
    for r in range(world):
        b, e = rank_range(n_pairs, r, world)
        if b <= global_index < e:
            return r
    raise IndexError(global_index)


def pairs_of_rank(n_pairs, rank, world):
    """How many of n_pairs pairs rank `rank` owns."""
    b, e = rank_range(n_pairs, rank, world)
    return e - b


def gather_results(local, dst=0, group=None, n_pairs=None, jobs_per_pair=None):
    """Gather per-rank result tensors on `dst`; returns the list (rank order) there, else None.

    Ranks hold different numbers of pairs whenever n_pairs % world != 0 while the gather needs equal
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
    """Gathered per-rank records (rank-major, trimmed to what each rank owns) in the global pair order of the level: with
    contiguous blocks that is their concatenation."""
    parts = []
    for r in range(world):
        n = pairs_of_rank(n_pairs, r, world) * jobs_per_pair
        if gathered[r].shape[0] != n:
            raise ValueError(f"rank {r}: {gathered[r].shape[0]} records, expected {n}")
        parts.append(gathered[r])
    return torch.cat(parts, 0)
