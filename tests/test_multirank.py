"""world_size-2 gloo test of the multi-GPU layout: pair sharding in contiguous blocks + final result gather (CPU)."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_pairs, jobs_per_pair, out_path):
    sys.path.insert(0, ROOT)
    sharding = importlib.import_module("jsvm-code-notes_b200.sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pairs = [(2 * i + 1, 2 * i) for i in range(n_pairs)]
    mine = sharding.shard_pairs(pairs, rank, world)
    # fake "results": one int32 record per job carrying (global pair index, job index)
    local = torch.cat([torch.zeros((0, 2), dtype=torch.int32)] +
                      [torch.stack([torch.full((jobs_per_pair,), gi, dtype=torch.int32),
                                    torch.arange(jobs_per_pair, dtype=torch.int32)], 1) for gi, _ in mine])
    gathered = sharding.gather_results(local, dst=0, n_pairs=n_pairs, jobs_per_pair=jobs_per_pair)
    if rank == 0:
        full = sharding.assemble_level(gathered, len(mine), jobs_per_pair, world, n_pairs)
        np.save(out_path, full.numpy())
    else:
        assert gathered is None
    dist.barrier()
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("n_pairs,port", [(8, 29531), (7, 29533), (1, 29535)])     # even, ragged (4 + 3 pairs) and one rank without work
def test_shard_and_gather_two_ranks(tmp_path, n_pairs, port):
    jobs_per_pair, world = 5, 2
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(world, port, n_pairs, jobs_per_pair, out), nprocs=world, join=True)
    full = np.load(out)
    assert full.shape == (n_pairs * jobs_per_pair, 2)
    want_pair = np.repeat(np.arange(n_pairs), jobs_per_pair)
    assert np.array_equal(full[:, 0], want_pair)
    assert np.array_equal(full[:, 1], np.tile(np.arange(jobs_per_pair), n_pairs))


def test_block_owner():
    sharding = importlib.import_module("jsvm-code-notes_b200.sharding")
    pairs = list(range(10))
    got = [sharding.shard_pairs(pairs, r, 4) for r in range(4)]
    assert [i for g in got for i, _ in g] == list(range(10))                      # contiguous blocks in rank order
    assert [len(g) for g in got] == [3, 3, 2, 2]
    assert all(sharding.pair_owner(i, 4, 10) == r for r, g in enumerate(got) for i, _ in g)
    # the two pairs of a current picture (2k, 2k + 1) stay on one rank whenever the blocks are even
    got = [sharding.shard_pairs(list(range(16)), r, 8) for r in range(8)]
    assert all([i for i, _ in g] == [2 * r, 2 * r + 1] for r, g in enumerate(got))


# ---- the same layout carrying real motion-estimation records (the CPU checker stands in for the GPU search) ----
def _level(synth, n_pairs):
    """Pictures and job lists of one temporal level on a 64x48 picture: pair g = (2g + 1, 2g)."""
    W, H = 64, 48
    seq = synth.Sequence(W, H, 321)
    pairs = [(2 * g + 1, 2 * g) for g in range(n_pairs)]
    return W, H, seq, pairs


def _search_pairs(ROOT_, synth, oracle_lib, W, H, seq, mine):
    """Results of the given [(global index, (cur, ref)), ...] pairs, concatenated in that order (int32 view of the 32-byte records)."""
    import subprocess
    lib = os.path.join(ROOT_, "oracle", "libme_oracle.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT_, "oracle"), "libme_oracle.so"])
    be = oracle_lib.CpuBackend(lib, "jsvm_ora_")
    out = []
    for g, (c, r) in mine:
        ses = be.session(W, H, 2, 8, 2)
        ses.set_plane(0, seq.frame(r))
        ses.set_plane(1, seq.frame(c))
        jobs = synth.make_jobs(W, H, synth.PARTS_9, cur_slot=1, ref_slot=0, qp=33, field="random", seed=700 + g)
        out.append(np.ascontiguousarray(ses.search(jobs)).view(np.int32).reshape(len(jobs), 8))
    return np.concatenate(out) if out else np.zeros((0, 8), np.int32)


def _me_worker(rank, world, port, n_pairs, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    sharding = importlib.import_module("jsvm-code-notes_b200.sharding")
    synth = importlib.import_module("jsvm-code-notes_b200.synth")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, H, seq, pairs = _level(synth, n_pairs)
    mine = sharding.shard_pairs(pairs, rank, world)
    jobs_per_pair = (W // 16) * (H // 16) * len(synth.PARTS_9)
    local = torch.from_numpy(_search_pairs(ROOT, synth, oracle_lib, W, H, seq, mine))
    gathered = sharding.gather_results(local, dst=0, n_pairs=n_pairs, jobs_per_pair=jobs_per_pair)
    if rank == 0:
        np.save(out_path, sharding.assemble_level(gathered, len(mine), jobs_per_pair, world, n_pairs).numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs,port", [(4, 29541), (5, 29543)])
def test_sharded_search_results_match_the_unsharded_level(tmp_path, n_pairs, port):
    """Every rank searches its block of pairs, the result RECORDS (MVs, SADs, costs, modes) are gathered and re-assembled on rank 0:
    byte-identical to one process searching the whole level."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    synth = importlib.import_module("jsvm-code-notes_b200.synth")
    out = str(tmp_path / "level.npy")
    mp.spawn(_me_worker, args=(2, port, n_pairs, out), nprocs=2, join=True)
    W, H, seq, pairs = _level(synth, n_pairs)
    want = _search_pairs(ROOT, synth, oracle_lib, W, H, seq, list(enumerate(pairs)))
    got = np.load(out)
    assert got.shape == want.shape and got.dtype == want.dtype
    assert np.array_equal(got, want)
    assert np.any(got[:, 0] != 0)                                             # real motion vectors, not padding
