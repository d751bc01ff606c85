"""Host-side logic that needs no GPU: the search-group planner, synthetic workloads, cost factors."""
import ctypes as C

import numpy as np
import pytest


def _windows(oracle, W, H, sr, jobs):
    ses = oracle.session(W, H, 1, sr, 0)
    fn = oracle.lib.jsvm_ora_window_positions
    fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_void_p]
    return np.array([fn(ses.ctx, jobs[i:i + 1].ctypes.data) for i in range(len(jobs))], np.int64)


@pytest.mark.parametrize("parts_name,field", [("PARTS_9", "zero"), ("PARTS_9", "random"), ("PARTS_41", "split")])
def test_planner_partitions_jobs_and_counts_positions(pkg, synth, oracle, parts_name, field):
    W, H, sr = 176, 144, 16
    parts = getattr(synth, parts_name)
    jobs = synth.make_jobs(W, H, parts, 1, 0, field=field, seed=5)
    groups, info = pkg.plan_host(W, H, 2, sr, jobs)
    g = groups.view(pkg.GROUP_DTYPE)
    assert len(g) == info.n_groups
    # every job sits in exactly one slot of exactly one group
    seen = np.zeros(len(jobs), np.int32)
    for rec in g:
        offs = rec["slot_job"][rec["slot_job"] >= 0].astype(np.int64)
        assert len(offs) == len(set(offs.tolist())) <= rec["n_jobs"]
        idx = rec["first_job"] + offs
        seen[idx] += 1
        jb = jobs[idx]
        assert (jb["mb_x"] == rec["mb_x"]).all() and (jb["mb_y"] == rec["mb_y"]).all()
        assert (jb["ref_slot"] == rec["ref_slot"]).all() and (jb["cur_slot"] == rec["cur_slot"]).all()
        assert rec["uw"] <= 145 and rec["uw"] * rec["uh"] <= 32768
    assert (seen == 1).all()
    # 8x8-granular groups come first and hold no sub-8x8 partition
    assert (g["gran4"][: info.n_groups_8x8] == 0).all() and (g["gran4"][info.n_groups_8x8:] == 1).all()
    # the position count the roofline uses equals the oracle's clipped window sizes
    assert info.n_positions == int(_windows(oracle, W, H, sr, jobs).sum())
    if field == "zero":
        assert info.n_groups == (W // 16) * (H // 16)      # one group per macroblock when all windows coincide
        assert info.n_group_positions * len(parts) == info.n_positions


def test_planner_rejects_bad_jobs(pkg, synth):
    jobs = synth.make_jobs(176, 144, synth.PARTS_9, 1, 0)
    bad = jobs.copy(); bad["mode"][3] = 7
    with pytest.raises(pkg.MotionEstimationError):
        pkg.plan_host(176, 144, 2, 16, bad)
    bad = jobs.copy(); bad["mb_x"][0] = 11
    with pytest.raises(pkg.MotionEstimationError):
        pkg.plan_host(176, 144, 2, 16, bad)
    bad = jobs.copy(); bad["blk"][1] = 1           # 16x8 cannot start at blk 1
    with pytest.raises(pkg.MotionEstimationError):
        pkg.plan_host(176, 144, 2, 16, bad)
    with pytest.raises(pkg.MotionEstimationError):
        pkg.plan_host(176, 144, 2, 65, jobs)       # search range beyond this build


def test_cost_factor_matches_oracle(pkg, synth, oracle):
    for qp in range(0, 52):
        lam = synth.lambda_for_qp(qp)
        assert pkg.cost_factor_sad(lam) == oracle.cost_factor(lam, 1) == synth.cost_factor_for_qp(qp)


def test_hierarchical_b_levels(synth):
    lv = synth.hierarchical_b_pairs(16)
    assert [len(l) for l in lv] == [1, 2, 4, 8, 16]
    coded = {0}
    for level in lv:
        for cur, ref in level:
            assert ref in coded and cur not in coded      # references are always coded earlier, at a lower level
        coded |= {c for c, _ in level}
    assert coded == set(range(17))
    assert [len(l) for l in synth.hierarchical_b_pairs(16, refs=4)][-1] > 16


def test_synthetic_sequence_is_deterministic(synth):
    a, b = synth.Sequence(64, 48, 3).frame(2), synth.Sequence(64, 48, 3).frame(2)
    assert a.dtype == np.uint8 and a.shape == (48, 64) and np.array_equal(a, b)
    assert a.std() > 10
