"""Full-frame GPU parity at BASELINE.json's own sizes: EVERY macroblock of a whole (current, reference) pair of C3 / C4 / C5
through both entry points -- jsvm_me_search (host buffers) and jsvm_me_plan + jsvm_me_search_device (resident buffers, what
bench.py times as `value`) -- against the CPU checker run on all host cores (the compiled unmodified reference when
oracle/_ref travelled to this box, else the C restatement).  All ten result fields and the prediction blocks, exact equality;
the full-search kernel that served the launch is asserted, so the benchmarked persistent kernel is the one compared."""
import os

import numpy as np
import pytest

import oracle_lib
from test_gpu_parity import assert_results_equal

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    # config, start/pred field, reference pictures, expected full-search kernel
    ("C4", "random", 1, "full_search_persistent_kernel<8,false>"),
    ("C4", "split", 1, "full_search_persistent_kernel<8,false>"),
    ("C2", "random", 2, "full_search_kernel<8,false>"),
    ("C3", "split", 1, "full_search_kernel<4,true>"),
    ("C5", "random", 4, "full_search_persistent_kernel<8,false>"),
]


def _explain_disagreement(got, got_dev, want, jobs, planes, path, prefix, W, H, n_refs, cfg):
    """On a mismatch, say which side moved: the macroblocks that differ are searched again by the checker in THIS process (no fork
    pool), and the report names the jobs for which the checker's second answer differs from its first."""
    raw = lambda a: np.ascontiguousarray(a).view(np.uint8).reshape(len(jobs), -1)
    bad = np.flatnonzero((raw(got) != raw(want)).any(axis=1) | (raw(got_dev) != raw(want)).any(axis=1))
    if not len(bad):
        return
    sel = np.flatnonzero(np.isin(jobs["mb_y"].astype(np.int64) * 4096 + jobs["mb_x"], np.unique(jobs["mb_y"][bad].astype(np.int64) * 4096 + jobs["mb_x"][bad]))
                         & np.isin(jobs["ref_slot"], np.unique(jobs["ref_slot"][bad])))[:4096]
    again = oracle_lib.parallel_search(path, prefix, W, H, n_refs + 1, cfg["search_range"], cfg["sub_dfunc"], planes, jobs[sel], n_proc=1)
    moved = int((raw(want)[sel] != np.ascontiguousarray(again).view(np.uint8).reshape(len(sel), -1)).any(axis=1).sum())
    print(f"{len(bad)} jobs differ (host-buffer path {int((raw(got) != raw(want)).any(axis=1).sum())}, resident path "
          f"{int((raw(got_dev) != raw(want)).any(axis=1).sum())}); the checker changed its own answer on {moved} of the {len(sel)} jobs re-run in-process; "
          f"host-buffer vs resident results differ on {int((raw(got) != raw(got_dev)).any(axis=1).sum())} jobs")


@pytest.mark.parametrize("cfg_name,field,n_refs,kernel", CASES, ids=[f"{c[0]}-{c[1]}-{c[2]}ref" for c in CASES])
def test_every_macroblock_of_a_pair_bit_exact(pkg, synth, cfg_name, field, n_refs, kernel):
    import torch
    cfg = synth.CONFIGS[cfg_name]
    W, H = cfg["width"], cfg["height"]
    seq = synth.Sequence(W, H, 1234 + int(cfg_name[1]))
    planes = {s: seq.frame(s) for s in range(n_refs + 1)}                      # slots 0 .. n_refs-1 = references, n_refs = current
    jobs = np.concatenate([synth.make_jobs(W, H, cfg["parts"], cur_slot=n_refs, ref_slot=r, qp=33 + (r & 1), field=field, seed=40 + r, bits_in=1)
                           for r in range(n_refs)])
    me = pkg.MotionEstimator(0)
    try:
        me.configure(W, H, n_refs + 1)
        me.set_params(cfg["search_range"], cfg["sub_dfunc"])
        for s, f in planes.items():
            me.upload_plane(s, f)
        me.interp_halfpel(list(range(n_refs)))

        # (1) host-buffer entry point
        got, gp = me.search(jobs, want_preds=True)
        assert me.last_search_kernel() == kernel
        # (2) resident entry point: the planner's groups and the job list live in HBM
        groups, info = me.plan(jobs)
        dev = torch.device("cuda", 0)
        d_jobs = torch.from_numpy(jobs.view(np.uint8).copy()).to(dev)
        d_groups = torch.from_numpy(groups).to(dev)
        d_res = torch.zeros(len(jobs) * 32, dtype=torch.uint8, device=dev)
        d_preds = torch.zeros(len(jobs) * 256, dtype=torch.uint8, device=dev)
        me.search_device(d_jobs.data_ptr(), len(jobs), d_groups.data_ptr(), info, d_res.data_ptr(), d_preds=d_preds.data_ptr())
        me.sync()
        torch.cuda.synchronize()
        assert me.last_search_kernel() == kernel
        got_dev = d_res.cpu().numpy().view(pkg.RESULT_DTYPE)
        gp_dev = d_preds.cpu().numpy().reshape(-1, 16, 16)
    finally:
        me.close()

    path, prefix, kind = oracle_lib.best_checker(ROOT)
    want, wp = oracle_lib.parallel_search(path, prefix, W, H, n_refs + 1, cfg["search_range"], cfg["sub_dfunc"], planes, jobs, want_preds=True)
    _explain_disagreement(got, got_dev, want, jobs, planes, path, prefix, W, H, n_refs, cfg)
    assert_results_equal(got, want, jobs)
    assert np.array_equal(gp, wp)
    assert_results_equal(got_dev, want, jobs)
    assert np.array_equal(gp_dev, wp)
    n_mb = (W // 16) * (H // 16)
    assert len(jobs) == n_mb * len(cfg["parts"]) * n_refs
    print(f"{cfg_name}/{field}: {len(jobs)} jobs ({n_mb} MBs x {len(cfg['parts'])} partitions x {n_refs} refs) equal to the {kind} checker")
