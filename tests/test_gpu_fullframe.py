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
    ("C3", "split", 1, "full_search_kernel<4,true>"),
    ("C5", "random", 4, "full_search_persistent_kernel<8,false>"),
]


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
    assert_results_equal(got, want, jobs)
    assert np.array_equal(gp, wp)
    assert_results_equal(got_dev, want, jobs)
    assert np.array_equal(gp_dev, wp)
    n_mb = (W // 16) * (H // 16)
    assert len(jobs) == n_mb * len(cfg["parts"]) * n_refs
    print(f"{cfg_name}/{field}: {len(jobs)} jobs ({n_mb} MBs x {len(cfg['parts'])} partitions x {n_refs} refs) equal to the {kind} checker")
