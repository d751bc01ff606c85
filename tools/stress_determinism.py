"""Repeat one full-frame search many times through both entry points and compare every run with the first one: any difference is a
race (the kernels are deterministic by construction).  python tools/stress_determinism.py [--cfg C5] [--refs 4] [--iters 40]"""
import argparse
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="C5")
ap.add_argument("--refs", type=int, default=4)
ap.add_argument("--iters", type=int, default=40)
ap.add_argument("--field", default="random")
ap.add_argument("--churn", type=int, default=1, help="allocate / scribble / free device memory between runs (stale-data reads show up)")
args = ap.parse_args()
pkg = importlib.import_module("jsvm-code-notes_b200")
synth = importlib.import_module("jsvm-code-notes_b200.synth")
cfg = synth.CONFIGS[args.cfg]
W, H, R = cfg["width"], cfg["height"], args.refs
seq = synth.Sequence(W, H, 1234 + int(args.cfg[1]))
planes = {s: seq.frame(s) for s in range(R + 1)}
jobs = np.concatenate([synth.make_jobs(W, H, cfg["parts"], cur_slot=R, ref_slot=r, qp=33 + (r & 1), field=args.field, seed=40 + r, bits_in=1) for r in range(R)])
dev = torch.device("cuda", 0)
base = None
bad = 0
for it in range(args.iters):
    if args.churn:
        junk = torch.randint(0, 255, (int(torch.randint(64, 512, (1,)).item()) << 20,), dtype=torch.uint8, device=dev)
        del junk
        torch.cuda.empty_cache()
    me = pkg.MotionEstimator(0)
    me.configure(W, H, R + 1)
    me.set_params(cfg["search_range"], cfg["sub_dfunc"])
    for s, f in planes.items():
        me.upload_plane(s, f)
    me.interp_halfpel(list(range(R)))
    got, gp = me.search(jobs, want_preds=True)
    groups, info = me.plan(jobs)
    d_jobs = torch.from_numpy(jobs.view(np.uint8).copy()).to(dev)
    d_groups = torch.from_numpy(groups).to(dev)
    d_res = torch.zeros(len(jobs) * 32, dtype=torch.uint8, device=dev)
    d_preds = torch.zeros(len(jobs) * 256, dtype=torch.uint8, device=dev)
    me.search_device(d_jobs.data_ptr(), len(jobs), d_groups.data_ptr(), info, d_res.data_ptr(), d_preds=d_preds.data_ptr())
    me.sync()
    torch.cuda.synchronize()
    got_dev = d_res.cpu().numpy().view(pkg.RESULT_DTYPE).copy()
    gp_dev = d_preds.cpu().numpy().reshape(-1, 16, 16).copy()
    me.close()
    cur = (got.copy(), gp.copy(), got_dev, gp_dev)
    if base is None:
        base = cur
        continue
    names = ["search results", "search preds", "search_device results", "search_device preds"]
    for nm, a, b in zip(names, base, cur):
        if not np.array_equal(a.view(np.uint8), b.view(np.uint8)):
            bad += 1
            diff = np.flatnonzero((a.view(np.uint8).reshape(len(jobs), -1) != b.view(np.uint8).reshape(len(jobs), -1)).any(axis=1))
            print(f"iter {it}: {nm} differ in {len(diff)} jobs, first {diff[:8]}", flush=True)
            for d in diff[:4]:
                print("   job", jobs[d], "\n   base", a[d] if a.dtype.names else "", "\n   now ", b[d] if a.dtype.names else "", flush=True)
    # the two entry points must agree with each other as well
    if not np.array_equal(cur[0].view(np.uint8), cur[2].view(np.uint8)):
        print(f"iter {it}: host-buffer and resident results differ", flush=True)
print("runs", args.iters, "differences", bad)
