"""Where the end-to-end (host-buffer) step spends its time: plane uploads + K1, then jsvm_me_search."""
import importlib, sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module("jsvm-code-notes_b200"); synth = importlib.import_module("jsvm-code-notes_b200.synth")
cfg = synth.CONFIGS["C4"]
frames, slot_of, pairs, jobs, ref_slots = bench.build_workload(synth, cfg, 16, "random", 0)
me = pkg.MotionEstimator(0); me.configure(cfg["width"], cfg["height"], len(frames)); me.set_params(64, 2)
hp = {p: torch.from_numpy(f).pin_memory() for p, f in frames.items()}
hj = torch.from_numpy(jobs.view(np.uint8).copy()).pin_memory().numpy().view(pkg.JOB_DTYPE)
hr = torch.empty(len(jobs) * 32, dtype=torch.uint8).pin_memory().numpy().view(pkg.RESULT_DTYPE)
for it in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for p, t in hp.items(): me.upload_plane_ptr(slot_of[p], t.data_ptr())
    me.interp_halfpel(ref_slots)
    me.sync(); t1 = time.perf_counter()
    me.search(hj, out=hr)
    t2 = time.perf_counter()
    g, info = me.plan(hj); t3 = time.perf_counter()
    print(f"upload+interp {1e3*(t1-t0):.2f} ms  search {1e3*(t2-t1):.2f} ms  (plan alone {1e3*(t3-t2):.2f} ms)")
