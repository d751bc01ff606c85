"""Where the end-to-end (host-buffer) step of bench.py spends its time: plane uploads, K1, then jsvm_me_search (host planner + job
upload + K2 + K3 + result copy-back, pipelined inside the library), against the resident step."""
import importlib, sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module("jsvm-code-notes_b200"); synth = importlib.import_module("jsvm-code-notes_b200.synth")
sharding = importlib.import_module("jsvm-code-notes_b200.sharding")
cfg = synth.CONFIGS["C4"]
seq, frames, slot_of, mine, jobs, ref_slots = bench.build_workload(synth, sharding, cfg, 16, "random", 0, 1)
me = pkg.MotionEstimator(0); me.configure(cfg["width"], cfg["height"], len(frames)); me.set_params(64, 2)
hp = {p: torch.from_numpy(f).pin_memory() for p, f in frames.items()}
hj = torch.from_numpy(jobs.view(np.uint8).copy()).pin_memory().numpy().view(pkg.JOB_DTYPE)
hr = torch.empty(len(jobs) * 32, dtype=torch.uint8).pin_memory().numpy().view(pkg.RESULT_DTYPE)
for it in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for p, t in hp.items(): me.upload_plane_ptr(slot_of[p], t.data_ptr())
    ta = time.perf_counter()
    me.sync(); tb = time.perf_counter()
    me.interp_halfpel(ref_slots)
    me.sync(); t1 = time.perf_counter()
    me.search(hj, out=hr)
    t2 = time.perf_counter()
    g, info = me.plan(hj); t3 = time.perf_counter()
    print(f"uploads: enqueue {1e3*(ta-t0):.2f} ms, done {1e3*(tb-t0):.2f} ms ({sum(t.numel() for t in hp.values())/1e6/(tb-t0)/1e3:.1f} GB/s); "
          f"K1 {1e3*(t1-tb):.2f} ms; search {1e3*(t2-t1):.2f} ms; whole step {1e3*(t2-t0):.2f} ms  (plan alone {1e3*(t3-t2):.2f} ms)")
# the same step without the intermediate syncs (what bench.py times)
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for p, t in hp.items(): me.upload_plane_ptr(slot_of[p], t.data_ptr())
    me.interp_halfpel(ref_slots)
    me.search(hj, out=hr)
    t2 = time.perf_counter()
    print(f"step without intermediate syncs {1e3*(t2-t0):.2f} ms")
# kernel time inside the pipelined host-buffer search (CUDA events per launch) against its wall time
for it in range(3):
    for p, t in hp.items(): me.upload_plane_ptr(slot_of[p], t.data_ptr())
    me.interp_halfpel(ref_slots); me.sync()
    me.timing_begin(); t0 = time.perf_counter()
    me.search(hj, out=hr)
    t1 = time.perf_counter(); kt = me.timing_end()
    print(f"search wall {1e3*(t1-t0):.2f} ms; kernels: K2 {kt['search_ms']:.2f} ms in {kt['search_launches']} launches, K3 {kt['subpel_ms']:.2f} ms in {kt['subpel_launches']} launches")
for env in ("JSVM_ME_NO_MERGE",):
    os.environ[env] = "1"
    for it in range(2):
        me.timing_begin(); t0 = time.perf_counter()
        me.search(hj, out=hr)
        t1 = time.perf_counter(); kt = me.timing_end()
        print(f"{env}: search wall {1e3*(t1-t0):.2f} ms; K2 {kt['search_ms']:.2f} ms in {kt['search_launches']} launches, K3 {kt['subpel_ms']:.2f} ms in {kt['subpel_launches']}")
