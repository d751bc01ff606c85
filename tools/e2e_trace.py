"""Timeline of one host-buffer search of the bench workload (JSVM_ME_TRACE): host time of every launch decision, device time of
every timed kernel."""
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
for p, t in hp.items(): me.upload_plane_ptr(slot_of[p], t.data_ptr())
me.interp_halfpel(ref_slots)
for it in range(3): me.search(hj, out=hr)
os.environ["JSVM_ME_TRACE"] = "1"
me.timing_begin(); t0 = time.perf_counter()
me.search(hj, out=hr)
t1 = time.perf_counter(); kt = me.timing_end()
print(f"search wall {1e3*(t1-t0):.2f} ms", kt)
if len(sys.argv) > 1 and sys.argv[1] == "more":
    os.environ["JSVM_ME_LATE_JOBS"] = "1"
    for it in range(2): me.search(hj, out=hr)
    print("---- each launch uploads its own jobs", flush=True)
    me.timing_begin(); t0 = time.perf_counter(); me.search(hj, out=hr); t1 = time.perf_counter(); kt = me.timing_end()
    print(f"search wall {1e3*(t1-t0):.2f} ms", kt, flush=True)
    del os.environ["JSVM_ME_LATE_JOBS"]
    per = len(jobs) // 16
    hj2 = torch.from_numpy(np.concatenate([jobs[i * per:(i + 1) * per] for i in reversed(range(16))]).view(np.uint8).copy()).pin_memory().numpy().view(pkg.JOB_DTYPE)
    for it in range(2): me.search(hj2, out=hr)
    print("---- pairs in reverse order", flush=True)
    me.timing_begin(); t0 = time.perf_counter(); me.search(hj2, out=hr); t1 = time.perf_counter(); kt = me.timing_end()
    print(f"search wall {1e3*(t1-t0):.2f} ms", kt, flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "benchlike":
    # the e2e step of bench.py: ctx stream = torch's current stream, plane uploads + K1 in front of every search
    me.set_stream(torch.cuda.current_stream().cuda_stream)
    def step():
        for p, t in hp.items(): me.upload_plane_ptr(slot_of[p], t.data_ptr())
        me.interp_halfpel(ref_slots)
        return me.search(hj, out=hr)
    for mode in ("early", "late"):
        if mode == "late": os.environ["JSVM_ME_LATE_JOBS"] = "1"
        for it in range(3): step()
        print("---- bench-like step,", mode, "job uploads", flush=True)
        torch.cuda.synchronize()
        me.timing_begin(); t0 = time.perf_counter(); step(); t1 = time.perf_counter(); kt = me.timing_end()
        print(f"step wall {1e3*(t1-t0):.2f} ms", kt, flush=True)
