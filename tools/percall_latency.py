"""Latency of one estimateBlockWithStart-sized call (jsvm_me_search with 1 job: the drop-in's per-call path) for a few search
ranges and block sizes; us per call over 2000 calls (ctypes overhead included, the same for every variant)."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("jsvm-code-notes_b200"); synth = importlib.import_module("jsvm-code-notes_b200.synth")
abi = importlib.import_module("jsvm-code-notes_b200.abi")
W, H = 176, 144
seq = synth.Sequence(W, H, 5)
me = pkg.MotionEstimator(0); me.configure(W, H, 2)
for s in range(2): me.upload_plane(s, seq.frame(s))
for sr in (8, 16, 32, 64):
    me.set_params(sr, 2)
    me.interp_halfpel([0])
    jobs = synth.make_jobs(W, H, synth.PARTS_9, cur_slot=1, ref_slot=0, qp=30, field="random", seed=1)
    for name, sel in (("16x16", jobs["mode"] == abi.MODE_16x16), ("8x8", jobs["mode"] == abi.BLK_8x8)):
        js = jobs[sel][40:41].copy()
        out = np.zeros(1, pkg.RESULT_DTYPE)
        for _ in range(200): me.search(js, out=out)
        t0 = time.perf_counter()
        for _ in range(2000): me.search(js, out=out)
        dt = (time.perf_counter() - t0) / 2000
        print(f"SR {sr:2d} {name}: {dt * 1e6:6.1f} us per call ({me.last_search_kernel()})", flush=True)
