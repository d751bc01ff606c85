#!/usr/bin/env python
"""Throughput of the picture-level motion search (jsvm_me_estimate_pictures) at the C4 geometry: pictures/s for batches of
independent pictures (a temporal level), resident planes, CUDA-event free wall clock around synchronised calls."""
import importlib, json, sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("jsvm-code-notes_b200")
synth = importlib.import_module("jsvm-code-notes_b200.synth")

def main():
    cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C4"]
    W, H, sr = cfg["width"], cfg["height"], cfg["search_range"]
    n_pics_list = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["1", "4", "16"])]
    n_refs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    me = pkg.MotionEstimator(0)
    n_slots = max(n_pics_list) + n_refs
    me.configure(W, H, n_slots)
    me.set_params(sr, 2)
    seq = synth.LocalMotionSequence(W, H, 77, tile=52)
    for s in range(n_slots):
        me.upload_plane(s, seq.frame(s % 6))
    me.interp_halfpel(list(range(n_slots)))
    for n in n_pics_list:
        tasks = synth.make_pic_tasks([n_refs + p for p in range(n)], [[(p + r) % (n_refs + p) for r in range(n_refs)] for p in range(n)], qp=32)
        reps = 2
        me.timing_begin()                                   # kernel timing on: plain launches
        me.estimate_pictures(tasks, download=False); me.sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            me.estimate_pictures(tasks, download=False)
        me.sync()
        dt_plain = (time.perf_counter() - t0) / reps
        tm = me.timing_end()
        me.estimate_pictures(tasks, download=False); me.sync()        # captures the graph
        l0 = me.launch_count()
        t0 = time.perf_counter()
        for _ in range(reps):
            me.estimate_pictures(tasks, download=False)
        me.sync()
        dt = (time.perf_counter() - t0) / reps
        print(json.dumps({"workload": cfg["name"], "n_pics": n, "n_refs": n_refs, "ms_per_batch_graph": dt * 1e3, "pictures_per_s": n / dt,
                          "ms_per_batch_plain_launches": dt_plain * 1e3, "kernels_per_batch": (me.launch_count() - l0) // reps,
                          "k2_ms": tm["search_ms"] / (reps + 1), "k3_ms": tm["subpel_ms"] / (reps + 1)}))

main()
