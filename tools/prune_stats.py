"""Prune statistics of the full search (library built with -DK2_PRUNE_STATS): share of (batch, slot) pairs whose rate / key code ran."""
import ctypes as C, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module("jsvm-code-notes_b200"); synth = importlib.import_module("jsvm-code-notes_b200.synth")
sharding = importlib.import_module("jsvm-code-notes_b200.sharding")
lib = C.CDLL(os.environ["JSVM_ME_LIB"])
for cfgn, field in (("C4", "random"), ("C4", "zero"), ("C2", "random")):
    cfg = synth.CONFIGS[cfgn]
    seq, frames, slot_of, mine, jobs, ref_slots = bench.build_workload(synth, sharding, cfg, 4, field, 0, 1)
    me = pkg.MotionEstimator(0); me.configure(cfg["width"], cfg["height"], len(frames)); me.set_params(cfg["search_range"], cfg["sub_dfunc"])
    for p, f in frames.items(): me.upload_plane(slot_of[p], f)
    me.interp_halfpel(ref_slots)
    out = (C.c_ulonglong * 4)()
    lib.jsvm_me_debug_prune_stats(out, 1)
    me.search(jobs)
    lib.jsvm_me_debug_prune_stats(out, 1)
    b, none_live, live, inside = [int(x) for x in out]
    print(cfgn, field, me.last_search_kernel(), f"batches {b}, fully pruned {none_live / b:.3f}, live slots per batch {live / b:.2f} of {inside / b:.2f} inside the window")
    me.close()
