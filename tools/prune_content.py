"""K2 time of one C4 pair x 8 on three kinds of content -- the synthetic texture of the bench, flat pictures (no SAD ever exceeds an
incumbent: the pruning test is pure overhead) and white noise (every SAD is large) -- for the library in JSVM_ME_LIB."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("jsvm-code-notes_b200"); synth = importlib.import_module("jsvm-code-notes_b200.synth")
cfg = synth.CONFIGS["C4"]; W, H = cfg["width"], cfg["height"]
rng = np.random.default_rng(3)
seq = synth.Sequence(W, H, 1238)
contents = {"texture": [seq.frame(0), seq.frame(1)],
            "flat": [np.full((H, W), 100, np.uint8)] * 2,
            "noise": [rng.integers(0, 256, (H, W)).astype(np.uint8) for _ in range(2)]}
jobs = np.concatenate([synth.make_jobs(W, H, cfg["parts"], cur_slot=1, ref_slot=0, qp=35, field="random", seed=100 + g) for g in range(8)])
me = pkg.MotionEstimator(0); me.configure(W, H, 2); me.set_params(64, 2)
for name, fr in contents.items():
    for s, f in enumerate(fr): me.upload_plane(s, f)
    me.interp_halfpel([0])
    me.search(jobs)
    me.timing_begin()
    for _ in range(3): me.search(jobs)
    kt = me.timing_end()
    print(f"{os.path.basename(os.environ.get('JSVM_ME_LIB', 'in-tree'))} {name}: K2 {kt['search_ms'] / 3:.3f} ms per 8 pairs, K3 {kt['subpel_ms'] / 3:.3f} ms", flush=True)
