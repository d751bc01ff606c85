"""CPU estimate (numpy, a few macroblocks of the bench content) of how many batches of 256 displacements an EXACT lower-bound
test could skip in the 8x8-granular full search, for two batch shapes -- the shipped one (8 rows x every 4th column across the
window, what a warp's 32 lanes cover) and a compact one (8 rows x 32 adjacent columns) -- and three bounds: the whole SAD (pass-1
pruning, shipped), the SAD of half of the rows of every quadrant (partial-distortion elimination of the SAD pass) and the
block-sum bound |sum(cur) - sum(ref)| on 4x4 sub-blocks (successive elimination, no SAD at all).  Incumbents = the final winners'
costs (what the centre-out order approaches after the first batches)."""
import importlib, sys, types, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name = "jsvm-code-notes_b200"
sys.modules[name] = types.ModuleType(name); sys.modules[name].__path__ = [os.path.join(ROOT, name)]      # synth only, no CUDA library
synth = importlib.import_module(name + ".synth")
W, H, R, pad = 1920, 1088, 64, 96
seq = synth.Sequence(W, H, 1238)
cur = seq.frame(1).astype(np.int32); refp = np.pad(seq.frame(0).astype(np.int32), pad, mode="edge")

def maps(bx, by):
    """per displacement: whole 8x8 SAD, SAD of rows 0-3, sum over the four 4x4 sub-blocks of |block-sum difference|"""
    blk = cur[by:by + 8, bx:bx + 8]
    full = np.zeros((2 * R + 1, 2 * R + 1), np.int64); half = full.copy(); sea = full.copy()
    bs = blk.reshape(2, 4, 2, 4).sum(axis=(1, 3))
    for dy in range(-R, R + 1):
        rows = refp[pad + by + dy:pad + by + dy + 8, pad + bx - R:pad + bx + R + 8]
        win = np.lib.stride_tricks.sliding_window_view(rows, (8, 8))[0]
        d = np.abs(win - blk)
        full[dy + R] = d.sum(axis=(1, 2)); half[dy + R] = d[:, :4, :].sum(axis=(1, 2))
        ws = win.reshape(-1, 2, 4, 2, 4).sum(axis=(2, 4))
        sea[dy + R] = np.abs(ws - bs).sum(axis=(1, 2))
    return full, half, sea

def slots(P):
    return [P[0] + P[1] + P[2] + P[3], P[0] + P[1], P[2] + P[3], P[0] + P[2], P[1] + P[3], P[0], P[1], P[2], P[3]]

rng = np.random.default_rng(0)
count = {(shape, b): 0 for shape in ("strided", "compact") for b in ("sad", "half", "sea")}
total = {"strided": 0, "compact": 0}
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    mbx, mby = rng.integers(8, W // 16 - 8), rng.integers(8, H // 16 - 8)
    M = [maps(16 * mbx + 8 * (q & 1), 16 * mby + 8 * (q >> 1)) for q in range(4)]
    S = {b: slots([M[q][i] for q in range(4)]) for i, b in enumerate(("sad", "half", "sea"))}
    bounds = [s.min() + 250 for s in S["sad"]]                       # winner's SAD + a typical rate term
    for r0 in range(0, 2 * R + 1, 8):
        for shape in ("strided", "compact"):
            colsets = [np.arange(s, 2 * R + 1, 4) for s in range(4)] if shape == "strided" else [np.arange(c, min(c + 32, 2 * R + 1)) for c in range(0, 2 * R + 1, 32)]
            for cols in colsets:
                total[shape] += 1
                for b in ("sad", "half", "sea"):
                    count[(shape, b)] += all(S[b][k][r0:r0 + 8][:, cols].min() > bounds[k] for k in range(9))
for shape in ("strided", "compact"):
    print(shape, {b: round(count[(shape, b)] / total[shape], 3) for b in ("sad", "half", "sea")}, "of", total[shape], "batches skip all nine slots")
