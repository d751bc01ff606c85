"""Synthetic pictures and job lists for the five BASELINE.json configurations.

Luma = box-blurred seeded uniform noise generated on a 4x finer grid and sampled with a per-frame
global translation given in QUARTER pels (so true motion has half- and quarter-pel parts), contrast
stretched, plus N(0, 2) noise.  Chroma is flat 128 and never enters this path.

Jobs mirror what MbEncoder asks of estimateBlockWithStart for one (current, reference) pair: for every
macroblock one call per partition of the configured partition set, with a start / predictor field that
is either all-zero or seeded uniform in [-32, +32] quarter-pel per block (both are benchmarked).
"""
import math

import numpy as np

from .abi import (PIC_TASK_DTYPE, MC_JOB_DTYPE, MC_WEIGHT_UNI, MC_WEIGHT_BI, JOB_DTYPE, MODE_16x16, MODE_16x8, MODE_8x16, BLK_8x8, BLK_8x4, BLK_4x8, BLK_4x4, DF_SAD, DF_HADAMARD)

# (mode, blk) of the 9-partition set {16x16, 2 x 16x8, 2 x 8x16, 4 x 8x8} and of all 41 partitions
PARTS_9 = [(MODE_16x16, 0), (MODE_16x8, 0), (MODE_16x8, 8), (MODE_8x16, 0), (MODE_8x16, 2),
           (BLK_8x8, 0), (BLK_8x8, 2), (BLK_8x8, 8), (BLK_8x8, 10)]
PARTS_41 = list(PARTS_9)
for _b8 in (0, 2, 8, 10):
    PARTS_41 += [(BLK_8x4, _b8), (BLK_8x4, _b8 + 4), (BLK_4x8, _b8), (BLK_4x8, _b8 + 1),
                 (BLK_4x4, _b8), (BLK_4x4, _b8 + 1), (BLK_4x4, _b8 + 4), (BLK_4x4, _b8 + 5)]

# The five configurations of BASELINE.json (picture size of the layer whose ME is run, search range,
# sub-pel distortion, partition set, pictures of one GOP / number of reference frames).
CONFIGS = {
    "C1": dict(name="QCIF 176x144 SR16 QPel", width=176, height=144, search_range=16, sub_dfunc=DF_HADAMARD, parts=PARTS_9, gop=8, refs=1),
    "C2": dict(name="CIF 352x288 (EL of QCIF->CIF) SR32 SAD+SATD", width=352, height=288, search_range=32, sub_dfunc=DF_HADAMARD, parts=PARTS_9, gop=8, refs=1),
    "C3": dict(name="720p 1280x720 SR32 all partitions", width=1280, height=720, search_range=32, sub_dfunc=DF_HADAMARD, parts=PARTS_41, gop=16, refs=1),
    "C4": dict(name="1080p 1920x1088 SR64 QPel", width=1920, height=1088, search_range=64, sub_dfunc=DF_HADAMARD, parts=PARTS_9, gop=16, refs=1),
    "C5": dict(name="1080p 1920x1088 (top layer of 540p->1080p) SR64 4 refs", width=1920, height=1088, search_range=64, sub_dfunc=DF_HADAMARD, parts=PARTS_9, gop=16, refs=4),
}


def lambda_for_qp(qp):
    """lambda = 0.85 * 2^(QP/3 - 4) (ENC/GOPEncoder.cpp:2871-2875 with gamma = 1)."""
    return 0.85 * 2.0 ** (qp / 3.0 - 4.0)


def cost_factor_for_qp(qp):
    return int(math.floor(65536.0 * math.sqrt(lambda_for_qp(qp)))) & 0xFFFFFFFF


def _box_blur(a, k):
    """k x k box mean with edge replication, via cumulative sums (fast for 4x-upsampled 1080p)."""
    p = k // 2
    a = np.pad(a, p, mode="edge").astype(np.float64)
    c = np.cumsum(np.cumsum(a, 0), 1)
    c = np.pad(c, ((1, 0), (1, 0)))
    return (c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k]) / (k * k)


class Sequence:
    """A seeded synthetic sequence; frame(n) is deterministic in (seed, n)."""

    def __init__(self, width, height, seed, motion_qpel=(9, 5), blur=21, guard=96):
        self.width, self.height, self.seed = width, height, seed
        self.motion = motion_qpel
        self.guard = guard
        rng = np.random.default_rng(seed)
        hh, ww = 4 * (height + 2 * guard), 4 * (width + 2 * guard)
        # noise on a coarse grid (one value per 2x2 pixels), upsampled 8x, then blurred on the fine grid
        coarse = rng.integers(0, 256, (hh // 8 + 1, ww // 8 + 1)).astype(np.float64)
        fine = np.kron(coarse, np.ones((8, 8)))[:hh, :ww]
        fine = _box_blur(fine, blur)
        lo, hi = fine.min(), fine.max()
        self.tex = ((fine - lo) / (hi - lo) * 255.0).astype(np.float32)

    def frame(self, n, gop=0):
        """8-bit luma of picture n: texture translated by n * motion quarter-pels, + N(0,2).  `gop` selects another noise
        realisation of the same motion (picture n of GOP `gop`: independent pictures for multi-GOP batches)."""
        g = 4 * self.guard
        dx, dy = n * self.motion[0], n * self.motion[1]
        if abs(dx) > g or abs(dy) > g:
            raise ValueError("picture index outside the guard band of the synthetic texture")
        y0, x0 = g + dy, g + dx
        f = self.tex[y0:y0 + 4 * self.height:4, x0:x0 + 4 * self.width:4].astype(np.float64)
        rng = np.random.default_rng(self.seed * 1000003 + n + 4099 * gop)
        f = f + rng.normal(0.0, 2.0, f.shape)
        return np.clip(np.rint(f), 0, 255).astype(np.uint8)


class LocalMotionSequence:
    """Two textures moving differently, stitched along a grid of `tile`-sized squares that is NOT aligned with the macroblock
    grid: macroblocks straddling a seam prefer split partitions, neighbouring macroblocks get different MVs (so that the MV
    predictors of the picture-level search are non-trivial)."""

    def __init__(self, width, height, seed, tile=20):
        self.a = Sequence(width, height, seed, motion_qpel=(9, 5))
        self.b = Sequence(width, height, seed + 1, motion_qpel=(-14, 11))
        yy, xx = np.mgrid[0:height, 0:width]
        self.mask = (((xx + 7) // tile + (yy + 3) // tile) & 1).astype(bool)

    def frame(self, n):
        return np.where(self.mask, self.a.frame(n), self.b.frame(n)).astype(np.uint8)


def chroma_planes(width, height, seed):
    """Seeded (Cb, Cr) planes of a 4:2:0 picture: smooth texture with full 8-bit range (YUV_SAD search, chroma MC)."""
    rng = np.random.default_rng(seed * 7919 + 17)
    out = []
    for _ in range(2):
        coarse = rng.integers(0, 256, (height // 8 + 2, width // 8 + 2)).astype(np.float64)
        fine = _box_blur(np.kron(coarse, np.ones((4, 4)))[:height // 2, :width // 2], 5)
        fine = (fine - fine.min()) / max(1e-9, fine.max() - fine.min()) * 255.0 + rng.normal(0.0, 2.0, fine.shape)
        out.append(np.clip(np.rint(fine), 0, 255).astype(np.uint8))
    return out[0], out[1]


def make_mc_jobs(width, height, n, n_slots, seed=0, kind="uni", mv_range=96, sizes=((16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8), (4, 4))):
    """n motion-compensation jobs with seeded random macroblocks, block sizes and quarter-pel MVs (every fractional phase;
    |mv| up to mv_range quarter-pels plus a few far outliers that the per-MB clamp catches).

    kind: "uni" (one list, plain), "bi" (two lists, (a+b+1)>>1), "uni_w" (explicit weights, one list),
          "bi_w" (explicit weights, two lists), "bi_implicit" (implicit weights: w0 + w1 = 64, denominators 5)."""
    rng = np.random.default_rng(seed)
    mbw, mbh = width // 16, height // 16
    jobs = np.zeros(n, MC_JOB_DTYPE)
    jobs["mb_x"] = rng.integers(0, mbw, n)
    jobs["mb_y"] = rng.integers(0, mbh, n)
    # corners and edges first
    fixed = [(0, 0), (mbw - 1, 0), (0, mbh - 1), (mbw - 1, mbh - 1)]
    for i, (x, y) in enumerate(fixed[:n]):
        jobs["mb_x"][i], jobs["mb_y"][i] = x, y
    for i in range(n):
        sx, sy = sizes[rng.integers(0, len(sizes))]
        if kind.startswith("bi") and (sx < 8 or sy < 8) and kind == "bi":
            sx, sy = 8, 8                      # the reference mixes bi-pred sub-8x8 blocks per 8x8 (compensateSubMb)
        bx = rng.integers(0, (16 - sx) // 4 + 1)
        by = rng.integers(0, (16 - sy) // 4 + 1)
        jobs["blk"][i] = by * 4 + bx
        jobs["size_x"][i], jobs["size_y"][i] = sx, sy
    mv = rng.integers(-mv_range, mv_range + 1, (n, 2, 2))
    far = rng.random(n) < 0.05
    mv[far] = rng.integers(-4000, 4001, (int(far.sum()), 2, 2))
    jobs["mv"] = mv.astype(np.int16)
    two = kind.startswith("bi")
    jobs["ref_slot"][:, 0] = rng.integers(0, n_slots, n)
    jobs["ref_slot"][:, 1] = rng.integers(0, n_slots, n) if two else -1
    if kind == "uni":
        flip = rng.random(n) < 0.3                                       # some blocks predicted from LIST_1 only
        jobs["ref_slot"][flip, 1] = jobs["ref_slot"][flip, 0]
        jobs["ref_slot"][flip, 0] = -1
    jobs["log_denom"] = 0
    if kind in ("uni_w", "bi_w"):
        jobs["weighting"] = MC_WEIGHT_UNI | MC_WEIGHT_BI if kind == "bi_w" else MC_WEIGHT_UNI
        jobs["log_denom"][:, 0] = rng.integers(0, 8, n)
        jobs["log_denom"][:, 1] = rng.integers(0, 8, n)
        for l in range(2):
            for comp in range(3):
                d = jobs["log_denom"][:, 1 if comp else 0].astype(np.int64)
                hi = np.minimum(127, (1 << d) * 2)
                # explicit weights: -128..127 in general; two lists: w0 + w1 <= 127 / 128 (COM/SampleWeighting.cpp:443-444)
                jobs["weight"][:, l, comp] = rng.integers(-hi // (2 if two else 1), hi // (2 if two else 1) + 1)
                jobs["offset"][:, l, comp] = rng.integers(-20, 21, n)
        jobs["weight_flag"] = rng.integers(0, 2, (n, 2))
    elif kind == "bi_implicit":
        jobs["weighting"] = MC_WEIGHT_BI
        jobs["log_denom"] = 5
        w1 = rng.integers(-64, 129, n)                                  # implicit: w1 = DistScaleFactor >> 2, w0 = 64 - w1
        for comp in range(3):
            jobs["weight"][:, 1, comp] = w1
            jobs["weight"][:, 0, comp] = 64 - w1
    return jobs


def make_jobs(width, height, parts, cur_slot, ref_slot, qp=30, field="random", seed=0, mb_subset=None, bits_in=0):
    """Job records for every macroblock x partition of one (current, reference) pair.

    field: "zero"   -> start = pred = (0, 0) for every block
           "random" -> start = pred = seeded uniform [-32, 32] quarter-pel per block
           "split"  -> independent random start and predictor (inter-layer motion prediction: the start is
                       the base-layer MV while the predictor is the spatial MVP, ENC/MbEncoder.cpp:4516-4529)
    """
    mbw, mbh = width // 16, height // 16
    mbs = [(y, x) for y in range(mbh) for x in range(mbw)] if mb_subset is None else list(mb_subset)
    n = len(mbs) * len(parts)
    jobs = np.zeros(n, JOB_DTYPE)
    mb = np.repeat(np.asarray(mbs, dtype=np.int16), len(parts), axis=0)
    jobs["mb_y"], jobs["mb_x"] = mb[:, 0], mb[:, 1]
    jobs["mode"] = np.tile(np.asarray([p[0] for p in parts], np.uint8), len(mbs))
    jobs["blk"] = np.tile(np.asarray([p[1] for p in parts], np.uint8), len(mbs))
    jobs["cur_slot"], jobs["ref_slot"] = cur_slot, ref_slot
    jobs["search_range"] = 0
    jobs["cost_factor"] = cost_factor_for_qp(qp)
    jobs["bits_in"] = bits_in
    jobs["org_index"] = -1
    rng = np.random.default_rng(seed)
    if field == "zero":
        pass
    elif field == "random":
        v = rng.integers(-32, 33, (n, 2)).astype(np.int16)
        jobs["start_hor"], jobs["start_ver"] = v[:, 0], v[:, 1]
        jobs["pred_hor"], jobs["pred_ver"] = v[:, 0], v[:, 1]
    elif field == "split":
        v = rng.integers(-32, 33, (n, 4)).astype(np.int16)
        jobs["start_hor"], jobs["start_ver"] = v[:, 0], v[:, 1]
        jobs["pred_hor"], jobs["pred_ver"] = v[:, 2], v[:, 3]
    else:
        raise ValueError(field)
    return jobs


def make_fast_search_extras(jobs, search_range, seed=0, p_el=0.2, p_bipred=0.25, p_iter=0.15, iter_range=8, far=0.1):
    """What the fast searches (SPIRAL / LOG / FAST / TZ) read beyond the plain job: returns (jobs', cands).
    jobs': a copy with the per-call flags in the top bits of `mode` (JSVM_JOB_EL, JSVM_JOB_BIPRED) on a seeded share of the jobs and
    a non-zero per-call search range (the bi-prediction refinement) on another share; cands: (n, 3, 2) MVs of the neighbours A, B, C
    near the predictor, a share of them far away (outside the search window, clamped per macroblock) -- but inside the span of
    the reference's bit-count table, |4x - pred| < (4 SR + 4) * 16 quarter-pels (ENC/MotionEstimationCost.cpp:33-41): beyond it
    the reference reads outside the table (undefined), while the device evaluates the Exp-Golomb length analytically."""
    rng = np.random.default_rng(seed)
    n = len(jobs)
    out = jobs.copy()
    flags = np.where(rng.random(n) < p_el, 0x40, 0) | np.where(rng.random(n) < p_bipred, 0x80, 0)
    out["mode"] = out["mode"] | flags.astype(np.uint8)
    out["search_range"] = np.where(rng.random(n) < p_iter, rng.integers(1, iter_range + 1, n), 0).astype(np.int16)
    base = np.stack([out["pred_hor"], out["pred_ver"]], -1).astype(np.int64)[:, None, :]
    cands = base + rng.integers(-24, 25, (n, 3, 2))
    wild = rng.random((n, 3)) < far
    span = min(600, (4 * search_range + 4) * 16 - 48)
    cands[wild] = (np.broadcast_to(base, cands.shape)[wild] + rng.integers(-span, span + 1, (int(wild.sum()), 2)))
    return out, np.clip(cands, -32768, 32767).astype(np.int16)


def make_pic_tasks(cur_slots, ref_slots, qp=30):
    """Picture-level search tasks: cur_slots[i] is searched in ref_slots[i] (a list of 1..8 reference slots, list-0 order)."""
    tasks = np.zeros(len(cur_slots), PIC_TASK_DTYPE)
    for i, (c, refs) in enumerate(zip(cur_slots, ref_slots)):
        tasks["cur_slot"][i], tasks["n_refs"][i] = c, len(refs)
        tasks["ref_slot"][i, :len(refs)] = refs
    tasks["cost_factor"] = cost_factor_for_qp(qp)
    return tasks


def make_resize_case(ref_w, ref_h, cur_w, cur_h, scaled_w=None, scaled_h=None, left=0, top=0, level_idc=40, seed=0, phases=(-1, 0, -1, 0)):
    """Inter-layer resampling inputs: ResizeParams + a seeded base-layer residual (luma, Cb, Cr: smooth signal with sharp steps and
    many zeros, range about +-255) + per-macroblock transform-size flags.  phases = (chroma_phase_x, chroma_phase_y,
    ref_chroma_phase_x, ref_chroma_phase_y), JSVM defaults -1 / 0."""
    from .abi import ResizeParams
    p = ResizeParams()
    p.level_idc = level_idc
    p.frame_width, p.frame_height, p.ref_frame_width, p.ref_frame_height = cur_w, cur_h, ref_w, ref_h
    p.scaled_ref_width, p.scaled_ref_height = scaled_w or cur_w, scaled_h or cur_h
    p.left_offset, p.top_offset = left, top
    p.chroma_phase_x, p.chroma_phase_y, p.ref_chroma_phase_x, p.ref_chroma_phase_y = phases
    rng = np.random.default_rng(seed)

    def plane(h, w):
        a = _box_blur(rng.normal(0, 90, (h, w)), 3)
        a[rng.random((h, w)) < 0.3] = 0
        a += (rng.random((h, w)) < 0.02) * rng.integers(-255, 256, (h, w))
        return np.clip(np.rint(a), -255, 255).astype(np.int16)
    t8 = (rng.random((ref_h // 16, ref_w // 16)) < 0.5).astype(np.uint8)
    return p, plane(ref_h, ref_w), plane(ref_h // 2, ref_w // 2), plane(ref_h // 2, ref_w // 2), t8


def make_motion_params(resize, slice_type, ref_layer_dq_id=0, direct_8x8_inference=1, check_residual_pred=1, mv_threshold=1):
    from .abi import MotionUpsampleParams
    p = MotionUpsampleParams()
    p.resize = resize
    p.slice_type, p.ref_layer_dq_id, p.direct_8x8_inference = slice_type, ref_layer_dq_id, direct_8x8_inference
    p.check_residual_pred, p.mv_threshold = check_residual_pred, mv_threshold
    return p


def make_base_motion(ref_w, ref_h, seed=0, slice_type=0, p_intra=0.15):
    """A seeded base-layer motion field (MB_MOTION_DTYPE, (ref_h/16, ref_w/16)): every macroblock mode and sub-macroblock mode,
    intra macroblocks (alone and in clusters), one or two lists (B: blocks predicted from one list only, B_Skip / direct), reference
    indices 1..2, MVs that are equal, 1 quarter-pel apart (merged by the G.8.6.1.2 threshold) or far apart between partitions."""
    from .abi import MB_MOTION_DTYPE
    rng = np.random.default_rng(seed)
    mbh, mbw = ref_h // 16, ref_w // 16
    f = np.zeros((mbh, mbw), MB_MOTION_DTYPE)
    f["slice_type"] = slice_type
    f["blk_mode"] = 8
    f["ref_idx"] = -1
    lists = 2 if slice_type == 1 else 1
    part8 = {1: [0, 0, 0, 0], 2: [0, 0, 1, 1], 3: [0, 1, 0, 1], 4: [0, 1, 2, 3], 5: [0, 1, 2, 3]}
    part4 = {8: [0, 0, 0, 0], 9: [0, 0, 1, 1], 10: [0, 1, 0, 1], 11: [0, 1, 2, 3], 0: [0, 0, 0, 0]}
    base_mv = rng.integers(-40, 41, 2)
    for y in range(mbh):
        for x in range(mbw):
            m = f[y, x]
            if rng.random() < p_intra or (x > 0 and f[y, x - 1]["mb_mode"] >= 6 and rng.random() < 0.4):
                m["mb_mode"] = rng.choice([6, 7, 31])
                continue
            mode = int(rng.choice([0, 1, 2, 3, 4, 4, 5]))
            if mode == 5 and slice_type == 1:
                mode = 4
            m["mb_mode"] = mode
            if mode in (4, 5):
                m["blk_mode"] = rng.choice([8, 9, 10, 11] + ([0] if slice_type == 1 else []), 4)
            if rng.random() < 0.3:
                base_mv = base_mv + rng.integers(-30, 31, 2)
            for l in range(lists):
                p8 = part8.get(mode, [0, 0, 0, 0])
                n8 = max(p8) + 1
                refs = rng.integers(1, 3, n8)
                if lists == 2:                              # some partitions use one list only
                    refs = np.where(rng.random(n8) < 0.25, -1, refs)
                if mode == 5:
                    refs[:] = 1
                mvs8 = base_mv + rng.choice([0, 0, 1, -1, 12], (n8, 2))
                for b8 in range(4):
                    m["ref_idx"][l][b8] = refs[p8[b8]]
                    p4 = part4[int(m["blk_mode"][b8])] if mode in (4, 5) else [0, 0, 0, 0]
                    sub = mvs8[p8[b8]] + rng.choice([0, 0, 1, -1, 9], (max(p4) + 1, 2))
                    for b4 in range(4):
                        blk = ((b8 >> 1) * 2 + (b4 >> 1)) * 4 + (b8 & 1) * 2 + (b4 & 1)
                        m["mv"][l][blk] = sub[p4[b4]] if refs[p8[b8]] > 0 else 0
            if lists == 2 and (m["ref_idx"] < 1).all():     # a B macroblock predicts from at least one list
                m["ref_idx"][0] = 1
    return f


def hierarchical_b_pairs(gop, refs=1):
    """(current, reference) picture pairs of one GOP, grouped by temporal level (coarsest first).

    Hierarchical-B coding order of ENC/GOPEncoder.cpp:530-640: the key picture `gop` references picture 0;
    level t >= 1 holds the pictures at odd multiples of gop / 2^t, each referencing its two neighbours at
    distance gop / 2^t (list 0 = past, list 1 = future).  With refs > 1 further already-coded pictures of
    lower levels are added, nearest first.  Pairs inside one level are independent of each other.
    """
    levels = []
    coded = [0]
    levels.append([(gop, 0)])
    coded.append(gop)
    step = gop // 2
    while step >= 1:
        lvl = []
        for cur in range(step, gop, 2 * step):
            past = sorted([p for p in coded if p < cur], key=lambda p: cur - p)[:refs]
            fut = sorted([p for p in coded if p > cur], key=lambda p: p - cur)[:refs]
            lvl += [(cur, r) for r in past] + [(cur, r) for r in fut]
        levels.append(lvl)
        coded += list(range(step, gop, 2 * step))
        step //= 2
    return levels
