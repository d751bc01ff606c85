"""Oracle (C restatement) against the compiled, unmodified reference on fresh seeded inputs.
Skipped where oracle/_ref was not built (the GPU box; the golden fixtures cover it there)."""
import numpy as np
import pytest


@pytest.mark.parametrize("sub_dfunc,parts_name,field,sr", [(2, "PARTS_9", "random", 16), (0, "PARTS_41", "split", 12), (2, "PARTS_9", "zero", 32)])
def test_search_and_planes_identical(oracle, reference, synth, sub_dfunc, parts_name, field, sr):
    W, H = 96, 80
    seq = synth.Sequence(W, H, 77 + sr)
    f0, f1 = seq.frame(0), seq.frame(2)
    a, b = oracle.session(W, H, 2, sr, sub_dfunc), reference.session(W, H, 2, sr, sub_dfunc)
    for s in (a, b):
        s.set_plane(0, f0); s.set_plane(1, f1)
    assert np.array_equal(a.halfpel(0), b.halfpel(0)) and np.array_equal(a.halfpel(1), b.halfpel(1))
    assert np.array_equal(a.fullpel(1), b.fullpel(1))
    jobs = synth.make_jobs(W, H, getattr(synth, parts_name), 1, 0, qp=31, field=field, seed=sr, bits_in=1)
    ra, pa = a.search(jobs, want_preds=True)
    rb, pb = b.search(jobs, want_preds=True)
    for f in ra.dtype.names:
        assert np.array_equal(ra[f], rb[f]), f
    assert np.array_equal(pa, pb)


def test_search_range_64_identical(oracle, reference, synth):
    """SR 64 (C4 / C5): 129 x 129 windows clipped by the picture on every side, CIF, a macroblock subset with corners and edges."""
    W, H, sr = 352, 288, 64
    seq = synth.Sequence(W, H, 141)
    f0, f1 = seq.frame(0), seq.frame(3)
    a, b = oracle.session(W, H, 2, sr, 2), reference.session(W, H, 2, sr, 2)
    for s in (a, b):
        s.set_plane(0, f0); s.set_plane(1, f1)
    mbs = [(0, 0), (0, 21), (17, 0), (17, 21), (0, 10), (17, 7), (9, 0), (6, 21), (3, 3), (9, 12), (14, 18), (7, 5)]
    for field, seed in (("random", 1), ("split", 2), ("zero", 3)):
        jobs = synth.make_jobs(W, H, synth.PARTS_9, 1, 0, qp=35, field=field, seed=seed, mb_subset=mbs, bits_in=2)
        ra, pa = a.search(jobs, want_preds=True)
        rb, pb = b.search(jobs, want_preds=True)
        for f in ra.dtype.names:
            assert np.array_equal(ra[f], rb[f]), (field, f)
        assert np.array_equal(pa, pb)


def test_extreme_content(oracle, reference, synth):
    """Flat, saturated and checkerboard pictures: SAD ties everywhere, clipping in the 6-tap filter."""
    W, H, sr = 64, 48, 8
    yy, xx = np.mgrid[0:H, 0:W]
    pics = [np.zeros((H, W), np.uint8), np.full((H, W), 255, np.uint8), (((xx + yy) & 1) * 255).astype(np.uint8),
            (((xx // 4 + yy // 4) & 1) * 255).astype(np.uint8)]
    for cur in pics:
        for ref in pics:
            a, b = oracle.session(W, H, 2, sr, 2), reference.session(W, H, 2, sr, 2)
            for s in (a, b):
                s.set_plane(0, ref); s.set_plane(1, cur)
            assert np.array_equal(a.halfpel(0), b.halfpel(0))
            jobs = synth.make_jobs(W, H, synth.PARTS_9, 1, 0, qp=26, field="random", seed=1)
            ra, rb = a.search(jobs), b.search(jobs)
            for f in ra.dtype.names:
                assert np.array_equal(ra[f], rb[f]), f


@pytest.mark.parametrize("full,sub", [(3, 2), (3, 1), (0, 1), (1, 2)])
def test_optional_distortion_functions_identical(oracle, reference, synth, full, sub):
    """a5: YUV-SAD (chroma planes, ( x>>1, y>>1 ) displacement) and the mixed cost-factor classes, fresh seeds."""
    import oracle_lib
    W, H, sr = 96, 80, 12
    seq = synth.Sequence(W, H, 300 + full)
    sessions = [oracle_lib.CpuSession(be, W, H, 2, sr, sub, full) for be in (oracle, reference)]
    f_sad, f_sse = reference.cost_factor(synth.lambda_for_qp(34), 1), reference.cost_factor(synth.lambda_for_qp(34), 0)
    for s in sessions:
        for slot in range(2):
            s.set_plane(slot, seq.frame(slot))
            s.set_chroma(slot, *synth.chroma_planes(W, H, 70 + slot))
        if (full == 1) != (sub == 1):
            s.map_cost_factor(*((f_sse, f_sad) if full == 1 else (f_sad, f_sse)))
    jobs = synth.make_jobs(W, H, synth.PARTS_41, 1, 0, qp=34, field="split", seed=9, bits_in=1)
    jobs["cost_factor"] = f_sse if full == 1 else f_sad
    (ra, pa), (rb, pb) = [s.search(jobs, want_preds=True) for s in sessions]
    for f in ra.dtype.names:
        assert np.array_equal(ra[f], rb[f]), f
    assert np.array_equal(pa, pb)


@pytest.mark.parametrize("kind", ["uni", "bi", "uni_w", "bi_w", "bi_implicit"])
def test_compensate_identical(oracle, reference, synth, kind):
    """f3: predBlk / chroma MC / SampleWeighting of the C port against the reference's MotionCompensation, fresh seeds."""
    W, H = 96, 80
    seq = synth.Sequence(W, H, 500)
    sessions = [be.session(W, H, 2, 8, 0) for be in (oracle, reference)]
    for s in sessions:
        for slot in range(2):
            s.set_plane(slot, seq.frame(slot))
            s.set_chroma(slot, *synth.chroma_planes(W, H, 80 + slot))
    jobs = synth.make_mc_jobs(W, H, 1500, 2, seed=17, kind=kind, mv_range=160)
    a, b = [s.compensate(jobs) for s in sessions]
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("search_mode", [1, 2, 3, 4])
@pytest.mark.parametrize("sr,fast_bi,full", [(16, 0, 0), (32, 1, 0), (8, 0, 3), (16, 1, 2)])
def test_fast_searches_identical(oracle, reference, synth, search_mode, sr, fast_bi, full):
    """f2: SPIRAL / LOG / FAST / TZ searches of the C port against the reference's, with neighbour-MV candidates, EL ranges,
    bi-predictive calls and per-call ranges; windows clipped by the picture on every side."""
    import oracle_lib
    W, H = 176, 144
    if search_mode == 1:
        W, H = 352, 288                              # the spiral search needs room for its un-clipped window
        if sr > 16:
            pytest.skip("spiral window larger than the clamp interval of a CIF macroblock row")
    if search_mode in (2, 3) and full == 3:
        pytest.skip("xPelLogSearch never sets pUSearch / pVSearch (ENC/MotionEstimation.cpp:472-709): YUV_SAD crashes the reference there")
    seq = synth.Sequence(W, H, 600 + sr)
    sessions = [oracle_lib.CpuSession(be, W, H, 2, sr, 2, full, search_mode) for be in (oracle, reference)]
    jobs = synth.make_jobs(W, H, synth.PARTS_41, 1, 0, qp=32, field="split", seed=search_mode, bits_in=1,
                           mb_subset=None if W == 176 else [(0, 0), (0, 21), (17, 0), (17, 21), (8, 10), (3, 15), (12, 4)])
    jobs, cands = synth.make_fast_search_extras(jobs, sr // 2, seed=sr + search_mode)
    outs = []
    for s in sessions:
        for slot in range(2):
            s.set_plane(slot, seq.frame(2 * slot))
            s.set_chroma(slot, *synth.chroma_planes(W, H, 90 + slot))
        s.set_fast_search_params(sr // 2, fast_bi)
        s.set_mv_candidates(cands)
        outs.append(s.search(jobs, want_preds=True))
    (ra, pa), (rb, pb) = outs
    for f in ra.dtype.names:
        bad = np.nonzero(ra[f] != rb[f])[0]
        assert len(bad) == 0, (f, len(bad), jobs[bad[0]], ra[bad[0]], rb[bad[0]])
    assert np.array_equal(pa, pb)


@pytest.mark.parametrize("W,H,sr,n_refs,sub", [(96, 80, 12, 1, 2), (176, 144, 16, 3, 0), (64, 48, 8, 2, 2)])
def test_picture_level_search_identical(oracle, reference, synth, W, H, sr, n_refs, sub):
    """f1: the C port's MV prediction + per-macroblock loops against the reference's MbDataAccess::getMvPredictor driven on a real
    MbDataCtrl (every partition's MV, reference index and cost, the macroblock mode and the final 4x4 motion field)."""
    seq = synth.LocalMotionSequence(W, H, 700 + W)
    sessions = [be.session(W, H, n_refs + 2, sr, sub) for be in (oracle, reference)]
    for s in sessions:
        for slot in range(n_refs + 2):
            s.set_plane(slot, seq.frame(slot))
    tasks = synth.make_pic_tasks([n_refs, n_refs + 1], [list(range(n_refs)), list(range(1, n_refs + 1))[::-1]], qp=33)
    a, b = [s.estimate_pictures(tasks) for s in sessions]
    for f in a.dtype.names:
        bad = np.argwhere(a[f] != b[f])
        assert len(bad) == 0, (f, len(bad), bad[0], a[tuple(bad[0][:3])], b[tuple(bad[0][:3])])
    assert len(np.unique(a["mode"])) >= 3 and (a["part_mv"] != 0).any()


RESIZE_CASES = [
    dict(ref_w=176, ref_h=144, cur_w=352, cur_h=288),                                               # dyadic QCIF -> CIF (C2)
    dict(ref_w=176, ref_h=144, cur_w=352, cur_h=288, level_idc=30),                                 # 16-bit position arithmetic
    dict(ref_w=176, ref_h=144, cur_w=272, cur_h=224, scaled_w=264, scaled_h=216, left=4, top=6),    # ESS ratio 3:2 with cropping window
    dict(ref_w=160, ref_h=128, cur_w=352, cur_h=288, scaled_w=300, scaled_h=250, left=32, top=16, phases=(0, 1, -1, -1)),   # arbitrary ratio, other chroma phases
    dict(ref_w=96, ref_h=80, cur_w=96, cur_h=80),                                                   # same resolution: identity
]


@pytest.mark.parametrize("case", RESIZE_CASES, ids=[str(i) for i in range(len(RESIZE_CASES))])
def test_residual_upsampling_identical(oracle, reference, synth, case):
    """f4: DownConvert::xResidualUpsampling of the C port against the reference's (luma with 4x4 / 8x8 transform blocks, chroma)."""
    p, y, cb, cr, t8 = synth.make_resize_case(seed=len(str(case)), **case)
    a = oracle.upsample_residual(p, y, cb, cr, t8)
    if case["cur_w"] == case["ref_w"]:
        # no resolution change and no cropping: DownConvert::residualUpsampling does nothing (COM/DownConvert.cpp:623-631, the caller
        # uses the base residual as it is); the resampling arithmetic degenerates to the identity
        assert np.array_equal(a[0], y) and np.array_equal(a[1], cb) and np.array_equal(a[2], cr)
        return
    b = reference.upsample_residual(p, y, cb, cr, t8)
    for k, (u, v) in enumerate(zip(a, b)):
        assert np.array_equal(u, v), (k, np.argwhere(u != v)[:3])
    assert (a[0] != 0).any()


MOTION_CASES = [
    (dict(ref_w=176, ref_h=144, cur_w=352, cur_h=288), 0, 0, 1),                                             # dyadic, P
    (dict(ref_w=176, ref_h=144, cur_w=352, cur_h=288), 1, 0, 1),                                             # dyadic, B
    (dict(ref_w=176, ref_h=144, cur_w=352, cur_h=288), 1, 16, 0),                                            # B, ref layer DQId != 0, no 8x8 inference
    (dict(ref_w=176, ref_h=144, cur_w=272, cur_h=224, scaled_w=264, scaled_h=216, left=4, top=6), 0, 0, 1),   # ESS 3:2 with cropping window
    (dict(ref_w=176, ref_h=144, cur_w=272, cur_h=224, scaled_w=264, scaled_h=216, left=4, top=6), 1, 0, 1),
    (dict(ref_w=160, ref_h=128, cur_w=352, cur_h=288, scaled_w=300, scaled_h=250, left=32, top=16, level_idc=30), 1, 0, 0),
    (dict(ref_w=176, ref_h=144, cur_w=176, cur_h=144), 1, 0, 1),                                             # same resolution (quality layer)
    (dict(ref_w=176, ref_h=144, cur_w=352, cur_h=288, scaled_w=352, scaled_h=288), 2, 0, 1),                 # I slice: no lists
]


@pytest.mark.parametrize("case,slice_type,dq,d8", MOTION_CASES, ids=[str(i) for i in range(len(MOTION_CASES))])
def test_motion_upsampling_identical(oracle, reference, synth, case, slice_type, dq, d8):
    """f4: MotionUpsampling::resample of the C port against MbDataCtrl::upsampleMotion (every output field of every macroblock)."""
    resize = synth.make_resize_case(seed=1, **case)[0]
    base = synth.make_base_motion(case["ref_w"], case["ref_h"], seed=len(str(case)) + slice_type, slice_type=min(slice_type, 1))
    p = synth.make_motion_params(resize, slice_type, dq, d8)
    a, b = oracle.upsample_motion(p, base), reference.upsample_motion(p, base)
    for f in a.dtype.names:
        bad = np.argwhere(a[f] != b[f])
        assert len(bad) == 0, (f, len(bad), bad[0], a[tuple(bad[0][:2])], b[tuple(bad[0][:2])])
    if slice_type < 2:
        assert len(np.unique(a["mb_mode"])) >= 4 and (a["mv"] != 0).any()
