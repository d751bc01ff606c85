"""GPU parity: the CUDA path (through the C ABI) against the CPU checker, bit-exact.

Every comparison is exact equality (integer / byte / index work): half-pel planes byte for byte,
and per job the whole result record -- quarter-pel MV, full-pel MV, full-pel SAD, min SAD incl. rate,
MV bits, bits, cost, best mode -- plus the 16x16 prediction block of compensateBlock.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIELDS = ["mv_hor", "mv_ver", "full_hor", "full_ver", "sad_fullpel", "min_sad", "mv_bits", "bits", "cost", "best_mode"]


def assert_results_equal(got, want, jobs=None):
    for f in FIELDS:
        bad = np.nonzero(got[f] != want[f])[0]
        if len(bad):
            i = int(bad[0])
            msg = f"field {f}: {len(bad)} of {len(got)} jobs differ; first job {i}: got {got[i]} want {want[i]}"
            if jobs is not None:
                msg += f" job {jobs[i]}"
            raise AssertionError(msg)


@pytest.fixture(scope="module")
def me(pkg):
    m = pkg.MotionEstimator(0)
    yield m
    m.close()


def _setup(me, oracle, synth, cfg, seed, n_frames=2, sub_dfunc=None):
    W, H = cfg["width"], cfg["height"]
    sub = cfg["sub_dfunc"] if sub_dfunc is None else sub_dfunc
    seq = synth.Sequence(W, H, seed)
    frames = [seq.frame(n) for n in range(n_frames)]
    me.configure(W, H, n_frames)
    me.set_params(cfg["search_range"], sub)
    me.set_fast_search_params(0, 0)
    ses = oracle.session(W, H, n_frames, cfg["search_range"], sub)
    for s, f in enumerate(frames):
        me.upload_plane(s, f)
        ses.set_plane(s, f)
    me.interp_halfpel(list(range(n_frames)))
    return ses, frames


@pytest.mark.parametrize("cfg_name", ["C1", "C2"])
def test_halfpel_plane_matches_filterFrame(me, oracle, synth, cfg_name):
    cfg = synth.CONFIGS[cfg_name]
    ses, _ = _setup(me, oracle, synth, cfg, seed=1234 + int(cfg_name[1]))
    for s in range(2):
        got, want = me.halfpel_plane(s), ses.halfpel(s)
        assert got.shape == want.shape
        assert np.array_equal(got, want), f"slot {s}: {np.count_nonzero(got != want)} samples differ, first at {np.argwhere(got != want)[0]}"


@pytest.mark.parametrize("field", ["zero", "random", "split"])
@pytest.mark.parametrize("sub_dfunc", [0, 2])
def test_c1_all_macroblocks_bit_exact(me, oracle, synth, field, sub_dfunc):
    """Config C1 (QCIF, SR 16, 9-partition set): every MB, every partition, SAD and Hadamard sub-pel."""
    cfg = synth.CONFIGS["C1"]
    ses, _ = _setup(me, oracle, synth, cfg, seed=1235, sub_dfunc=sub_dfunc)
    jobs = synth.make_jobs(cfg["width"], cfg["height"], cfg["parts"], cur_slot=1, ref_slot=0, qp=30, field=field, seed=7, bits_in=3)
    got, gp = me.search(jobs, want_preds=True)
    want, wp = ses.search(jobs, want_preds=True)
    assert_results_equal(got, want, jobs)
    assert np.array_equal(gp, wp)


def test_c1_all_41_partitions_bit_exact(me, oracle, synth):
    """4x4-granular kernel: all 41 partitions of every MB."""
    cfg = synth.CONFIGS["C1"]
    ses, _ = _setup(me, oracle, synth, cfg, seed=1236)
    jobs = synth.make_jobs(cfg["width"], cfg["height"], synth.PARTS_41, cur_slot=1, ref_slot=0, qp=33, field="split", seed=11)
    got, gp = me.search(jobs, want_preds=True)
    want, wp = ses.search(jobs, want_preds=True)
    assert_results_equal(got, want, jobs)
    assert np.array_equal(gp, wp)


# ---- golden vectors produced by the unmodified reference (tests/golden/make_golden.py) ----
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["tiny_all41_had", "tiny_9_sad", "qcif_9_had", "qcif_9_zero_sr32", "cif_9_sr64", "org_override_8bit",
                                  "org_override_16bit", "fullpel_had_had", "fullpel_had_sad_all41", "fullpel_ssd_ssd",
                                  "fullpel_ssd_ssd_16bit", "fullpel_had_had_16bit",
                                  "yuvsad_had_all41", "yuvsad_sad_org_chroma", "yuvsad_had_org_lumaonly", "yuvsad_ssd_mixed",
                                  "sad_ssd_mixed", "ssd_had_mixed",
                                  "fast_spiral_sr8", "fast_log_sr16", "fast_log_sr64_had", "fast_fme_sr32", "fast_tz_sr32",
                                  "fast_tz_sr64_fastbi", "fast_tz_sr16_yuv", "block_fastbi_sr16"])
def test_cuda_matches_reference_golden(me, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    W, H = int(g["width"]), int(g["height"])
    me.configure(W, H, 2)
    me.set_params(int(g["search_range"]), int(g["sub_dfunc"]), int(g["full_dfunc"]) if "full_dfunc" in g.files else 0,
                  int(g["search_mode"]) if "search_mode" in g.files else 0)
    if "search_mode" in g.files:                          # SPIRAL / LOG / FAST / TZ: neighbour-MV candidates, EL range, FastBiSearch
        me.set_fast_search_params(int(g["el_search_range"]), int(g["fast_bi_search"]))
        me.set_mv_candidates(g["cands"])
    else:
        me.set_fast_search_params(0, 0)
        me.set_mv_candidates(None)
    me.upload_plane(0, g["frame0"]); me.upload_plane(1, g["frame1"])
    me.interp_halfpel([0, 1])
    if "cb0" in g.files:                                  # YUV_SAD: chroma planes (+ chroma originals of prepared macroblocks)
        me.upload_chroma(0, g["cb0"], g["cr0"]); me.upload_chroma(1, g["cb1"], g["cr1"])
    if "factor_pair" in g.files:                          # the two stages use different cost-factor classes
        me.map_cost_factor(int(g["factor_pair"][0]), int(g["factor_pair"][1]))
    me.set_org_chroma(g["org_chroma"] if "org_chroma" in g.files else None)
    org = g["org_mbs"] if "org_mbs" in g.files else None
    res, preds = me.search(g["jobs"], org_mbs=org, want_preds=True)
    assert_results_equal(res, g["results"], g["jobs"])
    assert np.array_equal(preds, g["preds"])
    if "halfpel0" in g.files:
        assert np.array_equal(me.halfpel_plane(0), g["halfpel0"])
    me.set_fast_search_params(0, 0)
    me.set_mv_candidates(None)


@pytest.mark.parametrize("sub_dfunc", [0, 2])
@pytest.mark.parametrize("parts_name", ["PARTS_9", "PARTS_41"])
def test_signed_16bit_originals_bit_exact(me, oracle, synth, sub_dfunc, parts_name):
    """Residual-prediction originals (org - base-layer residual, un-clipped, N3): samples in [-255, 510].
    The SAD stages run on the clipped bytes plus the displacement-independent excess; the Hadamard uses the true
    16-bit differences.  Every field and the prediction blocks must equal the checker's."""
    cfg = synth.CONFIGS["C1"]
    W, H = cfg["width"], cfg["height"]
    ses, frames = _setup(me, oracle, synth, cfg, seed=1250, sub_dfunc=sub_dfunc)
    parts = getattr(synth, parts_name)
    mbs = [(0, 0), (0, 10), (8, 0), (8, 10), (4, 5), (3, 7), (5, 2), (7, 9)]
    rng = np.random.default_rng(77)
    cur = frames[1]
    org = np.zeros((len(mbs), 16, 16), np.int16)
    for i, (mby, mbx) in enumerate(mbs):
        blk = cur[16 * mby:16 * mby + 16, 16 * mbx:16 * mbx + 16].astype(np.int16)
        resid = rng.integers(-255, 256, (16, 16)).astype(np.int16)
        resid[rng.random((16, 16)) < 0.5] = 0                               # a mix of in-range and out-of-range samples
        org[i] = blk - resid
    assert org.min() < 0 and org.max() > 255
    jobs = synth.make_jobs(W, H, parts, cur_slot=1, ref_slot=0, qp=31, field="random", seed=13, mb_subset=mbs, bits_in=1)
    jobs["org_index"] = np.repeat(np.arange(len(mbs)), len(parts))
    got, gp = me.search(jobs, org_mbs=org, want_preds=True)
    want, wp = ses.search(jobs, org_mbs=org, want_preds=True)
    assert_results_equal(got, want, jobs)
    assert np.array_equal(gp, wp)


@pytest.mark.parametrize("parts_name", ["PARTS_9", "PARTS_41"])
def test_extreme_cost_factors_bit_exact(me, oracle, synth, parts_name):
    """Cost factors from 0 (lambda = 0) over QP 51 to values whose rate term wraps in 32 bits (N5), with per-partition starts so
    that every macroblock has strips straddling the edge of some job's window: the rate-table neutralisation of out-of-window
    rows (K2) must stay exact or step aside."""
    cfg = synth.CONFIGS["C1"]
    W, H = cfg["width"], cfg["height"]
    ses, _ = _setup(me, oracle, synth, cfg, seed=1260, sub_dfunc=0)
    parts = getattr(synth, parts_name)
    jobs = synth.make_jobs(W, H, parts, cur_slot=1, ref_slot=0, qp=30, field="random", seed=21, bits_in=2)
    factors = np.array([0, 1, 77, synth.cost_factor_for_qp(51), 7000000, 7000001, 123456789, 0x80000000, 0xFFFFFFFF], np.uint32)
    rng = np.random.default_rng(5)
    jobs["cost_factor"] = factors[rng.integers(0, len(factors), len(jobs))]
    got, gp = me.search(jobs, want_preds=True)
    want, wp = ses.search(jobs, want_preds=True)
    assert_results_equal(got, want, jobs)
    assert np.array_equal(gp, wp)


@pytest.mark.parametrize("full_dfunc,sub_dfunc", [(2, 2), (2, 0), (1, 1)])
def test_fullpel_hadamard_and_ssd_search_bit_exact(me, oracle, synth, full_dfunc, sub_dfunc):
    """SearchFuncFullPel 2 (Hadamard) and 1 (SSD) through the generic full-search kernel, all 41 partitions, picture corners and
    edges, SR 16, plane-sourced and signed 16-bit originals, against the checker (pinned against the reference for these
    modes by the fullpel_* golden vectors)."""
    cfg = synth.CONFIGS["C1"]
    W, H = cfg["width"], cfg["height"]
    seq = synth.Sequence(W, H, 1290)
    frames = [seq.frame(0), seq.frame(1)]
    me.configure(W, H, 2)
    me.set_params(cfg["search_range"], sub_dfunc, full_dfunc)
    ses = oracle.session(W, H, 2, cfg["search_range"], sub_dfunc, full_dfunc)
    for s, f in enumerate(frames):
        me.upload_plane(s, f); ses.set_plane(s, f)
    me.interp_halfpel([0, 1])
    mbs = [(0, 0), (0, 10), (8, 0), (8, 10), (4, 5), (3, 7)]
    jobs = synth.make_jobs(W, H, synth.PARTS_41, cur_slot=1, ref_slot=0, qp=30, field="split", seed=17, mb_subset=mbs, bits_in=2)
    got, gp = me.search(jobs, want_preds=True)
    want, wp = ses.search(jobs, want_preds=True)
    assert_results_equal(got, want, jobs)
    assert np.array_equal(gp, wp)
    assert me.last_search_kernel().startswith("full_search_generic_kernel")
    org = np.random.default_rng(4).integers(-255, 511, (len(mbs), 16, 16)).astype(np.int16)
    jobs["org_index"] = np.repeat(np.arange(len(mbs)), len(synth.PARTS_41))
    got, gp = me.search(jobs, org_mbs=org, want_preds=True)
    want, wp = ses.search(jobs, org_mbs=org, want_preds=True)
    assert_results_equal(got, want, jobs)
    assert np.array_equal(gp, wp)


def test_invalid_distortion_functions_are_refused(me, pkg):
    """MotionVectorSearchParams::check (ENC/CodingParameter.cpp:17-36): full-pel functions 0..3, sub-pel functions 0..2.  Loud errors."""
    for full, sub in [(0, 3), (3, 3), (4, 0), (-1, 0), (0, -1)]:
        with pytest.raises(pkg.MotionEstimationError):
            me.set_params(16, sub, full)
    me.set_params(16, 2, 0)


def test_extreme_start_vectors_and_per_job_ranges_bit_exact(me, oracle, synth):
    """Start vectors anywhere in the Short range (clamped to m_cMin / m_cMax per macroblock, N2, then '>> 2' on negatives) and
    per-call search ranges 1 .. 16 (the bi-prediction iterations pass uiSearchRange != 0): windows collapse to a few
    positions at the picture border, jobs of one macroblock get very different windows."""
    cfg = synth.CONFIGS["C1"]
    W, H = cfg["width"], cfg["height"]
    ses, _ = _setup(me, oracle, synth, cfg, seed=1280)
    jobs = synth.make_jobs(W, H, synth.PARTS_41, cur_slot=1, ref_slot=0, qp=28, field="zero", seed=1, bits_in=3)
    rng = np.random.default_rng(9)
    n = len(jobs)
    wild = rng.random(n) < 0.3
    jobs["start_hor"] = np.where(wild, rng.integers(-32768, 32768, n), rng.integers(-200, 201, n)).astype(np.int16)
    jobs["start_ver"] = np.where(wild, rng.integers(-32768, 32768, n), rng.integers(-200, 201, n)).astype(np.int16)
    # predictors stay near the clamped start, as the encoder's do (the reference's bit table only spans the search range)
    mbw, mbh = W // 16, H // 16
    min_h, max_h = (-16 * jobs["mb_x"].astype(np.int32) - 24) * 4, (16 * (mbw - jobs["mb_x"].astype(np.int32)) + 8) * 4
    min_v, max_v = (-16 * jobs["mb_y"].astype(np.int32) - 24) * 4, (16 * (mbh - jobs["mb_y"].astype(np.int32)) + 8) * 4
    jobs["pred_hor"] = (np.clip(jobs["start_hor"].astype(np.int32), min_h, max_h) + rng.integers(-64, 65, n)).astype(np.int16)
    jobs["pred_ver"] = (np.clip(jobs["start_ver"].astype(np.int32), min_v, max_v) + rng.integers(-64, 65, n)).astype(np.int16)
    jobs["search_range"] = rng.choice(np.array([0, 1, 2, 4, 7, 16], np.int16), n)
    got, gp = me.search(jobs, want_preds=True)
    want, wp = ses.search(jobs, want_preds=True)
    assert_results_equal(got, want, jobs)
    assert np.array_equal(gp, wp)


def test_originals_outside_the_supported_range_are_refused(me, pkg):
    """|sample| >= 1024 cannot come from a residual-prediction pass; loud error, no fallback."""
    g = np.load(os.path.join(GOLDEN, "org_override_16bit.npz"))
    me.configure(int(g["width"]), int(g["height"]), 2)
    me.set_params(int(g["search_range"]), int(g["sub_dfunc"]))
    me.upload_plane(0, g["frame0"]); me.upload_plane(1, g["frame1"])
    me.interp_halfpel([0, 1])
    org = g["org_mbs"].copy()
    org[0, 0, 0] = 2000
    with pytest.raises(pkg.MotionEstimationError):
        me.search(g["jobs"], org_mbs=org)


def test_reference_named_interface(me, pkg, oracle, synth):
    """initMb -> estimateBlockWithStart -> compensateBlock, one call at a time, as MbEncoder drives it."""
    cfg = synth.CONFIGS["C1"]
    ses, _ = _setup(me, oracle, synth, cfg, seed=1240)
    est = pkg.MotionEstimationQuarterPel.create(me)
    est.init(cfg["search_range"], cfg["sub_dfunc"])
    est.setMbQpLambda(synth.lambda_for_qp(30))
    est.loadOrgMbPelData(pkg.Frame(1))
    ref = pkg.Frame(0)
    for (mby, mbx, blk, mode) in [(0, 0, 0, 1), (4, 5, 8, 2), (8, 10, 10, 8), (3, 0, 5, 11)]:
        est.initMb(mby, mbx)
        mv, bits, cost = est.estimateBlockWithStart(ref, (5, -3), (4, -2), 2, blk, mode)
        pred = est.compensateBlock(blk, mode)
        job = synth.make_jobs(cfg["width"], cfg["height"], [(mode, blk)], 1, 0, qp=30, field="zero", mb_subset=[(mby, mbx)], bits_in=2)
        job["start_hor"], job["start_ver"], job["pred_hor"], job["pred_ver"] = 5, -3, 4, -2
        want, wp = ses.search(job, want_preds=True)
        assert mv == (want["mv_hor"][0], want["mv_ver"][0]) and bits == want["bits"][0] and cost == want["cost"][0]
        w, h = pkg.BLOCK_SIZE[mode]
        assert np.array_equal(pred, wp[0, :h, :w].astype(np.int16))


# ---- the larger BASELINE.json configurations: macroblock subsets against the oracle ----
def _edge_mbs(mbw, mbh, n_inner, seed):
    rng = np.random.default_rng(seed)
    mbs = {(0, 0), (0, mbw - 1), (mbh - 1, 0), (mbh - 1, mbw - 1), (0, mbw // 2), (mbh - 1, mbw // 3), (mbh // 2, 0), (mbh // 3, mbw - 1),
           (1, 1), (mbh - 2, mbw - 2), (3, 4), (4, mbw - 5)}
    while len(mbs) < 12 + n_inner:
        mbs.add((int(rng.integers(0, mbh)), int(rng.integers(0, mbw))))
    return sorted(mbs)


@pytest.mark.parametrize("cfg_name,n_inner,field,n_refs", [("C2", 60, "random", 1), ("C3", 24, "split", 1), ("C4", 20, "random", 1),
                                                           ("C4", 12, "zero", 1), ("C5", 10, "split", 4)])
def test_large_configs_subset_bit_exact(me, oracle, synth, cfg_name, n_inner, field, n_refs):
    cfg = synth.CONFIGS[cfg_name]
    W, H = cfg["width"], cfg["height"]
    n_frames = 1 + n_refs
    ses, _ = _setup(me, oracle, synth, cfg, seed=1234 + int(cfg_name[1]), n_frames=n_frames)
    mbs = _edge_mbs(W // 16, H // 16, n_inner, seed=int(cfg_name[1]))
    jobs = np.concatenate([synth.make_jobs(W, H, cfg["parts"], cur_slot=n_refs, ref_slot=r, qp=30 + r, field=field, seed=50 + r, mb_subset=mbs)
                           for r in range(n_refs)])
    got, gp = me.search(jobs, want_preds=True)
    want, wp = ses.search(jobs, want_preds=True)
    assert_results_equal(got, want, jobs)
    assert np.array_equal(gp, wp)
    # half-pel planes of the big pictures as well
    assert np.array_equal(me.halfpel_plane(0), ses.halfpel(0))


def test_reuploaded_slots_get_fresh_margins(me, oracle, synth):
    """Plane uploads only queue their copy; YuvPicBuffer::fillMargin runs in front of the first reader.  Slots that are uploaded
    again (new pictures in the same slots, then the two pictures swapped, with and without a new half-pel plane in between) must
    be searched with the margins of what they hold NOW: start vectors point across every picture edge."""
    cfg = synth.CONFIGS["C1"]
    W, H = cfg["width"], cfg["height"]
    seq = synth.Sequence(W, H, 77)
    me.configure(W, H, 2)
    me.set_params(cfg["search_range"], cfg["sub_dfunc"])
    me.set_fast_search_params(0, 0)
    for rnd, (a, b) in enumerate([(0, 1), (2, 3), (3, 2), (5, 2)]):
        fa, fb = seq.frame(a), seq.frame(b)
        me.upload_plane(0, fa)
        if rnd != 3:
            me.upload_plane(1, fb)                                           # round 3 keeps slot 1 and its half-pel plane
            me.interp_halfpel([1])
        me.upload_plane(0, fa)                                               # twice in a row: still one entry in the list
        ses = oracle.session(W, H, 2, cfg["search_range"], cfg["sub_dfunc"])
        ses.set_plane(0, fa)
        ses.set_plane(1, fb)
        jobs = synth.make_jobs(W, H, cfg["parts"], cur_slot=0, ref_slot=1, qp=30, field="random", seed=900 + rnd)
        edge = (jobs["mb_x"] == 0) | (jobs["mb_y"] == 0) | (jobs["mb_x"] == W // 16 - 1) | (jobs["mb_y"] == H // 16 - 1)
        jobs["start_hor"][edge] = np.where(jobs["mb_x"][edge] == 0, -200, 200).astype(np.int16)      # far outside: clamped to the MV limits
        jobs["start_ver"][edge] = np.where(jobs["mb_y"][edge] == 0, -200, 200).astype(np.int16)
        got = me.search(jobs)
        assert_results_equal(got, ses.search(jobs), jobs)
        assert np.array_equal(me.halfpel_plane(1), ses.halfpel(1))


def test_persistent_and_per_group_kernels_agree(me, oracle, synth, monkeypatch):
    """The same whole-picture job list through full_search_persistent_kernel (one CTA per SM, producer / consumer warps over two
    shared-memory stages) and through full_search_kernel (one CTA per group): identical bytes, and both equal to the checker."""
    cfg = synth.CONFIGS["C2"]                                              # (SR 32: the default choice here is the per-group kernel)
    W, H = cfg["width"], cfg["height"]
    ses, _ = _setup(me, oracle, synth, cfg, seed=1270)
    jobs = np.concatenate([synth.make_jobs(W, H, cfg["parts"], cur_slot=1, ref_slot=0, qp=32, field=f, seed=3) for f in ("random", "split")])
    out = {}
    for mode in ("persistent", "group"):
        monkeypatch.setenv("JSVM_ME_K2", mode)
        out[mode] = me.search(jobs)
        assert me.last_search_kernel().startswith("full_search_persistent_kernel" if mode == "persistent" else "full_search_kernel<")
    monkeypatch.delenv("JSVM_ME_K2")
    assert out["persistent"].tobytes() == out["group"].tobytes()
    sub = np.arange(0, len(jobs), 7)                                      # every 7th job against the checker
    want = ses.search(jobs[sub])
    assert_results_equal(out["persistent"][sub], want, jobs[sub])


@pytest.mark.parametrize("content", ["flat", "periodic", "steps"])
def test_tie_heavy_content_bit_exact(me, oracle, synth, monkeypatch, content):
    """Content on which many displacements have EXACTLY the same cost: flat pictures (every SAD is 0, the rate decides, then the
    raster order), an 8-pixel periodic pattern (zero-SAD ties one period apart), flat halves with one edge.  The full search
    visits strips centre-out and skips the rate / compare code of batches that cannot beat the incumbent; the winner must still
    be the reference's first strict minimum in raster order, through both kernels."""
    cfg = synth.CONFIGS["C2"]
    W, H = cfg["width"], cfg["height"]
    yy, xx = np.mgrid[0:H, 0:W]
    if content == "flat":
        frames = [np.full((H, W), 117, np.uint8), np.full((H, W), 117, np.uint8)]
    elif content == "periodic":
        base = (((xx // 4) + (yy // 4)) & 1) * 160 + 40
        frames = [base.astype(np.uint8), np.roll(base, (4, 8), axis=(0, 1)).astype(np.uint8)]
    else:
        frames = [np.where(xx < W // 2, 60, 190).astype(np.uint8), np.where(xx < W // 2 + 3, 60, 190).astype(np.uint8)]
    me.configure(W, H, 2)
    me.set_params(cfg["search_range"], cfg["sub_dfunc"])
    me.set_fast_search_params(0, 0)
    ses = oracle.session(W, H, 2, cfg["search_range"], cfg["sub_dfunc"])
    for s_, f in enumerate(frames):
        me.upload_plane(s_, f)
        ses.set_plane(s_, f)
    me.interp_halfpel([0, 1])
    jobs = np.concatenate([synth.make_jobs(W, H, cfg["parts"], cur_slot=1, ref_slot=0, qp=q, field=f, seed=5) for q, f in ((30, "random"), (44, "zero"))])
    out = {}
    for mode in ("persistent", "group"):
        monkeypatch.setenv("JSVM_ME_K2", mode)
        out[mode] = me.search(jobs)
    monkeypatch.delenv("JSVM_ME_K2")
    assert out["persistent"].tobytes() == out["group"].tobytes()
    sub = np.arange(0, len(jobs), 3)
    assert_results_equal(out["persistent"][sub], ses.search(jobs[sub]), jobs[sub])


# ---- full-size (1080p, SR 64, every macroblock) size-independent properties, no oracle needed ----
@pytest.fixture(scope="module")
def full_1080p(me, synth):
    cfg = synth.CONFIGS["C4"]
    W, H = cfg["width"], cfg["height"]
    seq = synth.Sequence(W, H, 4321, guard=96)
    f0 = seq.frame(0)
    big = seq.tex[::4, ::4]                                    # integer-pel texture, larger than the picture
    g = seq.guard
    shifted = np.clip(np.rint(big[g - 5:g - 5 + H, g + 7:g + 7 + W]), 0, 255).astype(np.uint8)    # content moved by (+7, -5) pixels
    base = np.clip(np.rint(big[g:g + H, g:g + W]), 0, 255).astype(np.uint8)
    me.configure(W, H, 4)
    me.set_params(cfg["search_range"], cfg["sub_dfunc"])
    for s, f in enumerate([f0, seq.frame(1), base, shifted]):
        me.upload_plane(s, f)
    me.interp_halfpel([0, 1, 2, 3])
    return cfg, W, H


def test_1080p_identity_search_finds_zero_motion(me, synth, full_1080p):
    """cur == ref, start = pred = 0: every job must return MV (0,0), SAD 0, no sub-pel winner."""
    cfg, W, H = full_1080p
    jobs = synth.make_jobs(W, H, cfg["parts"], cur_slot=0, ref_slot=0, qp=35, field="zero")
    res = me.search(jobs)
    assert len(res) == (W // 16) * (H // 16) * 9
    assert (res["mv_hor"] == 0).all() and (res["mv_ver"] == 0).all()
    assert (res["sad_fullpel"] == 0).all() and (res["best_mode"] == 0).all()
    f = int(jobs["cost_factor"][0])
    assert (res["min_sad"] == ((f * 2) >> 16)).all() and (res["mv_bits"] == 2).all()


def test_1080p_integer_translation_is_recovered(me, synth, full_1080p):
    """ref = cur translated by (+7, -5) pixels: interior macroblocks must find exactly that MV with SAD 0."""
    cfg, W, H = full_1080p
    jobs = synth.make_jobs(W, H, cfg["parts"], cur_slot=2, ref_slot=3, qp=20, field="zero")
    res = me.search(jobs)
    inner = (jobs["mb_x"] >= 1) & (jobs["mb_x"] < W // 16 - 1) & (jobs["mb_y"] >= 1) & (jobs["mb_y"] < H // 16 - 1)
    assert (res["sad_fullpel"][inner] == 0).all()
    assert (res["full_hor"][inner] == -7).all() and (res["full_ver"][inner] == 5).all()
    assert (res["mv_hor"][inner] == -28).all() and (res["mv_ver"][inner] == 20).all()


def test_1080p_grouping_does_not_change_results(me, synth, full_1080p):
    """Sharing one SAD pass across the 9 partitions of a macroblock (the planner's grouping) must give the
    same records as searching every partition on its own: interleave the jobs so that no two neighbours share a MB."""
    cfg, W, H = full_1080p
    jobs = synth.make_jobs(W, H, cfg["parts"], cur_slot=1, ref_slot=0, qp=33, field="random", seed=9)
    grouped = me.search(jobs)
    n_mb = len(jobs) // 9
    perm = np.arange(len(jobs)).reshape(n_mb, 9).T.reshape(-1)          # partition-major order: neighbours differ in MB
    alone = me.search(jobs[perm])
    back = np.empty_like(alone)
    back[perm] = alone
    assert_results_equal(grouped, back, jobs)
    # idempotence: a second identical call returns identical bytes
    again = me.search(jobs)
    assert grouped.tobytes() == again.tobytes()
