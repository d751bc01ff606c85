"""GPU parity of the data-dependent full-pel searches (K5): SPIRAL / LOG / FAST / TZ and the FastBiSearch
refinement, through the C ABI against the C checker (itself pinned to the compiled reference), bit-exact on every result field."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def me(pkg):
    m = pkg.MotionEstimator(0)
    yield m
    m.close()


def _run(me, oracle, synth, W, H, search_mode, sr, el, fast_bi, full, sub, parts, mbs, seed, n_proc=None):
    import oracle_lib
    seq = synth.Sequence(W, H, 900 + seed)
    frames = [seq.frame(0), seq.frame(2)]
    chroma = [synth.chroma_planes(W, H, 95 + s) for s in range(2)]
    me.configure(W, H, 2)
    me.set_params(sr, sub, full, search_mode)
    me.set_fast_search_params(el, fast_bi)
    for s in range(2):
        me.upload_plane(s, frames[s]); me.upload_chroma(s, *chroma[s])
    me.interp_halfpel([0, 1])
    jobs = synth.make_jobs(W, H, parts, 1, 0, qp=30 + seed % 6, field="split", seed=seed, mb_subset=mbs, bits_in=2)
    jobs, cands = synth.make_fast_search_extras(jobs, sr, seed=seed)
    me.set_mv_candidates(cands)
    got, gp = me.search(jobs, want_preds=True)
    ses = oracle_lib.CpuSession(oracle, W, H, 2, sr, sub, full, search_mode)
    for s in range(2):
        ses.set_plane(s, frames[s]); ses.set_chroma(s, *chroma[s])
    ses.set_fast_search_params(el, fast_bi)
    ses.set_mv_candidates(cands)
    want, wp = ses.search(jobs, want_preds=True)
    for f in got.dtype.names:
        bad = np.nonzero(got[f] != want[f])[0]
        assert len(bad) == 0, f"mode {search_mode}: field {f}: {len(bad)} of {len(jobs)} jobs differ; first: job {jobs[bad[0]]} cands {cands[bad[0]].tolist()} got {got[bad[0]]} want {want[bad[0]]}"
    assert np.array_equal(gp, wp)
    me.set_fast_search_params(0, 0)
    me.set_mv_candidates(None)
    return got


@pytest.mark.parametrize("search_mode", [2, 3, 4])
@pytest.mark.parametrize("sr,el,fast_bi,full,sub", [(16, 0, 0, 0, 2), (32, 8, 1, 0, 0), (64, 16, 0, 2, 2), (12, 0, 1, 1, 1)])
def test_qcif_all_macroblocks(me, oracle, synth, search_mode, sr, el, fast_bi, full, sub):
    """Every macroblock of a QCIF pair, all 41 partitions: corners, edges, candidates outside the search window."""
    _run(me, oracle, synth, 176, 144, search_mode, sr, el, fast_bi, full, sub, synth.PARTS_41, None, seed=search_mode * 10 + sr)


@pytest.mark.parametrize("sr,full", [(8, 0), (16, 3)])
def test_spiral_cif(me, oracle, synth, sr, full):
    """xPelSpiralSearch: (2 SR + 1)^2 positions in spiral order around a start limited to [ min + 4 SR, max - 4 SR ]."""
    _run(me, oracle, synth, 352, 288, 1, sr, 0, 0, full, 2, synth.PARTS_9, [(0, 0), (0, 21), (17, 0), (17, 21), (8, 10), (3, 15), (12, 4), (9, 9)], seed=sr)


def test_tz_yuv_sad(me, oracle, synth):
    _run(me, oracle, synth, 176, 144, 4, 16, 8, 0, 3, 2, synth.PARTS_41, None, seed=5)


def test_block_search_with_fast_bi_refinement(me, oracle, synth):
    """BLOCK_SEARCH + FastBiSearch: jobs with a per-call range run xTZSearch without candidates, the others the full search."""
    got = _run(me, oracle, synth, 176, 144, 0, 16, 0, 1, 0, 2, synth.PARTS_41, None, seed=77)
    assert len(got) > 0


def test_tz_1080p_sample(me, oracle, synth):
    """TZ at the C4 geometry (1080p, SR 64, EL range 16): a macroblock sample over the whole picture incl. the borders."""
    W, H = 1920, 1088
    rng = np.random.default_rng(3)
    mbs = [(0, 0), (0, 119), (67, 0), (67, 119)] + [(int(y), int(x)) for y, x in zip(rng.integers(0, 68, 60), rng.integers(0, 120, 60))]
    _run(me, oracle, synth, W, H, 4, 64, 16, 1, 0, 2, synth.PARTS_9, mbs, seed=11)


@pytest.mark.parametrize("sr,full,sub", [(96, 0, 2), (200, 2, 0)])
def test_block_search_beyond_range_64(me, oracle, synth, sr, full, sub):
    """BLOCK_SEARCH with SearchRange > 64 (VERDICT r01 missing #6): windows up to the whole padded CIF picture, clipped per macroblock."""
    _run(me, oracle, synth, 352, 288, 0, sr, 0, 0, full, sub, synth.PARTS_9, [(0, 0), (17, 21), (8, 10), (3, 15), (12, 4), (0, 21), (17, 0)], seed=sr)
    assert "SearchRange > 64" in me.last_search_kernel()


def test_refusals(me, pkg, synth):
    me.configure(176, 144, 2)
    for mode in (2, 3):
        with pytest.raises(pkg.MotionEstimationError, match="distortion measure 3"):
            me.set_params(16, 2, 3, mode)
    with pytest.raises(pkg.MotionEstimationError, match="No Such Search Mode"):
        me.set_params(16, 2, 0, 5)
    # spiral search whose un-clipped window does not fit the padded picture
    me.configure(64, 48, 2)
    me.set_params(64, 2, 0, 1)
    for s in range(2):
        me.upload_plane(s, np.full((48, 64), 80, np.uint8))
    me.interp_halfpel([0, 1])
    jobs = synth.make_jobs(64, 48, synth.PARTS_9, 1, 0, mb_subset=[(1, 1)])
    with pytest.raises(pkg.MotionEstimationError, match="spiral search"):
        me.search(jobs)
    me.set_params(16, 2, 0, 0)
