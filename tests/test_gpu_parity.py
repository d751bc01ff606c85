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
