"""GPU parity of the motion-compensated prediction (K4) and of the optional distortion functions (a5):
the CUDA path through the C ABI against the golden blocks of the compiled reference and against the C checker, bit-exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KINDS = ["uni", "bi", "uni_w", "bi_w", "bi_implicit"]


@pytest.fixture(scope="module")
def me(pkg):
    m = pkg.MotionEstimator(0)
    yield m
    m.close()


@pytest.mark.parametrize("kind", KINDS)
def test_compensate_matches_reference_golden(me, kind):
    g = np.load(os.path.join(GOLDEN, "mc_blocks.npz"))
    W, H = int(g["width"]), int(g["height"])
    me.configure(W, H, 3)
    for s in range(3):
        me.upload_plane(s, g["frames"][s])
        me.upload_chroma(s, g["cb"][s], g["cr"][s])
    me.interp_halfpel([0, 1, 2])
    py, pcb, pcr = me.compensate(g["jobs_" + kind])
    for nm, got, want in (("luma", py, g["y_" + kind]), ("cb", pcb, g["cb_" + kind]), ("cr", pcr, g["cr_" + kind])):
        bad = np.argwhere(got != want)
        assert len(bad) == 0, f"{kind} {nm}: {len(bad)} samples differ, first (job, y, x) = {bad[0]}, job {g['jobs_' + kind][bad[0][0]]}"


@pytest.mark.parametrize("cfg_name,n", [("C1", 4000), ("C4", 60000)])
def test_compensate_matches_checker(me, oracle, synth, cfg_name, n):
    """Seeded random blocks on QCIF and on full 1080p pictures (all phases, every block size, clamped far MVs, all weightings)."""
    cfg = synth.CONFIGS[cfg_name]
    W, H = cfg["width"], cfg["height"]
    seq = synth.Sequence(W, H, 4321)
    me.configure(W, H, 2)
    ses = oracle.session(W, H, 2, 8, 0)
    for s in range(2):
        f = seq.frame(s)
        cb, cr = synth.chroma_planes(W, H, 50 + s)
        me.upload_plane(s, f); me.upload_chroma(s, cb, cr)
        ses.set_plane(s, f); ses.set_chroma(s, cb, cr)
    me.interp_halfpel([0, 1])
    for k, kind in enumerate(KINDS):
        jobs = synth.make_mc_jobs(W, H, n // len(KINDS), 2, seed=100 + k, kind=kind, mv_range=200)
        got, want = me.compensate(jobs), ses.compensate(jobs)
        for a, b in zip(got, want):
            assert np.array_equal(a, b), kind
    # luma only
    jobs = synth.make_mc_jobs(W, H, 500, 2, seed=9, kind="bi")
    got, want = me.compensate(jobs, chroma=False), ses.compensate(jobs, chroma=False)
    assert got[1] is None and np.array_equal(got[0], want[0])


def test_compensate_equals_search_prediction(me, oracle, synth):
    """The prediction block estimateBlockWithStart leaves for compensateBlock is the motion-compensated block at the winning MV."""
    cfg = synth.CONFIGS["C1"]
    W, H = cfg["width"], cfg["height"]
    seq = synth.Sequence(W, H, 99)
    me.configure(W, H, 2)
    me.set_params(16, 2)
    for s in range(2):
        me.upload_plane(s, seq.frame(s))
    me.interp_halfpel([0, 1])
    jobs = synth.make_jobs(W, H, synth.PARTS_41, 1, 0, qp=30, field="random", seed=3)
    res, preds = me.search(jobs, want_preds=True)
    abi = synth.MC_JOB_DTYPE
    mc = np.zeros(len(jobs), abi)
    mc["ref_slot"][:, 0], mc["ref_slot"][:, 1] = 0, -1
    mc["mb_x"], mc["mb_y"], mc["blk"] = jobs["mb_x"], jobs["mb_y"], jobs["blk"]
    size = {1: (16, 16), 2: (16, 8), 3: (8, 16), 8: (8, 8), 9: (8, 4), 10: (4, 8), 11: (4, 4)}
    mc["size_x"] = [size[int(m)][0] for m in jobs["mode"]]
    mc["size_y"] = [size[int(m)][1] for m in jobs["mode"]]
    mc["mv"][:, 0, 0], mc["mv"][:, 0, 1] = res["mv_hor"], res["mv_ver"]
    py, _, _ = me.compensate(mc, chroma=False)
    # sub-pel candidates are not re-clamped (N10) while MotionCompensation clamps: compare where the MV is inside the MB's range
    mbw, mbh = W // 16, H // 16
    inside = ((res["mv_hor"] >= ((-16 * jobs["mb_x"].astype(int) - 24) * 4)) & (res["mv_hor"] <= ((16 * (mbw - jobs["mb_x"].astype(int)) + 8) * 4)) &
              (res["mv_ver"] >= ((-16 * jobs["mb_y"].astype(int) - 24) * 4)) & (res["mv_ver"] <= ((16 * (mbh - jobs["mb_y"].astype(int)) + 8) * 4)))
    assert inside.sum() > 0.9 * len(jobs)
    assert np.array_equal(py[inside], preds[inside])


def test_compensate_errors(me, synth):
    W, H = 64, 48
    me.configure(W, H, 2)
    me.upload_plane(0, np.zeros((H, W), np.uint8))
    jobs = synth.make_mc_jobs(W, H, 4, 1, seed=1, kind="uni")
    with pytest.raises(RuntimeError, match="half-pel"):
        me.compensate(jobs, chroma=False)
    me.interp_halfpel([0])
    with pytest.raises(RuntimeError, match="chroma"):
        me.compensate(jobs)
    bad = jobs.copy(); bad["size_x"][0] = 12
    with pytest.raises(RuntimeError, match="does not fit"):
        me.compensate(bad, chroma=False)


@pytest.mark.parametrize("full,sub", [(3, 2), (3, 0), (3, 1), (0, 1), (1, 0), (1, 2), (2, 1)])
def test_optional_distortion_functions_match_checker(me, oracle, synth, full, sub):
    """a5: YUV-SAD full-pel search and every mix of SAD-class and SSE-class stages, QCIF macroblock subset incl. corners."""
    import oracle_lib
    cfg = synth.CONFIGS["C1"]
    W, H = cfg["width"], cfg["height"]
    seq = synth.Sequence(W, H, 777)
    me.configure(W, H, 2)
    me.set_params(16, sub, full)
    ses = oracle_lib.CpuSession(oracle, W, H, 2, 16, sub, full)
    for s in range(2):
        f = seq.frame(s)
        cb, cr = synth.chroma_planes(W, H, 60 + s)
        me.upload_plane(s, f); me.upload_chroma(s, cb, cr)
        ses.set_plane(s, f); ses.set_chroma(s, cb, cr)
    me.interp_halfpel([0, 1])
    me.set_org_chroma(None); ses.set_org_chroma(None)
    mbs = [(0, 0), (0, 10), (8, 0), (8, 10), (4, 5), (3, 7)]
    jobs = synth.make_jobs(W, H, synth.PARTS_41, 1, 0, qp=32, field="split", seed=5, mb_subset=mbs, bits_in=1)
    f_sad, f_sse = oracle.cost_factor(synth.lambda_for_qp(32), 1), oracle.cost_factor(synth.lambda_for_qp(32), 0)
    jobs["cost_factor"] = f_sse if full == 1 else f_sad
    if (full == 1) != (sub == 1):
        pair = (f_sse, f_sad) if full == 1 else (f_sad, f_sse)
        me.map_cost_factor(*pair); ses.map_cost_factor(*pair)
    got, gp = me.search(jobs, want_preds=True)
    want, wp = ses.search(jobs, want_preds=True)
    for f in got.dtype.names:
        assert np.array_equal(got[f], want[f]), f
    assert np.array_equal(gp, wp)


def test_unmapped_cost_factor_is_refused(me, synth):
    cfg = synth.CONFIGS["C1"]
    W, H = cfg["width"], cfg["height"]
    me.configure(W, H, 2)
    me.set_params(16, 1, 0)                                  # SAD full-pel, SSD sub-pel
    for s in range(2):
        me.upload_plane(s, np.full((H, W), 90 + s, np.uint8))
    me.interp_halfpel([0, 1])
    jobs = synth.make_jobs(W, H, synth.PARTS_9, 1, 0, qp=20, mb_subset=[(1, 1)])
    jobs["cost_factor"] = 123457
    with pytest.raises(RuntimeError, match="jsvm_me_map_cost_factor"):
        me.search(jobs)
