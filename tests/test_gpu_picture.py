"""GPU parity of the picture-level motion search (K6): MV prediction, macroblock scheduling in wavefronts and
the search-group planning all run on the device; compared with the reference golden, with the C checker, and -- at 1080p -- with
the per-job search path fed with MV predictors recomputed on the host from the returned motion field."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def me(pkg):
    m = pkg.MotionEstimator(0)
    yield m
    m.close()


def assert_fields_equal(got, want):
    for f in got.dtype.names:
        bad = np.argwhere(got[f] != want[f])
        assert len(bad) == 0, f"{f}: {len(bad)} entries differ, first at (pic, mb_y, mb_x, ...) = {bad[0]}: got {got[tuple(bad[0][:3])]} want {want[tuple(bad[0][:3])]}"


def test_matches_reference_golden(me):
    g = np.load(os.path.join(GOLDEN, "picture_search_qcif.npz"))
    W, H = int(g["width"]), int(g["height"])
    me.configure(W, H, 4)
    me.set_params(int(g["search_range"]), int(g["sub_dfunc"]))
    me.set_fast_search_params(0, 0)
    for s in range(4):
        me.upload_plane(s, g["frames"][s])
    me.interp_halfpel([0, 1, 2, 3])
    assert_fields_equal(me.estimate_pictures(g["tasks"]), g["field"])
    assert_fields_equal(me.download_pictures(2), g["field"][:2])


@pytest.mark.parametrize("cfg_name,n_refs,sub,n_pics", [("C1", 1, 2, 3), ("C1", 4, 0, 2), ("C2", 2, 2, 2)])
def test_matches_checker(me, oracle, synth, cfg_name, n_refs, sub, n_pics):
    """Whole QCIF / CIF pictures with local motion, 1 to 4 reference frames, several pictures advancing together."""
    import oracle_lib
    from concurrent.futures import ThreadPoolExecutor
    cfg = synth.CONFIGS[cfg_name]
    W, H, sr = cfg["width"], cfg["height"], cfg["search_range"]
    n_slots = n_refs + n_pics
    seq = synth.LocalMotionSequence(W, H, 800 + n_refs)
    frames = [seq.frame(i) for i in range(n_slots)]
    me.configure(W, H, n_slots)
    me.set_params(sr, sub)
    for s, f in enumerate(frames):
        me.upload_plane(s, f)
    me.interp_halfpel(list(range(n_slots)))
    tasks = synth.make_pic_tasks([n_refs + p for p in range(n_pics)], [[(p + r) % (n_refs + p) for r in range(n_refs)] for p in range(n_pics)], qp=29 + n_refs)
    got = me.estimate_pictures(tasks)

    def check(p):                               # one checker session per picture (ctypes releases the GIL)
        ses = oracle.session(W, H, n_slots, sr, sub)
        for s, f in enumerate(frames):
            ses.set_plane(s, f)
        return ses.estimate_pictures(tasks[p:p + 1])[0]
    with ThreadPoolExecutor(n_pics) as ex:
        want = np.stack(list(ex.map(check, range(n_pics))))
    assert_fields_equal(got, want)
    assert len(np.unique(got["mode"])) >= 3


# ---- host restatement of the predictor derivation from a finished motion field (numpy / python, the third implementation) ----
def _neighbour(field, mx, my, b4):
    mbh, mbw = field.shape
    if mx < 0 or my < 0 or mx >= mbw or my >= mbh:
        return (0, 0, 0)
    m = field[my, mx]
    return (int(m["mv"][b4][0]), int(m["mv"][b4][1]), int(m["ref_idx"][((b4 >> 3) << 1) | ((b4 & 3) >> 1)]))


def _mvp(A, B, C, D, ref, typ):
    if C[2] == 0:
        C = D
    if (typ == 1 and A[2] == ref) or (A[2] == ref and B[2] != ref and C[2] != ref) or (B[2] == 0 and C[2] == 0):
        return A[:2]
    if (typ == 2 and B[2] == ref) or (A[2] != ref and B[2] == ref and C[2] != ref):
        return B[:2]
    if (typ == 3 and C[2] == ref) or (A[2] != ref and B[2] != ref and C[2] == ref):
        return C[:2]
    return (sorted((A[0], B[0], C[0]))[1], sorted((A[1], B[1], C[1]))[1])


def test_1080p_self_consistent_with_per_job_search(me, synth):
    """C4 geometry (1080p, SR 64): every partition result of the picture-level search must equal what the per-job search
    (bit-exact against the checker in test_gpu_fullframe.py) returns for the MV predictor derived on the host from the returned
    motion field; the mode must be the arg-min of the partition-cost sums."""
    cfg = synth.CONFIGS["C4"]
    W, H, sr = cfg["width"], cfg["height"], cfg["search_range"]
    seq = synth.LocalMotionSequence(W, H, 4242, tile=52)
    me.configure(W, H, 3)
    me.set_params(sr, 2)
    for s in range(3):
        me.upload_plane(s, seq.frame(s))
    me.interp_halfpel([0, 1, 2])
    tasks = synth.make_pic_tasks([2], [[1, 0]], qp=34)
    field = me.estimate_pictures(tasks)[0]
    mbh, mbw = field.shape
    part_mode = [1, 2, 2, 3, 3, 8, 8, 8, 8]
    part_blk = [0, 0, 8, 0, 2, 0, 2, 8, 10]
    part_bits = [1, 3, 3, 3, 3, 6, 1, 1, 1]
    rng = np.random.default_rng(1)
    sample = [(0, 0), (0, mbw - 1), (mbh - 1, 0), (mbh - 1, mbw - 1)] + [(int(y), int(x)) for y, x in zip(rng.integers(0, mbh, 400), rng.integers(0, mbw, 400))]
    jobs = np.zeros(len(sample) * 9 * 2, synth.JOB_DTYPE)
    k = 0
    for my, mx in sample:
        m = field[my, mx]
        cur = [(int(m["part_mv"][p][0]), int(m["part_mv"][p][1]), int(m["part_ref"][p])) for p in range(9)]
        for part in range(9):
            none = (0, 0, 0)
            if part in (0, 1, 3, 5):
                A, B = _neighbour(field, mx - 1, my, 3), _neighbour(field, mx, my - 1, 12)
                C = _neighbour(field, mx + 1, my - 1, 12) if part <= 1 else _neighbour(field, mx, my - 1, 14)
                D, typ = _neighbour(field, mx - 1, my - 1, 15), {1: 2, 3: 1}.get(part, 0)
            elif part == 2:
                A, B, C, D, typ = _neighbour(field, mx - 1, my, 11), cur[1], none, _neighbour(field, mx - 1, my, 7), 1
            elif part in (4, 6):
                A, B = cur[3 if part == 4 else 5], _neighbour(field, mx, my - 1, 14)
                C, D, typ = _neighbour(field, mx + 1, my - 1, 12), _neighbour(field, mx, my - 1, 13), 3 if part == 4 else 0
            elif part == 7:
                A, B, C, D, typ = _neighbour(field, mx - 1, my, 11), cur[5], cur[6], none, 0
            else:
                A, B, C, D, typ = cur[7], cur[6], none, cur[5], 0
            for r in (1, 2):
                j = jobs[k]; k += 1
                j["cur_slot"], j["ref_slot"], j["mb_x"], j["mb_y"] = 2, tasks["ref_slot"][0][r - 1], mx, my
                j["blk"], j["mode"] = part_blk[part], part_mode[part]
                pv = _mvp(A, B, C, D, r, typ)
                j["start_hor"], j["start_ver"], j["pred_hor"], j["pred_ver"] = pv[0], pv[1], pv[0], pv[1]
                j["cost_factor"], j["bits_in"], j["org_index"] = tasks["cost_factor"][0], part_bits[part] + 1, -1
    res = me.search(jobs).reshape(len(sample), 9, 2)
    for i, (my, mx) in enumerate(sample):
        m = field[my, mx]
        for part in range(9):
            r = 0 if res[i, part, 0]["cost"] <= res[i, part, 1]["cost"] else 1          # strict '<' in reference order: the first wins ties
            assert int(m["part_ref"][part]) == r + 1, (my, mx, part)
            assert int(m["part_cost"][part]) == int(res[i, part, r]["cost"]), (my, mx, part)
            assert (int(m["part_mv"][part][0]), int(m["part_mv"][part][1])) == (int(res[i, part, r]["mv_hor"]), int(res[i, part, r]["mv_ver"])), (my, mx, part)
        c = m["part_cost"].astype(np.int64)
        sums = [c[0], c[1] + c[2], c[3] + c[4], c[5:].sum()]
        assert int(m["mode"]) == int(np.argmin(sums)) + 1 and int(m["cost"]) == int(min(sums))
    assert len(np.unique(field["mode"])) == 4


def test_refusals(me, pkg, synth):
    me.configure(176, 144, 3)
    me.set_params(16, 2)
    me.upload_plane(0, np.zeros((144, 176), np.uint8))
    with pytest.raises(pkg.MotionEstimationError, match="half-pel"):
        me.estimate_pictures(synth.make_pic_tasks([0], [[1]]))
    me.set_params(16, 2, 2)                                    # full-pel Hadamard: not the packed-byte path
    with pytest.raises(pkg.MotionEstimationError, match="SearchFuncFullPel 0"):
        me.estimate_pictures(synth.make_pic_tasks([0], [[0]]))
    me.set_params(16, 2)
