"""Pins the C oracle (oracle/me_oracle.c) against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py ran oracle/_ref/libjsvm_ref.so in the build container).  Runs on CPU."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SEARCH_CASES = ["tiny_all41_had", "tiny_9_sad", "qcif_9_had", "qcif_9_zero_sr32", "cif_9_sr64", "org_override_8bit", "org_override_16bit",
                "fullpel_had_had", "fullpel_had_sad_all41", "fullpel_ssd_ssd", "fullpel_ssd_ssd_16bit", "fullpel_had_had_16bit",
                "yuvsad_had_all41", "yuvsad_sad_org_chroma", "yuvsad_had_org_lumaonly", "yuvsad_ssd_mixed", "sad_ssd_mixed", "ssd_had_mixed"]
FAST_CASES = ["fast_spiral_sr8", "fast_log_sr16", "fast_log_sr64_had", "fast_fme_sr32", "fast_tz_sr32", "fast_tz_sr64_fastbi", "fast_tz_sr16_yuv",
              "block_fastbi_sr16"]
MC_KINDS = ["uni", "bi", "uni_w", "bi_w", "bi_implicit"]


def prepare_session(ses, g):
    """Planes, chroma planes, chroma originals and cost-factor pairs of a search fixture (shared with the GPU tests)."""
    ses.set_plane(0, g["frame0"])
    ses.set_plane(1, g["frame1"])
    if "cb0" in g.files:
        ses.set_chroma(0, g["cb0"], g["cr0"])
        ses.set_chroma(1, g["cb1"], g["cr1"])
    if "factor_pair" in g.files:
        ses.map_cost_factor(int(g["factor_pair"][0]), int(g["factor_pair"][1]))
    ses.set_org_chroma(g["org_chroma"] if "org_chroma" in g.files else None)
    if "search_mode" in g.files:
        ses.set_fast_search_params(int(g["el_search_range"]), int(g["fast_bi_search"]))
        ses.set_mv_candidates(g["cands"])


def load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


def test_fixtures_present():
    assert len(glob.glob(os.path.join(HERE, "golden", "*.npz"))) >= 30


@pytest.mark.parametrize("name", SEARCH_CASES + FAST_CASES)
def test_oracle_search_matches_reference_golden(oracle, name):
    g = load(name)
    W, H = int(g["width"]), int(g["height"])
    ses = oracle.session(W, H, 2, int(g["search_range"]), int(g["sub_dfunc"]), int(g["full_dfunc"]) if "full_dfunc" in g.files else 0,
                         int(g["search_mode"]) if "search_mode" in g.files else 0)
    prepare_session(ses, g)
    org = g["org_mbs"] if "org_mbs" in g.files else None
    res, preds = ses.search(g["jobs"], org_mbs=org, want_preds=True)
    for f in res.dtype.names:
        assert np.array_equal(res[f], g["results"][f]), f"{name}: field {f} differs from the reference golden"
    assert np.array_equal(preds, g["preds"])
    if "halfpel0" in g.files:
        assert np.array_equal(ses.halfpel(0), g["halfpel0"])       # QuarterPelFilter::filterFrame
        assert np.array_equal(ses.fullpel(0), g["fullpel0"])       # YuvPicBuffer::fillMargin


@pytest.mark.parametrize("kind", MC_KINDS)
def test_oracle_compensate_matches_reference_golden(oracle, kind):
    """MotionCompensation (predBlk, chroma bilinear, SampleWeighting) of the C port against the compiled reference."""
    g = load("mc_blocks")
    W, H = int(g["width"]), int(g["height"])
    ses = oracle.session(W, H, 3, 8, 0)
    for s in range(3):
        ses.set_plane(s, g["frames"][s])
        ses.set_chroma(s, g["cb"][s], g["cr"][s])
    py, pcb, pcr = ses.compensate(g["jobs_" + kind])
    assert np.array_equal(py, g["y_" + kind]) and np.array_equal(pcb, g["cb_" + kind]) and np.array_equal(pcr, g["cr_" + kind])


def test_oracle_picture_search_matches_reference_golden(oracle):
    """f1: MV prediction + per-macroblock loops of the C port against the reference's getMvPredictor / estimateBlockWithStart."""
    g = load("picture_search_qcif")
    W, H = int(g["width"]), int(g["height"])
    ses = oracle.session(W, H, 4, int(g["search_range"]), int(g["sub_dfunc"]))
    for s in range(4):
        ses.set_plane(s, g["frames"][s])
    out = ses.estimate_pictures(g["tasks"])
    for f in out.dtype.names:
        assert np.array_equal(out[f], g["field"][f]), f
    assert len(np.unique(g["field"]["mode"])) == 4 and len(np.unique(g["field"]["part_ref"])) == 2


def resize_params_from(abi, arr):
    p = abi.ResizeParams()
    for (n, _), v in zip(p._fields_, arr):
        setattr(p, n, int(v))
    return p


def test_oracle_residual_upsampling_matches_reference_golden(oracle):
    import oracle_lib
    g = load("residual_upsampling")
    for i in range(int(g["n"])):
        p = resize_params_from(oracle_lib.abi, g[f"params{i}"])
        oy, ocb, ocr = oracle.upsample_residual(p, g[f"y{i}"], g[f"cb{i}"], g[f"cr{i}"], g[f"t8_{i}"])
        assert np.array_equal(oy, g[f"oy{i}"]) and np.array_equal(ocb, g[f"ocb{i}"]) and np.array_equal(ocr, g[f"ocr{i}"]), i


def motion_case(synth, abi, g, i):
    resize = resize_params_from(abi, g[f"resize{i}"])
    st, dq, d8, chk, thr = [int(v) for v in g[f"switches{i}"]]
    p = abi.MotionUpsampleParams()
    import ctypes
    ctypes.memmove(ctypes.addressof(p.resize), ctypes.addressof(resize), ctypes.sizeof(resize))
    p.slice_type, p.ref_layer_dq_id, p.direct_8x8_inference, p.check_residual_pred, p.mv_threshold = st, dq, d8, chk, thr
    return p


def test_oracle_motion_upsampling_matches_reference_golden(oracle, synth):
    import oracle_lib
    g = load("motion_upsampling")
    for i in range(int(g["n"])):
        pred = oracle.upsample_motion(motion_case(synth, oracle_lib.abi, g, i), g[f"base{i}"])
        for f in pred.dtype.names:
            assert np.array_equal(pred[f], g[f"pred{i}"][f]), (i, f)


def test_mc_golden_covers_all_phases_and_clamps():
    g = load("mc_blocks")
    j = g["jobs_uni"]
    l = np.where(j["ref_slot"][:, 0] >= 0, 0, 1)
    mv = j["mv"][np.arange(len(j)), l]
    assert len({(int(a) & 3, int(b) & 3) for a, b in mv}) == 16 and len({(int(a) & 7, int(b) & 7) for a, b in mv}) >= 60
    assert (np.abs(mv) > 1000).any()                                  # MVs the per-macroblock clamp limits
    assert (g["jobs_uni_w"]["weight_flag"] == 0).any() and (g["jobs_uni_w"]["weight_flag"] == 1).any()


def test_oracle_distortion_table_and_factors(oracle):
    g = load("distortion_and_factors")
    orgs, refs = g["orgs"], g["refs"]
    for i, mode, blk, df, want in g["rows"]:
        bx, by = 4 * (int(blk) & 3), 4 * (int(blk) >> 2)
        r = np.ascontiguousarray(refs[i])
        got = oracle.distortion(np.ascontiguousarray(orgs[i]).ctypes.data, r.ctypes.data + 2 * (by * 24 + bx), 24, int(mode), int(df), int(blk))
        assert got == want, (i, mode, blk, df)
    synth_lambda = lambda q: 0.85 * 2.0 ** (q / 3.0 - 4.0)
    for q in range(52):
        assert oracle.cost_factor(synth_lambda(q), 1) == g["factors"][q, 0]
        assert oracle.cost_factor(synth_lambda(q), 0) == g["factors"][q, 1]


def test_golden_results_exercise_subpel_and_edges():
    """The fixtures are only useful if they hit half-pel and quarter-pel winners and clipped windows."""
    modes = np.concatenate([load(n)["results"]["best_mode"] for n in SEARCH_CASES])
    assert (modes == 0).any() and ((modes > 0) & (modes < 16)).any() and (modes >= 16).any()
    g = load("qcif_9_had")
    assert (g["jobs"]["mb_x"] == 0).any() and (g["jobs"]["mb_x"] == 10).any() and (g["jobs"]["mb_y"] == 8).any()
