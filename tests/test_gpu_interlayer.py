"""GPU parity of the inter-layer resampling (K7) through the C ABI: against the golden planes of the compiled
reference and against the C checker on the C2 / C5 layer geometries."""
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


def test_residual_upsampling_matches_reference_golden(me):
    from test_golden_oracle import resize_params_from
    import importlib
    abi = importlib.import_module("jsvm-code-notes_b200.abi")
    g = np.load(os.path.join(GOLDEN, "residual_upsampling.npz"))
    for i in range(int(g["n"])):
        p = resize_params_from(abi, g[f"params{i}"])
        oy, ocb, ocr = me.upsample_residual(p, g[f"y{i}"], g[f"cb{i}"], g[f"cr{i}"], g[f"t8_{i}"])
        for nm, got, want in (("luma", oy, g[f"oy{i}"]), ("cb", ocb, g[f"ocb{i}"]), ("cr", ocr, g[f"ocr{i}"])):
            bad = np.argwhere(got != want)
            assert len(bad) == 0, f"case {i} {nm}: {len(bad)} samples differ, first at {bad[0]}: got {got[tuple(bad[0])]} want {want[tuple(bad[0])]}"


@pytest.mark.parametrize("case", [
    dict(ref_w=176, ref_h=144, cur_w=352, cur_h=288),                                                  # C2: QCIF -> CIF
    dict(ref_w=960, ref_h=544, cur_w=1920, cur_h=1088),                                                # C5: 540p -> 1080p
    dict(ref_w=1280, ref_h=720, cur_w=1920, cur_h=1088, scaled_w=1920, scaled_h=1080, top=4),          # 720p -> 1080p, ratio 3:2, cropping window
    dict(ref_w=960, ref_h=544, cur_w=1920, cur_h=1088, level_idc=30, phases=(0, 0, 0, 0)),
], ids=["C2", "C5", "ess-3-2", "C5-level30"])
def test_residual_upsampling_matches_checker(me, oracle, synth, case):
    p, y, cb, cr, t8 = synth.make_resize_case(seed=7, **case)
    got = me.upsample_residual(p, y, cb, cr, t8)
    want = oracle.upsample_residual(p, y, cb, cr, t8)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    oy, _, _ = me.upsample_residual(p, y, base_transform8x8=t8)             # luma only
    assert np.array_equal(oy, want[0])


def test_residual_upsampling_refusals(me, pkg, synth):
    p, y, cb, cr, t8 = synth.make_resize_case(176, 144, 352, 288)
    p.scaled_ref_width = 400
    with pytest.raises(pkg.MotionEstimationError, match="does not fit"):
        me.upsample_residual(p, y, cb, cr, t8)


def test_motion_upsampling_matches_reference_golden(me, synth):
    from test_golden_oracle import motion_case
    import importlib
    abi = importlib.import_module("jsvm-code-notes_b200.abi")
    g = np.load(os.path.join(GOLDEN, "motion_upsampling.npz"))
    for i in range(int(g["n"])):
        pred = me.upsample_motion(motion_case(synth, abi, g, i), g[f"base{i}"])
        for f in pred.dtype.names:
            bad = np.argwhere(pred[f] != g[f"pred{i}"][f])
            assert len(bad) == 0, f"case {i} field {f}: {len(bad)} differ, first at {bad[0]}"


@pytest.mark.parametrize("case,slice_type", [
    (dict(ref_w=960, ref_h=544, cur_w=1920, cur_h=1088), 1),                                                    # C5: 540p -> 1080p, B
    (dict(ref_w=960, ref_h=544, cur_w=1920, cur_h=1088), 0),
    (dict(ref_w=1280, ref_h=720, cur_w=1920, cur_h=1088, scaled_w=1920, scaled_h=1080, top=4), 1),              # ratio 3:2
    (dict(ref_w=176, ref_h=144, cur_w=352, cur_h=288), 1),                                                      # C2
    (dict(ref_w=1920, ref_h=1088, cur_w=1920, cur_h=1088), 1),                                                  # quality layer on top of 1080p
], ids=["C5-B", "C5-P", "ess-3-2", "C2", "same-res"])
def test_motion_upsampling_matches_checker(me, oracle, synth, case, slice_type):
    resize = synth.make_resize_case(seed=1, **case)[0]
    base = synth.make_base_motion(case["ref_w"], case["ref_h"], seed=5 + slice_type, slice_type=slice_type)
    p = synth.make_motion_params(resize, slice_type)
    got, want = me.upsample_motion(p, base), oracle.upsample_motion(p, base)
    for f in got.dtype.names:
        bad = np.argwhere(got[f] != want[f])
        assert len(bad) == 0, f"field {f}: {len(bad)} differ, first at {bad[0]}: got {got[tuple(bad[0][:2])]} want {want[tuple(bad[0][:2])]}"
