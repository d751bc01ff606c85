"""System-level drop-in parity: the unmodified reference encoder CLI linked with the B200 drop-in must write the same
bitstream and the same reconstruction as the stock build (the check of R/data/runAndCheckSVC.bat:45-50).
Both binaries are built in the container that has /root/reference and travel to the GPU box."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("args", [
    ["--frames", "9", "--gop", "8", "--sr", "16", "--subpel", "2"],                                   # C1: QCIF, SR 16, hier-B, SATD sub-pel
    ["--frames", "5", "--gop", "4", "--sr", "32", "--subpel", "0", "--all-partitions", "--refs", "2"],  # all 41 partitions, 2 refs, SAD sub-pel
    # C2: two spatial layers QCIF -> CIF, adaptive inter-layer mode / motion / residual prediction (signed 16-bit originals, N3)
    ["--width", "352", "--height", "288", "--layers", "2", "--frames", "5", "--gop", "4", "--sr", "32", "--subpel", "2", "--saturate"],
    # C5 in small: spatial + quality (CGS) layers, 2 reference frames, inter-layer motion prediction
    ["--width", "352", "--height", "288", "--layers", "2", "--quality-layers", "1", "--refs", "2", "--frames", "5", "--gop", "4",
     "--sr", "16", "--subpel", "2"],
    # the other distortion functions of XDistortion: Hadamard for both stages, SSD for both stages (SSE cost factor)
    ["--frames", "5", "--gop", "4", "--sr", "8", "--subpel", "2", "--fullpel", "2"],
    ["--frames", "5", "--gop", "4", "--sr", "8", "--subpel", "1", "--fullpel", "1"],
    # weighted prediction on a fading sequence (ENC/MotionEstimation.cpp:163-247): explicit weights for P and B pictures, and
    # implicit bi-prediction weights (hierarchical B with unequal temporal distances needs refs > 1)
    ["--frames", "9", "--gop", "4", "--sr", "8", "--subpel", "2", "--wp", "1", "--wbp", "1", "--fade", "0.04"],
    ["--frames", "9", "--gop", "4", "--sr", "8", "--subpel", "2", "--wp", "1", "--wbp", "2", "--fade", "0.04", "--refs", "2"],
    # the fast searches: the shipped encoder.cfg defaults (TZ, EL range 4, FastBiSearch, 4 bi-pred iterations over +-8) on two
    # spatial layers; FAST (log search seeded with the neighbours' MVs); LOG; SPIRAL; and BLOCK_SEARCH with the FastBiSearch refinement
    ["--width", "352", "--height", "288", "--layers", "2", "--frames", "5", "--gop", "4", "--sr", "32", "--subpel", "2",
     "--searchmode", "4", "--elsr", "4", "--fastbi", "1", "--bipreditr", "4", "--itersr", "8"],
    ["--frames", "9", "--gop", "8", "--sr", "16", "--subpel", "2", "--searchmode", "3", "--all-partitions"],
    ["--frames", "5", "--gop", "4", "--sr", "32", "--subpel", "0", "--searchmode", "2", "--refs", "2"],
    ["--width", "352", "--height", "288", "--frames", "3", "--gop", "2", "--sr", "8", "--subpel", "2", "--searchmode", "1"],
    ["--frames", "5", "--gop", "4", "--sr", "16", "--subpel", "2", "--fastbi", "1", "--itersr", "8"],
    # YUV-SAD full-pel search on textured chroma (block search and TZ), and the mixed cost-factor classes
    ["--frames", "5", "--gop", "4", "--sr", "8", "--subpel", "2", "--fullpel", "3", "--chroma"],
    ["--frames", "5", "--gop", "4", "--sr", "16", "--subpel", "2", "--fullpel", "3", "--chroma", "--searchmode", "4", "--wp", "1", "--wbp", "1", "--fade", "0.04"],
    ["--frames", "5", "--gop", "4", "--sr", "8", "--subpel", "1", "--fullpel", "0"],
    ["--frames", "5", "--gop", "4", "--sr", "8", "--subpel", "2", "--fullpel", "1"],
], ids=["C1-qcif-hierB-satd", "all41-2refs-sad", "C2-two-spatial-layers", "C5-spatial+quality-2refs", "fullpel-hadamard", "ssd-ssd",
        "weighted-explicit", "weighted-implicit", "tz-shipped-defaults-2-layers", "fast-all41", "log-2refs", "spiral-cif", "block-fastbi",
        "yuvsad-block", "yuvsad-tz-weighted", "sad-then-ssd", "ssd-then-satd"])
def test_encoder_with_dropin_is_bit_identical(args):
    import system_parity
    if not (os.path.exists(system_parity.STOCK) and os.path.exists(system_parity.B200)):
        pytest.skip("reference encoder binaries not built (need /root/reference at build time)")
    res = system_parity.compare(args + ["--timeout", "600"])
    assert res["stock"]["bitstream_md5"] == res["b200"]["bitstream_md5"], res
    assert res["stock"]["recon_md5"] == res["b200"]["recon_md5"], res
    assert res["stock"]["bytes"] > 1000
    assert res["b200"]["device_path"] and "estimateBlockWithStart 0 " not in res["b200"]["device_path"]     # the search really ran on the device
    if "--saturate" in args:                                   # residual prediction reached the search with signed 16-bit originals (N3)
        assert "signed 16-bit originals 0," not in res["b200"]["device_path"], res["b200"]["device_path"]
    if "--wp" in args:                                         # the weighted originals (a14) really went through the device search
        assert "weighted uni 0," not in res["b200"]["device_path"], res["b200"]["device_path"]
        which = "weighted bi explicit 0," if args[args.index("--wbp") + 1] == "1" else "weighted bi implicit 0)"
        assert which not in res["b200"]["device_path"], res["b200"]["device_path"]
    print(res["b200"]["device_path"])
