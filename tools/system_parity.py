#!/usr/bin/env python
"""System-level parity: the UNMODIFIED reference encoder CLI (oracle/_ref/H264AVCEncoderLibTestStatic) against the
same CLI linked with the B200 drop-in (oracle/_ref/H264AVCEncoderLibTestB200, recipe in oracle/Makefile) on a synthetic
sequence: bitstream and reconstruction must be byte-identical (the check of R/data/runAndCheckSVC.bat:45-50, applied
across the two builds).  Needs a GPU for the drop-in leg.

    python tools/system_parity.py [--width 176 --height 144 --frames 9 --gop 8 --sr 16 --subpel 2] [--only stock|b200]
"""
import argparse
import hashlib
import importlib
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
STOCK = os.path.join(ROOT, "oracle", "_ref", "H264AVCEncoderLibTestStatic")
B200 = os.path.join(ROOT, "oracle", "_ref", "H264AVCEncoderLibTestB200")

MAIN_CFG = """OutputFile {out}
FrameRate 30.0
MaxDelay 1200.0
FramesToBeEncoded {frames}
NonRequiredEnable 0
CgsSnrRefinement 1
EncodeKeyPictures 1
MGSControl 0
GOPSize {gop}
IntraPeriod -1
NumberReferenceFrames {refs}
BaseLayerMode 2
ConstrainedIntraUps 1
SearchMode {searchmode}
SearchFuncFullPel {fullpel}
SearchFuncSubPel {subpel}
SearchRange {sr}
ELSearchRange {elsr}
FastBiSearch {fastbi}
BiPredIter {bipreditr}
IterSearchRange {itersr}
LoopFilterDisable 0
LoopFilterAlphaC0Offset 0
LoopFilterBetaOffset 0
InterLayerLoopFilterDisable 2
InterLayerLoopFilterAlphaC0Offset 0
InterLayerLoopFilterBetaOffset 0
NumLayers {nlayers}
{layercfgs}WeightedPrediction {wp}
WeightedBiprediction {wbp}
LARDO 0
MultiLayerLambdaSel 0
PreAndSuffixUnitEnable 0
"""

LAYER_CFG = """SourceWidth {w}
SourceHeight {h}
FrameRateIn 30
FrameRateOut 30
InputFile {yuv}
ReconFile {rec}
SymbolMode 1
IDRPeriod 0
IntraPeriod 0
MbAff 0
PAff 0
BottomFieldFirst 0
LowComplexityMbMode 0
ProfileIdc {profile}
MinLevelIdc 13
UseLongTerm 0
MMCOEnable 0
MMCOBaseEnable 0
Enable8x8Transform 1
ConstrainedIntraPred 0
CbQPIndexOffset 0
CrQPIndexOffset 0
ScalingMatricesPresent 0
IPCMRate 0
BiPredLT8x8Disable {lt8}
MCBlocksLT8x8Disable {lt8}
DisableBSlices 0
MaxDeltaQP 0
QP {qp}
ForceReOrdering 0
InterLayerPred {ilpred}
ILModePred {ilpred}
ILMotionPred {ilpred}
ILResidualPred {ilpred}
SliceSkip 0
UseESS 0
"""


def md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def write_yuv(path, frames_u8, textured_chroma=False):
    h, w = frames_u8[0].shape
    chroma = np.full((h // 2) * (w // 2) * 2, 128, np.uint8).tobytes()
    with open(path, "wb") as f:
        for fr in frames_u8:
            f.write(np.ascontiguousarray(fr).tobytes())
            if textured_chroma:                       # chroma that follows the luma motion: 2x2 means of the luma, Cr inverted
                c = ((fr[0::2, 0::2].astype(np.uint16) + fr[0::2, 1::2] + fr[1::2, 0::2] + fr[1::2, 1::2] + 2) >> 2).astype(np.uint8)
                f.write(c.tobytes()); f.write((255 - c).tobytes())
            else:
                f.write(chroma)


def layer_plan(a):
    """Spatial layers from the top resolution down: --layers 2 --quality-layers 1 at 352x288 gives QCIF -> CIF -> CIF(CGS).
    Each dyadic layer halves the picture (sizes stay multiples of 16 at the layers used here)."""
    sizes = [(a.width >> (a.layers - 1 - i), a.height >> (a.layers - 1 - i)) for i in range(a.layers)]
    plan = [{"w": w, "h": h, "qp": a.qp + 2.0 * (a.layers - 1 - i), "spatial": i} for i, (w, h) in enumerate(sizes)]
    for q in range(a.quality_layers):
        plan.append({"w": a.width, "h": a.height, "qp": a.qp - 4.0 * (q + 1), "spatial": a.layers - 1})
    for l in plan:
        if l["w"] % 16 or l["h"] % 16:
            raise ValueError(f"layer size {l['w']}x{l['h']} is not a multiple of 16")
    return plan


def write_inputs(workdir, a):
    synth = importlib.import_module("jsvm-code-notes_b200.synth")
    seq = synth.Sequence(a.width, a.height, a.seed)
    top = [seq.frame(n) for n in range(a.frames)]
    if a.fade:
        # linear fade to black / brightening: picture n is scaled by (1 - fade * n) so that the encoder's weight estimation
        # (explicit WeightedPrediction / WeightedBiprediction) finds non-default weights and offsets
        top = [np.clip(np.rint(f.astype(np.float64) * (1.0 - a.fade * n)), 0, 255).astype(np.uint8) for n, f in enumerate(top)]
    if a.saturate:
        # contrast x3 around mid-grey with hard clipping: large areas at 0 and 255, so that "original minus base-layer
        # residual" leaves [0,255] in the residual-prediction passes (N3)
        top = [np.clip((f.astype(np.int32) - 128) * 3 + 128, 0, 255).astype(np.uint8) for f in top]
    by_spatial = {a.layers - 1: top}
    cur = top
    for s in range(a.layers - 2, -1, -1):                    # 2x2 box down-sampling for the lower spatial layers
        cur = [((f[0::2, 0::2].astype(np.uint16) + f[0::2, 1::2] + f[1::2, 0::2] + f[1::2, 1::2] + 2) >> 2).astype(np.uint8) for f in cur]
        by_spatial[s] = cur
    for s, frs in by_spatial.items():
        write_yuv(os.path.join(workdir, f"in_s{s}.yuv"), frs, a.chroma)


def run(exe, workdir, tag, a):
    out = os.path.join(workdir, f"{tag}.264")
    main_cfg = os.path.join(workdir, f"{tag}_encoder.cfg")
    plan = layer_plan(a)
    lines, recs = "", []
    for i, l in enumerate(plan):
        layer = os.path.join(workdir, f"{tag}_layer{i}.cfg")
        rec = os.path.join(workdir, f"{tag}_rec{i}.yuv")
        recs.append(rec)
        open(layer, "w").write(LAYER_CFG.format(w=l["w"], h=l["h"], yuv=os.path.join(workdir, f"in_s{l['spatial']}.yuv"), rec=rec, qp=l["qp"],
                                                lt8=0 if a.all_partitions else 1, ilpred=0 if i == 0 else a.ilpred,
                                                profile=100 if i == 0 else 86))
        lines += f"LayerCfg {layer}\n"
    open(main_cfg, "w").write(MAIN_CFG.format(out=out, frames=a.frames, gop=a.gop, refs=a.refs, sr=a.sr, subpel=a.subpel, fullpel=a.fullpel,
                                              nlayers=len(plan), layercfgs=lines, wp=a.wp, wbp=a.wbp,
                                              searchmode=a.searchmode, elsr=a.elsr, fastbi=a.fastbi, bipreditr=a.bipreditr, itersr=a.itersr))
    t0 = time.perf_counter()
    p = subprocess.run([exe, "-pf", main_cfg], cwd=workdir, capture_output=True, text=True, timeout=a.timeout,
                       env=dict(os.environ, JSVM_ME_STATS="1"))
    dt = time.perf_counter() - t0
    if p.returncode != 0 or not os.path.exists(out):
        sys.stderr.write(p.stdout[-3000:] + "\n" + p.stderr[-3000:] + "\n")
        raise RuntimeError(f"{os.path.basename(exe)} failed with code {p.returncode}")
    h = hashlib.md5()
    for r in recs:
        h.update(md5(r).encode())
    stats = [ln for ln in p.stderr.splitlines() if ln.startswith("jsvm_me_b200 stats:")]
    return {"bitstream_md5": md5(out), "recon_md5": h.hexdigest(), "bytes": os.path.getsize(out), "seconds": round(dt, 2), "layers": len(plan),
            "device_path": stats[-1][len("jsvm_me_b200 stats: "):] if stats else None}


def compare(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=176)
    ap.add_argument("--height", type=int, default=144)
    ap.add_argument("--frames", type=int, default=9)
    ap.add_argument("--gop", type=int, default=8)
    ap.add_argument("--refs", type=int, default=1)
    ap.add_argument("--sr", type=int, default=16)
    ap.add_argument("--subpel", type=int, default=2)
    ap.add_argument("--fullpel", type=int, default=0, help="SearchFuncFullPel: 0 SAD, 1 SSD, 2 HADAMARD, 3 YUV-SAD")
    ap.add_argument("--searchmode", type=int, default=0, help="SearchMode: 0 block, 1 spiral, 2 log, 3 fast, 4 TZ (the shipped default)")
    ap.add_argument("--elsr", type=int, default=0, help="ELSearchRange")
    ap.add_argument("--fastbi", type=int, default=0, help="FastBiSearch")
    ap.add_argument("--bipreditr", type=int, default=2, help="BiPredIter")
    ap.add_argument("--itersr", type=int, default=4, help="IterSearchRange")
    ap.add_argument("--chroma", action="store_true", help="textured chroma planes instead of flat 128 (YUV-SAD)")
    ap.add_argument("--qp", type=float, default=32.0)
    ap.add_argument("--all-partitions", action="store_true")
    ap.add_argument("--seed", type=int, default=1235)
    ap.add_argument("--saturate", action="store_true", help="clip the synthetic luma hard at 0 / 255")
    ap.add_argument("--layers", type=int, default=1, help="dyadic spatial layers; --width/--height name the TOP layer")
    ap.add_argument("--quality-layers", type=int, default=0, help="extra CGS quality layers at the top resolution")
    ap.add_argument("--ilpred", type=int, default=2, help="InterLayerPred / ILModePred / ILMotionPred / ILResidualPred of enhancement layers")
    ap.add_argument("--wp", type=int, default=0, help="WeightedPrediction (P pictures): 0 off, 1 explicit")
    ap.add_argument("--wbp", type=int, default=0, help="WeightedBiprediction (B pictures): 0 off, 1 explicit, 2 implicit")
    ap.add_argument("--fade", type=float, default=0.0, help="per-picture luma fade factor (exercises weighted prediction)")
    ap.add_argument("--only", choices=["stock", "b200"], default=None)
    ap.add_argument("--timeout", type=int, default=1500)
    a = ap.parse_args(argv)
    res = {}
    with tempfile.TemporaryDirectory() as wd:
        write_inputs(wd, a)
        if a.only != "b200":
            res["stock"] = run(STOCK, wd, "stock", a)
        if a.only != "stock":
            res["b200"] = run(B200, wd, "b200", a)
    return res


def main(argv=None):
    res = compare(argv)
    for k, v in res.items():
        print(k, v)
    if len(res) == 2:
        same = res["stock"]["bitstream_md5"] == res["b200"]["bitstream_md5"] and res["stock"]["recon_md5"] == res["b200"]["recon_md5"]
        print("IDENTICAL" if same else "MISMATCH")
        return 0 if same else 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
