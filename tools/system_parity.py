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
SearchMode 0
SearchFuncFullPel 0
SearchFuncSubPel {subpel}
SearchRange {sr}
ELSearchRange 0
FastBiSearch 0
BiPredIter 2
IterSearchRange 4
LoopFilterDisable 0
LoopFilterAlphaC0Offset 0
LoopFilterBetaOffset 0
InterLayerLoopFilterDisable 2
InterLayerLoopFilterAlphaC0Offset 0
InterLayerLoopFilterBetaOffset 0
NumLayers 1
LayerCfg {layer}
WeightedPrediction 0
WeightedBiprediction 0
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
ProfileIdc 100
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
InterLayerPred 0
ILModePred 0
ILMotionPred 0
ILResidualPred 0
SliceSkip 0
UseESS 0
"""


def md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def write_yuv(path, width, height, frames, seed):
    synth = importlib.import_module("jsvm-code-notes_b200.synth")
    seq = synth.Sequence(width, height, seed)
    chroma = np.full((height // 2) * (width // 2) * 2, 128, np.uint8).tobytes()
    with open(path, "wb") as f:
        for n in range(frames):
            f.write(seq.frame(n).tobytes())
            f.write(chroma)


def run(exe, workdir, tag, a):
    out = os.path.join(workdir, f"{tag}.264")
    rec = os.path.join(workdir, f"{tag}_rec.yuv")
    layer = os.path.join(workdir, f"{tag}_layer0.cfg")
    main_cfg = os.path.join(workdir, f"{tag}_encoder.cfg")
    open(layer, "w").write(LAYER_CFG.format(w=a.width, h=a.height, yuv=os.path.join(workdir, "in.yuv"), rec=rec, qp=a.qp,
                                            lt8=0 if a.all_partitions else 1))
    open(main_cfg, "w").write(MAIN_CFG.format(out=out, frames=a.frames, gop=a.gop, refs=a.refs, sr=a.sr, subpel=a.subpel, layer=layer))
    t0 = time.perf_counter()
    p = subprocess.run([exe, "-pf", main_cfg], cwd=workdir, capture_output=True, text=True, timeout=a.timeout)
    dt = time.perf_counter() - t0
    if p.returncode != 0 or not os.path.exists(out):
        sys.stderr.write(p.stdout[-3000:] + "\n" + p.stderr[-3000:] + "\n")
        raise RuntimeError(f"{os.path.basename(exe)} failed with code {p.returncode}")
    return {"bitstream_md5": md5(out), "recon_md5": md5(rec), "bytes": os.path.getsize(out), "seconds": round(dt, 2)}


def compare(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=176)
    ap.add_argument("--height", type=int, default=144)
    ap.add_argument("--frames", type=int, default=9)
    ap.add_argument("--gop", type=int, default=8)
    ap.add_argument("--refs", type=int, default=1)
    ap.add_argument("--sr", type=int, default=16)
    ap.add_argument("--subpel", type=int, default=2)
    ap.add_argument("--qp", type=float, default=32.0)
    ap.add_argument("--all-partitions", action="store_true")
    ap.add_argument("--seed", type=int, default=1235)
    ap.add_argument("--only", choices=["stock", "b200"], default=None)
    ap.add_argument("--timeout", type=int, default=1500)
    a = ap.parse_args(argv)
    res = {}
    with tempfile.TemporaryDirectory() as wd:
        write_yuv(os.path.join(wd, "in.yuv"), a.width, a.height, a.frames, a.seed)
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
