#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libjsvm_ref.so, built from
/root/reference by `make -C oracle ref`).  Run in the build container only; the fixtures are committed so
that the GPU box (which has no /root/reference) can pin the oracle and the CUDA path against them.

    python tests/golden/make_golden.py [group ...]      groups: search, org, fullpel, table, yuv, mc, fast, pic (default: all)
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402

synth = importlib.import_module("jsvm-code-notes_b200.synth")
abi = importlib.import_module("jsvm-code-notes_b200.abi")

CASES = {
    # name: (W, H, search_range, sub_dfunc, parts, field, seed, mb_subset, qp)
    "tiny_all41_had": (64, 48, 12, 2, "PARTS_41", "split", 21, None, 28),
    "tiny_9_sad": (64, 48, 16, 0, "PARTS_9", "random", 22, None, 36),
    "qcif_9_had": (176, 144, 16, 2, "PARTS_9", "random", 23, [(0, 0), (0, 10), (8, 0), (8, 10), (4, 5), (3, 7), (0, 5), (8, 4)], 30),
    "qcif_9_zero_sr32": (176, 144, 32, 2, "PARTS_9", "zero", 24, [(0, 0), (8, 10), (4, 5), (2, 9)], 33),
    # SR 64 (the search range of C4 / C5): CIF, all four corners, edges and interior macroblocks, windows clipped on every side
    "cif_9_sr64": (352, 288, 64, 2, "PARTS_9", "random", 25,
                   [(0, 0), (0, 21), (17, 0), (17, 21), (0, 9), (17, 12), (8, 0), (9, 21), (4, 4), (8, 11), (13, 17), (5, 16)], 35),
}
PLANE_FREE = {"cif_9_sr64"}          # fixtures that do not carry the half-pel / margin planes (size)


def gen_search(ref):
    for name, (W, H, sr, sub, parts, field, seed, subset, qp) in CASES.items():
        seq = synth.Sequence(W, H, seed)
        f0, f1 = seq.frame(0), seq.frame(1)
        ses = ref.session(W, H, 2, sr, sub)
        ses.set_plane(0, f0)
        ses.set_plane(1, f1)
        jobs = synth.make_jobs(W, H, getattr(synth, parts), 1, 0, qp=qp, field=field, seed=seed, mb_subset=subset, bits_in=2)
        res, preds = ses.search(jobs, want_preds=True)
        planes = {} if name in PLANE_FREE else dict(halfpel0=ses.halfpel(0), fullpel0=ses.fullpel(0))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), width=W, height=H, search_range=sr, sub_dfunc=sub,
                            frame0=f0, frame1=f1, jobs=jobs, results=res, preds=preds, **planes)
        print(name, len(jobs), "jobs")


def gen_org(ref):
    # weighted / bi-pred style originals handed in by the caller (8-bit range) and a signed 16-bit one
    W, H, sr = 64, 48, 8
    seq = synth.Sequence(W, H, 31)
    f0, f1 = seq.frame(0), seq.frame(1)
    rng = np.random.default_rng(31)
    org8 = rng.integers(0, 256, (3, 16, 16)).astype(np.int16)
    org16 = rng.integers(-255, 511, (3, 16, 16)).astype(np.int16)
    for nm, org in (("org_override_8bit", org8), ("org_override_16bit", org16)):
        ses = ref.session(W, H, 2, sr, 2)
        ses.set_plane(0, f0); ses.set_plane(1, f1)
        jobs = synth.make_jobs(W, H, synth.PARTS_9, 1, 0, qp=30, field="random", seed=5, mb_subset=[(0, 0), (1, 2), (2, 3)])
        jobs["org_index"] = np.repeat(np.arange(3), 9)
        res, preds = ses.search(jobs, org_mbs=org, want_preds=True)
        np.savez_compressed(os.path.join(HERE, nm + ".npz"), width=W, height=H, search_range=sr, sub_dfunc=2,
                            frame0=f0, frame1=f1, jobs=jobs, results=res, preds=preds, org_mbs=org)
        print(nm, len(jobs), "jobs")


def gen_fullpel(ref):
    # full-pel Hadamard / SSD searches (generic full-search kernel), 8-bit and signed 16-bit originals
    W, H, sr = 64, 48, 8
    seq = synth.Sequence(W, H, 33)
    f0, f1 = seq.frame(0), seq.frame(1)
    org16b = np.random.default_rng(33).integers(-255, 511, (12, 16, 16)).astype(np.int16)
    for nm, full, sub, parts, org in (("fullpel_had_had", 2, 2, "PARTS_9", None), ("fullpel_had_sad_all41", 2, 0, "PARTS_41", None),
                                      ("fullpel_ssd_ssd", 1, 1, "PARTS_9", None), ("fullpel_ssd_ssd_16bit", 1, 1, "PARTS_9", org16b),
                                      ("fullpel_had_had_16bit", 2, 2, "PARTS_9", org16b)):
        ses = oracle_lib.CpuSession(ref, W, H, 2, sr, sub, full)
        ses.set_plane(0, f0); ses.set_plane(1, f1)
        jobs = synth.make_jobs(W, H, getattr(synth, parts), 1, 0, qp=29, field="random", seed=6, bits_in=1)
        extra = {}
        if org is not None:
            jobs["org_index"] = np.repeat(np.arange(12), len(getattr(synth, parts)))
            extra["org_mbs"] = org
        res, preds = ses.search(jobs, org_mbs=org, want_preds=True)
        np.savez_compressed(os.path.join(HERE, nm + ".npz"), width=W, height=H, search_range=sr, sub_dfunc=sub, full_dfunc=full,
                            frame0=f0, frame1=f1, jobs=jobs, results=res, preds=preds, **extra)
        print(nm, len(jobs), "jobs")


def gen_table(ref):
    # distortion function table (XDistortion SAD / SSD / HADAMARD for all 7 block modes) and cost factors
    rng = np.random.default_rng(31)
    rng.integers(0, 256, (3, 16, 16)); rng.integers(-255, 511, (3, 16, 16))      # (the draws gen_org makes from the same stream)
    orgs = rng.integers(0, 256, (8, 16, 16)).astype(np.int16)
    refs = rng.integers(0, 256, (8, 20, 24)).astype(np.int16)
    rows = []
    for i in range(8):
        for mode, blk in synth.PARTS_41[::3]:
            for df in (0, 1, 2):
                bx, by = 4 * (blk & 3), 4 * (blk >> 2)
                sub = np.ascontiguousarray(refs[i])
                v = ref.distortion(orgs[i].ctypes.data, sub[by:, bx:].ctypes.data if False else sub.ctypes.data + 2 * (by * 24 + bx), 24, mode, df, blk)
                rows.append((i, mode, blk, df, v))
    factors = np.array([[ref.cost_factor(synth.lambda_for_qp(q), 1), ref.cost_factor(synth.lambda_for_qp(q), 0)] for q in range(52)], np.uint32)
    np.savez_compressed(os.path.join(HERE, "distortion_and_factors.npz"), orgs=orgs, refs=refs,
                        rows=np.array(rows, np.int64), factors=factors)
    print("distortion rows", len(rows))


def gen_yuv(ref):
    # SearchFuncFullPel = YUV_SAD (chroma planes), and the mixed cost-factor classes (SAD-class full-pel + SSD sub-pel and back)
    W, H, sr = 64, 48, 8
    seq = synth.Sequence(W, H, 35)
    f0, f1 = seq.frame(0), seq.frame(1)
    c0, c1 = synth.chroma_planes(W, H, 35), synth.chroma_planes(W, H, 36)
    rng = np.random.default_rng(35)
    org = rng.integers(0, 256, (12, 16, 16)).astype(np.int16)
    orgc = rng.integers(0, 256, (12, 2, 8, 8)).astype(np.int16)
    fs, fe = synth.cost_factor_for_qp(29), ref.cost_factor(synth.lambda_for_qp(29), 0)
    cases = (("yuvsad_had_all41", 3, 2, "PARTS_41", None, None), ("yuvsad_sad_org_chroma", 3, 0, "PARTS_9", org, orgc),
             ("yuvsad_had_org_lumaonly", 3, 2, "PARTS_9", org, None), ("yuvsad_ssd_mixed", 3, 1, "PARTS_9", None, None),
             ("sad_ssd_mixed", 0, 1, "PARTS_41", None, None), ("ssd_had_mixed", 1, 2, "PARTS_9", None, None))
    for nm, full, sub, parts, o, oc in cases:
        ses = oracle_lib.CpuSession(ref, W, H, 2, sr, sub, full)
        ses.set_plane(0, f0); ses.set_plane(1, f1)
        ses.set_chroma(0, *c0); ses.set_chroma(1, *c1)
        jobs = synth.make_jobs(W, H, getattr(synth, parts), 1, 0, qp=29, field="split", seed=7, bits_in=1)
        mixed = (full == 1) != (sub == 1)
        if full == 1:
            jobs["cost_factor"] = fe
        if mixed:
            ses.map_cost_factor(fe if full == 1 else fs, fs if full == 1 else fe)
        extra = dict(factor_pair=np.array([fe if full == 1 else fs, fs if full == 1 else fe], np.uint32)) if mixed else {}
        if o is not None:
            jobs["org_index"] = np.repeat(np.arange(12), len(getattr(synth, parts)))
            extra["org_mbs"] = o
        if oc is not None:
            ses.set_org_chroma(oc)
            extra["org_chroma"] = oc
        res, preds = ses.search(jobs, org_mbs=o, want_preds=True)
        np.savez_compressed(os.path.join(HERE, nm + ".npz"), width=W, height=H, search_range=sr, sub_dfunc=sub, full_dfunc=full,
                            frame0=f0, frame1=f1, cb0=c0[0], cr0=c0[1], cb1=c1[0], cr1=c1[1], jobs=jobs, results=res, preds=preds, **extra)
        print(nm, len(jobs), "jobs")


def gen_mc(ref):
    # MotionCompensation: predBlk (all 16 fractional phases), chroma 1/8-sample filter, SampleWeighting variants
    W, H = 64, 48
    seq = synth.Sequence(W, H, 37)
    frames = [seq.frame(i) for i in range(3)]
    chroma = [synth.chroma_planes(W, H, 40 + i) for i in range(3)]
    ses = ref.session(W, H, 3, 8, 0)
    for s in range(3):
        ses.set_plane(s, frames[s]); ses.set_chroma(s, *chroma[s])
    out = dict(width=W, height=H, frames=np.stack(frames), cb=np.stack([c[0] for c in chroma]), cr=np.stack([c[1] for c in chroma]))
    for kind in ("uni", "bi", "uni_w", "bi_w", "bi_implicit"):
        jobs = synth.make_mc_jobs(W, H, 400, 3, seed=len(kind), kind=kind)
        py, pcb, pcr = ses.compensate(jobs)
        out["jobs_" + kind], out["y_" + kind], out["cb_" + kind], out["cr_" + kind] = jobs, py, pcb, pcr
        print("mc", kind, len(jobs), "jobs")
    np.savez_compressed(os.path.join(HERE, "mc_blocks.npz"), **out)


FAST_CASES = {
    # name: (search_mode, SR, EL range, FastBiSearch, full dfunc, sub dfunc, parts)
    "fast_spiral_sr8": (1, 8, 0, 0, 0, 2, "PARTS_9"),
    "fast_log_sr16": (2, 16, 0, 0, 0, 2, "PARTS_41"),
    "fast_log_sr64_had": (2, 64, 0, 1, 2, 0, "PARTS_9"),
    "fast_fme_sr32": (3, 32, 0, 0, 0, 2, "PARTS_41"),
    "fast_tz_sr32": (4, 32, 8, 0, 0, 2, "PARTS_41"),
    "fast_tz_sr64_fastbi": (4, 64, 16, 1, 0, 0, "PARTS_9"),
    "fast_tz_sr16_yuv": (4, 16, 4, 0, 3, 2, "PARTS_9"),
    "block_fastbi_sr16": (0, 16, 0, 1, 0, 2, "PARTS_41"),
}


def gen_fast(ref):
    # SPIRAL / LOG / FAST / TZ searches (and the FastBiSearch refinement under BLOCK_SEARCH) on QCIF: neighbour-MV candidates,
    # EL ranges, bi-predictive calls, per-call ranges; macroblock subset with the corners and edges
    W, H = 176, 144
    mbs = [(0, 0), (0, 10), (8, 0), (8, 10), (4, 5), (3, 7), (0, 5), (8, 4), (5, 0), (2, 10)]
    for name, (mode, sr, el, fastbi, full, sub, parts) in FAST_CASES.items():
        seq = synth.Sequence(W, H, 44 + mode)
        f0, f1 = seq.frame(0), seq.frame(2)
        c0, c1 = synth.chroma_planes(W, H, 45), synth.chroma_planes(W, H, 46)
        ses = oracle_lib.CpuSession(ref, W, H, 2, sr, sub, full, mode)
        ses.set_plane(0, f0); ses.set_plane(1, f1)
        ses.set_chroma(0, *c0); ses.set_chroma(1, *c1)
        subset = [(4, 5), (3, 7), (5, 4), (4, 6)] if mode == 1 else mbs       # spiral: interior only (un-clipped window)
        jobs = synth.make_jobs(W, H, getattr(synth, parts), 1, 0, qp=31, field="split", seed=mode + sr, mb_subset=subset, bits_in=1)
        jobs, cands = synth.make_fast_search_extras(jobs, sr, seed=sr + mode)
        ses.set_fast_search_params(el, fastbi)
        ses.set_mv_candidates(cands)
        res, preds = ses.search(jobs, want_preds=True)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), width=W, height=H, search_range=sr, sub_dfunc=sub, full_dfunc=full,
                            search_mode=mode, el_search_range=el, fast_bi_search=fastbi, frame0=f0, frame1=f1,
                            cb0=c0[0], cr0=c0[1], cb1=c1[0], cr1=c1[1], jobs=jobs, cands=cands, results=res, preds=preds)
        print(name, len(jobs), "jobs")


def gen_pic(ref):
    # picture-level search: the reference's getMvPredictor on a real MbDataCtrl + estimateBlockWithStart per partition and reference
    W, H, sr = 176, 144, 16
    seq = synth.LocalMotionSequence(W, H, 51)
    frames = [seq.frame(i) for i in range(4)]
    ses = ref.session(W, H, 4, sr, 2)
    for s, f in enumerate(frames):
        ses.set_plane(s, f)
    tasks = synth.make_pic_tasks([2, 3, 3], [[1, 0], [2, 1], [0, 2]], qp=31)
    out = ses.estimate_pictures(tasks)
    np.savez_compressed(os.path.join(HERE, "picture_search_qcif.npz"), width=W, height=H, search_range=sr, sub_dfunc=2,
                        frames=np.stack(frames), tasks=tasks, field=out)
    print("picture_search_qcif", out.shape, "modes", np.bincount(out["mode"].ravel()))


def gen_resample(ref):
    # residual upsampling (DownConvert::xResidualUpsampling): dyadic, level <= 30 arithmetic, ESS with cropping window, other phases
    cases = [dict(ref_w=176, ref_h=144, cur_w=352, cur_h=288), dict(ref_w=176, ref_h=144, cur_w=352, cur_h=288, level_idc=30),
             dict(ref_w=176, ref_h=144, cur_w=272, cur_h=224, scaled_w=264, scaled_h=216, left=4, top=6),
             dict(ref_w=160, ref_h=128, cur_w=352, cur_h=288, scaled_w=300, scaled_h=250, left=32, top=16, phases=(0, 1, -1, -1))]
    out = {}
    for i, case in enumerate(cases):
        p, y, cb, cr, t8 = synth.make_resize_case(seed=60 + i, **case)
        oy, ocb, ocr = ref.upsample_residual(p, y, cb, cr, t8)
        out.update({f"params{i}": np.array([getattr(p, n) for n, _ in p._fields_], np.int32), f"y{i}": y, f"cb{i}": cb, f"cr{i}": cr, f"t8_{i}": t8,
                    f"oy{i}": oy, f"ocb{i}": ocb, f"ocr{i}": ocr})
        print("residual upsampling", case, int(np.abs(oy).sum()))
    np.savez_compressed(os.path.join(HERE, "residual_upsampling.npz"), n=len(cases), **out)


MOTION_GOLDEN = [
    (dict(ref_w=176, ref_h=144, cur_w=352, cur_h=288), 0, 0, 1), (dict(ref_w=176, ref_h=144, cur_w=352, cur_h=288), 1, 0, 1),
    (dict(ref_w=176, ref_h=144, cur_w=352, cur_h=288), 1, 16, 0),
    (dict(ref_w=176, ref_h=144, cur_w=272, cur_h=224, scaled_w=264, scaled_h=216, left=4, top=6), 1, 0, 1),
    (dict(ref_w=160, ref_h=128, cur_w=352, cur_h=288, scaled_w=300, scaled_h=250, left=32, top=16, level_idc=30), 0, 0, 0),
    (dict(ref_w=176, ref_h=144, cur_w=176, cur_h=144), 1, 0, 1),
]


def gen_motion(ref):
    # inter-layer motion prediction (MbDataCtrl::upsampleMotion): dyadic P / B, ESS, arbitrary ratio, same resolution
    out = {}
    for i, (case, st, dq, d8) in enumerate(MOTION_GOLDEN):
        resize = synth.make_resize_case(seed=1, **case)[0]
        base = synth.make_base_motion(case["ref_w"], case["ref_h"], seed=70 + i, slice_type=st)
        p = synth.make_motion_params(resize, st, dq, d8)
        pred = ref.upsample_motion(p, base)
        out.update({f"resize{i}": np.array([getattr(resize, n) for n, _ in resize._fields_], np.int32), f"switches{i}": np.array([st, dq, d8, 1, 1], np.int32),
                    f"base{i}": base, f"pred{i}": pred})
        print("motion upsampling", case, "modes", np.bincount(pred["mb_mode"].ravel()))
    np.savez_compressed(os.path.join(HERE, "motion_upsampling.npz"), n=len(MOTION_GOLDEN), **out)


GROUPS = {"motion": gen_motion, "resample": gen_resample, "pic": gen_pic, "fast": gen_fast, "yuv": gen_yuv, "mc": gen_mc, "search": gen_search, "org": gen_org, "fullpel": gen_fullpel, "table": gen_table}


def main():
    ref = oracle_lib.CpuBackend(os.path.join(ROOT, "oracle", "_ref", "libjsvm_ref.so"), "jsvm_ref_")
    for g in (sys.argv[1:] or list(GROUPS)):
        GROUPS[g](ref)


if __name__ == "__main__":
    main()
