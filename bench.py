#!/usr/bin/env python
"""bench.py -- 1080p full-search ME frame-pairs/s (SR+-64, QPel) on N B200s, with roofline and CPU baseline.

One "step" = one pass of the hot path over one batch: the top temporal level of a GOP-16 hierarchical-B
sequence at 1920x1088 (BASELINE.json configs[3]): 8 B pictures x 2 lists = 16 independent (current,
reference) pairs per GPU -- K1 half-pel planes of the 9 reference pictures, K2 full search (SR +-64) and
K3 quarter-pel refinement for all 8160 macroblocks x 9 partitions of every pair (1.17 M
estimateBlockWithStart results per step per GPU).  The batch is ONE global list of pairs (N x 16 pairs = N GOPs for weak scaling,
16 pairs in total with --scaling strong) dealt in contiguous blocks to the ranks by sharding.shard_pairs: no data-path collective; for N > 1
the result field is gathered to rank 0 over NCCL by the library itself (jsvm_me_gather_results_device) inside the timed region.

`value`   : pairs/s with planes, jobs and search groups already resident in HBM (CUDA events, max over ranks).
`e2e`     : pairs/s through the host-buffer C ABI (jsvm_me_upload_plane_u8 / jsvm_me_search): pinned host
            planes and jobs -> H2D, planning, kernels, results D2H, all inside the timed region.
`roofline`: K2 (dominant kernel) algorithmic INT32 lane-ops / measured launch time vs the INT32 ALU-pipe peak
            measured by tools/int32_peak on this pool; `roofline_k1` is the HBM roofline of the half-pel kernel.
`cpu_baseline` / `--impl reference`: the unmodified reference (oracle/_ref) on the host cores, bounded sample.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "jsvm-code-notes_b200"
METRIC = "1080p full-search ME frame-pairs/s (SR+-64, QPel)"
UNIT = "frame-pairs/s"
OPS_PER_POS_9 = 87       # 64 packed abs-diff-acc + 5 tree adds + 9 rate adds + 9 compare/selects
OPS_PER_POS_41 = 171
# K3: INT32 lane-ops per (candidate, 4x4 block).  SAD: 4 packed abs-diff-accumulate (pixels / 4).  Hadamard
# (xCalcHadamard4x4, ENC/Distortion.cpp:291-340, in its cheapest exact form): 16 horizontal dot products on packed
# bytes + 16 vertical first-stage adds + 16 abs + 8 max (= last butterfly + halving) + 8 accumulates = 64.
OPS_4x4_SAD = 4
OPS_4x4_HADAMARD = 64
INT32_PEAK_FALLBACK = 1.855e13   # lane-ops/s, tools/int32_peak on this pool's B200 (profiles/int32_peak_r01.json)


def light_module(name):
    """`synth` / `abi` of the package WITHOUT running the package __init__ (which dlopens the CUDA library): the reference arm
    and the CPU checker must not map any of the product's native code."""
    import types
    alias = "jsvm_me_b200_light"
    if alias not in sys.modules:
        m = types.ModuleType(alias)
        m.__path__ = [os.path.join(ROOT, PKG)]
        sys.modules[alias] = m
    return importlib.import_module(f"{alias}.{name}")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C4")
    ap.add_argument("--field", default="random", choices=["zero", "random", "split"])
    ap.add_argument("--pairs", type=int, default=16, help="(current, reference) pairs per GPU per step")
    ap.add_argument("--cpu-sample-mbs", type=int, default=0, help="macroblocks in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: --pairs per GPU (N GOPs); strong: --pairs in total")
    ap.add_argument("--no-extras", action="store_true", help="skip the short measurements of the other kernels (TZ, picture-level, MC, resampling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--parity-fraction", type=float, default=0.05, help="fraction of the macroblocks of a step that is re-computed by "
                    "the CPU checker after the timed legs (0 = skip)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def global_pairs(n_pairs):
    """The batch: pair g = (gop, current, reference).  Every GOP contributes the 16 pairs of its top temporal level: pictures
    1, 3, .. 15 each referencing their two even neighbours."""
    out = []
    for g in range(n_pairs):
        gop, k = divmod(g, 16)
        c = 2 * (k // 2) + 1
        out.append((gop, c, c - 1 if k % 2 == 0 else c + 1))
    return out


def pair_jobs(synth, cfg, g, cur_slot, ref_slot, field):
    return synth.make_jobs(cfg["width"], cfg["height"], cfg["parts"], cur_slot, ref_slot, qp=35, field=field, seed=100 + g)   # QP of the top level (DQP4TLevel)


def build_workload(synth, sharding, cfg, n_pairs_global, field, rank, world):
    """Planes and jobs of this rank's share of the batch (contiguous blocks of whole pairs, sharding.shard_pairs)."""
    W, H = cfg["width"], cfg["height"]
    seq = synth.Sequence(W, H, 1234 + 4)
    mine = sharding.shard_pairs(global_pairs(n_pairs_global), rank, world)
    pics = sorted({(gop, p) for _, (gop, c, r) in mine for p in (c, r)})
    frames = {k: seq.frame(k[1], gop=k[0]) for k in pics}
    slot_of = {k: i for i, k in enumerate(pics)}
    jobs = [pair_jobs(synth, cfg, g, slot_of[(gop, c)], slot_of[(gop, r)], field) for g, (gop, c, r) in mine]
    jobs = np.concatenate(jobs) if jobs else np.zeros(0, synth.JOB_DTYPE)
    ref_slots = sorted({slot_of[(gop, r)] for _, (gop, c, r) in mine})
    return seq, frames, slot_of, mine, jobs, ref_slots


# ------------------------------------------------------------------------------------------------
# CPU baseline: the unmodified reference (oracle/_ref) or, if it was not built, the C oracle
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, path, prefix, W, H, sr, sub, cur, ref, jobs_bytes = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    abi = light_module("abi")
    be = oracle_lib.CpuBackend(path, prefix)
    ses = be.session(W, H, 2, sr, sub)
    ses.set_plane(0, ref)
    ses.set_plane(1, cur)
    jobs = np.frombuffer(jobs_bytes, dtype=abi.JOB_DTYPE).copy()
    jobs["ref_slot"], jobs["cur_slot"] = 0, 1
    t0 = time.perf_counter()
    ses.search(jobs)
    return time.perf_counter() - t0


def cpu_baseline(synth, cfg, field, sample_mbs, steps=1):
    """Times estimateBlockWithStart of the reference on a bounded sample of the same workload."""
    import multiprocessing as mp
    ref_path = os.path.join(ROOT, "oracle", "_ref", "libjsvm_ref.so")
    ora_path = os.path.join(ROOT, "oracle", "libme_oracle.so")
    if os.path.exists(ref_path):
        kind, path, prefix = "reference", ref_path, "jsvm_ref_"
    else:
        kind, path, prefix = "port", ora_path, "jsvm_ora_"
        if not os.path.exists(ora_path):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    W, H = cfg["width"], cfg["height"]
    cores = os.cpu_count() or 1
    seq = synth.Sequence(W, H, 1234 + 4)
    cur, ref = seq.frame(1), seq.frame(0)
    mbw, mbh = W // 16, H // 16
    if sample_mbs <= 0:
        # the reference needs ~16 ms per 1080p SR64 macroblock (9 partitions) per core: ~15 s of work per core
        sample_mbs = min(mbw * mbh, 900 * cores if kind == "reference" else 150 * cores)
    rng = np.random.default_rng(99)
    pick = rng.choice(mbw * mbh, size=sample_mbs, replace=False)
    mbs = [(int(i) // mbw, int(i) % mbw) for i in sorted(pick)]
    chunks = [mbs[i::cores] for i in range(cores)]
    chunks = [c for c in chunks if c]
    values, walls = [], []
    for _ in range(steps):
        tasks = []
        for i, c in enumerate(chunks):
            jobs = synth.make_jobs(W, H, cfg["parts"], 1, 0, qp=35, field=field, seed=7 + i, mb_subset=c)
            tasks.append((kind, path, prefix, W, H, cfg["search_range"], cfg["sub_dfunc"], cur, ref, jobs.tobytes()))
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(len(chunks)) as pool:
            times = pool.map(_cpu_worker, tasks)
        wall = max(times)                                   # search time only (plane set-up excluded), slowest worker
        values.append((sample_mbs / (mbw * mbh)) / wall)
        walls.append(time.perf_counter() - t0)                # the whole sample step incl. process start-up and plane set-up
    v = float(np.mean(values))
    return {"value": v, "unit": UNIT, "cores": len(chunks), "kind": kind,
            "sample": f"{sample_mbs} of {mbw * mbh} macroblocks of one 1080p pair x {len(cfg['parts'])} partitions, SR+-{cfg['search_range']}, "
                      f"{len(chunks)} processes x 1 thread, search time only"}, values, walls


def parity_check(synth, cfg, seq, field, legs, fraction, seed=4242):
    """Re-computes a seeded sample of the batch's macroblocks (all their partitions) with the CPU checker -- the compiled unmodified
    reference when oracle/_ref is present, else the C restatement -- and compares every result field of every timed leg with it.
    legs: {name: (global pair indices the leg covers, results in that pair order)} -- for N > 1 the gathered field of ALL ranks.
    Runs after the timed regions; the checker is never the thing measured."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    W, H = cfg["width"], cfg["height"]
    per_mb = len(cfg["parts"])
    n_mb = (W // 16) * (H // 16)
    per_pair = n_mb * per_mb
    path, prefix, kind = oracle_lib.best_checker(ROOT)
    rng = np.random.default_rng(seed)
    bad, n_jobs_checked, n_units = {}, 0, 0
    for name, (pair_ids, got) in legs.items():
        pairs = global_pairs(max(pair_ids) + 1)
        units = len(pair_ids) * n_mb
        pick = np.sort(rng.choice(units, size=max(1, int(round(fraction * units))), replace=False))
        pics = sorted({(pairs[pair_ids[u // n_mb]][0], p) for u in pick for p in pairs[pair_ids[u // n_mb]][1:]})
        slot = {k: i for i, k in enumerate(pics)}
        planes = {slot[k]: seq.frame(k[1], gop=k[0]) for k in pics}
        want_jobs, got_rows = [], []
        for li in np.unique(pick // n_mb):
            g = pair_ids[li]
            gop, c, r = pairs[g]
            jb = pair_jobs(synth, cfg, g, slot[(gop, c)], slot[(gop, r)], field)
            mbs = pick[pick // n_mb == li] % n_mb
            idx = (mbs[:, None] * per_mb + np.arange(per_mb)[None, :]).reshape(-1)
            want_jobs.append(jb[idx]); got_rows.append(got[li * per_pair + idx])
        want = oracle_lib.parallel_search(path, prefix, W, H, len(pics), cfg["search_range"], cfg["sub_dfunc"], planes, np.concatenate(want_jobs))
        g_all = np.concatenate(got_rows)
        n_bad = int(sum(np.count_nonzero(g_all[f] != want[f]) for f in want.dtype.names))
        if n_bad:
            bad[name] = n_bad
        n_jobs_checked += len(g_all); n_units += len(pick)
    return {"jobs": int(n_jobs_checked), "macroblocks": int(n_units), "fraction": float(fraction), "checker": kind,
            "legs": sorted(legs), "fields": list(oracle_lib.RESULT_DTYPE.names), "mismatches": bad}


# ------------------------------------------------------------------------------------------------
def clocks_start():
    try:
        return subprocess.Popen(
            ["nvidia-smi", "--query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
             "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return None


def clocks_stop(proc, dev_index):
    if proc is None:
        return None
    proc.terminate()
    try:
        out, _ = proc.communicate(timeout=5)
    except Exception:
        proc.kill()
        return None
    sm, mx, reasons = [], 0, set()
    for line in out.splitlines():
        f = [x.strip() for x in line.split(",")]
        if len(f) < 9 or not f[0].isdigit() or int(f[0]) != dev_index:
            continue
        try:
            sm.append(float(f[1])); mx = max(mx, float(f[2]))
        except ValueError:
            continue
        for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
            if val.lower().startswith("active"):
                reasons.add(name)
    if not sm:
        return None
    busy = [s for s in sm if s > 0.5 * mx] or sm
    return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(pairs):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
    (profiles/r01_traffic.json, written by tools/summarize_profile.py); only valid for the batch size that was profiled."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        return d["dram_bytes_per_launch"] if int(d.get("pairs", -1)) == int(pairs) else {}
    except Exception:
        return {}


def int32_peak():
    """INT32 ALU-pipe peak (lane-ops/s): measured live with tools/int32_peak if present, else the recorded value."""
    exe = os.path.join(ROOT, "tools", "int32_peak")
    if os.path.exists(exe):
        try:
            out = subprocess.run([exe, "0.5"], capture_output=True, text=True, timeout=120).stdout
            d = json.loads(out.strip().splitlines()[-1])
            return float(d["burst"]["vabsdiff4_acc"]), "measured live (tools/int32_peak, VABSDIFF4.U8.ACC chains on all SMs)"
        except Exception:
            pass
    return INT32_PEAK_FALLBACK, "recorded (profiles/int32_peak_r01.json)"


def measure_extras(pkg, synth, me, cfg, torch, dev, hbm_peak, peak):
    """Short device-timed measurements of the kernels beside the headline path (K4 motion compensation, K5 TZ search, K6
    picture-level search, K7 inter-layer resampling) at the same picture size; reported under `extras`, not part of `value`."""
    W, H, sr = cfg["width"], cfg["height"], cfg["search_range"]
    n_mb = (W // 16) * (H // 16)
    out = {}
    seq = synth.LocalMotionSequence(W, H, 99, tile=52)
    n_slots = 10
    me.configure(W, H, n_slots)
    for s_ in range(n_slots):
        me.upload_plane(s_, seq.frame(s_ % 5))
        me.upload_chroma(s_, *synth.chroma_planes(W, H, 200 + s_ % 3))
    me.interp_halfpel(list(range(n_slots)))

    def timed(fn, reps=3):
        fn(); me.sync()
        me.timing_begin()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        me.sync()
        dt = (time.perf_counter() - t0) / reps
        kt = me.timing_end()
        return dt, {k: v / reps for k, v in kt.items()}

    # K6: picture-level search, 8 independent P pictures (a temporal level), one reference each
    me.set_params(sr, cfg["sub_dfunc"])
    tasks = synth.make_pic_tasks(list(range(1, 9)), [[p - 1] for p in range(1, 9)], qp=35)
    _, kt = timed(lambda: me.estimate_pictures(tasks, download=False), reps=1)          # kernel times: plain launches (events)
    l0 = me.launch_count()
    me.estimate_pictures(tasks, download=False); me.sync()                              # captures the CUDA graph
    n_launch = me.launch_count() - l0
    t0 = time.perf_counter()
    for _ in range(3):
        me.estimate_pictures(tasks, download=False)
    me.sync()
    dt = (time.perf_counter() - t0) / 3
    one = synth.make_pic_tasks([1], [[0]], qp=35)
    me.estimate_pictures(one, download=False); me.sync()
    t0 = time.perf_counter()
    me.estimate_pictures(one, download=False); me.sync()
    dt1 = time.perf_counter() - t0
    out["picture_level_search"] = {"pictures_per_s": len(tasks) / dt, "pictures_per_call": len(tasks), "refs": 1, "ms_per_call": dt * 1e3,
                                   "ms_single_picture": dt1 * 1e3, "kernels_per_call": int(n_launch), "launch": "one CUDA graph per call",
                                   "k2_ms": kt["search_ms"], "k3_ms": kt["subpel_ms"],
                                   "what": "jsvm_me_estimate_pictures: MV prediction + wavefront scheduling + planning on the device, 9 partitions x 8160 MBs per picture"}
    # K5: TZ search with the shipped encoder.cfg switches (ELSearchRange 8, FastBiSearch 1) on 4 pairs
    jobs = np.concatenate([synth.make_jobs(W, H, cfg["parts"], 2 * i + 1, 2 * i, qp=35, field="random", seed=300 + i) for i in range(4)])
    jobs, cands = synth.make_fast_search_extras(jobs, sr, seed=5, p_el=0.0, p_bipred=0.0, p_iter=0.0)
    d_jobs = torch.from_numpy(jobs.view(np.uint8)).to(dev)
    d_cands = torch.from_numpy(cands.view(np.uint8).reshape(-1)).to(dev)
    d_res = torch.zeros(len(jobs) * 32, dtype=torch.uint8, device=dev)
    me.set_params(sr, cfg["sub_dfunc"], 0, 4)
    me.set_fast_search_params(8, 1)
    me._check(me._lib.jsvm_me_set_mv_candidates_device(me._ctx, d_cands.data_ptr()))
    info = pkg.PlanInfo()
    dt, kt = timed(lambda: me.search_device(d_jobs.data_ptr(), len(jobs), 0, info, d_res.data_ptr()))
    out["tz_search"] = {"pairs_per_s": 4 / dt, "ms_per_pair": dt * 1e3 / 4, "k5_ms_per_pair": kt["search_ms"] / 4, "k3_ms_per_pair": kt["subpel_ms"] / 4,
                        "kernel": me.last_search_kernel(), "what": "SearchMode 4 (xTZSearch), one warp per job, neighbour-MV candidates, resident inputs"}
    me._check(me._lib.jsvm_me_set_mv_candidates_device(me._ctx, None))
    me.set_fast_search_params(0, 0)
    me.set_params(sr, cfg["sub_dfunc"])
    # K4: bi-predictive 16x16 blocks with chroma, one job per macroblock of 8 pictures
    mc = synth.make_mc_jobs(W, H, 8 * n_mb, n_slots, seed=3, kind="bi", sizes=((16, 16),))
    dt, kt = timed(lambda: me.compensate(mc))
    mc_bytes = len(mc) * (2 * (256 + 128) + 384 + 48)
    out["motion_compensation"] = {"blocks_per_s": len(mc) / (kt["other_ms"] * 1e-3), "kernel_ms": kt["other_ms"], "call_ms": dt * 1e3,
                                  "roofline": {"bound": "hbm", "achieved": mc_bytes / (kt["other_ms"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                               "frac": mc_bytes / (kt["other_ms"] * 1e-3) / 1e9 / hbm_peak, "alg_bytes_per_launch": mc_bytes},
                                  "what": "jsvm_me_compensate: 16x16 bi-predictive luma + chroma blocks (predBlk / xPredChroma / xMixB)"}
    # K7: residual upsampling half size -> full size (C5: 540p -> 1080p), luma + chroma; inter-layer motion prediction
    rp, y, cb, cr, t8 = synth.make_resize_case(W // 2 // 16 * 16, H // 2 // 16 * 16, W, H, seed=2)
    dt, kt = timed(lambda: me.upsample_residual(rp, y, cb, cr, t8))
    rs_bytes = 2 * (y.size + cb.size + cr.size) + 2 * (W * H * 3 // 2)
    out["residual_upsampling"] = {"frames_per_s_kernel": 1.0 / (kt["other_ms"] * 1e-3), "kernel_ms": kt["other_ms"], "call_ms_with_copies": dt * 1e3,
                                  "roofline": {"bound": "hbm", "achieved": rs_bytes / (kt["other_ms"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                               "frac": rs_bytes / (kt["other_ms"] * 1e-3) / 1e9 / hbm_peak, "alg_bytes_per_launch": rs_bytes}}
    base = synth.make_base_motion(rp.ref_frame_width, rp.ref_frame_height, seed=4, slice_type=1)
    mp_ = synth.make_motion_params(rp, 1)
    dt, kt = timed(lambda: me.upsample_motion(mp_, base))
    out["motion_upsampling"] = {"frames_per_s_kernel": 1.0 / (kt["other_ms"] * 1e-3), "kernel_ms": kt["other_ms"], "call_ms_with_copies": dt * 1e3}
    return out


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # (the reference arm never imports the package itself: its __init__ dlopens the CUDA library)
    synth = light_module("synth") if args.impl == "reference" else importlib.import_module(PKG + ".synth")
    cfg = synth.CONFIGS[args.config]
    per = "GPU" if args.scaling == "weak" else "step in total"
    workload = f"{args.config}: {cfg['name']}, top temporal level of GOP 16, {args.pairs} (cur,ref) pairs/{per}/step, " \
               f"{len(cfg['parts'])} partitions/MB, start=pred field '{args.field}'"
    n_mb = (cfg["width"] // 16) * (cfg["height"] // 16)
    config = {"workload": workload, "jobs_per_step_per_gpu": args.pairs * n_mb * len(cfg["parts"])}      # the same dict in both arms

    if args.impl == "reference":
        # One step of this arm = ONE bounded sample of the workload (the macroblocks of about one pair, spread over all host
        # cores); `value` extrapolates it to pairs/s, `ms_per_step` is the measured wall time of a sample step (so that
        # steps x ms_per_step is what this run really took), `ms_per_workload_step` what a whole 16-pair step would take.
        if rank != 0:
            return
        for _ in range(max(0, min(args.warmup, 1))):
            cpu_baseline(synth, cfg, args.field, args.cpu_sample_mbs, steps=1)
        base, values, walls = cpu_baseline(synth, cfg, args.field, args.cpu_sample_mbs, steps=max(1, args.steps))
        line = {"metric": METRIC, "value": base["value"], "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
                "steps": len(values), "warmup": max(0, min(args.warmup, 1)), "ms_per_step": 1e3 * float(np.mean(walls)),
                "ms_per_workload_step": 1e3 * args.pairs / base["value"] if base["value"] else None,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config, "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    pkg = importlib.import_module(PKG)      # fails loudly if the CUDA library is missing
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream()

    W, H = cfg["width"], cfg["height"]
    sharding = importlib.import_module(PKG + ".sharding")
    n_pairs_global = args.pairs * world if args.scaling == "weak" else args.pairs
    seq, frames, slot_of, mine, jobs, ref_slots = build_workload(synth, sharding, cfg, n_pairs_global, args.field, rank, world)
    n_jobs = len(jobs)
    per_pair = n_mb * len(cfg["parts"])
    n_per_rank = np.array([sharding.pairs_of_rank(n_pairs_global, r, world) * per_pair for r in range(world)], np.int32)
    n_max = int(n_per_rank.max())
    me = pkg.MotionEstimator(local_rank)
    me.configure(W, H, max(1, len(frames)))
    me.set_params(cfg["search_range"], cfg["sub_dfunc"])
    me.set_stream(stream.cuda_stream)

    # pinned host copies (e2e leg) and device-resident inputs (value leg)
    host_planes = {p: torch.from_numpy(f).pin_memory() for p, f in frames.items()}
    for p, t in host_planes.items():
        me.upload_plane_ptr(slot_of[p], t.data_ptr())
    groups, info = me.plan(jobs)
    d_jobs = torch.from_numpy(jobs.view(np.uint8)).to(dev)
    d_groups = torch.from_numpy(groups).to(dev)
    d_results = torch.zeros(max(1, n_jobs) * 32, dtype=torch.uint8, device=dev)
    d_all = torch.zeros(world * n_max * 32, dtype=torch.uint8, device=dev) if (world > 1 and rank == 0) else None
    if world > 1:
        # the gather runs inside the library (NCCL behind the C ABI); torch.distributed only carries the 128-byte unique id
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(me.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, src=0)
        me.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def step_resident():
        me.interp_halfpel(ref_slots)
        me.search_device(d_jobs.data_ptr(), n_jobs, d_groups.data_ptr(), info, d_results.data_ptr())
        if world > 1:
            me.gather_results_device(d_results.data_ptr(), n_jobs, n_per_rank, d_all.data_ptr() if d_all is not None else 0, n_max, 0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    barrier()

    clk = clocks_start() if rank == 0 else None
    launches0 = me.launch_count()
    me.timing_begin()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for e0, e1 in ev:
        flush.fill_(1)                         # L2 flush between timed iterations (outside the event pair)
        e0.record(stream)
        step_resident()
        e1.record(stream)
    barrier()
    total_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    kt = me.timing_end()
    launches = me.launch_count() - launches0
    clocks = clocks_stop(clk, local_rank) if rank == 0 else None
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = n_pairs_global / (ms_per_step * 1e-3)

    legs = {}
    if rank == 0 and args.parity_fraction > 0:
        if world > 1:
            # the gathered field of ALL ranks in global pair order: rank r holds the contiguous block sharding.rank_range( r )
            allr = d_all.cpu().numpy().view(pkg.RESULT_DTYPE).reshape(world, n_max)
            glob = np.concatenate([allr[r, :n_per_rank[r]] for r in range(world)])
            legs["value (jsvm_me_search_device + jsvm_me_gather_results_device, all ranks)"] = (list(range(n_pairs_global)), glob)
        else:
            legs["value (jsvm_me_search_device)"] = ([g for g, _ in mine], d_results.cpu().numpy().view(pkg.RESULT_DTYPE)[:n_jobs].copy())

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        host_jobs = torch.from_numpy(jobs.view(np.uint8).copy()).pin_memory()
        hj = host_jobs.numpy().view(pkg.JOB_DTYPE)
        host_res = torch.empty(n_jobs * 32, dtype=torch.uint8).pin_memory()
        hr = host_res.numpy().view(pkg.RESULT_DTYPE)
        def step_e2e():
            for p, tns in host_planes.items():
                me.upload_plane_ptr(slot_of[p], tns.data_ptr())
            me.interp_halfpel(ref_slots)
            return me.search(hj, out=hr)
        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e()
        barrier()
        n_e2e = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            res = step_e2e()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        h2d = sum(int(tns.numel()) for tns in host_planes.values()) + n_jobs * 32 + len(groups)
        if rank == 0 and args.parity_fraction > 0:
            legs["e2e (jsvm_me_search), rank 0"] = ([g for g, _ in mine], hr.copy())
        e2e = {"value": n_pairs_global * n_e2e / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": n_jobs * 32, "steps": n_e2e, "ms_per_step": dt / n_e2e * 1e3}

    if rank != 0:
        if world > 1:
            dist.barrier()                       # rank 0 still reads the gathered buffer
            me.comm_destroy()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K2) and of K1 ----
    peak, peak_how = int32_peak()
    per_pos = OPS_PER_POS_41 if len(cfg["parts"]) == 41 else OPS_PER_POS_9
    alg_ops_per_launch = info.n_positions / len(cfg["parts"]) * per_pos
    k2_ms = kt["search_ms"] / max(1, kt["search_launches"])
    achieved = alg_ops_per_launch / (k2_ms * 1e-3)
    traffic = ncu_traffic(len(mine)) if args.config == "C4" else {}
    roofline = {"bound": "int32", "kernel": me.last_search_kernel(), "achieved": achieved / 1e12, "peak": peak / 1e12,
                "unit": "Tlane-op/s", "frac": achieved / peak, "traffic": traffic.get("k2_full_search"),
                "alg_ops_per_launch": alg_ops_per_launch, "ms_per_launch": k2_ms, "peak_source": peak_how,
                "positions_evaluated_per_launch": int(info.n_group_positions),
                "share_of_step": kt["search_ms"] / total_ms,
                "note": "achieved = ALGORITHMIC lane-ops (SURVEY 8d) / launch time: every position's SAD is evaluated; the 8x8-granular "
                        "kernels skip the rate / compare code of a batch of displacements when an exact lower bound of every partition "
                        "SAD exceeds the incumbent's cost (tools/prune_stats.py: 81 % of the batches on C4), results unchanged"}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    k1_bytes = len(ref_slots) * (W * H + 4 * (W + 56) * (H + 56))
    k1_ms = kt["interp_ms"] / max(1, kt["interp_launches"])
    roofline_k1 = {"bound": "hbm", "kernel": "halfpel_kernel", "achieved": k1_bytes / (k1_ms * 1e-3) / 1e9, "peak": hbm_peak,
                   "unit": "GB/s", "frac": k1_bytes / (k1_ms * 1e-3) / 1e9 / hbm_peak, "traffic": traffic.get("k1_halfpel"),
                   "alg_bytes_per_launch": k1_bytes, "ms_per_launch": k1_ms,
                   "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                   "share_of_step": kt["interp_ms"] / total_ms}

    # ---- what was timed is what the reference computes: a seeded sample of the step against the CPU checker ----
    parity = None
    if legs:
        parity = parity_check(synth, cfg, seq, args.field, legs, args.parity_fraction / (world if args.scaling == "weak" else 1))

    # K3 (sub-pel refinement): algorithmic INT32 lane-ops per job = 17 candidates x (4x4 blocks of the partition) x OPS_4x4, plus
    # the 8 quarter-pel candidates' byte-wise averages (one lane-op per 4 samples) -- DESIGN.md section 3
    part_px = {1: 256, 2: 128, 3: 128, 8: 64, 9: 32, 10: 32, 11: 16}
    ops_4x4 = OPS_4x4_HADAMARD if cfg["sub_dfunc"] == 2 else OPS_4x4_SAD
    k3_ops_per_mb = sum(17 * (part_px[m] // 16) * ops_4x4 + 8 * (part_px[m] // 4) for m, _ in cfg["parts"])
    k3_ops = k3_ops_per_mb * (n_jobs // len(cfg["parts"]))
    k3_ms = kt["subpel_ms"] / max(1, kt["subpel_launches"])
    roofline_k3 = {"bound": "int32", "kernel": "subpel_kernel", "achieved": k3_ops / (k3_ms * 1e-3) / 1e12, "peak": peak / 1e12,
                   "unit": "Tlane-op/s", "frac": k3_ops / (k3_ms * 1e-3) / peak, "traffic": traffic.get("k3_subpel"),
                   "alg_ops_per_launch": k3_ops, "alg_ops_per_4x4_candidate": ops_4x4, "ms_per_launch": k3_ms, "peak_source": peak_how,
                   "share_of_step": kt["subpel_ms"] / total_ms}

    extras = None
    if world == 1 and not args.no_extras:
        try:
            extras = measure_extras(pkg, synth, me, cfg, torch, dev, hbm_peak, peak)
        except Exception as e:                          # the headline line must not depend on the side measurements
            extras = {"error": repr(e)}

    base = None
    if not args.no_cpu_baseline and world == 1:
        base, _, _ = cpu_baseline(synth, cfg, args.field, args.cpu_sample_mbs)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": config,
            "method": {"l2": "256 MiB L2 flush between timed iterations",
                       "timing": "CUDA events per step on the launching stream, summed; max over ranks"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": roofline, "roofline_k1": roofline_k1, "roofline_k3": roofline_k3, "cpu_baseline": base,
            "parity_checked": parity["jobs"] if parity else 0, "parity": parity,
            "multi_gpu": {"pairs_global": n_pairs_global, "sharding": "sharding.shard_pairs (contiguous blocks of whole pairs: neighbouring pairs share planes)",
                          "collective": "jsvm_me_gather_results_device (NCCL send/recv inside the library)" if world > 1 else None},
            "extras": extras,
            "kernel_ms_per_step": {"k1_halfpel": kt["interp_ms"] / args.steps, "k2_full_search": kt["search_ms"] / args.steps,
                                   "k3_subpel": kt["subpel_ms"] / args.steps}}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        me.comm_destroy()
        dist.destroy_process_group()
    if parity and parity["mismatches"]:
        sys.stderr.write(f"PARITY FAILURE: {parity['mismatches']} result fields differ from the {parity['checker']} checker\n")
        sys.exit(3)


if __name__ == "__main__":
    main()
