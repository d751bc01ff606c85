"""Host-side Python mirror of the reference's motion-search interface over the C ABI.

Two layers:

* ``MotionEstimator`` -- thin object wrapper over ``jsvm_me_*`` (bulk / batched use: bench, tests).
* ``QuarterPelFilter`` / ``MotionEstimationQuarterPel`` -- the names, argument meaning and call order
  of the reference classes (ENC/MotionEstimation.h:84-115, ENC/MotionEstimationQuarterPel.h,
  INCC/QuarterPelFilter.h:24-35) so parity tests read like reference call sites:
  ``initMb`` -> ``estimateBlockWithStart`` -> ``compensateBlock``.  Errors raise
  ``MotionEstimationError`` where the reference returns ``Err::m_nERR``.

The C++ drop-in that the unmodified H264AVCEncoderLib links against lives in ``dropin/`` and calls
the same C ABI; nothing here or there computes motion search on the CPU.
"""
import ctypes as C

import numpy as np

from . import abi
from .abi import JOB_DTYPE, RESULT_DTYPE, MC_JOB_DTYPE, PIC_TASK_DTYPE, PIC_MB_DTYPE, PlanInfo


class MotionEstimationError(RuntimeError):
    """A jsvm_me_* call returned Err::m_nERR."""


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def plan_host(width, height, n_slots, search_range, jobs):
    """jsvm_me_plan_host: group jobs into search groups without a device (host logic only)."""
    lib = abi.load_library()
    jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
    gsz = lib.jsvm_me_group_size()
    groups = np.empty(max(1, len(jobs)) * gsz, np.uint8)
    info = PlanInfo()
    if lib.jsvm_me_plan_host(width, height, n_slots, search_range, _ptr(jobs), len(jobs), _ptr(groups), C.byref(info)) != 0:
        raise MotionEstimationError(lib.jsvm_me_last_error().decode())
    return groups[: info.n_groups * gsz].copy(), info


# numpy view of the 80-byte search-group record (csrc/me_common.cuh: struct Group)
GROUP_DTYPE = np.dtype([("first_job", "<i4"), ("n_jobs", "<i2"), ("gran4", "<i2"), ("cur_slot", "<i2"), ("ref_slot", "<i2"),
                        ("mb_x", "<i2"), ("mb_y", "<i2"), ("org_index", "<i4"), ("ux0", "<i2"), ("uy0", "<i2"),
                        ("uw", "<i2"), ("uh", "<i2"), ("slot_job", "i1", (41,)), ("pad", "i1", (11,))])
assert GROUP_DTYPE.itemsize == 80


class MotionEstimator:
    """One device context = one picture resolution (one spatial layer) on one GPU."""

    def __init__(self, device=0):
        self._lib = abi.load_library()
        self._ctx = C.c_void_p()
        self._check(self._lib.jsvm_me_create(C.byref(self._ctx), int(device)))
        self.width = self.height = self.n_slots = 0

    def _check(self, rc):
        if rc != 0:
            raise MotionEstimationError(self._lib.jsvm_me_last_error().decode())

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self._lib.jsvm_me_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration -------------------------------------------------------------------
    def configure(self, width, height, n_slots):
        self._check(self._lib.jsvm_me_configure(self._ctx, width, height, n_slots))
        self.width, self.height, self.n_slots = width, height, n_slots

    def set_params(self, search_range, sub_pel_dfunc=abi.DF_SAD, full_pel_dfunc=abi.DF_SAD, search_mode=abi.SEARCH_BLOCK):
        self._check(self._lib.jsvm_me_set_params(self._ctx, search_mode, full_pel_dfunc, sub_pel_dfunc, search_range))

    def set_stream(self, cuda_stream):
        self._check(self._lib.jsvm_me_set_stream(self._ctx, C.c_void_p(int(cuda_stream))))

    # ---- planes --------------------------------------------------------------------------
    def upload_plane(self, slot, luma):
        """luma: (H, W) uint8 ndarray (no margins) -- or a raw host pointer via upload_plane_ptr."""
        luma = np.ascontiguousarray(luma, dtype=np.uint8)
        if luma.shape != (self.height, self.width):
            raise MotionEstimationError(f"plane shape {luma.shape} != {(self.height, self.width)}")
        self._check(self._lib.jsvm_me_upload_plane_u8(self._ctx, slot, _ptr(luma), self.width, self.height, self.width))

    def upload_plane_ptr(self, slot, host_ptr, stride=None):
        self._check(self._lib.jsvm_me_upload_plane_u8(self._ctx, slot, C.c_void_p(int(host_ptr)), self.width, self.height, stride or self.width))

    def set_plane_device(self, slot, dev_ptr, stride=None):
        self._check(self._lib.jsvm_me_set_plane_device_u8(self._ctx, slot, C.c_void_p(int(dev_ptr)), self.width, self.height, stride or self.width))

    def upload_plane_xpel(self, slot, plane_with_margin, v_margin=None):
        """plane_with_margin: (H + 2*v_margin, W + 64) int16 ndarray laid out like a YuvPicBuffer luma plane.  v_margin is the
        number of rows above the picture: 64 for a real YuvPicBuffer dump (YUV_Y_MARGIN rows are allocated), 32 for the rows the
        reference actually fills; inferred from the shape when omitted.  Only the inner 32 rows are read."""
        p = np.ascontiguousarray(plane_with_margin, dtype=np.int16)
        s = self.width + 2 * abi.MARGIN
        vm = (p.shape[0] - self.height) // 2 if v_margin is None else int(v_margin)
        if p.ndim != 2 or p.shape[1] != s or vm < abi.MARGIN or p.shape[0] != self.height + 2 * vm:
            raise MotionEstimationError(f"plane shape {p.shape}: expected ({self.height} + 2*v_margin, {s}) with v_margin >= {abi.MARGIN}")
        origin = p.ctypes.data + 2 * (vm * s + abi.MARGIN)
        self._check(self._lib.jsvm_me_upload_plane_xpel(self._ctx, slot, C.c_void_p(origin), self.width, self.height, s))

    def upload_chroma(self, slot, cb, cr):
        """cb, cr: (H/2, W/2) uint8 ndarrays (no margins): the slot's chroma planes (YUV_SAD search, compensate)."""
        cb = np.ascontiguousarray(cb, dtype=np.uint8)
        cr = np.ascontiguousarray(cr, dtype=np.uint8)
        shape = (self.height // 2, self.width // 2)
        if cb.shape != shape or cr.shape != shape:
            raise MotionEstimationError(f"chroma plane shapes {cb.shape} / {cr.shape} != {shape}")
        self._check(self._lib.jsvm_me_upload_chroma_u8(self._ctx, slot, _ptr(cb), _ptr(cr), shape[1]))

    def upload_chroma_xpel(self, slot, cb_with_margin, cr_with_margin):
        """(H/2 + 32, W/2 + 32) int16 planes laid out like YuvPicBuffer chroma planes with their 16-sample margins filled."""
        planes = [np.ascontiguousarray(p, dtype=np.int16) for p in (cb_with_margin, cr_with_margin)]
        s = self.width // 2 + abi.MARGIN
        for p in planes:
            if p.shape != (self.height // 2 + abi.MARGIN, s):
                raise MotionEstimationError(f"chroma plane shape {p.shape}: expected {(self.height // 2 + abi.MARGIN, s)}")
        org = [p.ctypes.data + 2 * ((abi.MARGIN // 2) * s + abi.MARGIN // 2) for p in planes]
        self._check(self._lib.jsvm_me_upload_chroma_xpel(self._ctx, slot, C.c_void_p(org[0]), C.c_void_p(org[1]), s))

    def set_org_chroma(self, org_chroma_mbs):
        """(n, 2, 8, 8) int16 chroma originals of the caller-prepared macroblocks (or None); kept alive until replaced."""
        if org_chroma_mbs is None:
            self._org_chroma = None
            self._check(self._lib.jsvm_me_set_org_chroma(self._ctx, None, 0))
            return
        self._org_chroma = np.ascontiguousarray(org_chroma_mbs, dtype=np.int16).reshape(-1, 2, 8, 8)
        self._check(self._lib.jsvm_me_set_org_chroma(self._ctx, _ptr(self._org_chroma), len(self._org_chroma)))

    def map_cost_factor(self, factor_full_pel, factor_sub_pel):
        self._check(self._lib.jsvm_me_map_cost_factor(self._ctx, int(factor_full_pel), int(factor_sub_pel)))

    def set_fast_search_params(self, el_search_range=0, fast_bi_search=0):
        self._check(self._lib.jsvm_me_set_fast_search_params(self._ctx, int(el_search_range), int(fast_bi_search)))

    def set_mv_candidates(self, abc):
        """(n_jobs, 3, 2) int16 MVs of the neighbours A, B, C per job for FAST / TZ searches (or None); kept alive until replaced."""
        if abc is None:
            self._cands = None
            self._check(self._lib.jsvm_me_set_mv_candidates(self._ctx, None, 0))
            return
        self._cands = np.ascontiguousarray(abc, dtype=np.int16).reshape(-1, 3, 2)
        self._check(self._lib.jsvm_me_set_mv_candidates(self._ctx, _ptr(self._cands), len(self._cands)))

    def compensate(self, mc_jobs, chroma=True):
        """MotionCompensation (predBlk / xPredChromaPel / SampleWeighting) for every MC_JOB_DTYPE record.
        Returns (pred_y (n,16,16), pred_cb (n,8,8), pred_cr (n,8,8)) uint8 (chroma None when chroma=False)."""
        jobs = np.ascontiguousarray(mc_jobs, dtype=MC_JOB_DTYPE)
        n = len(jobs)
        py = np.zeros((n, 16, 16), np.uint8)
        pcb = np.zeros((n, 8, 8), np.uint8) if chroma else None
        pcr = np.zeros((n, 8, 8), np.uint8) if chroma else None
        self._check(self._lib.jsvm_me_compensate(self._ctx, _ptr(jobs), n, _ptr(py), _ptr(pcb), _ptr(pcr)))
        return py, pcb, pcr

    def interp_halfpel(self, slots):
        """QuarterPelFilter::filterFrame for one slot or a list of slots (one launch)."""
        slots = np.atleast_1d(np.asarray(slots, dtype=np.int32))
        self._check(self._lib.jsvm_me_interp_halfpel_batch(self._ctx, _ptr(slots), len(slots)))

    def halfpel_plane(self, slot):
        """The region filterFrame writes, as (2(H+56), 2(W+56)) uint8."""
        out = np.empty((2 * (self.height + 56), 2 * (self.width + 56)), np.uint8)
        self._check(self._lib.jsvm_me_download_halfpel_u8(self._ctx, slot, _ptr(out), out.shape[1]))
        return out

    def halfpel_plane_xpel(self, slot, v_margin=abi.MARGIN):
        """Half-pel YuvPicBuffer layout: (2(H + 2*v_margin), 2(W+64)) int16 with origin at (2*v_margin, 64); v_margin = 64 gives
        the allocation of a real half-pel YuvPicBuffer (see upload_plane_xpel)."""
        if v_margin < abi.MARGIN:
            raise MotionEstimationError(f"v_margin {v_margin} < {abi.MARGIN}")
        s = 2 * (self.width + 2 * abi.MARGIN)
        out = np.zeros((2 * (self.height + 2 * v_margin), s), np.int16)
        origin = out.ctypes.data + 2 * (2 * v_margin * s + 2 * abi.MARGIN)
        self._check(self._lib.jsvm_me_download_halfpel_xpel(self._ctx, slot, C.c_void_p(origin), s))
        return out

    # ---- search --------------------------------------------------------------------------
    def search(self, jobs, org_mbs=None, want_preds=False, out=None):
        """estimateBlockWithStart for every record of `jobs` (JOB_DTYPE ndarray).
        `out`: optional preallocated RESULT_DTYPE array (e.g. a view of pinned memory) that receives the results."""
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        n = len(jobs)
        res = np.empty(n, RESULT_DTYPE) if out is None else out
        if res.dtype != RESULT_DTYPE or len(res) != n or not res.flags.c_contiguous:
            raise MotionEstimationError("out must be a contiguous RESULT_DTYPE array with one record per job")
        preds = np.zeros((n, 16, 16), np.uint8) if want_preds else None
        n_org = 0
        if org_mbs is not None:
            org_mbs = np.ascontiguousarray(org_mbs, dtype=np.int16).reshape(-1, 16, 16)
            n_org = len(org_mbs)
        self._check(self._lib.jsvm_me_search(self._ctx, _ptr(jobs), n, _ptr(org_mbs), n_org, _ptr(res), _ptr(preds)))
        return (res, preds) if want_preds else res

    def estimate_pictures(self, tasks, download=True):
        """Picture-level motion search (MV prediction and macroblock scheduling on the device) for every PIC_TASK_DTYPE record.
        Returns (n_pics, H/16, W/16) PIC_MB_DTYPE records, or None with download=False (results stay on the device)."""
        tasks = np.ascontiguousarray(tasks, dtype=PIC_TASK_DTYPE)
        out = np.zeros((len(tasks), self.height // 16, self.width // 16), PIC_MB_DTYPE) if download else None
        self._check(self._lib.jsvm_me_estimate_pictures(self._ctx, _ptr(tasks), len(tasks), _ptr(out)))
        return out

    def download_pictures(self, n_pics):
        out = np.zeros((n_pics, self.height // 16, self.width // 16), PIC_MB_DTYPE)
        self._check(self._lib.jsvm_me_download_pictures(self._ctx, _ptr(out), n_pics))
        return out

    def upsample_residual(self, params, base_y, base_cb=None, base_cr=None, base_transform8x8=None):
        """Frame::residualUpsampling for progressive layers.  params: abi.ResizeParams; base planes int16 (ref_h, ref_w) and
        (ref_h/2, ref_w/2); base_transform8x8: (ref_h/16, ref_w/16) flags.  Returns (y, cb, cr) int16 at the enhancement size."""
        by = np.ascontiguousarray(base_y, np.int16)
        chroma = base_cb is not None
        bcb = np.ascontiguousarray(base_cb, np.int16) if chroma else None
        bcr = np.ascontiguousarray(base_cr, np.int16) if chroma else None
        t8 = np.ascontiguousarray(base_transform8x8 if base_transform8x8 is not None else np.zeros((by.shape[0] // 16, by.shape[1] // 16)), np.uint8)
        oy = np.zeros((params.frame_height, params.frame_width), np.int16)
        ocb = np.zeros((params.frame_height // 2, params.frame_width // 2), np.int16) if chroma else None
        ocr = np.zeros((params.frame_height // 2, params.frame_width // 2), np.int16) if chroma else None
        self._check(self._lib.jsvm_me_upsample_residual(self._ctx, C.byref(params), _ptr(by), _ptr(bcb), _ptr(bcr), _ptr(t8), _ptr(oy), _ptr(ocb), _ptr(ocr)))
        return oy, ocb, ocr

    def upsample_motion(self, params, base):
        """MbDataCtrl::upsampleMotion for progressive layers.  params: abi.MotionUpsampleParams; base: (ref_h/16, ref_w/16)
        MB_MOTION_DTYPE records.  Returns the (frame_h/16, frame_w/16) prediction records."""
        base = np.ascontiguousarray(base, dtype=abi.MB_MOTION_DTYPE)
        out = np.zeros((params.resize.frame_height // 16, params.resize.frame_width // 16), abi.MB_MOTION_DTYPE)
        self._check(self._lib.jsvm_me_upsample_motion(self._ctx, C.byref(params), _ptr(base), _ptr(out)))
        return out

    # ---- multi-GPU gather (NCCL inside the library) --------------------------------------------
    def comm_unique_id(self):
        buf = (C.c_uint8 * 128)()
        self._check(self._lib.jsvm_me_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, rank, world, unique_id):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self._lib.jsvm_me_comm_init(self._ctx, rank, world, buf))

    def gather_results_device(self, d_local, n_local, n_per_rank, d_all, n_max, root=0):
        npr = np.ascontiguousarray(n_per_rank, dtype=np.int32)
        self._check(self._lib.jsvm_me_gather_results_device(self._ctx, C.c_void_p(int(d_local)), n_local, _ptr(npr),
                                                           C.c_void_p(int(d_all)) if d_all else None, n_max, root))

    def comm_destroy(self):
        self._check(self._lib.jsvm_me_comm_destroy(self._ctx))

    def plan(self, jobs):
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        gsz = self._lib.jsvm_me_group_size()
        groups = np.empty(len(jobs) * gsz, np.uint8)
        info = PlanInfo()
        self._check(self._lib.jsvm_me_plan(self._ctx, _ptr(jobs), len(jobs), _ptr(groups), C.byref(info)))
        return groups[: info.n_groups * gsz].copy(), info

    def search_device(self, d_jobs, n_jobs, d_groups, info, d_results, d_org_mbs=0, n_org_mbs=0, d_preds=0):
        """Asynchronous search on device-resident buffers (raw device pointers as ints)."""
        self._check(self._lib.jsvm_me_search_device(
            self._ctx, C.c_void_p(int(d_jobs)), n_jobs, C.c_void_p(int(d_groups)), C.byref(info),
            C.c_void_p(int(d_org_mbs)) if d_org_mbs else None, n_org_mbs,
            C.c_void_p(int(d_results)), C.c_void_p(int(d_preds)) if d_preds else None))

    def fetch_pred(self, job_idx, w, h):
        out = np.zeros((h, w), np.int16)
        self._check(self._lib.jsvm_me_fetch_pred(self._ctx, job_idx, _ptr(out), w))
        return out

    def sync(self):
        self._check(self._lib.jsvm_me_sync(self._ctx))

    # ---- measurement ---------------------------------------------------------------------
    def launch_count(self):
        return int(self._lib.jsvm_me_launch_count(self._ctx))

    def last_search_kernel(self):
        return self._lib.jsvm_me_last_search_kernel(self._ctx).decode()

    def timing_begin(self):
        self._check(self._lib.jsvm_me_timing_begin(self._ctx))

    def timing_end(self):
        ms = [C.c_double() for _ in range(3)]
        n = [C.c_uint64() for _ in range(3)]
        self._check(self._lib.jsvm_me_timing_end(self._ctx, *[C.byref(x) for x in ms], *[C.byref(x) for x in n]))
        mc_ms, mc_n = C.c_double(), C.c_uint64()
        self._check(self._lib.jsvm_me_timing_compensate(self._ctx, C.byref(mc_ms), C.byref(mc_n)))
        return {"interp_ms": ms[0].value, "search_ms": ms[1].value, "subpel_ms": ms[2].value,
                "interp_launches": n[0].value, "search_launches": n[1].value, "subpel_launches": n[2].value,
                "other_ms": mc_ms.value, "other_launches": mc_n.value}      # K4 (compensate) / K7 (inter-layer resampling)


# ---------------------------------------------------------------------------------------------
# Reference-named mirror
# ---------------------------------------------------------------------------------------------
BLOCK_SIZE = {abi.MODE_16x16: (16, 16), abi.MODE_16x8: (16, 8), abi.MODE_8x16: (8, 16),
              abi.BLK_8x8: (8, 8), abi.BLK_8x4: (8, 4), abi.BLK_4x8: (4, 8), abi.BLK_4x4: (4, 4)}  # (w, h), ENC/Distortion.cpp:50-74


def cost_factor_sad(lam):
    """RateDistortion::setMbQpLambda: m_uiCostFactorMotionSAD = floor(65536 * sqrt(lambda)) (ENC/RateDistortion.cpp:59-70).
    IEEE double on the host, exactly as the reference computes it (N11); never recomputed on the device."""
    import math
    return int(math.floor(65536.0 * math.sqrt(float(lam)))) & 0xFFFFFFFF


class Frame:
    """A frame-store slot: the stand-in for the reference's Frame (full-pel + half-pel YuvPicBuffer)."""

    def __init__(self, slot):
        self.slot = slot


class QuarterPelFilter:
    """INCC/QuarterPelFilter.h:24-35 -- only filterFrame is on the hot path."""

    def __init__(self, estimator):
        self._me = estimator

    def filterFrame(self, frame):
        self._me.interp_halfpel([frame.slot])


class MotionEstimationQuarterPel:
    """ENC/MotionEstimationQuarterPel.h + ENC/MotionEstimation.h:84-115, one call at a time."""

    def __init__(self, estimator):
        self._me = estimator
        self._mb = None
        self._cost_factor = 0
        self._org_frame = None
        self._last = None

    @staticmethod
    def create(estimator):
        return MotionEstimationQuarterPel(estimator)

    def init(self, search_range, sub_pel_dfunc=abi.DF_SAD, full_pel_dfunc=abi.DF_SAD, search_mode=abi.SEARCH_BLOCK):
        self._me.set_params(search_range, sub_pel_dfunc, full_pel_dfunc, search_mode)

    def setMbQpLambda(self, lam):
        self._cost_factor = cost_factor_sad(lam)

    def loadOrgMbPelData(self, org_frame):
        """XDistortion::loadOrgMbPelData: the picture the original macroblock is taken from."""
        self._org_frame = org_frame

    def initMb(self, uiMbPosY, uiMbPosX):
        self._mb = (uiMbPosY, uiMbPosX)

    def estimateBlockWithStart(self, rcRefFrame, rcMv, rcMvPred, ruiBits, uiBlk, uiMode, uiSearchRange=0, org_mb=None):
        """Returns (mv, bits, cost) = (rcMv, ruiBits, ruiCost) on exit."""
        if self._mb is None:
            raise MotionEstimationError("initMb has not been called")
        job = np.zeros(1, JOB_DTYPE)
        j = job[0]
        j["cur_slot"] = self._org_frame.slot if self._org_frame is not None else rcRefFrame.slot
        j["ref_slot"] = rcRefFrame.slot
        j["mb_y"], j["mb_x"] = self._mb
        j["blk"], j["mode"], j["search_range"] = uiBlk, uiMode, uiSearchRange
        j["start_hor"], j["start_ver"] = rcMv
        j["pred_hor"], j["pred_ver"] = rcMvPred
        j["cost_factor"], j["bits_in"] = self._cost_factor, ruiBits
        j["org_index"] = -1 if org_mb is None else 0
        res = self._me.search(job, org_mbs=org_mb)
        r = res[0]
        self._last = (uiBlk, uiMode)
        return (int(r["mv_hor"]), int(r["mv_ver"])), int(r["bits"]), int(r["cost"])

    def compensateBlock(self, uiBlk, uiMode):
        if self._last != (uiBlk, uiMode):
            raise MotionEstimationError("compensateBlock must follow the matching estimateBlockWithStart")
        w, h = BLOCK_SIZE[uiMode]
        return self._me.fetch_pred(0, w, h)
