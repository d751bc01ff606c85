"""ctypes access to the two CPU checkers (TEST INFRASTRUCTURE ONLY -- never imported by the product).

CpuBackend(path, prefix) wraps either oracle/libme_oracle.so (prefix jsvm_ora_, the C restatement)
or oracle/_ref/libjsvm_ref.so (prefix jsvm_ref_, the unmodified reference behind our harness).
Both take the product's job / result records (include/jsvm_me_b200.h).
"""
import ctypes as C
import importlib.util
import os

import numpy as np

# the record layouts only: abi.py is loaded by path so that the checker never runs the product package's __init__ (which dlopens
# the CUDA library) -- bench.py's reference arm must not map any of the product's native code
_spec = importlib.util.spec_from_file_location(
    "jsvm_me_abi_records", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "jsvm-code-notes_b200", "abi.py"))
abi = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(abi)
JOB_DTYPE, RESULT_DTYPE, MC_JOB_DTYPE = abi.JOB_DTYPE, abi.RESULT_DTYPE, abi.MC_JOB_DTYPE
PIC_TASK_DTYPE, PIC_MB_DTYPE = abi.PIC_TASK_DTYPE, abi.PIC_MB_DTYPE


class CpuSession:
    def __init__(self, be, width, height, n_slots, search_range, sub_dfunc, full_dfunc=0, search_mode=0):
        self.be, self.width, self.height = be, width, height
        self.ctx = C.c_void_p(be.create(width, height, n_slots, search_mode, full_dfunc, sub_dfunc, search_range))
        if not self.ctx:
            raise RuntimeError("checker create() failed")

    def close(self):
        if self.ctx:
            self.be.destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        self.close()

    def set_plane(self, slot, luma):
        luma = np.ascontiguousarray(luma, np.uint8)
        assert luma.shape == (self.height, self.width)
        assert self.be.set_plane_u8(self.ctx, slot, luma.ctypes.data, self.width) == 0

    def set_chroma(self, slot, cb, cr):
        cb = np.ascontiguousarray(cb, np.uint8)
        cr = np.ascontiguousarray(cr, np.uint8)
        assert cb.shape == cr.shape == (self.height // 2, self.width // 2)
        assert self.be.set_chroma_u8(self.ctx, slot, cb.ctypes.data, cr.ctypes.data, self.width // 2) == 0

    def set_org_chroma(self, org_chroma_mbs):
        self._org_chroma = None if org_chroma_mbs is None else np.ascontiguousarray(org_chroma_mbs, np.int16).reshape(-1, 2, 8, 8)
        assert self.be.set_org_chroma(self.ctx, None if self._org_chroma is None else self._org_chroma.ctypes.data,
                                      0 if self._org_chroma is None else len(self._org_chroma)) == 0

    def map_cost_factor(self, factor_full_pel, factor_sub_pel):
        assert self.be.map_cost_factor(self.ctx, int(factor_full_pel), int(factor_sub_pel)) == 0

    def set_fast_search_params(self, el_search_range=0, fast_bi_search=0):
        assert self.be.set_fast_search_params(self.ctx, int(el_search_range), int(fast_bi_search)) == 0

    def set_mv_candidates(self, abc):
        """(n_jobs, 3, 2) int16: MVs of the neighbours A, B, C per job (FAST / TZ searches); None clears."""
        self._cands = None if abc is None else np.ascontiguousarray(abc, np.int16).reshape(-1, 3, 2)
        assert self.be.set_mv_candidates(self.ctx, None if self._cands is None else self._cands.ctypes.data,
                                         0 if self._cands is None else len(self._cands)) == 0

    def estimate_pictures(self, tasks):
        tasks = np.ascontiguousarray(tasks, PIC_TASK_DTYPE)
        out = np.zeros((len(tasks), self.height // 16, self.width // 16), PIC_MB_DTYPE)
        assert self.be.estimate_pictures(self.ctx, tasks.ctypes.data, len(tasks), out.ctypes.data) == 0, "checker estimate_pictures() failed"
        return out

    def compensate(self, mc_jobs, chroma=True):
        jobs = np.ascontiguousarray(mc_jobs, MC_JOB_DTYPE)
        n = len(jobs)
        py = np.zeros((n, 16, 16), np.uint8)
        pcb = np.zeros((n, 8, 8), np.uint8) if chroma else None
        pcr = np.zeros((n, 8, 8), np.uint8) if chroma else None
        rc = self.be.compensate(self.ctx, jobs.ctypes.data, n, py.ctypes.data,
                                pcb.ctypes.data if chroma else None, pcr.ctypes.data if chroma else None)
        assert rc == 0, "checker compensate() failed"
        return py, pcb, pcr

    def fullpel(self, slot):
        out = np.zeros((self.height + 64, self.width + 64), np.uint8)
        assert self.be.get_fullpel_u8(self.ctx, slot, out.ctypes.data, out.shape[1]) == 0
        return out

    def halfpel(self, slot):
        out = np.zeros((2 * (self.height + 56), 2 * (self.width + 56)), np.uint8)
        assert self.be.get_halfpel_u8(self.ctx, slot, out.ctypes.data, out.shape[1]) == 0
        return out

    def search(self, jobs, org_mbs=None, want_preds=False):
        jobs = np.ascontiguousarray(jobs, JOB_DTYPE)
        n = len(jobs)
        res = np.zeros(n, RESULT_DTYPE)
        preds = np.zeros((n, 16, 16), np.uint8) if want_preds else None
        n_org, org_ptr = 0, None
        if org_mbs is not None:
            org_mbs = np.ascontiguousarray(org_mbs, np.int16).reshape(-1, 16, 16)
            n_org, org_ptr = len(org_mbs), org_mbs.ctypes.data
        rc = self.be.search(self.ctx, jobs.ctypes.data, n, org_ptr, n_org, res.ctypes.data,
                            preds.ctypes.data if want_preds else None)
        assert rc == 0, "checker search() failed"
        return (res, preds) if want_preds else res


class CpuBackend:
    def __init__(self, path, prefix):
        self.path, self.prefix = path, prefix
        lib = C.CDLL(path)
        self.lib = lib
        protos = [("create", C.c_void_p, [C.c_int] * 7),
                  ("destroy", None, [C.c_void_p]),
                  ("set_plane_u8", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
                  ("get_fullpel_u8", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
                  ("get_halfpel_u8", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
                  ("search", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
                  ("set_chroma_u8", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
                  ("set_org_chroma", C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
                  ("map_cost_factor", C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
                  ("set_fast_search_params", C.c_int, [C.c_void_p, C.c_int, C.c_int]),
                  ("set_mv_candidates", C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
                  ("estimate_pictures", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
                  ("compensate", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
                  ("cost_factor", C.c_uint32, [C.c_double, C.c_int]),
                  ("distortion", C.c_uint32, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int])]
        for name, rt, at in protos:
            fn = getattr(lib, prefix + name)
            fn.restype, fn.argtypes = rt, at
            setattr(self, name, fn)

    def upsample_residual(self, params, base_y, base_cb, base_cr, base_transform8x8):
        by, bcb, bcr = [np.ascontiguousarray(a, np.int16) for a in (base_y, base_cb, base_cr)]
        t8 = np.ascontiguousarray(base_transform8x8, np.uint8)
        oy = np.zeros((params.frame_height, params.frame_width), np.int16)
        ocb = np.zeros((params.frame_height // 2, params.frame_width // 2), np.int16)
        ocr = np.zeros_like(ocb)
        fn = getattr(self.lib, self.prefix + "upsample_residual")
        fn.restype, fn.argtypes = C.c_int, [C.c_void_p] * 8
        assert fn(C.addressof(params), by.ctypes.data, bcb.ctypes.data, bcr.ctypes.data, t8.ctypes.data, oy.ctypes.data, ocb.ctypes.data, ocr.ctypes.data) == 0
        return oy, ocb, ocr

    def upsample_motion(self, params, base):
        base = np.ascontiguousarray(base, abi.MB_MOTION_DTYPE)
        out = np.zeros((params.resize.frame_height // 16, params.resize.frame_width // 16), abi.MB_MOTION_DTYPE)
        fn = getattr(self.lib, self.prefix + "upsample_motion")
        fn.restype, fn.argtypes = C.c_int, [C.c_void_p] * 3
        assert fn(C.addressof(params), base.ctypes.data, out.ctypes.data) == 0
        return out

    def session(self, width, height, n_slots, search_range, sub_dfunc, full_dfunc=0, search_mode=0):
        return CpuSession(self, width, height, n_slots, search_range, sub_dfunc, full_dfunc, search_mode)


# ---- the checker on many host cores: full-frame comparisons (tests) and bench.py's parity sample ----
_POOL_STATE = {}


def _pool_worker(task):
    begin, end = task
    st = _POOL_STATE
    be = CpuBackend(st["path"], st["prefix"])
    ses = CpuSession(be, st["W"], st["H"], st["n_slots"], st["sr"], st["sub"], st["full"], st.get("search_mode", 0))
    for slot, luma in st["planes"].items():
        ses.set_plane(slot, luma)
    out = ses.search(st["jobs"][begin:end], org_mbs=st["org"], want_preds=st["want_preds"])
    ses.close()
    return (begin, end) + (out if st["want_preds"] else (out, None))


def parallel_search(path, prefix, W, H, n_slots, sr, sub, planes, jobs, full=0, org_mbs=None, want_preds=False, n_proc=None, search_mode=0):
    """estimateBlockWithStart of the CPU checker for every job, spread over host cores (fork pool; job list cut at
    macroblock boundaries).  `planes`: {slot: (H, W) uint8}.  Returns results (and prediction blocks)."""
    import multiprocessing as mp
    import os
    jobs = np.ascontiguousarray(jobs, JOB_DTYPE)
    n = len(jobs)
    n_proc = n_proc or min(os.cpu_count() or 1, 64)
    used = sorted({int(s) for s in np.unique(jobs["ref_slot"])} | {int(s) for s in np.unique(jobs["cur_slot"])})
    _POOL_STATE.update(path=path, prefix=prefix, W=W, H=H, n_slots=n_slots, sr=sr, sub=sub, full=full, search_mode=search_mode,
                       planes={s: planes[s] for s in used if s in planes}, jobs=jobs, org=org_mbs, want_preds=want_preds)
    n_tasks = max(1, min(n // 64, n_proc * 6))
    cuts = [0]
    for t in range(1, n_tasks):
        i = max(cuts[-1], n * t // n_tasks)
        while 0 < i < n and (jobs["mb_x"][i] == jobs["mb_x"][i - 1] and jobs["mb_y"][i] == jobs["mb_y"][i - 1]
                             and jobs["ref_slot"][i] == jobs["ref_slot"][i - 1]):
            i += 1
        cuts.append(i)
    cuts.append(n)
    tasks = [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    res = np.zeros(n, RESULT_DTYPE)
    preds = np.zeros((n, 16, 16), np.uint8) if want_preds else None
    if n_proc == 1 or len(tasks) == 1:
        outs = [_pool_worker(t) for t in tasks]
    else:
        with mp.get_context("fork").Pool(min(n_proc, len(tasks))) as pool:
            outs = pool.map(_pool_worker, tasks, chunksize=1)
    for a, b, r, p in outs:
        res[a:b] = r
        if want_preds:
            preds[a:b] = p
    _POOL_STATE.clear()
    return (res, preds) if want_preds else res


def best_checker(root):
    """(path, prefix, kind): the compiled unmodified reference when oracle/_ref was built (it travels to the GPU box), else the C port."""
    import os
    ref = os.path.join(root, "oracle", "_ref", "libjsvm_ref.so")
    if os.path.exists(ref):
        return ref, "jsvm_ref_", "reference"
    return os.path.join(root, "oracle", "libme_oracle.so"), "jsvm_ora_", "port"
