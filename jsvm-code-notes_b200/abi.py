"""ctypes mirror of include/jsvm_me_b200.h -- the C ABI of the B200 motion-estimation path.

The library (libjsvm_me_b200.so, built in-tree by __graft_entry__.build()) is the product; this
module only declares its records and prototypes.  Loading fails loudly if the library is missing:
there is no Python / CPU fallback for any compute entry point.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JSVM_ME_LIB", os.path.join(_HERE, "libjsvm_me_b200.so"))   # override only for A/B kernel experiments

# INCC/CommonDefs.h:169-190, 208-223
MODE_16x16, MODE_16x8, MODE_8x16 = 1, 2, 3
BLK_8x8, BLK_8x4, BLK_4x8, BLK_4x4 = 8, 9, 10, 11
DF_SAD, DF_SSD, DF_HADAMARD, DF_YUV_SAD = 0, 1, 2, 3
SEARCH_BLOCK, SEARCH_SPIRAL, SEARCH_LOG, SEARCH_FAST, SEARCH_TZ = 0, 1, 2, 3, 4
JOB_MODE_MASK, JOB_EL, JOB_BIPRED = 0x3F, 0x40, 0x80
MARGIN = 32


class Mv(C.Structure):
    _fields_ = [("hor", C.c_int16), ("ver", C.c_int16)]


class Job(C.Structure):
    _fields_ = [("cur_slot", C.c_int16), ("ref_slot", C.c_int16), ("mb_x", C.c_int16), ("mb_y", C.c_int16),
                ("blk", C.c_uint8), ("mode", C.c_uint8), ("search_range", C.c_int16),
                ("start", Mv), ("pred", Mv), ("cost_factor", C.c_uint32), ("bits_in", C.c_uint32),
                ("org_index", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("mv", Mv), ("mv_fullpel", Mv), ("sad_fullpel", C.c_uint32), ("min_sad", C.c_uint32),
                ("mv_bits", C.c_uint32), ("bits", C.c_uint32), ("cost", C.c_uint32), ("best_mode", C.c_uint32)]


class PlanInfo(C.Structure):
    _fields_ = [("n_groups", C.c_int32), ("n_groups_8x8", C.c_int32), ("max_uh_8x8", C.c_int32),
                ("max_uh_4x4", C.c_int32), ("n_positions", C.c_int64), ("n_group_positions", C.c_int64)]


assert C.sizeof(Job) == 32 and C.sizeof(Result) == 32

# numpy views of the same records (structured dtypes are the bulk interface used by bench / tests)
JOB_DTYPE = np.dtype([("cur_slot", "<i2"), ("ref_slot", "<i2"), ("mb_x", "<i2"), ("mb_y", "<i2"),
                      ("blk", "u1"), ("mode", "u1"), ("search_range", "<i2"),
                      ("start_hor", "<i2"), ("start_ver", "<i2"), ("pred_hor", "<i2"), ("pred_ver", "<i2"),
                      ("cost_factor", "<u4"), ("bits_in", "<u4"), ("org_index", "<i4")])
RESULT_DTYPE = np.dtype([("mv_hor", "<i2"), ("mv_ver", "<i2"), ("full_hor", "<i2"), ("full_ver", "<i2"),
                         ("sad_fullpel", "<u4"), ("min_sad", "<u4"), ("mv_bits", "<u4"), ("bits", "<u4"),
                         ("cost", "<u4"), ("best_mode", "<u4")])
assert JOB_DTYPE.itemsize == 32 and RESULT_DTYPE.itemsize == 32

# jsvm_mc_job (motion-compensated prediction of one block, COM/MotionCompensation.cpp:1200-1405)
MC_DEFAULT, MC_WEIGHT_UNI, MC_WEIGHT_BI = 0, 1, 2
MC_JOB_DTYPE = np.dtype([("ref_slot", "<i2", (2,)), ("mb_x", "<i2"), ("mb_y", "<i2"), ("blk", "u1"), ("size_x", "u1"), ("size_y", "u1"),
                         ("weighting", "u1"), ("mv", "<i2", (2, 2)),            # mv[list] = (hor, ver)
                         ("weight", "<i2", (2, 3)), ("offset", "<i2", (2, 3)),  # [list][Y, Cb, Cr]
                         ("log_denom", "u1", (2,)), ("weight_flag", "u1", (2,))])
assert MC_JOB_DTYPE.itemsize == 48

# jsvm_pic_task / jsvm_pic_mb (picture-level motion search)
PIC_MAX_REFS = 8
PIC_TASK_DTYPE = np.dtype([("cur_slot", "<i2"), ("n_refs", "<i2"), ("ref_slot", "<i2", (PIC_MAX_REFS,)), ("cost_factor", "<u4")])
PIC_MB_DTYPE = np.dtype([("mode", "u1"), ("pad", "u1", (3,)), ("ref_idx", "i1", (4,)), ("mv", "<i2", (16, 2)), ("cost", "<u4"),
                         ("part_mv", "<i2", (9, 2)), ("part_ref", "i1", (9,)), ("pad2", "i1", (3,)), ("part_cost", "<u4", (9,))])
assert PIC_TASK_DTYPE.itemsize == 24 and PIC_MB_DTYPE.itemsize == 160


MB_MOTION_DTYPE = np.dtype([("mb_mode", "u1"), ("slice_type", "u1"), ("blk_mode", "u1", (4,)), ("in_crop_window", "u1"), ("res_pred_safe", "u1"),
                            ("fwd_bwd", "<u2"), ("base_intra", "u1"), ("pad", "u1"), ("ref_idx", "i1", (2, 4)), ("mv", "<i2", (2, 16, 2))])
assert MB_MOTION_DTYPE.itemsize == 148


class ResizeParams(C.Structure):
    """jsvm_resize_params: the fields of ResizeParameters the progressive inter-layer resampling reads."""
    _fields_ = [(n, C.c_int32) for n in ("level_idc", "frame_width", "frame_height", "ref_frame_width", "ref_frame_height",
                                         "scaled_ref_width", "scaled_ref_height", "left_offset", "top_offset",
                                         "chroma_phase_x", "chroma_phase_y", "ref_chroma_phase_x", "ref_chroma_phase_y")]


class MotionUpsampleParams(C.Structure):
    """jsvm_motion_upsample_params"""
    _fields_ = [("resize", ResizeParams)] + [(n, C.c_int32) for n in ("slice_type", "ref_layer_dq_id", "direct_8x8_inference",
                                                                      "check_residual_pred", "mv_threshold")]


# every symbol the header declares: (name, restype, argtypes)
PROTOTYPES = [
    ("jsvm_me_create", C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    ("jsvm_me_destroy", C.c_int, [C.c_void_p]),
    ("jsvm_me_last_error", C.c_char_p, []),
    ("jsvm_me_configure", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    ("jsvm_me_set_params", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    ("jsvm_me_set_stream", C.c_int, [C.c_void_p, C.c_void_p]),
    ("jsvm_me_upload_plane_xpel", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    ("jsvm_me_upload_plane_u8", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    ("jsvm_me_set_plane_device_u8", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    ("jsvm_me_set_fast_search_params", C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    ("jsvm_me_set_mv_candidates", C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    ("jsvm_me_set_mv_candidates_device", C.c_int, [C.c_void_p, C.c_void_p]),
    ("jsvm_me_map_cost_factor", C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    ("jsvm_me_upload_chroma_u8", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    ("jsvm_me_upload_chroma_xpel", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    ("jsvm_me_set_org_chroma", C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    ("jsvm_me_compensate", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("jsvm_me_compensate_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("jsvm_me_timing_compensate", C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    ("jsvm_me_interp_halfpel", C.c_int, [C.c_void_p, C.c_int]),
    ("jsvm_me_interp_halfpel_batch", C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    ("jsvm_me_download_halfpel_xpel", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    ("jsvm_me_download_halfpel_u8", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    ("jsvm_me_search", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    ("jsvm_me_group_size", C.c_size_t, []),
    ("jsvm_me_plan", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(PlanInfo)]),
    ("jsvm_me_plan_host", C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(PlanInfo)]),
    ("jsvm_me_search_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(PlanInfo),
                                        C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    ("jsvm_me_estimate_pictures", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    ("jsvm_me_download_pictures", C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    ("jsvm_me_upsample_motion", C.c_int, [C.c_void_p, C.POINTER(MotionUpsampleParams), C.c_void_p, C.c_void_p]),
    ("jsvm_me_upsample_residual", C.c_int, [C.c_void_p, C.POINTER(ResizeParams)] + [C.c_void_p] * 7),
    ("jsvm_me_comm_unique_id", C.c_int, [C.c_void_p]),
    ("jsvm_me_comm_init", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    ("jsvm_me_gather_results_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    ("jsvm_me_comm_destroy", C.c_int, [C.c_void_p]),
    ("jsvm_me_fetch_pred", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    ("jsvm_me_sync", C.c_int, [C.c_void_p]),
    ("jsvm_me_launch_count", C.c_uint64, [C.c_void_p]),
    ("jsvm_me_last_search_kernel", C.c_char_p, [C.c_void_p]),
    ("jsvm_me_timing_begin", C.c_int, [C.c_void_p]),
    ("jsvm_me_timing_end", C.c_int, [C.c_void_p] + [C.POINTER(C.c_double)] * 3 + [C.POINTER(C.c_uint64)] * 3),
]

_lib = None


def load_library():
    """dlopen the in-tree CUDA library and bind every prototype.  No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "The motion-estimation path has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in PROTOTYPES:
        fn = getattr(lib, name)          # AttributeError here = the library does not export the header's symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib
