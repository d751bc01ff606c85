// me_common.cuh -- shared definitions of the B200 motion-estimation kernels.
//
// Plane geometry mirrors the reference (R = /root/reference/JSVM/H264Extension):
//   full-pel luma plane : (H+64) x (W+64) bytes, picture origin at (32,32): the 32 margin rows / columns the reference fills and
//                         reads (INCC/CommonDefs.h:246-247 YUV_X_MARGIN 32; YUV_Y_MARGIN 64 is the ALLOCATED vertical margin, of
//                         which frame pictures use m_uiYMargin = 64/2, COM/YuvBufferCtrl.cpp:134-157)
//   half-pel luma plane : 2(H+64) x 2(W+64) bytes, picture origin at (64,64)
//                         (YuvBufferCtrl with uiResolution = 1, ENC/ControlMngH264AVCEncoder.cpp:56-57)
// Samples are stored as bytes on the device: reference planes hold reconstructed pictures
// (0..255) and QuarterPelFilter::filterFrame gClips every output (COM/QuarterPelFilter.cpp:139-140).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/jsvm_me_b200.h"

namespace jsvm {

constexpr int kMargin   = JSVM_ME_MARGIN;   // 32
constexpr int kHpOrigin = 2 * kMargin;      // 64
constexpr int kCMargin  = kMargin / 2;      // chroma plane margin: 16 samples on every side (INCC/YuvPicBuffer.h:68-69)

// Partition slots of one macroblock (41 = 1 + 2 + 2 + 4 + 8 + 8 + 16).
//   0        16x16
//   1..2     16x8   (blk 0, 8)
//   3..4     8x16   (blk 0, 2)
//   5..8     8x8    (blk 0, 2, 8, 10)
//   9..16    8x4    (8x8 index * 2 + {upper, lower})
//   17..24   4x8    (8x8 index * 2 + {left, right})
//   25..40   4x4    (8x8 index * 4 + {0,1,2,3} raster inside the 8x8)
constexpr int kSlots     = 41;
constexpr int kSlotsGran8 = 9;

// One search group = the jobs of one macroblock that are evaluated from one SAD pass
// (planner output; 64 bytes, see jsvm_me_plan in jsvm_me_capi.cu).
struct alignas(16) Group {
  int32_t first_job;       // jobs[first_job .. first_job + n_jobs)
  int16_t n_jobs;
  int16_t gran4;           // 1 -> some job is smaller than 8x8: 4x4 SAD accumulators needed
  int16_t cur_slot, ref_slot;
  int16_t mb_x, mb_y;
  int32_t org_index;
  int16_t ux0, uy0;        // union search window: displacement of its top-left corner (integer pel)
  int16_t uw, uh;          // union window size (positions)
  int8_t  slot_job[kSlots];// job offset (relative to first_job) sitting in each partition slot, -1 = none
  int8_t  pad[3];
};
static_assert(sizeof(Group) == 80, "Group layout");

// How K3 decodes the packed K2 key of a job: origin and width of the group's union window.
struct JobAux { int16_t ux0, uy0, uw, pad; };

// Per-job search window in displacement space, the restatement of
//   Mv::limitComponents + '>>= 2'            (ENC/MotionEstimation.cpp:373-374, INCC/Mv.h:80-85,118-122)
//   SearchRect::init                         (ENC/MotionEstimation.h:44-56)
//   MotionCompensation::initMb               (COM/MotionCompensation.cpp:152-167)
// including the Short casts (N2).  x in [xlo, xhi], y in [ylo, yhi].
struct Window { int xlo, xhi, ylo, yhi; };

__host__ __device__ inline int16_t sat16(int v) { return (int16_t)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

__host__ __device__ inline Window job_window(const jsvm_me_job& j, int mb_w, int mb_h, int default_range) {
  const int16_t min_h = sat16(((-(int)j.mb_x << 4) - 24) << 2), min_v = sat16(((-(int)j.mb_y << 4) - 24) << 2);
  const int16_t max_h = sat16((((mb_w - j.mb_x) << 4) + 8) << 2), max_v = sat16((((mb_h - j.mb_y) << 4) + 8) << 2);
  int h = j.start.hor, v = j.start.ver;
  h = h < min_h ? min_h : h; h = h > max_h ? max_h : h;
  v = v < min_v ? min_v : v; v = v > max_v ? max_v : v;
  h >>= 2; v >>= 2;                                   // arithmetic shift on the (Short) value
  const int mnh = min_h >> 2, mnv = min_v >> 2, mxh = max_h >> 2, mxv = max_v >> 2;
  const int16_t lim = (int16_t)(j.search_range ? j.search_range : default_range);
  const int16_t pos_v = (int16_t)(mxv - v), neg_v = (int16_t)(v - mnv);
  const int16_t pos_h = (int16_t)(mxh - h), neg_h = (int16_t)(h - mnh);
  Window w;
  w.ylo = -((lim < neg_v ? lim : neg_v) - v);
  w.yhi =  (lim < pos_v ? lim : pos_v) + v;
  w.xlo = -((lim < neg_h ? lim : neg_h) - h);
  w.xhi =  (lim < pos_h ? lim : pos_h) + h;
  return w;
}

// signed Exp-Golomb length of a quarter-pel MV difference, MotionEstimationCost::xGetComponentBits
// (ENC/MotionEstimationCost.cpp:75-87): 1 + 2*floor(log2(2|d|+1)).
__host__ __device__ inline uint32_t mv_component_bits(int d) {
  const uint32_t t = ((uint32_t)(d < 0 ? -d : d) << 1) + 1u;
#if defined(__CUDA_ARCH__)
  return 1u + 2u * (31u - (uint32_t)__clz((int)t));
#else
  return 1u + 2u * (31u - (uint32_t)__builtin_clz(t));
#endif
}

// xGetCost( bits ) = ( m_uiCostFactor * bits ) >> 16, unsigned 32-bit wrap (ENC/MotionEstimationCost.h:27, N5)
__host__ __device__ inline uint32_t rate_cost(uint32_t factor, uint32_t bits) { return (factor * bits) >> 16; }

__host__ __device__ inline bool mode_size(int mode, int& w, int& h) {
  switch (mode & JSVM_JOB_MODE_MASK) {          // XDistortion::init m_aiRows / m_aiCols, ENC/Distortion.cpp:50-74
    case JSVM_MODE_16x16: w = 16; h = 16; return true;
    case JSVM_MODE_16x8:  w = 16; h = 8;  return true;
    case JSVM_MODE_8x16:  w = 8;  h = 16; return true;
    case JSVM_BLK_8x8:    w = 8;  h = 8;  return true;
    case JSVM_BLK_8x4:    w = 8;  h = 4;  return true;
    case JSVM_BLK_4x8:    w = 4;  h = 8;  return true;
    case JSVM_BLK_4x4:    w = 4;  h = 4;  return true;
    default: return false;
  }
}

// (mode, blk) -> partition slot, -1 if blk is not a legal origin for the mode.
__host__ __device__ inline int partition_slot(int mode, int blk) {
  const int bx = blk & 3, by = blk >> 2;
  const int b8 = (by >> 1) * 2 + (bx >> 1);           // which 8x8
  switch (mode & JSVM_JOB_MODE_MASK) {
    case JSVM_MODE_16x16: return blk == 0 ? 0 : -1;
    case JSVM_MODE_16x8:  return (bx == 0 && (by == 0 || by == 2)) ? 1 + (by >> 1) : -1;
    case JSVM_MODE_8x16:  return (by == 0 && (bx == 0 || bx == 2)) ? 3 + (bx >> 1) : -1;
    case JSVM_BLK_8x8:    return (!(bx & 1) && !(by & 1)) ? 5 + b8 : -1;
    case JSVM_BLK_8x4:    return !(bx & 1) ? 9 + b8 * 2 + (by & 1) : -1;
    case JSVM_BLK_4x8:    return !(by & 1) ? 17 + b8 * 2 + (bx & 1) : -1;
    case JSVM_BLK_4x4:    return 25 + b8 * 4 + (by & 1) * 2 + (bx & 1);
    default: return -1;
  }
}

#if defined(__CUDACC__)
// c + dot product of four unsigned bytes (a) with four signed bytes (b): IDP.4A.U8.S8
__device__ __forceinline__ int dp4a_u8s8(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// four int16 samples -> four bytes, each clipped to [0,255] (gClip, INCC/GlobalFunctions.h:37-44)
__device__ __forceinline__ uint32_t pack_clip8(const short4 v) {
  const int a = min(max((int)v.x, 0), 255), b = min(max((int)v.y, 0), 255), c = min(max((int)v.z, 0), 255), d = min(max((int)v.w, 0), 255);
  return (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | ((uint32_t)d << 24);
}
#endif

// packed argmin key: cost in the high 17 bits, raster index inside the union window in the low 15
// (N12: SAD16x16 <= 65280 and rate << 2^16 in the 8-bit domain; uw*uh <= 32768 is enforced by the planner).
constexpr int kIdxBits = 15;
constexpr uint32_t kKeyInvalid = 0xffffffffu;
// rate term of a union-window row that lies outside a job's own window (see k2_build_stage): above every valid cost of a
// partition of at most 128 samples, and 65280 + kRowOutCost + rate still fits the 17-bit cost field
constexpr uint32_t kRowOutCost = 40000u;

}  // namespace jsvm
