// k4_mc.cuh -- K4: motion-compensated prediction of luma and chroma blocks.
//
// Replaces (R = /root/reference/JSVM/H264Extension)
//   MotionCompensation::xGetMbPredData (MV clamp)          COM/MotionCompensation.cpp:1363-1405 (limitComponents :1390)
//   MotionCompensation::xPredLuma -> QuarterPelFilter::predBlk and its nine position classes
//                                                          COM/MotionCompensation.cpp:171-204, COM/QuarterPelFilter.cpp:266-619
//   MotionCompensation::xPredChroma / xPredChromaPel       COM/MotionCompensation.cpp:206-290
//   SampleWeighting::weightLumaSamples / weightChromaSamples and xMixB / xMixBWeight / xWeight
//                                                          COM/SampleWeighting.cpp:159-265, 382-461
//
// predBlk computes, from the full-pel plane, exactly the samples of QuarterPelFilter::filterFrame's half-pel plane
//   (dx,dy) = (2,0) / (0,2): gClip( (6-tap + 16) / 32 );   (2,2): gClip( (6-tap of the unrounded 6-taps + 512) / 1024 )
// and rounds quarter positions as the average ( p + q + 1 ) >> 1 of two half-grid samples:
//   one odd coordinate : the two neighbours along that axis (xPredDy0Dx13, xPredDx0Dy13, xPredDx2Dy13, xPredDy2Dx13)
//   both odd           : the two diagonal neighbours that are HALF samples (b/h/m/s of H.264 8.4.2.2.1), i.e. the corners of
//                        the half-grid cell whose coordinate sum is odd (xPredElse)
// so the luma block is gathered from the slot's half-pel plane (K1 output, bit-exact against filterFrame), never re-filtered.
// Chroma is the 1/8-sample bilinear filter; its four special cases in xPredChromaPel are the general formula
//   ( (8-xF)(8-yF) A + xF (8-yF) B + (8-xF) yF C + xF yF D + 32 ) >> 6
// evaluated with fewer terms (same integers).
//
// One CTA of 64 threads per job: thread t owns luma row t / 4, columns 4 (t % 4) .. +3, and chroma sample t of each plane.
#pragma once

#include "me_common.cuh"

namespace jsvm {

constexpr int kK4Threads = 64;

__device__ __forceinline__ int clip255_dev(int v) { return min(max(v, 0), 255); }

// one luma sample at quarter-pel position (qx, qy) (picture coordinates * 4) from the half-pel plane `hp` (origin at (0,0))
__device__ __forceinline__ int luma_qpel(const uint8_t* __restrict__ hp, int half_stride, int qx, int qy) {
  const int hx = qx >> 1, hy = qy >> 1;                      // floor: the half-grid cell
  const uint8_t* p = hp + (ptrdiff_t)hy * half_stride + hx;
  const int ox = qx & 1, oy = qy & 1;
  if (!ox && !oy) return p[0];
  if (ox && !oy) return (p[0] + p[1] + 1) >> 1;
  if (!ox && oy) return (p[0] + p[half_stride] + 1) >> 1;
  // both odd: the two corners of the cell with odd coordinate sum
  if ((hx + hy) & 1) return (p[0] + p[half_stride + 1] + 1) >> 1;
  return (p[1] + p[half_stride] + 1) >> 1;
}

__device__ __forceinline__ int chroma_8th(const uint8_t* __restrict__ cp, int cstride, int cx8, int cy8) {
  const int xF = cx8 & 7, yF = cy8 & 7;
  const uint8_t* p = cp + (ptrdiff_t)(cy8 >> 3) * cstride + (cx8 >> 3);
  const int A = p[0], B = p[1], C = p[cstride], D = p[cstride + 1];
  return ((8 - xF) * (8 - yF) * A + xF * (8 - yF) * B + (8 - xF) * yF * C + xF * yF * D + 32) >> 6;
}

// SampleWeighting: combine the predictions of the two lists (or weight the one that is there)
__device__ __forceinline__ int weight_sample(const jsvm_mc_job& j, int comp /* 0 Y, 1 Cb, 2 Cr */, int a, int b, bool has0, bool has1) {
  const int d = j.log_denom[comp ? 1 : 0];
  if (has0 && has1) {
    if (!(j.weighting & JSVM_MC_WEIGHT_BI)) return (a + b + 1) >> 1;                                       // xMixB
    const int off = (j.offset[0][comp] + j.offset[1][comp] + 1) >> 1;                                  // xMixBWeight
    return clip255_dev(((j.weight[0][comp] * a + j.weight[1][comp] * b + (1 << d)) >> (d + 1)) + off);
  }
  const int l = has0 ? 0 : 1;
  const int v = has0 ? a : b;
  if ((j.weighting & JSVM_MC_WEIGHT_UNI) && j.weight_flag[comp ? 1 : 0]) {                                // xWeight
    const int add = ((1 + j.offset[l][comp] * 2) << d) >> 1;
    return clip255_dev((j.weight[l][comp] * v + add) >> d);
  }
  return v;
}

__global__ void __launch_bounds__(kK4Threads)
compensate_kernel(const uint8_t* __restrict__ half_planes, size_t half_bytes, int half_stride,
                  const uint8_t* __restrict__ chroma_planes, size_t chroma_bytes /* one plane */, int chroma_stride,
                  const jsvm_mc_job* __restrict__ jobs, int n_jobs, int mb_w, int mb_h,
                  uint8_t* __restrict__ pred_y, uint8_t* __restrict__ pred_cb, uint8_t* __restrict__ pred_cr)
{
  const int job_idx = blockIdx.x;
  if (job_idx >= n_jobs) return;
  const jsvm_mc_job j = jobs[job_idx];
  const int t = threadIdx.x;
  const int w = j.size_x, h = j.size_y;
  const int bx0 = 16 * j.mb_x + 4 * (j.blk & 3), by0 = 16 * j.mb_y + 4 * (j.blk >> 2);
  // m_cMin / m_cMax of the macroblock (COM/MotionCompensation.cpp:152-167), Mv::limitComponents
  const int min_h = sat16(((-(int)j.mb_x << 4) - 24) << 2), min_v = sat16(((-(int)j.mb_y << 4) - 24) << 2);
  const int max_h = sat16((((mb_w - j.mb_x) << 4) + 8) << 2), max_v = sat16((((mb_h - j.mb_y) << 4) + 8) << 2);
  int mvx[2], mvy[2];
  bool has[2];
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    has[l] = j.ref_slot[l] >= 0;
    mvx[l] = min(max((int)j.mv[l].hor, min_h), max_h);
    mvy[l] = min(max((int)j.mv[l].ver, min_v), max_v);
  }

  // ---- luma: row t / 4, four samples ----
  {
    const int y = t >> 2, x0 = (t & 3) * 4;
    uint32_t word = 0;
    if (y < h && x0 < w) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int v[2] = {0, 0};
#pragma unroll
        for (int l = 0; l < 2; ++l)
          if (has[l]) {
            const uint8_t* hp = half_planes + (size_t)j.ref_slot[l] * half_bytes + (size_t)kHpOrigin * half_stride + kHpOrigin;
            v[l] = luma_qpel(hp, half_stride, 4 * (bx0 + x0 + k) + mvx[l], 4 * (by0 + y) + mvy[l]);
          }
        word |= (uint32_t)weight_sample(j, 0, v[0], v[1], has[0], has[1]) << (8 * k);
      }
    }
    reinterpret_cast<uint32_t*>(pred_y + (size_t)job_idx * 256)[t] = word;
  }

  // ---- chroma: sample t of the (w/2) x (h/2) block of each plane ----
  if (pred_cb && pred_cr) {
    const int cw = w >> 1, ch = h >> 1;
    const int y = t >> 3, x = t & 7;
    int out[2] = {0, 0};
    if (y < ch && x < cw) {
#pragma unroll
      for (int comp = 0; comp < 2; ++comp) {
        int v[2] = {0, 0};
#pragma unroll
        for (int l = 0; l < 2; ++l)
          if (has[l]) {
            const uint8_t* cp = chroma_planes + ((size_t)j.ref_slot[l] * 2 + comp) * chroma_bytes + (size_t)kCMargin * chroma_stride + kCMargin;
            // xPredChroma: chroma block at ( bx0 / 2, by0 / 2 ), MV in 1/8 chroma samples = the luma quarter-pel MV
            v[l] = chroma_8th(cp, chroma_stride, 8 * (bx0 / 2 + x) + mvx[l], 8 * (by0 / 2 + y) + mvy[l]);
          }
        out[comp] = weight_sample(j, 1 + comp, v[0], v[1], has[0], has[1]);
      }
    }
    pred_cb[(size_t)job_idx * 64 + t] = (uint8_t)out[0];
    pred_cr[(size_t)job_idx * 64 + t] = (uint8_t)out[1];
  }
}

// YuvPicBuffer::xFillPlaneMargin for any plane geometry (chroma: margin 16): one block row per padded row
__global__ void plane_margin_fill_kernel(uint8_t* __restrict__ plane, int stride, int W, int H, int margin)
{
  const int row = blockIdx.y;                                    // 0 .. H + 2 margin - 1
  const int src_row = min(max(row - margin, 0), H - 1) + margin;
  const uint8_t* sr = plane + (size_t)src_row * stride;
  uint8_t* dr = plane + (size_t)row * stride;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < W + 2 * margin; x += gridDim.x * blockDim.x) {
    const bool inside = (row == src_row) && x >= margin && x < margin + W;
    if (!inside) dr[x] = sr[min(max(x - margin, 0), W - 1) + margin];
  }
}

}  // namespace jsvm
