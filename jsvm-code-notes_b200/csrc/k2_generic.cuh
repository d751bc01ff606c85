// k2_generic.cuh -- full search with the distortion functions that the packed-byte kernel does not cover:
// SearchFuncFullPel = 1 (SSD, XDistortion::xGetSSE16x/8x/4x, ENC/Distortion.cpp:512-560, 653-686, 769-794),
// SearchFuncFullPel = 2 (HADAMARD, xGetHAD16x/8x/4x, :798-855) and SearchFuncFullPel = 3 (YUV_SAD, xGetYuvSAD16x/8x/4x,
// :435-505, 593-647, 718-764: luma SAD plus the SADs of the two (w/2) x (h/2) chroma blocks displaced by ( y>>1, x>>1 ),
// ENC/MotionEstimation.cpp:386-387) in MotionEstimation::xPelBlockSearch (ENC/MotionEstimation.cpp:351-412).
//
// Correctness first: one CTA per job, the reference window staged once as bytes in shared memory, one displacement per
// thread and pass, the distortion evaluated sample by sample on the job's TRUE original (signed 16-bit originals included:
// the clipping identity of the SAD kernel does not hold for squares or Hadamard sums), 64-bit keys
// (cost << 32 | raster index) so that costs of any size keep the reference's tie-break (lowest y, then lowest x, N1).
// The SAD search -- the path BASELINE.json measures -- never comes here (full_search_persistent_kernel / full_search_kernel).
#pragma once

#include "me_common.cuh"

namespace jsvm {

constexpr int kK2GThreads = 256;
constexpr int kK2GMaxSide = 129 + 15;       // (2 * 64 + 1) displacements + block size - 1
constexpr int kK2GMaxCSide = 66 + 7;        // chroma: displacements ( x>>1 ) of 129 luma displacements (at most 66) + block size - 1

// XDistortion::xCalcHadamard4x4 on plain differences (ENC/Distortion.cpp:291-340; per-4x4 halving, N8): the last butterfly stage,
// the absolute values and the halving collapse into max(|a|, |b|) (see satd4x4 in k3_subpel.cuh)
__device__ __forceinline__ uint32_t hadamard4x4_diff(const int (&d)[16]) {
  int m[16];
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    const int t0 = d[0 + n] + d[12 + n], t1 = d[4 + n] + d[8 + n];
    const int t2 = d[4 + n] - d[8 + n],  t3 = d[0 + n] - d[12 + n];
    m[0 + n] = t0 + t1; m[8 + n] = t0 - t1; m[4 + n] = t3 + t2; m[12 + n] = t3 - t2;
  }
  uint32_t sum = 0;
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    int a = m[4 * n] + m[4 * n + 3], b = m[4 * n + 1] + m[4 * n + 2];
    sum += (uint32_t)max(abs(a), abs(b));
    a = m[4 * n + 1] - m[4 * n + 2]; b = m[4 * n] - m[4 * n + 3];
    sum += (uint32_t)max(abs(a), abs(b));
  }
  return sum;
}

template <int DFUNC>      // JSVM_DF_SAD (cross-check only), JSVM_DF_SSD, JSVM_DF_HADAMARD, JSVM_DF_YUV_SAD
__global__ void __launch_bounds__(kK2GThreads)
full_search_generic_kernel(const uint8_t* __restrict__ planes, size_t plane_bytes, int plane_stride,
                           const uint8_t* __restrict__ chroma_planes, size_t chroma_bytes, int chroma_stride,   // YUV_SAD only
                           const jsvm_me_job* __restrict__ jobs, int n_jobs,
                           const int16_t* __restrict__ org_mbs,
                           const int16_t* __restrict__ org_chroma_mbs,     // YUV_SAD: Cb 8x8 + Cr 8x8 per prepared macroblock (or NULL)
                           int mb_w, int mb_h, int default_range,
                           uint32_t* __restrict__ keys,            // per job: raster index of the winner inside the job's window
                           uint32_t* __restrict__ full_costs,      // per job: its cost (distortion + rate)
                           JobAux* __restrict__ aux)               // per job: window origin and width
{
  __shared__ uint8_t s_ref[kK2GMaxSide * kK2GMaxSide];
  __shared__ int16_t s_org[16 * 16];
  __shared__ unsigned long long s_best;
  constexpr bool YUV_FN = DFUNC == JSVM_DF_YUV_SAD;
  __shared__ uint8_t s_cref[YUV_FN ? 2 * kK2GMaxCSide * kK2GMaxCSide : 1];
  __shared__ int16_t s_corg[YUV_FN ? 2 * 64 : 1];

  const int job_idx = blockIdx.x;
  if (job_idx >= n_jobs) return;
  const jsvm_me_job j = jobs[job_idx];
  // ENC/MotionEstimation.cpp:250-253: with a bi-predictive or unequally weighted original the YUV-SAD search falls back to the
  // luma SAD; those are the jobs whose original is caller-prepared (org_index >= 0) WITHOUT chroma originals
  // (jsvm_me_set_org_chroma), and every job flagged JSVM_JOB_BIPRED.  Equal-weight originals come with their chroma
  // (weightInverseChromaSamples, :236-241).
  const bool YUV = YUV_FN && !(j.org_index >= 0 && org_chroma_mbs == nullptr) && !(j.mode & JSVM_JOB_BIPRED);
  const int tid = threadIdx.x;
  int w = 16, h = 16;
  mode_size(j.mode, w, h);
  const Window win = job_window(j, mb_w, mb_h, default_range);
  const int ww = win.xhi - win.xlo + 1, wh = win.yhi - win.ylo + 1;     // displacements
  const int cols = ww + w - 1, rows = wh + h - 1;                        // staged reference samples
  const int bx0 = 16 * j.mb_x + 4 * (j.blk & 3), by0 = 16 * j.mb_y + 4 * (j.blk >> 2);

  if (tid == 0) s_best = ~0ull;
  {
    const uint8_t* src = planes + (size_t)j.ref_slot * plane_bytes + (size_t)(kMargin + by0 + win.ylo) * plane_stride + kMargin + bx0 + win.xlo;
    for (int i = tid; i < rows * cols; i += kK2GThreads) {
      const int r = i / cols, c = i - r * cols;
      s_ref[r * kK2GMaxSide + c] = __ldg(src + (size_t)r * plane_stride + c);
    }
    for (int i = tid; i < w * h; i += kK2GThreads) {
      const int y = i / w, x = i - y * w;
      s_org[y * 16 + x] = j.org_index >= 0
          ? __ldg(org_mbs + (size_t)j.org_index * 256 + (4 * (j.blk >> 2) + y) * 16 + 4 * (j.blk & 3) + x)
          : (int16_t)__ldg(planes + (size_t)j.cur_slot * plane_bytes + (size_t)(kMargin + by0 + y) * plane_stride + kMargin + bx0 + x);
    }
    if (YUV) {
      // chroma reference windows: displacement ( x>>1, y>>1 ) of every luma displacement, block at ( bx0 / 2, by0 / 2 )
      const int cxlo = win.xlo >> 1, cylo = win.ylo >> 1;
      const int ccols = (win.xhi >> 1) - cxlo + (w >> 1), crows = (win.yhi >> 1) - cylo + (h >> 1);
      for (int i = tid; i < 2 * crows * ccols; i += kK2GThreads) {
        const int comp = i / (crows * ccols), k = i - comp * crows * ccols, r = k / ccols, c = k - r * ccols;
        const uint8_t* src = chroma_planes + ((size_t)j.ref_slot * 2 + comp) * chroma_bytes +
                             (size_t)(kCMargin + by0 / 2 + cylo + r) * chroma_stride + kCMargin + bx0 / 2 + cxlo + c;
        s_cref[(comp * kK2GMaxCSide + r) * kK2GMaxCSide + c] = __ldg(src);
      }
      for (int i = tid; i < 2 * (w >> 1) * (h >> 1); i += kK2GThreads) {
        const int comp = i / ((w >> 1) * (h >> 1)), k = i - comp * (w >> 1) * (h >> 1), y = k / (w >> 1), x = k - y * (w >> 1);
        const int cy = 2 * (j.blk >> 2) + y, cx = 2 * (j.blk & 3) + x;             // inside the macroblock's 8x8 chroma block
        s_corg[comp * 64 + y * 8 + x] = (j.org_index >= 0 && org_chroma_mbs)
            ? __ldg(org_chroma_mbs + (size_t)j.org_index * 128 + comp * 64 + cy * 8 + cx)
            : (int16_t)__ldg(chroma_planes + ((size_t)j.cur_slot * 2 + comp) * chroma_bytes +
                             (size_t)(kCMargin + 8 * j.mb_y + cy) * chroma_stride + kCMargin + 8 * j.mb_x + cx);
      }
    }
  }
  __syncthreads();

  const uint32_t f = j.cost_factor;
  unsigned long long best = ~0ull;
  for (int p = tid; p < ww * wh; p += kK2GThreads) {
    const int py = p / ww, px = p - py * ww;
    const uint8_t* rp = s_ref + py * kK2GMaxSide + px;
    uint32_t dist = 0;
    if (DFUNC == JSVM_DF_HADAMARD) {
      for (int by = 0; by < h; by += 4)
        for (int bx = 0; bx < w; bx += 4) {
          int d[16];
#pragma unroll
          for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int x = 0; x < 4; ++x) d[y * 4 + x] = (int)s_org[(by + y) * 16 + bx + x] - (int)rp[(by + y) * kK2GMaxSide + bx + x];
          dist += hadamard4x4_diff(d);
        }
    } else {
      for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
          const int d = (int)s_org[y * 16 + x] - (int)rp[y * kK2GMaxSide + x];
          dist += DFUNC == JSVM_DF_SSD ? (uint32_t)(d * d) : (uint32_t)abs(d);
        }
      if (YUV) {
        // pUSearch / pVSearch = chroma block origin + ( y>>1 ) * iCStride + ( x>>1 ) with the arithmetic shift of the
        // displacement (ENC/MotionEstimation.cpp:386-387)
        const int cpx = ((win.xlo + px) >> 1) - (win.xlo >> 1), cpy = ((win.ylo + py) >> 1) - (win.ylo >> 1);
        for (int comp = 0; comp < 2; ++comp) {
          const uint8_t* cp = s_cref + (comp * kK2GMaxCSide + cpy) * kK2GMaxCSide + cpx;
          for (int y = 0; y < (h >> 1); ++y)
            for (int x = 0; x < (w >> 1); ++x) dist += (uint32_t)abs((int)s_corg[comp * 64 + y * 8 + x] - (int)cp[y * kK2GMaxCSide + x]);
        }
      }
    }
    // xGetCost( x, y ) with m_uiMvScaleShift = 2 (N6), 32-bit wrap-around (N5)
    const uint32_t cost = dist + rate_cost(f, mv_component_bits(((win.xlo + px) << 2) - j.pred.hor) + mv_component_bits(((win.ylo + py) << 2) - j.pred.ver));
    const unsigned long long key = ((unsigned long long)cost << 32) | (uint32_t)p;
    best = key < best ? key : best;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, off);
    best = o < best ? o : best;
  }
  if ((tid & 31) == 0) atomicMin(&s_best, best);
  __syncthreads();
  if (tid == 0) {
    keys[job_idx] = (uint32_t)(s_best & 0xffffffffu);
    full_costs[job_idx] = (uint32_t)(s_best >> 32);
    JobAux a; a.ux0 = (int16_t)win.xlo; a.uy0 = (int16_t)win.ylo; a.uw = (int16_t)ww; a.pad = 1;
    aux[job_idx] = a;
  }
}

}  // namespace jsvm
