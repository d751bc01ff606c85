// k3_subpel.cuh -- K3: half-pel + quarter-pel refinement and the final cost of one job.
//
// Replaces (R = /root/reference/JSVM/H264Extension)
//   MotionEstimationQuarterPel::xSubPelSearch          ENC/MotionEstimationQuarterPel.cpp:94-167
//   MotionEstimationQuarterPel::xCompensateBlocksHalf  ENC/MotionEstimationQuarterPel.cpp:226-263
//   QuarterPelFilter::filterBlock -> xXFilter1..4      COM/QuarterPelFilter.cpp:152-262
//   XDistortion::xGetSAD* / xGetHAD* / xCalcHadamard4x4  ENC/Distortion.cpp:291-340, 397-429, 798-855
//   the tail of estimateBlockWithStart                 ENC/MotionEstimation.cpp:313-337
//   MotionEstimationQuarterPel::compensateBlock        ENC/MotionEstimationQuarterPel.cpp:266-328
//
// One warp per job.  A lane task is (candidate, 4x4 sub-block): the 9 half-pel candidates are
// the half-pel plane sampled at every second sample with an offset of (dx/2, dy/2); the 8
// quarter-pel candidates are (p + q + 1) >> 1 of two half-pel-plane samples chosen by the
// filter class of the half-pel winner.  Candidate order and strict '<' replicate the reference:
// 64-bit keys (cost << 8 | n) are min-reduced, so ties go to the lowest n and the quarter-pel
// stage can only strictly beat the half-pel winner (N1).  ~0.4 % of K2's work.
#pragma once

#include "me_common.cuh"

namespace jsvm {

constexpr int kK3WarpsPerCta = 8;


__constant__ int8_t c_half_mv[9][2]    = { {0,0}, {0,-2}, {0,2}, {-2,0}, {2,0}, {-2,-2}, {2,-2}, {-2,2}, {2,2} };   // g_acHPSearchMv
__constant__ int8_t c_quarter_mv[9][2] = { {0,0}, {0,-1}, {0,1}, {-1,-1}, {1,-1}, {-1,0}, {1,0}, {-1,1}, {1,1} };   // g_acQPSearchMv
__constant__ uint8_t c_filter_of_half[9] = { 0, 2, 2, 1, 1, 3, 3, 3, 3 };                                          // g_aucFilter

// XDistortion::xCalcHadamard4x4 on a 4x4 difference block (per-4x4 halving, N8)
__device__ __forceinline__ uint32_t hadamard4x4(const int (&d)[16]) {
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
    sum += (uint32_t)abs(a + b) + (uint32_t)abs(a - b);
    a = m[4 * n + 1] - m[4 * n + 2]; b = m[4 * n] - m[4 * n + 3];
    sum += (uint32_t)abs(a + b) + (uint32_t)abs(a - b);
  }
  return sum >> 1;
}

// one sample of sub-pel candidate: stage 0 = half-pel candidate n, stage 1 = quarter-pel candidate n
__device__ __forceinline__ int subpel_sample(const uint8_t* base, int hs, int stage, int dx, int dy, bool axis_diag, int px, int py) {
  const uint8_t* o = base + 2 * py * hs + 2 * px;
  if (stage == 0) return o[(dy / 2) * hs + (dx / 2)];
  if (axis_diag && dx != 0 && dy != 0) return (o[dx] + o[dy * hs] + 1) >> 1;
  return (o[0] + o[dy * hs + dx] + 1) >> 1;
}

__global__ void __launch_bounds__(32 * kK3WarpsPerCta)
subpel_kernel(const uint8_t* __restrict__ planes, size_t plane_bytes, int plane_stride,
              const uint8_t* __restrict__ half_planes, size_t half_bytes, int half_stride,
              const jsvm_me_job* __restrict__ jobs, const JobAux* __restrict__ aux, int n_jobs,
              const int16_t* __restrict__ org_mbs,
              const uint32_t* __restrict__ keys, int sub_dfunc,
              jsvm_me_result* __restrict__ results, uint8_t* __restrict__ preds)
{
  const int job_idx = blockIdx.x * kK3WarpsPerCta + (threadIdx.x >> 5);
  if (job_idx >= n_jobs) return;
  const int lane = threadIdx.x & 31;
  const jsvm_me_job j = jobs[job_idx];
  const JobAux a = aux[job_idx];
  int w = 16, h = 16;
  mode_size(j.mode, w, h);
  const int nbx = w >> 2, nb = nbx * (h >> 2);
  const uint32_t f = j.cost_factor;

  // ---- decode the full-pel winner (xPelBlockSearch result) ----
  const uint32_t key = keys[job_idx];
  const uint32_t idx = key & ((1u << kIdxBits) - 1);
  const int fx = a.ux0 + (int)(idx % (uint32_t)a.uw), fy = a.uy0 + (int)(idx / (uint32_t)a.uw);
  const uint32_t full_cost = key >> kIdxBits;
  const uint32_t full_rate = rate_cost(f, mv_component_bits((fx << 2) - j.pred.hor) + mv_component_bits((fy << 2) - j.pred.ver));

  const int bx0 = 16 * j.mb_x + 4 * (j.blk & 3), by0 = 16 * j.mb_y + 4 * (j.blk >> 2);
  const uint8_t* hp0 = half_planes + (size_t)j.ref_slot * half_bytes + (size_t)(kHpOrigin + 2 * by0) * half_stride + kHpOrigin + 2 * bx0;
  const uint8_t* cur0 = planes + (size_t)j.cur_slot * plane_bytes + (size_t)(kMargin + by0) * plane_stride + kMargin + bx0;
  const int16_t* org0 = j.org_index >= 0 ? org_mbs + (size_t)j.org_index * 256 + (4 * (j.blk >> 2)) * 16 + 4 * (j.blk & 3) : nullptr;

  int mvh = fx << 2, mvv = fy << 2;                       // cMv <<= 2
  unsigned long long best = ~0ull;
  uint32_t best_mode = 0;
  const uint8_t* base = hp0 + (mvh >> 1) + (mvv >> 1) * half_stride;
  bool axis_diag = false;

  for (int stage = 0; stage < 2; ++stage) {
    const int n0 = stage, ncand = 9 - stage;
    unsigned long long stage_best = ~0ull;
    for (int t0 = 0; t0 < ncand * nb; t0 += 32) {
      const int t = t0 + lane;
      const bool active = t < ncand * nb;
      const int n = n0 + (active ? t / nb : 0);
      const int b = active ? t % nb : 0;
      const int px0 = (b % nbx) * 4, py0 = (b / nbx) * 4;
      const int dx = stage == 0 ? c_half_mv[n][0] : c_quarter_mv[n][0];
      const int dy = stage == 0 ? c_half_mv[n][1] : c_quarter_mv[n][1];
      uint32_t dist = 0;
      if (active) {
        int d[16];
#pragma unroll
        for (int y = 0; y < 4; ++y)
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const int o = org0 ? (int)org0[(py0 + y) * 16 + px0 + x] : (int)cur0[(py0 + y) * plane_stride + px0 + x];
            d[y * 4 + x] = o - subpel_sample(base, half_stride, stage, dx, dy, axis_diag, px0 + x, py0 + y);
          }
        if (sub_dfunc == JSVM_DF_HADAMARD) dist = hadamard4x4(d);
        else {
#pragma unroll
          for (int i = 0; i < 16; ++i) dist += (uint32_t)abs(d[i]);
        }
      }
      // sum the 4x4 blocks of one candidate (nb is a power of two, groups are lane-aligned)
      for (int off = 1; off < nb; off <<= 1) dist += __shfl_xor_sync(0xffffffffu, dist, off);
      unsigned long long k = ~0ull;
      if (active && b == 0) {
        const int th = mvh + dx, tv = mvv + dy;
        const uint32_t c = dist + rate_cost(f, mv_component_bits(th - j.pred.hor) + mv_component_bits(tv - j.pred.ver));
        k = ((unsigned long long)c << 8) | (unsigned)n;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, k, off);
        k = o < k ? o : k;
      }
      stage_best = k < stage_best ? k : stage_best;
    }
    if (stage == 0) {
      best = stage_best & ~0xffull;                       // incumbent for the quarter-pel stage (n field 0 wins ties)
      best_mode = (uint32_t)(stage_best & 0xff);
      mvh += c_half_mv[best_mode][0]; mvv += c_half_mv[best_mode][1];
      const int filt = c_filter_of_half[best_mode];
      axis_diag = (filt == 0 || filt == 3);
      base = hp0 + (mvh >> 1) + (mvv >> 1) * half_stride;   // ENC/MotionEstimationQuarterPel.cpp:137-139
    } else if (stage_best < best) {
      const uint32_t n = (uint32_t)(stage_best & 0xff);
      best = stage_best;
      best_mode = n << 4;
      mvh += c_quarter_mv[n][0]; mvv += c_quarter_mv[n][1];
    }
  }

  const uint32_t min_sad = (uint32_t)(best >> 8);
  if (lane == 0) {
    // ENC/MotionEstimation.cpp:331-337 with fWeight = 1.0
    const uint32_t mv_bits = mv_component_bits(mvh - j.pred.hor) + mv_component_bits(mvv - j.pred.ver);
    jsvm_me_result r;
    r.mv.hor = (int16_t)mvh; r.mv.ver = (int16_t)mvv;
    r.mv_fullpel.hor = (int16_t)fx; r.mv_fullpel.ver = (int16_t)fy;
    r.sad_fullpel = full_cost - full_rate;
    r.min_sad = min_sad;
    r.mv_bits = mv_bits;
    r.bits = j.bits_in + mv_bits;
    r.cost = (min_sad - rate_cost(f, mv_bits)) + rate_cost(f, j.bits_in + mv_bits);
    r.best_mode = best_mode;
    results[job_idx] = r;
  }

  // ---- compensateBlock: the winner's prediction block (stride 16, zero padded) ----
  if (preds) {
    uint8_t* dp = preds + (size_t)job_idx * 256;
    const int stage = best_mode < 0x10 ? 0 : 1;
    const int n = stage ? (int)(best_mode >> 4) : (int)best_mode;
    const int dx = stage ? c_quarter_mv[n][0] : c_half_mv[n][0];
    const int dy = stage ? c_quarter_mv[n][1] : c_half_mv[n][1];
    // the half-pel buffers were filled around the FULL-PEL mv, the quarter-pel ones around the half-pel winner
    const uint8_t* pbase = stage ? base : hp0 + ((fx << 2) >> 1) + ((fy << 2) >> 1) * half_stride;
    for (int i = lane; i < 256; i += 32) {
      const int x = i & 15, y = i >> 4;
      dp[i] = (x < w && y < h) ? (uint8_t)subpel_sample(pbase, half_stride, stage, dx, dy, axis_diag, x, y) : 0;
    }
  }
}

}  // namespace jsvm
