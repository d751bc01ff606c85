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
// One warp per job.
//   * The (2w+3) x (2h+3) patch of the half-pel plane around the full-pel winner is read ONCE with
//     aligned 32-bit loads and de-interleaved into four phase planes in shared memory
//     (phase = parity of the half-pel coordinates: integer / H-half / V-half / diagonal samples), so
//     that the block a candidate predicts is 4 CONTIGUOUS bytes per 4x4 row -- exactly the four
//     buffers xCompensateBlocksHalf gathers, kept for all 17 candidates.
//   * A lane task is (candidate, 4x4 sub-block).  Half-pel candidates read one phase plane; quarter-pel
//     candidates are the byte-wise rounded average (p + q + 1) >> 1 of two phase-plane reads chosen by the
//     filter class of the half-pel winner (xXFilter1/4: axis neighbours for the diagonal positions).
//   * SAD = VABSDIFF4 per row; SATD = the reference's 4x4 Hadamard with the per-4x4 halving (N8).
//   * Candidate order and strict '<' are reproduced with 64-bit keys (cost << 8 | n): ties go to the lowest
//     n, and the quarter-pel stage can only strictly beat the half-pel winner (N1).
#pragma once

#include <type_traits>
#include "me_common.cuh"

namespace jsvm {

#ifndef K3_MIN_CTAS
#define K3_MIN_CTAS 4
#endif
constexpr int kK3WarpsPerCta = 8;
// phase plane: (2*16+3+1)/2 rows; 32 bytes of a row are staged (16-byte aligned patch origin: up to 17 + 34 half-pel columns), and the
// pitch of 40 keeps the rows of the four 4x4 sub-block rows (4 rows = 160 bytes apart) in different shared-memory banks
constexpr int kK3PlaneRows = 18, kK3PlanePitch = 40;

__constant__ int8_t c_half_mv[9][2]    = { {0,0}, {0,-2}, {0,2}, {-2,0}, {2,0}, {-2,-2}, {2,-2}, {-2,2}, {2,2} };   // g_acHPSearchMv
__constant__ int8_t c_quarter_mv[9][2] = { {0,0}, {0,-1}, {0,1}, {-1,-1}, {1,-1}, {-1,0}, {1,0}, {-1,1}, {1,1} };   // g_acQPSearchMv
__constant__ uint8_t c_filter_of_half[9] = { 0, 2, 2, 1, 1, 3, 3, 3, 3 };                                          // g_aucFilter

// XDistortion::xCalcHadamard4x4 (ENC/Distortion.cpp:291-340) of (original - prediction) on one 4x4 block, with the per-4x4
// halving (N8).  The 2-D Hadamard transform is linear and separable in exact integers, so
//   * the HORIZONTAL 4-point transform of a prediction row is four dot products of its packed bytes with (+-1) patterns
//     (IDP.4A, u8 x s8) -- no unpacking -- and the transformed ORIGINAL row, computed once per job, goes in negated as the
//     accumulator: e[row][k] = H_k( prediction row ) - H_k( original row );
//   * the vertical transform of the e columns ends with the butterflies (a + b, a - b) whose absolute values are summed;
//     |a + b| + |a - b| = 2 max(|a|, |b|), so the block sum is even, the reference's "/ 2" is exact, and the SATD is the
//     sum of max(|a|, |b|) over the 8 final pairs: the last butterfly stage and the halving disappear.
// (The reference transforms vertically first; the 16 coefficients are the same set.)
// horizontal transform of one original row, negated: -( o0+o1+o2+o3, o0+o1-o2-o3, o0-o1-o2+o3, o0-o1+o2-o3 )
__device__ __forceinline__ void neg_hrow(int o0, int o1, int o2, int o3, int (&out)[4]) {
  const int s01 = o0 + o1, s23 = o2 + o3, d01 = o0 - o1, d23 = o2 - o3;
  out[0] = -(s01 + s23); out[1] = -(s01 - s23); out[2] = -(d01 - d23); out[3] = -(d01 + d23);
}
__device__ __forceinline__ uint32_t satd4x4(const uint32_t (&crow)[4], const int (&nho)[16]) {
  constexpr uint32_t kPat[4] = { 0x01010101u, 0xffff0101u, 0x01ffff01u, 0xff01ff01u };   // bytes (x0, x1, x2, x3) of the four patterns
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e0 = dp4a_u8s8(crow[0], kPat[k], nho[0 + k]), e1 = dp4a_u8s8(crow[1], kPat[k], nho[4 + k]);
    const int e2 = dp4a_u8s8(crow[2], kPat[k], nho[8 + k]), e3 = dp4a_u8s8(crow[3], kPat[k], nho[12 + k]);
    const int t0 = e0 + e3, t1 = e1 + e2, t2 = e1 - e2, t3 = e0 - e3;
    sum += (uint32_t)max(abs(t0), abs(t1)) + (uint32_t)max(abs(t3), abs(t2));
  }
  return sum;
}

struct K3Warp {
  uint8_t  plane[4][kK3PlaneRows][kK3PlanePitch];   // [phase = (Y&1)*2 + (X&1)][(Y - Y0) >> 1][(X - X0) >> 1]
};

// Where the block predicted at half-pel offset (ox, oy) from the full-pel winner starts inside the phase planes:
// half-pel coordinates X = BX + ox + 2x, Y = BY + oy + 2y; relx = BX - X0 (2..17), Y0 = BY - 2.  Returns the byte
// offset of (block column x, block row y) inside K3Warp::plane; consecutive block rows are kK3PlanePitch bytes apart
// and the pitch is a multiple of 4, so the alignment of a 4-sample read is the same for every row.
__device__ __forceinline__ int cand_offset(int relx, int ox, int oy, int x, int y) {
  const int hx = relx + ox, hy = 2 + oy;                       // >= 0 by construction
  return (((hy & 1) * 2 + (hx & 1)) * kK3PlaneRows + (hy >> 1) + y) * kK3PlanePitch + (hx >> 1) + x;
}
__device__ __forceinline__ uint32_t fetch4(const uint8_t* planes, int off) {
  const uint32_t* q = reinterpret_cast<const uint32_t*>(planes + (off & ~3));
  return __funnelshift_r(q[0], q[1], 8 * (off & 3));
}
// The two phase-plane reads whose rounded average predicts candidate n of a stage, packed as offa | offb << 16 (block column 0,
// row 0).  Stage 0 (half-pel candidates, g_acHPSearchMv) reads one plane: offa == offb.  Stage 1 (quarter-pel candidates around
// the half-pel winner at (cx, cy), g_acQPSearchMv): the filter class of the winner decides -- xXFilter1/4 average the two AXIS
// neighbours for diagonal positions (COM/QuarterPelFilter.cpp:152-262).
template <int STAGE>
__device__ __forceinline__ uint32_t cand_reads(int relx, int n, int cx, int cy, bool axis_diag) {
  int offa, offb;
  if (STAGE == 0) {
    offa = offb = cand_offset(relx, c_half_mv[n][0] / 2, c_half_mv[n][1] / 2, 0, 0);
  } else {
    const int dx = c_quarter_mv[n][0], dy = c_quarter_mv[n][1];
    if (axis_diag && dx != 0 && dy != 0) { offa = cand_offset(relx, cx + dx, cy, 0, 0); offb = cand_offset(relx, cx, cy + dy, 0, 0); }
    else { offa = cand_offset(relx, cx, cy, 0, 0); offb = cand_offset(relx, cx + dx, cy + dy, 0, 0); }
  }
  return (uint32_t)offa | ((uint32_t)offb << 16);
}

template <int DFUNC>     // sub-pel distortion: JSVM_DF_SAD (SearchFuncSubPel 0), JSVM_DF_SSD (1) or JSVM_DF_HADAMARD (2)
__global__ void __launch_bounds__(32 * kK3WarpsPerCta, K3_MIN_CTAS)
subpel_kernel(const uint8_t* __restrict__ planes, size_t plane_bytes, int plane_stride,
              const uint8_t* __restrict__ half_planes, size_t half_bytes, int half_stride,
              const jsvm_me_job* __restrict__ jobs, const JobAux* __restrict__ aux, int n_jobs,
              const int16_t* __restrict__ org_mbs,
              const uint32_t* __restrict__ keys,
              const uint32_t* __restrict__ full_costs,   // jobs whose aux.pad is 0: keys hold (cost << 15 | index) of the packed-byte SAD
                                                         // search; aux.pad 1 (generic and fast searches): keys = index, full_costs = cost
              jsvm_me_result* __restrict__ results, uint8_t* __restrict__ preds,
              const uint2* __restrict__ factor_map, int n_factor_map)    // (full-pel factor, sub-pel factor) pairs; 0: one factor for both stages
{
  constexpr bool HAD = DFUNC == JSVM_DF_HADAMARD, SSD = DFUNC == JSVM_DF_SSD;
  constexpr uint32_t kAll = 0xffffffffu;
  __shared__ __align__(16) K3Warp s_warp[kK3WarpsPerCta];
  const int job_idx = blockIdx.x * kK3WarpsPerCta + (threadIdx.x >> 5);
  if (job_idx >= n_jobs) return;
  K3Warp& sw = s_warp[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const jsvm_me_job j = jobs[job_idx];
  const JobAux a = aux[job_idx];
  // block size (XDistortion::init m_aiRows / m_aiCols, ENC/Distortion.cpp:50-74; me_common.cuh mode_size) as log2 of the number of 4x4
  // columns / rows, two bits per mode value: modes 1, 2, 3 (16x16, 16x8, 8x16) and 8 .. 11 (8x8, 8x4, 4x8, 4x4); validated by the host
  constexpr uint32_t kLgCols = (2u << 2) | (2u << 4) | (1u << 6) | (1u << 16) | (1u << 18) | (0u << 20) | (0u << 22);
  constexpr uint32_t kLgRows = (2u << 2) | (1u << 4) | (2u << 6) | (1u << 16) | (0u << 18) | (1u << 20) | (0u << 22);
  const int mode2 = 2 * (j.mode & JSVM_JOB_MODE_MASK);
  const int lg_nbx = (int)((kLgCols >> mode2) & 3u), lg_nb = lg_nbx + (int)((kLgRows >> mode2) & 3u);
  const int nbx = 1 << lg_nbx, nb = 1 << lg_nb;              // 4x4 sub-blocks per row / per block
  const int w = 4 << lg_nbx, h = 4 << (lg_nb - lg_nbx);
  // The job carries the factor of the full-pel stage; the sub-pel stage and the final cost use the factor xSetMEPars( 0, ... )
  // selects (ENC/MotionEstimation.cpp:325): the same one unless the two stages are of different classes (SAD-type vs SSE),
  // in which case the host registered the pair (jsvm_me_map_cost_factor).  An unmapped factor was rejected before the launch.
  const uint32_t f_full = j.cost_factor;
  uint32_t f = f_full;
  if (n_factor_map > 0) {
    uint32_t hit = kAll;
    for (int i = lane; i < n_factor_map; i += 32)
      if (factor_map[i].x == f_full) hit = (uint32_t)i;
    hit = __reduce_min_sync(kAll, hit);
    if (hit != kAll) f = factor_map[hit].y;
  }

  // ---- decode the full-pel winner (xPelBlockSearch result) ----
  const uint32_t key = keys[job_idx];
  const bool wide = full_costs != nullptr && a.pad != 0;
  const uint32_t idx = wide ? key : key & ((1u << kIdxBits) - 1);
  // idx / uw without the integer-division sequence: idx < 2^15 and uw <= 145 (the planner bounds uw * uh by 2^15; the fast searches
  // write idx 0), so (idx + 0.5) / uw lies at least 0.5 / 145 away from an integer -- four orders of magnitude more than the 2 ulp
  // of the approximate division
  const int qy = (int)__fdividef((float)idx + 0.5f, (float)a.uw);
  const int fx = a.ux0 + (int)idx - qy * a.uw, fy = a.uy0 + qy;
  const uint32_t full_cost = wide ? full_costs[job_idx] : key >> kIdxBits;
  const uint32_t full_rate = rate_cost(f_full, mv_component_bits((fx << 2) - j.pred.hor) + mv_component_bits((fy << 2) - j.pred.ver));

  const int bx0 = 16 * j.mb_x + 4 * (j.blk & 3), by0 = 16 * j.mb_y + 4 * (j.blk >> 2);

  // ---- issue the loads of the half-pel patch: 16 bytes per lane, 4 lanes per row, 8 rows per pass ----
  const int BX = 2 * (bx0 + fx), BY = 2 * (by0 + fy);         // half-pel coordinates of the block at the full-pel winner
  const int X0 = (BX - 2) & ~15, Y0 = BY - 2;                 // patch origin: 16-byte aligned column, even row
  const int relx = BX - X0;                                   // 2..17
  const int nbytes = relx + 2 * w + 2;                        // columns X0 .. BX + 2w + 1 (<= 51): inside the plane row for every legal MV
  const int nrows = 2 * h + 3;
  const int rr = lane >> 2, q = lane & 3;
  uint4 v[5];
  {
    // offsets inside one half-pel plane fit 32 bits (4 (W + 64)(H + 64) bytes): one 64-bit base, 32-bit row steps
    // (the empty asm keeps the compiler from re-deriving the 64-bit address under the predicate of every load)
    const uint8_t* src = half_planes + (size_t)j.ref_slot * half_bytes + ((uint32_t)(kHpOrigin + Y0 + rr) * (uint32_t)half_stride + (uint32_t)(kHpOrigin + X0 + 16 * q));
    asm volatile("" : "+l"(src));
    const uint32_t step = 8u * (uint32_t)half_stride;
    const bool col_ok = 16 * q < nbytes;
#pragma unroll
    for (int it = 0; it < 5; ++it)
      v[it] = (col_ok && rr + 8 * it < nrows) ? __ldg(reinterpret_cast<const uint4*>(src + (uint32_t)it * step)) : make_uint4(0u, 0u, 0u, 0u);
  }

  // each lane always works on the same 4x4 sub-block (32 % nb == 0): keep its original rows in registers
  const int blk_b = lane & (nb - 1);
  const int px0 = (blk_b & (nbx - 1)) * 4, py0 = (blk_b >> lg_nbx) * 4;
  // The lane's 4x4 of the original: packed bytes for the SAD, its (negated) row transforms for the Hadamard.  Originals supplied by the
  // caller (org_mbs) may be signed 16-bit (residual-prediction passes, N3): every prediction sample p lies in [0,255], so
  // |o - p| = |clip(o) - p| + |o - clip(o)|; the SAD runs on the clipped bytes and `excess` (a per-4x4 constant) is added.
  uint32_t orow[4];
  int nho[16];                                               // Hadamard variant: negated horizontal transform of the original rows
  int osq[16];                                               // SSD variant: the original samples themselves
  uint32_t excess = 0, excess_block = 0;
  if (j.org_index >= 0) {
    const int16_t* op = org_mbs + (size_t)j.org_index * 256 + (4 * (j.blk >> 2) + py0) * 16 + 4 * (j.blk & 3) + px0;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const short4 s4 = __ldg(reinterpret_cast<const short4*>(op + y * 16));
      const int o[4] = { s4.x, s4.y, s4.z, s4.w };
      orow[y] = pack_clip8(s4);
#pragma unroll
      for (int x = 0; x < 4; ++x) excess += (uint32_t)abs(o[x] - (int)((orow[y] >> (8 * x)) & 0xff));
      if (HAD) {
        int hr[4];
        neg_hrow(o[0], o[1], o[2], o[3], hr);
#pragma unroll
        for (int k = 0; k < 4; ++k) nho[4 * y + k] = hr[k];
      }
      if (SSD) {
#pragma unroll
        for (int k = 0; k < 4; ++k) osq[4 * y + k] = o[k];
      }
    }
    excess_block = excess;                                   // the whole partition's constant (full-pel SAD)
    for (int off = 1; off < nb; off <<= 1) excess_block += __shfl_xor_sync(kAll, excess_block, off);
  } else {
    const uint8_t* cp = planes + (size_t)j.cur_slot * plane_bytes + (size_t)(kMargin + by0 + py0) * plane_stride + kMargin + bx0 + px0;
    constexpr uint32_t kNegPat[4] = { 0xffffffffu, 0x0101ffffu, 0xff0101ffu, 0x01ff01ffu };   // minus the four patterns of satd4x4
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      orow[y] = __ldg(reinterpret_cast<const uint32_t*>(cp + (size_t)y * plane_stride));
      if (HAD) {
#pragma unroll
        for (int k = 0; k < 4; ++k) nho[4 * y + k] = dp4a_u8s8(orow[y], kNegPat[k], 0);
      }
      if (SSD) {
#pragma unroll
        for (int k = 0; k < 4; ++k) osq[4 * y + k] = (int)((orow[y] >> (8 * k)) & 0xff);
      }
    }
  }

  // ---- de-interleave the patch into the four phase planes: even / odd columns of a row go to neighbouring planes ----
  {
#pragma unroll
    for (int it = 0; it < 5; ++it) {
      const int r = rr + 8 * it;
      if (r < nrows) {                                        // columns past nbytes hold zeros that no candidate reads
        uint8_t* d0 = &sw.plane[(r & 1) * 2][r >> 1][8 * q];
        *reinterpret_cast<uint2*>(d0) = make_uint2(__byte_perm(v[it].x, v[it].y, 0x6420), __byte_perm(v[it].z, v[it].w, 0x6420));
        *reinterpret_cast<uint2*>(d0 + kK3PlaneRows * kK3PlanePitch) = make_uint2(__byte_perm(v[it].x, v[it].y, 0x7531), __byte_perm(v[it].z, v[it].w, 0x7531));
      }
    }
  }
  __syncwarp();

  const uint8_t* pl = &sw.plane[0][0][0];
  const uint32_t blk_off = (uint32_t)(py0 * kK3PlanePitch + px0) * 0x10001u;     // the lane's 4x4 inside the block, for both reads
  int mvh = fx << 2, mvv = fy << 2;                           // cMv <<= 2
  // stage keys (cost << 8) | n: 32 bits for SAD / Hadamard costs (< 2^23), 64 bits for squared differences (up to 2^28)
  typedef typename std::conditional<SSD, unsigned long long, uint32_t>::type Key;
  constexpr Key kNoKey = ~(Key)0;

  // One stage of xSubPelSearch: candidates n0 .. 8 (n0 = STAGE: the quarter-pel stage skips its centre).  Lane i first prepares
  // candidate n0 + i (MV-rate term and the two phase-plane reads); a lane task is (candidate, 4x4 sub-block), 32 tasks per pass.
  // Lanes past the last task repeat candidate n0 on valid addresses and are masked out of the minimum.
  auto run_stage = [&](auto stage_tag, int cx, int cy, bool axis_diag, uint32_t& reads_of_best) -> Key {
    constexpr int STAGE = decltype(stage_tag)::value;
    constexpr int ncand = 9 - STAGE;
    uint32_t my_rate = 0, my_reads = 0;
    {
      const int n = STAGE + (lane < ncand ? lane : 0);
      const int th = mvh + (STAGE == 0 ? c_half_mv[n][0] : c_quarter_mv[n][0]), tv = mvv + (STAGE == 0 ? c_half_mv[n][1] : c_quarter_mv[n][1]);
      my_rate = rate_cost(f, mv_component_bits(th - j.pred.hor) + mv_component_bits(tv - j.pred.ver));
      my_reads = cand_reads<STAGE>(relx, n, cx, cy, axis_diag);
    }
    Key stage_best = kNoKey;
    const int ntask = ncand << lg_nb;
    for (int t0 = 0; t0 < ntask; t0 += 32) {
      const int t = t0 + lane;
      const bool active = t < ntask;
      const int ci = active ? (t >> lg_nb) : 0;
      const uint32_t reads = __shfl_sync(kAll, my_reads, ci) + blk_off;
      const uint32_t rate = __shfl_sync(kAll, my_rate, ci);
      const int offa = (int)(reads & 0xffffu), offb = (int)(reads >> 16);
      uint32_t crow[4];
#pragma unroll
      for (int y = 0; y < 4; ++y) {
        const uint32_t pa = fetch4(pl, offa + y * kK3PlanePitch);
        crow[y] = STAGE == 0 ? pa : __vavgu4(pa, fetch4(pl, offb + y * kK3PlanePitch));   // per byte (p + q + 1) >> 1
      }
      uint32_t dist = 0;
      if (HAD) {
        dist = satd4x4(crow, nho);
      } else if (SSD) {                                       // xGetSSE*: sum of squared differences on the true original
#pragma unroll
        for (int y = 0; y < 4; ++y)
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const int d = osq[4 * y + x] - (int)__byte_perm(crow[y], 0, 0x4440 + x);
            dist += (uint32_t)(d * d);
          }
      } else {
#pragma unroll
        for (int y = 0; y < 4; ++y) dist = __vsadu4(orow[y], crow[y]) + dist;
        dist += excess;
      }
      // sum the 4x4 blocks of one candidate (nb is a power of two, groups are lane-aligned)
#pragma unroll
      for (int off = 1; off < 16; off <<= 1) {
        const uint32_t o = __shfl_xor_sync(kAll, dist, off);
        if (off < nb) dist += o;
      }
      // warp minimum of (cost << 8) | n: lowest cost, then lowest candidate number.  SAD / Hadamard costs stay below 2^23
      // (2^17 in the 8-bit domain, N12; originals in [-1024, 1023] are enforced by jsvm_me_search): one 32-bit REDUX.
      // Squared differences reach 2^28: minimum cost first, then the lowest n among the lanes that hold it.
      const bool valid = active && blk_b == 0;
      const uint32_t c = dist + rate;
      const uint32_t n = (uint32_t)(STAGE + ci);
      Key k;
      if (!SSD) {
        k = (Key)__reduce_min_sync(kAll, valid ? ((c << 8) | n) : kAll);
      } else {
        const uint32_t cmin = __reduce_min_sync(kAll, valid ? c : kAll);
        const uint32_t nmin = __reduce_min_sync(kAll, (valid && c == cmin) ? n : 0xffu);
        k = cmin == kAll ? kNoKey : (Key)(((unsigned long long)cmin << 8) | nmin);
      }
      stage_best = k < stage_best ? k : stage_best;
    }
    reads_of_best = __shfl_sync(kAll, my_reads, (int)(stage_best & 0xff) - STAGE);
    return stage_best;
  };

  // half-pel stage, then the quarter-pel stage around its winner, which can only strictly beat it (N1)
  uint32_t best_reads = 0, quarter_reads = 0;
  const Key half_best = run_stage(std::integral_constant<int, 0>(), 0, 0, false, best_reads);
  Key best = half_best & ~(Key)0xff;                          // incumbent for the quarter-pel stage (n field 0 wins ties)
  uint32_t best_mode = (uint32_t)(half_best & 0xff);
  mvh += c_half_mv[best_mode][0]; mvv += c_half_mv[best_mode][1];
  const int cx = c_half_mv[best_mode][0] / 2, cy = c_half_mv[best_mode][1] / 2;   // ENC/MotionEstimationQuarterPel.cpp:137-139
  const int filt = c_filter_of_half[best_mode];
  const Key quarter_best = run_stage(std::integral_constant<int, 1>(), cx, cy, filt == 0 || filt == 3, quarter_reads);
  if (quarter_best < best) {
    const uint32_t n = (uint32_t)(quarter_best & 0xff);
    best = quarter_best;
    best_mode = n << 4;
    best_reads = quarter_reads;
    mvh += c_quarter_mv[n][0]; mvv += c_quarter_mv[n][1];
  }

  const uint32_t min_sad = (uint32_t)(best >> 8);
  if (lane == 0) {
    // ENC/MotionEstimation.cpp:331-337 with fWeight = 1.0
    const uint32_t mv_bits = mv_component_bits(mvh - j.pred.hor) + mv_component_bits(mvv - j.pred.ver);
    jsvm_me_result r;
    r.mv.hor = (int16_t)mvh; r.mv.ver = (int16_t)mvv;
    r.mv_fullpel.hor = (int16_t)fx; r.mv_fullpel.ver = (int16_t)fy;
    r.sad_fullpel = full_cost - full_rate + (wide ? 0u : excess_block);    // the generic search works on the true original
    r.min_sad = min_sad;
    r.mv_bits = mv_bits;
    r.bits = j.bits_in + mv_bits;
    r.cost = (min_sad - rate_cost(f, mv_bits)) + rate_cost(f, j.bits_in + mv_bits);
    r.best_mode = best_mode;
    *reinterpret_cast<uint4*>(&results[job_idx]) = *reinterpret_cast<const uint4*>(&r);
    *(reinterpret_cast<uint4*>(&results[job_idx]) + 1) = *(reinterpret_cast<const uint4*>(&r) + 1);
  }

  // ---- compensateBlock: the winner's prediction block (stride 16, zero padded); a half-pel winner has offa == offb ----
  if (preds) {
    uint32_t* dp = reinterpret_cast<uint32_t*>(preds + (size_t)job_idx * 256);
    const int offa = (int)(best_reads & 0xffffu), offb = (int)(best_reads >> 16);
#pragma unroll
    for (int i0 = 0; i0 < 64; i0 += 32) {
      const int i = i0 + lane;
      const int y = i >> 2, k = i & 3;
      uint32_t pv = 0;
      if (4 * k < w && y < h)
        pv = __vavgu4(fetch4(pl, offa + y * kK3PlanePitch + 4 * k), fetch4(pl, offb + y * kK3PlanePitch + 4 * k));
      dp[i] = pv;
    }
  }
}

}  // namespace jsvm
