// k5_fastsearch.cuh -- K5: the data-dependent full-pel searches, one warp per job.
//
// Replaces (R = /root/reference/JSVM/H264Extension, ENC = src/lib/H264AVCEncoderLib)
//   MotionEstimation::xPelSpiralSearch      ENC/MotionEstimation.cpp:415-468 (table built in init, :69-113)
//   MotionEstimation::xPelLogSearch         :472-709   (LOG_SEARCH; FAST_SEARCH = the same seeded with rcMvPred and the neighbours' MVs)
//   MotionEstimation::xTZSearch             :1106-1292 with xTZCheckPoint :733-757, xTZ2PointSearch :760-880,
//                                           xTZ8PointDiamondSearch :929-1103 and the TZ_SEARCH_CONFIGURATION constants :713-730
//   the strategy switch of estimateBlockWithStart :262-307, incl. the refinement searches of the bi-prediction iterations
//   (uiSearchRange != 0: xPelBlockSearch over that range, or xTZSearch with an empty candidate list when FastBiSearch is set)
//
// How the sequential reference maps to a warp.  Every search is a chain of "rounds": a short ORDERED list of points, each of which
// the reference tests with  if( best > cost ) { best = cost; remember the point }.  After one round the state is therefore
// "the FIRST point of the list that attains the round's minimum, if that minimum is below the incumbent" -- which does not depend
// on evaluating the points one after the other.  A round is evaluated 8 points at a time (4 lanes per point, each lane a share of the
// block's 4x4 sub-blocks), followed by a warp argmin with the list index as tie-break; rounds follow each other as in the reference.
// The point lists, their order and the ONE-SIDED border tests are the reference's: a start point outside the search rectangle
// (a neighbour's MV, the zero vector) is legal and its diamond points are tested against one border only, exactly as there.
//
// Distortion: every SearchFuncFullPel except YUV_SAD under LOG / FAST (xPelLogSearch never sets the chroma search pointers, so the
// reference itself is undefined there; jsvm_me_set_params refuses that pair).  Samples are read from the full-pel plane through
// L1 / L2 (aligned 32-bit loads + funnel shift): a fast search touches a few hundred points, staging a window would cost more.
#pragma once

#include "me_common.cuh"

namespace jsvm {

constexpr int kK5WarpsPerCta = 4;

struct K5Rect { int neg_h, pos_h, neg_v, pos_v; };        // SearchRect (ENC/MotionEstimation.h:42-62)

struct K5Job {
  const uint8_t* ref0;          // reference luma sample under the block's top-left corner at displacement (0,0)
  const uint8_t* cref0[2];      // the same for Cb / Cr (YUV_SAD)
  const int16_t* org;           // shared: block-relative original, stride 16
  const uint32_t* orgp;         // shared: the same clipped to [0,255] and packed, 4 words per row (SAD / YUV_SAD)
  const int16_t* corg;          // shared: chroma originals [2][64], stride 8
  int w, h;
  int plane_stride, chroma_stride;
  int nbx, nb;                  // 4x4 sub-blocks per row / in the block
  int pred_h, pred_v;
  uint32_t f;
  bool yuv;
};

struct K5Best { uint32_t cost; int x, y, dist, nr; };

// four consecutive bytes at an arbitrary address (two aligned words + funnel shift)
__device__ __forceinline__ uint32_t ldg_u8x4(const uint8_t* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(a & 3) * 8;
  const uint32_t lo = __ldg(w);
  if (sh == 0) return lo;
  return __funnelshift_r(lo, __ldg(w + 1), sh);
}

__device__ __forceinline__ uint32_t had4x4_from_diffs(const int (&d)[16]) {     // xCalcHadamard4x4 (ENC/Distortion.cpp:291-340), N8
  int m[16];
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    const int t0 = d[n] + d[12 + n], t1 = d[4 + n] + d[8 + n], t2 = d[4 + n] - d[8 + n], t3 = d[n] - d[12 + n];
    m[n] = t0 + t1; m[8 + n] = t0 - t1; m[4 + n] = t3 + t2; m[12 + n] = t3 - t2;
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

// this lane's share of the distortion at displacement (x, y).
// SAD (and the luma part of YUV_SAD): rows sub, sub + 4, ... of the block on packed bytes -- the reference row is assembled from
// aligned words with funnel shifts and meets the clipped, packed original in VABSDIFF4 (4 abs-diffs + add per instruction).
// Originals outside [0,255] (residual-prediction passes, N3) are searched as their clip: |o - p| = |clip(o) - p| + |o - clip(o)|
// for p in [0,255], and the second term does not depend on the displacement (the kernel adds it to the cost it reports).
// SSD / HADAMARD: sub-blocks sub, sub + 4, ... on the true 16-bit differences.
template <int DFUNC>
__device__ __forceinline__ uint32_t k5_partial(const K5Job& J, int x, int y, int sub, int nsub) {
  uint32_t dist = 0;
  if (DFUNC == JSVM_DF_SAD || DFUNC == JSVM_DF_YUV_SAD) {
    const int nw = J.w >> 2;
    for (int r = sub; r < J.h; r += nsub) {
      const uint8_t* rp = J.ref0 + (ptrdiff_t)(y + r) * J.plane_stride + x;
      const uintptr_t a = reinterpret_cast<uintptr_t>(rp);
      const uint32_t* wp = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
      const uint32_t sh = (uint32_t)(a & 3) * 8;
      const uint32_t* op = J.orgp + r * 4;
      uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1);
      dist = sad4(op[0], __funnelshift_r(w0, w1, sh), dist);
      if (nw >= 2) {
        w0 = __ldg(wp + 2);
        dist = sad4(op[1], __funnelshift_r(w1, w0, sh), dist);
        if (nw == 4) {
          w1 = __ldg(wp + 3);
          dist = sad4(op[2], __funnelshift_r(w0, w1, sh), dist);
          w0 = __ldg(wp + 4);
          dist = sad4(op[3], __funnelshift_r(w1, w0, sh), dist);
        }
      }
    }
    if (DFUNC == JSVM_DF_YUV_SAD && J.yuv) {
      // chroma: (w/2) x (h/2) samples of Cb and Cr displaced by ( x>>1, y>>1 ) (ENC/MotionEstimation.cpp:743-744); rows sub, sub + 4, ..
      for (int comp = 0; comp < 2; ++comp)
        for (int r = sub; r < (J.h >> 1); r += nsub) {
          const uint8_t* cp = J.cref0[comp] + (ptrdiff_t)((y >> 1) + r) * J.chroma_stride + (x >> 1);
          const int16_t* co = J.corg + comp * 64 + r * 8;
          for (int c = 0; c < (J.w >> 1); ++c) dist += (uint32_t)abs((int)co[c] - (int)__ldg(cp + c));
        }
    }
  } else {
  for (int b = sub; b < J.nb; b += nsub) {
    const int bx = (b % J.nbx) * 4, by = (b / J.nbx) * 4;
    const uint8_t* rp = J.ref0 + (ptrdiff_t)(y + by) * J.plane_stride + x + bx;
    int d[16];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint32_t v = ldg_u8x4(rp + (ptrdiff_t)r * J.plane_stride);
      const int2 o = *reinterpret_cast<const int2*>(J.org + (by + r) * 16 + bx);      // four int16 originals
      d[4 * r + 0] = (int)(short)(o.x & 0xffff) - (int)(v & 0xff);
      d[4 * r + 1] = (o.x >> 16) - (int)((v >> 8) & 0xff);
      d[4 * r + 2] = (int)(short)(o.y & 0xffff) - (int)((v >> 16) & 0xff);
      d[4 * r + 3] = (o.y >> 16) - (int)(v >> 24);
    }
    if (DFUNC == JSVM_DF_HADAMARD) dist += had4x4_from_diffs(d);
    else {
#pragma unroll
      for (int i = 0; i < 16; ++i) dist += DFUNC == JSVM_DF_SSD ? (uint32_t)(d[i] * d[i]) : (uint32_t)abs(d[i]);
    }
  }
  }
  return dist;
}

struct K5Point { int x, y, nr, dist; bool valid; };

// One round: `count` points produced by gen(k) in list order, 8 at a time.  All lanes return the same state.
template <int DFUNC, typename Gen>
__device__ __forceinline__ int k5_round(const K5Job& J, K5Best& st, int count, Gen gen) {
  const int lane = threadIdx.x & 31, g = lane >> 2, sub = lane & 3;
  int winner = -1;                                            // list index of the point that replaced the incumbent, if any
  int k0 = 0;
  // long lists (raster, spiral, block search): 32 points at a time, one lane per point -- no partial sums to combine
  for (; k0 + 32 <= count; k0 += 32) {
    const int k = k0 + lane;
    const K5Point p = gen(k);
    uint32_t cost = 0xffffffffu;
    if (p.valid)
      cost = k5_partial<DFUNC>(J, p.x, p.y, 0, 1) + rate_cost(J.f, mv_component_bits((p.x << 2) - J.pred_h) + mv_component_bits((p.y << 2) - J.pred_v));
    const uint32_t m = __reduce_min_sync(0xffffffffu, cost);
    if (m < st.cost) {
      const int src = __ffs(__ballot_sync(0xffffffffu, p.valid && cost == m)) - 1;
      st.cost = m;
      st.x = __shfl_sync(0xffffffffu, p.x, src); st.y = __shfl_sync(0xffffffffu, p.y, src);
      st.dist = __shfl_sync(0xffffffffu, p.dist, src); st.nr = __shfl_sync(0xffffffffu, p.nr, src);
      winner = k0 + src;
    }
  }
  for (; k0 < count; k0 += 8) {
    const int k = k0 + g;
    K5Point p = gen(k);
    p.valid = p.valid && k < count;
    uint32_t cost = 0xffffffffu;
    uint32_t part = p.valid ? k5_partial<DFUNC>(J, p.x, p.y, sub, 4) : 0u;
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    if (p.valid)    // xGetCost( x, y ) with m_uiMvScaleShift = 2 (N6), 32-bit wrap-around (N5)
      cost = part + rate_cost(J.f, mv_component_bits((p.x << 2) - J.pred_h) + mv_component_bits((p.y << 2) - J.pred_v));
    const uint32_t m = __reduce_min_sync(0xffffffffu, cost);
    if (m < st.cost) {                                        // strict: the incumbent survives ties
      const int src = __ffs(__ballot_sync(0xffffffffu, p.valid && cost == m)) - 1;      // first point of the list with that cost
      st.cost = m;
      st.x = __shfl_sync(0xffffffffu, p.x, src); st.y = __shfl_sync(0xffffffffu, p.y, src);
      st.dist = __shfl_sync(0xffffffffu, p.dist, src); st.nr = __shfl_sync(0xffffffffu, p.nr, src);
      winner = k0 + (src >> 2);
    }
  }
  return winner;
}

// xTZ8PointDiamondSearch: point k of the list for distance d around (sx, sy); list lengths 4 (d == 1), 8 (d <= 8), 16
__device__ __forceinline__ K5Point k5_diamond_point(int k, int sx, int sy, int d, const K5Rect& r) {
  K5Point p; p.nr = 0; p.dist = d; p.valid = false; p.x = sx; p.y = sy;
  const int top = sy - d, bot = sy + d, left = sx - d, right = sx + d;
  if (d == 1) {
    switch (k) {
      case 0: p.y = top;   p.nr = 2; p.valid = top >= -r.neg_v; break;
      case 1: p.x = left;  p.nr = 4; p.valid = left >= -r.neg_h; break;
      case 2: p.x = right; p.nr = 5; p.valid = right <= r.pos_h; break;
      case 3: p.y = bot;   p.nr = 7; p.valid = bot <= r.pos_v; break;
      default: break;
    }
  } else if (d <= 8) {
    const int h = d >> 1, top2 = sy - h, bot2 = sy + h, left2 = sx - h, right2 = sx + h;
    switch (k) {
      case 0: p.y = top; p.nr = 2; p.valid = top >= -r.neg_v; break;
      case 1: p.x = left2;  p.y = top2; p.nr = 1; p.dist = h; p.valid = top2 >= -r.neg_v && left2 >= -r.neg_h; break;
      case 2: p.x = right2; p.y = top2; p.nr = 3; p.dist = h; p.valid = top2 >= -r.neg_v && right2 <= r.pos_h; break;
      case 3: p.x = left;  p.nr = 4; p.valid = left >= -r.neg_h; break;
      case 4: p.x = right; p.nr = 5; p.valid = right <= r.pos_h; break;
      case 5: p.x = left2;  p.y = bot2; p.nr = 6; p.dist = h; p.valid = bot2 <= r.pos_v && left2 >= -r.neg_h; break;
      case 6: p.x = right2; p.y = bot2; p.nr = 8; p.dist = h; p.valid = bot2 <= r.pos_v && right2 <= r.pos_h; break;
      case 7: p.y = bot; p.nr = 7; p.valid = bot <= r.pos_v; break;
      default: break;
    }
  } else {
    if (k < 4) {
      switch (k) {
        case 0: p.y = top;   p.valid = top >= -r.neg_v; break;
        case 1: p.x = left;  p.valid = left >= -r.neg_h; break;
        case 2: p.x = right; p.valid = right <= r.pos_h; break;
        default: p.y = bot;  p.valid = bot <= r.pos_v; break;
      }
    } else if (k < 16) {
      const int idx = ((k - 4) >> 2) + 1, m = (k - 4) & 3, q = (d >> 2) * idx;
      const bool up = m < 2, lft = (m & 1) == 0;
      p.y = up ? top + q : bot - q;
      p.x = lft ? sx - q : sx + q;
      p.valid = (up ? p.y >= -r.neg_v : p.y <= r.pos_v) && (lft ? p.x >= -r.neg_h : p.x <= r.pos_h);
    }
  }
  return p;
}

// xTZ2PointSearch: the two untested neighbours of the best point found at distance 1 (point numbers 1..8)
__device__ __forceinline__ K5Point k5_two_point(int k, int x, int y, int nr, const K5Rect& r) {
  // offsets (dx, dy) of the two points per ucPointNr, 2 bits each biased by 1
  constexpr int8_t off[9][4] = { {0, 0, 0, 0}, {-1, 0, 0, -1}, {-1, -1, 1, -1}, {0, -1, 1, 0}, {-1, 1, -1, -1},
                                 {1, -1, 1, 1}, {-1, 0, 0, 1}, {-1, 1, 1, 1}, {1, 0, 0, 1} };
  K5Point p; p.nr = 0; p.dist = 2; p.valid = false; p.x = x; p.y = y;
  if (k < 2 && nr >= 1 && nr <= 8) {
    const int dx = off[nr][2 * k], dy = off[nr][2 * k + 1];
    p.x = x + dx; p.y = y + dy;
    p.valid = (dx < 0 ? p.x >= -r.neg_h : dx > 0 ? p.x <= r.pos_h : true) && (dy < 0 ? p.y >= -r.neg_v : dy > 0 ? p.y <= r.pos_v : true);
  }
  return p;
}

// spiral table entry k (MotionEstimation::init, ENC/MotionEstimation.cpp:86-111): rings of growing offset, inside a ring first the
// top / bottom rows (without corners, interleaved), then the left / right columns (interleaved)
__device__ __forceinline__ void k5_spiral_offset(int k, int& dx, int& dy) {
  if (k == 0) { dx = 0; dy = 0; return; }
  int s = (int)sqrtf((float)k);
  while (s * s > k) --s;
  while ((s + 1) * (s + 1) <= k) ++s;
  const int off = (s + 1) >> 1;                                // (2 off - 1)^2 <= k < (2 off + 1)^2
  const int r = k - (2 * off - 1) * (2 * off - 1);
  const int n_rows = 2 * (2 * off - 1);
  if (r < n_rows) { dx = -off + 1 + (r >> 1); dy = (r & 1) ? off : -off; }
  else { const int r2 = r - n_rows; dy = -off + (r2 >> 1); dx = (r2 & 1) ? off : -off; }
}

template <int DFUNC>
__global__ void __launch_bounds__(32 * kK5WarpsPerCta)
fast_search_kernel(const uint8_t* __restrict__ planes, size_t plane_bytes, int plane_stride,
                   const uint8_t* __restrict__ chroma_planes, size_t chroma_bytes, int chroma_stride,
                   const jsvm_me_job* __restrict__ jobs, int n_jobs,
                   const int16_t* __restrict__ org_mbs, const int16_t* __restrict__ org_chroma_mbs,
                   const jsvm_mv* __restrict__ cands,          // 3 per job (A, B, C) or NULL (all zero)
                   int mb_w, int mb_h, int search_mode, int default_range, int el_range, int fast_bi,
                   int only_refinement,                        // BLOCK_SEARCH + FastBiSearch: only the jobs with a per-call range
                   uint32_t* __restrict__ keys, uint32_t* __restrict__ full_costs, JobAux* __restrict__ aux)
{
  __shared__ __align__(16) int16_t s_org[kK5WarpsPerCta][256];
  __shared__ __align__(16) uint32_t s_orgp[kK5WarpsPerCta][64];
  __shared__ __align__(16) int16_t s_corg[kK5WarpsPerCta][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int job_idx = blockIdx.x * kK5WarpsPerCta + warp;
  if (job_idx >= n_jobs) return;
  const jsvm_me_job j = jobs[job_idx];
  if (only_refinement && j.search_range == 0) return;
  int w = 16, h = 16;
  mode_size(j.mode, w, h);
  const int bx0 = 16 * j.mb_x + 4 * (j.blk & 3), by0 = 16 * j.mb_y + 4 * (j.blk >> 2);
  const bool bipred = (j.mode & JSVM_JOB_BIPRED) != 0, el = (j.mode & JSVM_JOB_EL) != 0;

  K5Job J;
  J.plane_stride = plane_stride; J.chroma_stride = chroma_stride;
  J.ref0 = planes + (size_t)j.ref_slot * plane_bytes + (size_t)(kMargin + by0) * plane_stride + kMargin + bx0;
  J.org = s_org[warp]; J.orgp = s_orgp[warp]; J.corg = s_corg[warp];
  J.w = w; J.h = h;
  J.nbx = w >> 2; J.nb = (w >> 2) * (h >> 2);
  J.pred_h = j.pred.hor; J.pred_v = j.pred.ver;
  J.f = j.cost_factor;
  // ENC/MotionEstimation.cpp:250-253: bi-predictive and unequally weighted originals are searched with the luma SAD
  J.yuv = DFUNC == JSVM_DF_YUV_SAD && !(j.org_index >= 0 && org_chroma_mbs == nullptr) && !bipred;
  J.cref0[0] = J.cref0[1] = nullptr;
  for (int i = lane; i < 256; i += 32) {
    const int y = i >> 4, x = i & 15;
    int16_t v = 0;
    if (x < w && y < h)
      v = j.org_index >= 0
          ? __ldg(org_mbs + (size_t)j.org_index * 256 + (4 * (j.blk >> 2) + y) * 16 + 4 * (j.blk & 3) + x)
          : (int16_t)__ldg(planes + (size_t)j.cur_slot * plane_bytes + (size_t)(kMargin + by0 + y) * plane_stride + kMargin + bx0 + x);
    s_org[warp][i] = v;
  }
  // clipped, packed copy for the byte-wise SAD and the displacement-independent excess of originals outside [0,255]
  uint32_t excess = 0;
  if (DFUNC == JSVM_DF_SAD || DFUNC == JSVM_DF_YUV_SAD) {
    __syncwarp();
    for (int i = lane; i < 64; i += 32) {
      const short4 q = *reinterpret_cast<const short4*>(&s_org[warp][4 * i]);
      s_orgp[warp][i] = pack_clip8(q);
      const int o[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) excess += (uint32_t)abs(o[k] - min(max(o[k], 0), 255));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) excess += __shfl_xor_sync(0xffffffffu, excess, off);
  }
  if (DFUNC == JSVM_DF_YUV_SAD && J.yuv) {
    for (int comp = 0; comp < 2; ++comp)
      J.cref0[comp] = chroma_planes + ((size_t)j.ref_slot * 2 + comp) * chroma_bytes + (size_t)(kCMargin + by0 / 2) * chroma_stride + kCMargin + bx0 / 2;
    for (int i = lane; i < 128; i += 32) {
      const int comp = i >> 6, y = (i >> 3) & 7, x = i & 7;
      int16_t v = 0;
      if (x < (w >> 1) && y < (h >> 1)) {
        const int cy = 2 * (j.blk >> 2) + y, cx = 2 * (j.blk & 3) + x;
        v = j.org_index >= 0
            ? __ldg(org_chroma_mbs + (size_t)j.org_index * 128 + comp * 64 + cy * 8 + cx)
            : (int16_t)__ldg(chroma_planes + ((size_t)j.cur_slot * 2 + comp) * chroma_bytes +
                             (size_t)(kCMargin + 8 * j.mb_y + cy) * chroma_stride + kCMargin + 8 * j.mb_x + cx);
      }
      s_corg[warp][i] = v;
    }
  }
  __syncwarp();

  // m_cMin / m_cMax (COM/MotionCompensation.cpp:152-167) and Mv::limitComponents + '>>= 2'
  const int min_h = sat16(((-(int)j.mb_x << 4) - 24) << 2), min_v = sat16(((-(int)j.mb_y << 4) - 24) << 2);
  const int max_h = sat16((((mb_w - j.mb_x) << 4) + 8) << 2), max_v = sat16((((mb_h - j.mb_y) << 4) + 8) << 2);
  auto pel_h = [&](int v) { return (int)(int16_t)min(max(v, min_h), max_h) >> 2; };
  auto pel_v = [&](int v) { return (int)(int16_t)min(max(v, min_v), max_v) >> 2; };
  const int start_x = pel_h(j.start.hor), start_y = pel_v(j.start.ver);

  // which search (ENC/MotionEstimation.cpp:262-307)
  enum { ALG_BLOCK, ALG_SPIRAL, ALG_LOG, ALG_FAST, ALG_TZ };
  int alg, range;
  bool with_cands = false;
  if (j.search_range) { alg = fast_bi ? ALG_TZ : ALG_BLOCK; range = j.search_range; }
  else {
    alg = search_mode; range = default_range;
    if (search_mode == JSVM_SEARCH_TZ) { with_cands = true; range = el ? el_range : default_range; }
    if (search_mode == JSVM_SEARCH_FAST) with_cands = true;
  }
  // SearchRect::init( range, clipped start, m_cMin, m_cMax ) with its Short casts (ENC/MotionEstimation.h:44-56)
  K5Rect R;
  {
    const int16_t lim = (int16_t)range;
    const int16_t pos_v = (int16_t)((max_v >> 2) - start_y), neg_v = (int16_t)(start_y - (min_v >> 2));
    const int16_t pos_h = (int16_t)((max_h >> 2) - start_x), neg_h = (int16_t)(start_x - (min_h >> 2));
    R.neg_v = min((int)lim, (int)neg_v) - start_y; R.pos_v = min((int)lim, (int)pos_v) + start_y;
    R.neg_h = min((int)lim, (int)neg_h) - start_x; R.pos_h = min((int)lim, (int)pos_h) + start_x;
  }
  // candidate list: rcMvPred, A, B, C (each limited to the macroblock's range, then >> 2)
  int cand_x[4] = {0, 0, 0, 0}, cand_y[4] = {0, 0, 0, 0};
  if (with_cands) {
    cand_x[0] = pel_h(j.pred.hor); cand_y[0] = pel_v(j.pred.ver);
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      const jsvm_mv m = cands ? cands[3 * (size_t)job_idx + n] : jsvm_mv{0, 0};
      cand_x[n + 1] = pel_h(m.hor); cand_y[n + 1] = pel_v(m.ver);
    }
  }

  K5Best st; st.cost = 0xffffffffu; st.x = start_x; st.y = start_y; st.dist = 0; st.nr = 0;

  if (alg == ALG_BLOCK) {
    // xPelBlockSearch over the clipped window, raster order (N1)
    const int ww = R.pos_h + R.neg_h + 1, wh = R.pos_v + R.neg_v + 1;
    if (ww > 0 && wh > 0)
      k5_round<DFUNC>(J, st, ww * wh, [&](int k) { K5Point p; p.x = -R.neg_h + k % ww; p.y = -R.neg_v + k / ww; p.nr = 0; p.dist = 0; p.valid = true; return p; });
  } else if (alg == ALG_SPIRAL) {
    // start limited to [ m_cMin + 4 SR, m_cMax - 4 SR ] (the planner refuses macroblocks where that interval is empty)
    const int lo_h = (int16_t)(min_h + 4 * default_range), hi_h = (int16_t)(max_h - 4 * default_range);
    const int lo_v = (int16_t)(min_v + 4 * default_range), hi_v = (int16_t)(max_v - 4 * default_range);
    const int cx = (int)(int16_t)min(max((int)j.start.hor, lo_h), hi_h) >> 2, cy = (int)(int16_t)min(max((int)j.start.ver, lo_v), hi_v) >> 2;
    st.x = cx; st.y = cy;
    const int side = 2 * default_range + 1;
    k5_round<DFUNC>(J, st, side * side, [&](int k) { K5Point p; int dx, dy; k5_spiral_offset(k, dx, dy); p.x = cx + dx; p.y = cy + dy; p.nr = 0; p.dist = 0; p.valid = true; return p; });
  } else if (alg == ALG_LOG || alg == ALG_FAST) {
    int step0;
    if (alg == ALG_LOG) {                                     // m_iMaxLogStep << ( pcBSP ? 0 : 1 ) (:50-58, :289)
      unsigned lim = (unsigned)default_range >> 4; step0 = 1;
      while (lim) { lim >>= 1; step0 <<= 1; }
      step0 <<= bipred ? 0 : 1;
    } else step0 = bipred ? 1 : 2;
    // the start, then the candidates (strict improvement only)
    k5_round<DFUNC>(J, st, with_cands ? 5 : 1, [&](int k) {
      K5Point p; p.nr = 0; p.dist = 0; p.valid = true;
      p.x = k == 0 ? start_x : cand_x[(k - 1) & 3]; p.y = k == 0 ? start_y : cand_y[(k - 1) & 3];
      return p; });
    int x = st.x, y = st.y, dx_old = 0, dy_old = 0;
    for (int step = step0; step != 0; step >>= 1) {
      dx_old = 0; dy_old = 0;
      for (;;) {
        const int win = k5_round<DFUNC>(J, st, 4, [&](int k) {
          K5Point p; p.nr = 0; p.dist = 0; p.x = x; p.y = y; p.valid = false;
          switch (k) {
            case 0: p.x = x - step; p.valid = dx_old <= 0 && -x <= R.neg_h - step; break;
            case 1: p.x = x + step; p.valid = dx_old >= 0 && x < R.pos_h - step; break;
            case 2: p.y = y - step; p.valid = dy_old <= 0 && -y <= R.neg_v - step; break;
            case 3: p.y = y + step; p.valid = dy_old >= 0 && y < R.pos_v - step; break;
            default: break;
          }
          return p; });
        if (win < 0) break;
        dx_old = win == 0 ? -step : win == 1 ? step : 0;
        dy_old = win == 2 ? -step : win == 3 ? step : 0;
        x += dx_old; y += dy_old;
      }
    }
    k5_round<DFUNC>(J, st, 4, [&](int k) {                    // the four diagonal neighbours, once
      K5Point p; p.nr = 0; p.dist = 0;
      const bool lft = k < 2, up = (k & 1) == 0;
      p.x = lft ? x - 1 : x + 1; p.y = up ? y - 1 : y + 1;
      p.valid = k < 4 && (lft ? (dx_old != 1 && -x <= R.neg_h - 1) : (dx_old != -1 && x < R.pos_h - 1)) &&
                (up ? (dy_old != 1 && -y <= R.neg_v - 1) : (dy_old != -1 && y < R.pos_v - 1));
      return p; });
  } else {
    // ---- xTZSearch ----
    k5_round<DFUNC>(J, st, 6, [&](int k) {                    // start, candidates, zero vector
      K5Point p; p.nr = 0; p.dist = 0;
      p.valid = k == 0 || k == 5 || (with_cands && k < 5);
      p.x = k == 0 ? start_x : k == 5 ? 0 : cand_x[(k - 1) & 3];
      p.y = k == 0 ? start_y : k == 5 ? 0 : cand_y[(k - 1) & 3];
      return p; });
    int sx = st.x, sy = st.y;
    for (int d = 1; d <= range; d *= 2)
      k5_round<DFUNC>(J, st, d == 1 ? 4 : d <= 8 ? 8 : 16, [&](int k) { return k5_diamond_point(k, sx, sy, d, R); });
    if (st.dist == 1) {
      st.dist = 0;
      const int x = st.x, y = st.y, nr = st.nr;
      k5_round<DFUNC>(J, st, 2, [&](int k) { return k5_two_point(k, x, y, nr, R); });
    }
    if (st.dist > 3) {                                        // raster search, step iRaster = 3
      st.dist = 3;
      const int nx = (R.pos_h + R.neg_h) / 3 + 1, ny = (R.pos_v + R.neg_v) / 3 + 1;
      if (R.pos_h + R.neg_h >= 0 && R.pos_v + R.neg_v >= 0)
        k5_round<DFUNC>(J, st, nx * ny, [&](int k) { K5Point p; p.x = -R.neg_h + 3 * (k % nx); p.y = -R.neg_v + 3 * (k / nx); p.nr = 0; p.dist = 3; p.valid = true; return p; });
    }
    while (st.dist > 0) {                                     // star refinement
      sx = st.x; sy = st.y; st.dist = 0; st.nr = 0;
      for (int d = 1; d < range + 1; d *= 2)
        k5_round<DFUNC>(J, st, d == 1 ? 4 : d <= 8 ? 8 : 16, [&](int k) { return k5_diamond_point(k, sx, sy, d, R); });
      if (st.dist == 1) {
        st.dist = 0;
        if (st.nr != 0) {
          const int x = st.x, y = st.y, nr = st.nr;
          k5_round<DFUNC>(J, st, 2, [&](int k) { return k5_two_point(k, x, y, nr, R); });
        }
      }
    }
  }

  if (lane == 0) {
    keys[job_idx] = 0;                                        // K3: winner = (ux0, uy0) + index 0 of a 1-wide window
    full_costs[job_idx] = st.cost + excess;
    JobAux a; a.ux0 = (int16_t)st.x; a.uy0 = (int16_t)st.y; a.uw = 1; a.pad = 1;
    aux[job_idx] = a;
  }
}

}  // namespace jsvm
