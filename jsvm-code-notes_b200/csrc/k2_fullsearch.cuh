// k2_fullsearch.cuh -- K2: exhaustive integer-pel block matching for all partitions of a macroblock.
//
// Replaces the hot loop of the reference,
//   MotionEstimation::xPelBlockSearch            ENC/MotionEstimation.cpp:351-412
//   x XDistortion::xGetSAD16x / 8x / 4x          ENC/Distortion.cpp:397-429, 563-587, 691-711
//   + MotionEstimationCost::xGetCost( x, y )     ENC/MotionEstimationCost.h:26-28
// with one SAD pass per (macroblock, displacement) that is shared by every partition searched
// on that macroblock (64 packed abs-diff-accumulate per position instead of 256*7).
//
// One CTA per search group (see Group in me_common.cuh):
//   * the reference window (union of the jobs' clipped windows + 15 rows/cols) is staged in shared
//     memory as FOUR byte-shifted copies (copy s starts at window_x + s; built once per CTA from
//     16-byte global loads + funnel shifts -- TMA cannot do it: its innermost start coordinate
//     must be 16-byte aligned), so that a thread at
//     displacement x = 4m + s reads its 16 reference bytes of a row as the four aligned words
//     m..m+3 of copy s -- no funnel shifts in the hot loop, and lanes with consecutive m hit
//     consecutive banks;
//   * the current macroblock lives in 64 registers per thread (16 rows x 4 packed words);
//   * a thread owns a strip of V vertically adjacent displacements: every reference row it loads
//     is reused for up to V displacements (VABSDIFF4.U8.ACC against different current rows);
//   * per displacement the 4 (8x8 granularity) or 16 (4x4 granularity) SADs are combined into the
//     partition SADs, the lambda-weighted MV rate is added, and (cost << 15 | raster index) keys are
//     min-reduced: lowest cost first, then lowest raster index = the reference's "first strict
//     minimum in y-outer / x-inner scan order" tie-break (N1).
#pragma once

#include <cuda.h>
#include <type_traits>
#include "me_common.cuh"

namespace jsvm {

constexpr int kK2Threads   = 256;
constexpr int kWinRowBytes = 160;                 // bytes per staged window row (max 145 positions + 15)
constexpr int kWinRowWords = kWinRowBytes / 4;    // 40 words: 40 mod 32 = 8 keeps successive rows on different banks
constexpr int kMaxUW       = 145;                 // widest union window handled by one CTA
constexpr int kMaxM        = (kMaxUW + 3) / 4;    // 37
constexpr int kMaxRows     = 184;                 // window rows covered by the strips (uh <= 176, rounded up to a strip)

__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// ---- mbarrier / TMA helpers (sm_90+ PTX; SASS: SYNCS.*, UTMALDG) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 3D tiled load: tensor = [slot][row][byte] of the full-pel plane pool
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

// Which SAD accumulators make up a partition slot.
//   8x8 granularity : acc q = (row >= 8) * 2 + (col >= 8)                       (4 accumulators)
//   4x4 granularity : acc a = (row / 4) * 4 + (col / 4)                         (16 accumulators)
__host__ __device__ constexpr uint32_t slot_mask_gran8(int sl) {
  return sl == 0 ? 0xFu : sl == 1 ? 0x3u : sl == 2 ? 0xCu : sl == 3 ? 0x5u : sl == 4 ? 0xAu : (1u << (sl - 5));
}
__host__ __device__ constexpr uint32_t slot_mask_gran4(int sl) {
  if (sl == 0) return 0xFFFFu;
  if (sl <= 2) return sl == 1 ? 0x00FFu : 0xFF00u;
  if (sl <= 4) return sl == 3 ? 0x3333u : 0xCCCCu;
  if (sl <= 8) { int b = sl - 5;  return 0x0033u << ((b >> 1) * 8 + (b & 1) * 2); }
  if (sl <= 16) { int i = sl - 9,  b = i >> 1; return 0x0003u << ((b >> 1) * 8 + (b & 1) * 2 + (i & 1) * 4); }
  if (sl <= 24) { int i = sl - 17, b = i >> 1; return 0x0011u << ((b >> 1) * 8 + (b & 1) * 2 + (i & 1)); }
  { int i = sl - 25, b = i >> 2, k = i & 3;    return 0x0001u << ((b >> 1) * 8 + (b & 1) * 2 + (k >> 1) * 4 + (k & 1)); }
}

// sad + ((factor * bits) >> 16), packed with the raster index: (cost << 15) + idx.
//   t    = f * by + fbx          f * (bits_x + bits_y) in 32-bit wrap-around arithmetic (N5); FMA pipe (IMAD)
//   cost = sad + (t >> 16)       xGetCost, ENC/MotionEstimationCost.h:27
//   key  = cost * 32768 + idx    FMA pipe (IMAD); the min itself is the only ALU-pipe op besides the shift-add
// Written as PTX so that the compiler cannot re-associate (cost << 15) into one multiply per SAD term.
__device__ __forceinline__ uint32_t rate_key(uint32_t f, uint32_t by, uint32_t fbx, uint32_t sad, uint32_t idx) {
  uint32_t key;
  asm("{\n"
      ".reg .u32 t, c;\n"
      "mad.lo.u32 t, %1, %2, %3;\n"
      "shr.u32 t, t, 16;\n"
      "add.u32 c, t, %4;\n"
      "mad.lo.u32 %0, c, 32768, %5;\n"
      "}\n" : "=r"(key) : "r"(f), "r"(by), "r"(fbx), "r"(sad), "r"(idx));
  return key;
}

template <int V, bool GRAN4>
struct K2Smem {
  static constexpr int NS   = GRAN4 ? kSlots : kSlotsGran8;
  static constexpr int NACC = GRAN4 ? 16 : 4;
  static constexpr int kMaxStrips = kMaxRows / V;
  typedef typename std::conditional<GRAN4, unsigned long long, uint32_t>::type Mask;   // one bit per partition slot
  // dynamic shared memory layout (bytes); the four window copies come first
  uint32_t factor[NS];
  int16_t  xlo[NS], xhi[NS], ylo[NS], yhi[NS];    // relative to the union window origin (empty slot: xhi < xlo)
  int32_t  job[NS];                               // absolute job index or -1
  uint32_t best[NS];
  uint8_t  bx[NS][4 * kMaxM + 12];                // bits of the horizontal difference, permuted [s][m]
  Mask     xmask[4 * kMaxM];                      // slots whose window contains column x, permuted [s][m]
  Mask     yfull[kMaxStrips];                     // slots whose window contains ALL rows of the strip
  Mask     yany[kMaxStrips];                      // slots whose window contains SOME row of the strip
  // followed by: uint32_t by[NS][nstrips * V]    bits of the vertical MV difference per window row (16-byte aligned)
};

// Reference rows staged per copy: only the uh + 15 rows that valid displacements touch.  The last strip may reach up to
// V-1 rows further; those reads belong to displacements outside every window (masked by yany / yfull), land in the
// shared memory that follows the copies (tables) and are never used.
__host__ __device__ inline int k2_rows(int uh) { return uh + 15; }
// words between two window copies: padded so that (copy stride mod 32) == (M rounded up to even) mod 32 -- a warp whose
// lanes run off the end of copy s (m = M-1) continues in copy s+1 (m = 0) on the NEXT bank: no shared-memory conflicts.
__host__ __device__ inline int k2_copy_words(int rows, int m_units) {
  const int base = rows * kWinRowWords;
  return base + (((m_units - base) % 32) + 32) % 32;
}

template <int V, bool GRAN4>
__global__ void __launch_bounds__(kK2Threads, GRAN4 ? 1 : 2)
full_search_kernel(const uint8_t* __restrict__ planes,            // full-pel plane pool [slot][H+64][W+64]
                   size_t plane_bytes, int plane_stride,
                   const Group* __restrict__ groups,
                   const jsvm_me_job* __restrict__ jobs,
                   const int16_t* __restrict__ org_mbs,
                   int mb_w, int mb_h, int default_range,
                   uint32_t* __restrict__ keys,                   // per job: packed (cost << 15 | idx)
                   JobAux* __restrict__ aux)                      // per job: how K3 decodes the key
{
  using SM = K2Smem<V, GRAN4>;
  using Mask = typename SM::Mask;
  constexpr int NS = SM::NS, NACC = SM::NACC;
  extern __shared__ __align__(128) uint8_t smem_raw[];

  const Group& g = groups[blockIdx.x];
  const int tid = threadIdx.x;
  const int uw = g.uw, uh = g.uh;
  const int nstrips = (uh + V - 1) / V;
  const int rows = k2_rows(uh);
  const int M = ((uw + 3) >> 2) + (((uw + 3) >> 2) & 1);          // column groups per copy, rounded up to even
  const int copy_words = k2_copy_words(rows, M);
  const int by_stride = nstrips * V;
  uint32_t* winw = reinterpret_cast<uint32_t*>(smem_raw);
  uint8_t* after = smem_raw + (((size_t)4 * copy_words * 4 + 15) & ~(size_t)15);
  SM& sm = *reinterpret_cast<SM*>(after);
  uint32_t* by_tab = reinterpret_cast<uint32_t*>(after + ((sizeof(SM) + 15) & ~(size_t)15));

  // ---- prologue 1: per-slot job parameters ----
  if (tid < NS) {
    const int jo = g.slot_job[tid];
    sm.job[tid] = jo < 0 ? -1 : g.first_job + jo;
    sm.best[tid] = kKeyInvalid;
    if (jo >= 0) {
      const jsvm_me_job j = jobs[g.first_job + jo];
      const Window w = job_window(j, mb_w, mb_h, default_range);
      sm.factor[tid] = j.cost_factor;
      sm.xlo[tid] = (int16_t)(w.xlo - g.ux0); sm.xhi[tid] = (int16_t)(w.xhi - g.ux0);
      sm.ylo[tid] = (int16_t)(w.ylo - g.uy0); sm.yhi[tid] = (int16_t)(w.yhi - g.uy0);
    } else {
      sm.factor[tid] = 0; sm.xlo[tid] = 0; sm.xhi[tid] = -1; sm.ylo[tid] = 0; sm.yhi[tid] = -1;
    }
  }

  // ---- prologue 2: the four byte-shifted copies of the reference window ----
  {
    const int px = kMargin + 16 * g.mb_x + g.ux0;               // window origin in plane coordinates
    const int py = kMargin + 16 * g.mb_y + g.uy0;
    const int a16 = px & 15;
    const uint8_t* src0 = planes + (size_t)g.ref_slot * plane_bytes + (size_t)py * plane_stride + (px - a16);   // 16-byte aligned
    constexpr int kGroups = (kWinRowWords + 4 + 3) / 4;         // 11 groups of 4 source words cover every destination word
    for (int i = tid; i < rows * kGroups; i += kK2Threads) {
      const int r = i / kGroups, gq = i - r * kGroups;
      const uint8_t* sp = src0 + (size_t)r * plane_stride + 16 * gq;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(sp));
      const uint32_t w4 = __ldg(reinterpret_cast<const uint32_t*>(sp + 16));
      const uint32_t w[5] = { v.x, v.y, v.z, v.w, w4 };
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int sh = a16 + s, q = sh >> 2, rb = (sh & 3) * 8;
        uint32_t* dst = winw + s * copy_words + r * kWinRowWords;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int m = 4 * gq + k - q;                         // destination word of copy s
          if (m >= 0 && m < kWinRowWords) dst[m] = __funnelshift_r(w[k], w[k + 1], rb);
        }
      }
    }
  }

  // ---- prologue 3: the current macroblock into registers (16 rows x 4 packed words) ----
  uint32_t cur[16][4];
  if (g.org_index < 0) {
    const uint8_t* cp = planes + (size_t)g.cur_slot * plane_bytes + (size_t)(kMargin + 16 * g.mb_y) * plane_stride + kMargin + 16 * g.mb_x;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(cp + (size_t)r * plane_stride));
      cur[r][0] = v.x; cur[r][1] = v.y; cur[r][2] = v.z; cur[r][3] = v.w;
    }
  } else {
    const int16_t* op = org_mbs + (size_t)g.org_index * 256;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const short4 v = __ldg(reinterpret_cast<const short4*>(op + r * 16 + 4 * k));
        cur[r][k] = (uint32_t)(v.x & 0xff) | ((uint32_t)(v.y & 0xff) << 8) | ((uint32_t)(v.z & 0xff) << 16) | ((uint32_t)(v.w & 0xff) << 24);
      }
    }
  }
  __syncthreads();            // slot parameters visible

  // ---- prologue 4: MV-rate tables and validity masks ----
  for (int i = tid; i < NS * by_stride; i += kK2Threads) {
    const int sl = i / by_stride, y = i - sl * by_stride;
    uint32_t b = 0;
    // xGetBits with m_uiMvScaleShift = 2: table index (y << 2) - pred (N6)
    if (sm.job[sl] >= 0 && y < uh) b = mv_component_bits(((g.uy0 + y) << 2) - jobs[sm.job[sl]].pred.ver);
    by_tab[sl * by_stride + y] = b;
  }
  for (int i = tid; i < NS * 4 * kMaxM; i += kK2Threads) {
    const int sl = i / (4 * kMaxM), r = i - sl * 4 * kMaxM;
    const int s = r / kMaxM, m = r - s * kMaxM;
    const int x = 4 * m + s;
    uint32_t b = 0;
    if (sm.job[sl] >= 0 && x < uw) b = mv_component_bits(((g.ux0 + x) << 2) - jobs[sm.job[sl]].pred.hor);
    sm.bx[sl][s * kMaxM + m] = (uint8_t)b;
  }
  for (int i = tid; i < 4 * kMaxM; i += kK2Threads) {
    const int s = i / kMaxM, m = i - s * kMaxM;
    const int x = 4 * m + s;
    Mask mk = 0;
    for (int sl = 0; sl < NS; ++sl) if (x >= sm.xlo[sl] && x <= sm.xhi[sl]) mk |= (Mask)1 << sl;
    sm.xmask[i] = mk;
  }
  for (int i = tid; i < SM::kMaxStrips; i += kK2Threads) {
    const int y0 = i * V, y1 = y0 + V - 1;
    Mask full = 0, any = 0;
    for (int sl = 0; sl < NS; ++sl) {
      if (sm.xhi[sl] < sm.xlo[sl]) continue;
      if (sm.ylo[sl] <= y0 && y1 <= sm.yhi[sl]) full |= (Mask)1 << sl;
      if (sm.ylo[sl] <= y1 && y0 <= sm.yhi[sl]) any |= (Mask)1 << sl;
    }
    sm.yfull[i] = full; sm.yany[i] = any;
  }
  __syncthreads();            // window copies and tables visible

  uint32_t best[NS];
#pragma unroll
  for (int sl = 0; sl < NS; ++sl) best[sl] = kKeyInvalid;

  const int units_per_strip = 4 * M;
  const int nunits = nstrips * units_per_strip;

  for (int u = tid; u < nunits; u += kK2Threads) {
    const int strip = u / units_per_strip;
    const int rem = u - strip * units_per_strip;
    const int s = rem / M, m = rem - s * M;
    const int x = 4 * m + s;
    const int ys = strip * V;
    const int cidx = s * kMaxM + m;

    // ---- SAD pass over the strip ----
    uint32_t acc[V][NACC];
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
      for (int a = 0; a < NACC; ++a) acc[v][a] = 0;

    const uint32_t* rp = winw + s * copy_words + ys * kWinRowWords + m;
#pragma unroll
    for (int r = 0; r < 15 + V; ++r) {
      const uint32_t w0 = rp[r * kWinRowWords + 0], w1 = rp[r * kWinRowWords + 1];
      const uint32_t w2 = rp[r * kWinRowWords + 2], w3 = rp[r * kWinRowWords + 3];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int c = r - v;                    // current-MB row matched by this reference row at displacement ys+v
        if (c < 0 || c > 15) continue;
        if (GRAN4) {
          const int a = (c >> 2) * 4;
          acc[v][a + 0] = sad4(cur[c][0], w0, acc[v][a + 0]);
          acc[v][a + 1] = sad4(cur[c][1], w1, acc[v][a + 1]);
          acc[v][a + 2] = sad4(cur[c][2], w2, acc[v][a + 2]);
          acc[v][a + 3] = sad4(cur[c][3], w3, acc[v][a + 3]);
        } else {
          const int q = (c >> 3) * 2;
          acc[v][q + 0] = sad4(cur[c][0], w0, acc[v][q + 0]);
          acc[v][q + 0] = sad4(cur[c][1], w1, acc[v][q + 0]);
          acc[v][q + 1] = sad4(cur[c][2], w2, acc[v][q + 1]);
          acc[v][q + 1] = sad4(cur[c][3], w3, acc[v][q + 1]);
        }
      }
    }

    // ---- per partition: SAD + rate, packed-key min ----
    const Mask mx = sm.xmask[cidx];
    const Mask mfull = mx & sm.yfull[strip], many = mx & sm.yany[strip];
    if (many == 0) continue;                                      // dead lane (x >= uw) or outside every job's window
    uint32_t idxv[V];
#pragma unroll
    for (int v = 0; v < V; ++v) idxv[v] = (uint32_t)((ys + v) * uw + x);

    auto update = [&](const int sl, const uint32_t (&sadv)[V]) {
      if (!((many >> sl) & 1)) return;
      const uint32_t f = sm.factor[sl];
      const uint32_t fbx = f * (uint32_t)sm.bx[sl][cidx];
      uint32_t byv[V];
#pragma unroll
      for (int v4 = 0; v4 < V; v4 += 4) {
        const uint4 q = *reinterpret_cast<const uint4*>(&by_tab[sl * by_stride + ys + v4]);   // warp-uniform address: broadcast LDS.128
        byv[v4] = q.x; byv[v4 + 1] = q.y; byv[v4 + 2] = q.z; byv[v4 + 3] = q.w;
      }
      uint32_t b = best[sl];
      if ((mfull >> sl) & 1) {
#pragma unroll
        for (int v = 0; v < V; ++v) b = min(b, rate_key(f, byv[v], fbx, sadv[v], idxv[v]));
      } else {
        const int lo = sm.ylo[sl] - ys, hi = sm.yhi[sl] - ys;       // valid strip rows: lo <= v <= hi
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const uint32_t key = rate_key(f, byv[v], fbx, sadv[v], idxv[v]);
          if (v >= lo && v <= hi) b = min(b, key);
        }
      }
      best[sl] = b;
    };

    if (!GRAN4) {
      uint32_t t0[V], t1[V];
      // 8x8 partitions read the accumulators directly
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int v = 0; v < V; ++v) t0[v] = acc[v][q];
        update(5 + q, t0);
      }
      // 16x8 (upper, lower) and, from their sum, 16x16
#pragma unroll
      for (int v = 0; v < V; ++v) { t0[v] = acc[v][0] + acc[v][1]; t1[v] = acc[v][2] + acc[v][3]; }
      update(1, t0); update(2, t1);
#pragma unroll
      for (int v = 0; v < V; ++v) t0[v] += t1[v];
      update(0, t0);
      // 8x16 (left, right)
#pragma unroll
      for (int v = 0; v < V; ++v) { t0[v] = acc[v][0] + acc[v][2]; t1[v] = acc[v][1] + acc[v][3]; }
      update(3, t0); update(4, t1);
    } else {
#pragma unroll
      for (int sl = 0; sl < NS; ++sl) {
        const uint32_t mask = slot_mask_gran4(sl);                 // folds after unrolling
        uint32_t t0[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          uint32_t sad = 0;
#pragma unroll
          for (int a = 0; a < NACC; ++a) if (mask & (1u << a)) sad += acc[v][a];
          t0[v] = sad;
        }
        update(sl, t0);
      }
    }
  }

  // ---- CTA reduction: warp min, then shared atomicMin, then one store per job ----
#pragma unroll
  for (int sl = 0; sl < NS; ++sl) {
    const uint32_t w = __reduce_min_sync(0xffffffffu, best[sl]);
    if ((tid & 31) == 0 && w != kKeyInvalid) atomicMin(&sm.best[sl], w);
  }
  __syncthreads();
  if (tid < NS && sm.job[tid] >= 0) {
    keys[sm.job[tid]] = sm.best[tid];
    JobAux a; a.ux0 = g.ux0; a.uy0 = g.uy0; a.uw = g.uw; a.pad = 0;
    aux[sm.job[tid]] = a;
  }
}

template <int V, bool GRAN4>
inline size_t k2_smem_bytes(int max_uh) {
  using SM = K2Smem<V, GRAN4>;
  const size_t win = (((size_t)4 * (k2_rows(max_uh) * kWinRowWords + 31) * 4) + 15) & ~(size_t)15;
  const size_t by = (size_t)SM::NS * (((max_uh + V - 1) / V) * V) * 4;
  return win + ((sizeof(SM) + 15) & ~(size_t)15) + by;
}

}  // namespace jsvm
