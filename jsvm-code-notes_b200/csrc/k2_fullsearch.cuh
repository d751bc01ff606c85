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

#ifndef K2_THREADS
#define K2_THREADS 256              // threads of the one-CTA-per-group kernel
#endif
#ifndef K2_COPY_UNROLL
#define K2_COPY_UNROLL 4            // window-copy items a staging thread keeps in flight
#endif
constexpr int kK2Threads   = K2_THREADS;
#ifndef K2_PRUNE
#define K2_PRUNE 1
#endif
#ifdef K2_PRUNE_STATS
__device__ unsigned long long g_prune_stats[4];     // batches, batches with no live slot, live slots, slots inside the window
#endif
constexpr int kWinRowBytes = 160;                 // bytes per staged window row (max 145 positions + 15)
constexpr int kWinRowWords = kWinRowBytes / 4;    // 40 words: 40 mod 32 = 8 keeps successive rows on different banks
constexpr int kMaxUW       = 145;                 // widest union window handled by one CTA
constexpr int kMaxM        = 38;                  // column groups per copy: (145 + 3 alignment columns + 3) / 4 = 37, rounded up to even
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
// the same for waiters that have nothing urgent to do (the staging warps of the persistent kernel): they must not steal issue
// slots from the search warps.  try_wait with a suspend-time hint parks the warp in hardware until the phase completes or the
// hint elapses; the nanosleep between attempts covers implementations that return early.  (Round 1 polled with nanosleep(200),
// which returned after ~30 ns: 82 warp-instructions of polling per search-loop iteration, 6 % of all issue slots.)
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  for (;;) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"(100000u) : "memory");
    if (done) break;
    __nanosleep(2000);
  }
}
// shared-memory atomics issued by ONE lane as plain PTX: atomicAdd / atomicMin in a lane-0 branch make the compiler emit its
// warp-aggregation sequence (VOTE, FLO, POPC, S2R LTMASK, SHFL: ~14 instructions per call) around a single active lane
__device__ __forceinline__ int smem_fetch_add(int32_t* p, int v) {
  int old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void smem_min(uint32_t* p, uint32_t v) {
  asm volatile("red.shared.min.u32 [%0], %1;" :: "r"(smem_u32(p)), "r"(v) : "memory");
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
//   cost = sad + (t >> 16)       xGetCost, ENC/MotionEstimationCost.h:27; one LEA.HI on the ALU pipe
//   key  = cost * 32768 + idx    FMA pipe (IMAD)
// Written as PTX so that the compiler cannot re-associate (cost << 15) into one multiply per SAD term.
// (Measured on B200: the shift-add as hi32(t * 65536) + sad -- IMAD.HI with the constant hidden in a kernel argument, which
// keeps it on the FMA pipe -- is 16 % slower: IMAD.HI wants a 64-bit addend pair and runs at a fraction of the IMAD rate.)
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
  // Running minima a thread carries through a group.  8x8 granularity: one register per slot (per-thread minima, reduced
  // over the warp once per group).  4x4 granularity (41 slots): the minimum of every unit is reduced over the warp at once
  // (REDUX) and lane l keeps the warp's minima of slots l and l + 32 -- 2 registers instead of 41.
  static constexpr int NBEST = GRAN4 ? (NS + 31) / 32 : NS;
  typedef typename std::conditional<GRAN4, unsigned long long, uint32_t>::type Mask;   // one bit per partition slot
  // one stage of dynamic shared memory: the four window copies come first, then this record, then the by table
  uint32_t cur[16][4];                            // the current macroblock, 16 packed bytes per row (16-byte aligned)
  uint32_t factor[NS];
  int16_t  xlo[NS], xhi[NS], ylo[NS], yhi[NS];    // relative to the union window origin (empty slot: xhi < xlo)
  int16_t  ph[NS], pv[NS];                        // rcMvPred of the slot's job (quarter pel)
  uint32_t inv[NS];                               // "bits" value of rows outside the job's window (0: exact row-by-row pass instead)
  int32_t  job[NS];                               // absolute job index or -1
  int32_t  next_unit;                             // work counter: warps pull batches of 32 units
  int16_t  ux0, uy0, uw, uh, xa, pad;             // geometry of the staged group (persistent kernel: consumers read it here)
  uint32_t best[NS];
  uint8_t  bx[NS][4 * kMaxM + 12];                // bits of the horizontal difference, permuted [s][m]
  Mask     xmask[4 * kMaxM];                      // slots whose window contains column x, permuted [s][m]
  Mask     yfull[kMaxStrips];                     // slots whose window contains ALL rows of the strip
  Mask     yany[kMaxStrips];                      // slots whose window contains SOME row of the strip
  // followed by: uint32_t by[NS][nstrips * V]    bits of the vertical MV difference per window row (16-byte aligned)
};

// Reference rows staged per copy: only the uh + 15 rows that valid displacements touch.  The last strip may reach up to
// V-1 rows further; those reads belong to displacements outside every window (masked by yany / yfull), land in the
// shared memory that follows the copies (by table, slack) and are never used.
__host__ __device__ inline int k2_rows(int uh) { return uh + 15; }
// words between two window copies: padded so that (copy stride mod 32) == (M rounded up to even) mod 32 -- a warp whose
// lanes run off the end of copy s (m = M-1) continues in copy s+1 (m = 0) on the NEXT bank: no shared-memory conflicts.
__host__ __device__ inline int k2_copy_words(int rows, int m_units) {
  const int base = rows * kWinRowWords;
  return base + (((m_units - base) % 32) + 32) % 32;
}

// Geometry of one staged group.  The staged window starts on a 4-byte boundary of the reference plane: `xa` (0..3) extra
// columns on the left; column x' of the staged window is displacement ux0 - xa + x' (the extra columns belong to no window).
template <int V>
struct K2Geo {
  int xa, uw, uh, uwx, nstrips, rows, M, copy_words, by_stride;
  __device__ __forceinline__ K2Geo(int xa_, int uw_, int uh_) {
    xa = xa_; uw = uw_; uh = uh_; uwx = uw + xa;
    nstrips = (uh + V - 1) / V;
    rows = k2_rows(uh);
    M = ((uwx + 3) >> 2) + (((uwx + 3) >> 2) & 1);                // column groups per copy, rounded up to even
    copy_words = k2_copy_words(rows, M);
    by_stride = nstrips * V;
  }
};

// One stage of dynamic shared memory: [K2Smem record][four window copies][by table][slack].  The record sits at a fixed
// offset so that its position does not move with the size of the staged group (the persistent kernel writes a finished
// group's keys out of the record while the next group's window is already being staged behind it).
template <int V, bool GRAN4>
struct K2Stage {
  using SM = K2Smem<V, GRAN4>;
  static constexpr size_t kRecBytes = (sizeof(SM) + 127) & ~(size_t)127;
  uint32_t* winw; SM* sm; uint32_t* by_tab;
  __device__ __forceinline__ K2Stage(uint8_t* base, const K2Geo<V>& geo) {
    sm = reinterpret_cast<SM*>(base);
    winw = reinterpret_cast<uint32_t*>(base + kRecBytes);
    by_tab = reinterpret_cast<uint32_t*>(base + kRecBytes + (((size_t)4 * geo.copy_words * 4 + 15) & ~(size_t)15));
  }
};

// ---- staging of one group: job parameters, the four window copies, the current macroblock, rate tables, masks ----
// Executed by kThreads threads (whole warps; ptid = index among them), which meet once at named barrier 1: the first
// warp derives the per-slot job parameters (lane l owns slots l and l + 32) and stores them in the stage record; after the
// barrier every thread builds its share of the (slot, row) / (slot, column) table entries from the record.
template <int V, bool GRAN4, int kThreads>
__device__ __forceinline__ void k2_build_stage(const Group& g, const K2Geo<V>& geo, const K2Stage<V, GRAN4>& st,
                                               const uint8_t* __restrict__ planes, size_t plane_bytes, int plane_stride,
                                               const jsvm_me_job* __restrict__ jobs, const int16_t* __restrict__ org_mbs,
                                               int mb_w, int mb_h, int default_range,
                                               const int ptid, const int first_unit)
{
  using SM = K2Smem<V, GRAN4>;
  using Mask = typename SM::Mask;
  constexpr int NS = SM::NS;
  constexpr int pthreads = kThreads;
  SM& sm = *st.sm;
  uint32_t* winw = st.winw;
  uint32_t* by_tab = st.by_tab;
  const int xa = geo.xa, uwx = geo.uwx, rows = geo.rows, copy_words = geo.copy_words, by_stride = geo.by_stride, nstrips = geo.nstrips;
  const int px = kMargin + 16 * g.mb_x + g.ux0;                 // union window origin in plane coordinates
  const int py = kMargin + 16 * g.mb_y + g.uy0;

  // ---- 1a: the first warp fetches the job records of its slots (lane l owns slots l and l + 32); they are turned into
  //      per-slot parameters in step 1b, after the window copies have been issued (both are chains of L2 round trips) ----
  constexpr int kSlotRegs = (NS + 31) / 32;
  jsvm_me_job slot_jobs[kSlotRegs];
  int slot_jo[kSlotRegs];
#pragma unroll
  for (int q = 0; q < kSlotRegs; ++q) {
    const int sl = ptid + 32 * q;
    slot_jo[q] = (ptid < 32 && sl < NS) ? (int)g.slot_job[sl] : -1;
    if (slot_jo[q] >= 0) slot_jobs[q] = jobs[g.first_job + slot_jo[q]];
  }
  if (ptid == 0) {
    sm.next_unit = first_unit;
    sm.ux0 = g.ux0; sm.uy0 = g.uy0; sm.uw = g.uw; sm.uh = g.uh; sm.xa = (int16_t)xa;
    // groups that hold a single 8x8 job (the picture-level search, k6_picture.cuh): which quadrant (read by the QUAD search loop)
    int quad = 0;
    if (!GRAN4) for (int q = 0; q < 4; ++q) if (g.slot_job[5 + q] >= 0) quad = q;
    sm.pad = (int16_t)quad;
  }

  // ---- 2: the four byte-shifted copies of the reference window ----
  // item = (row, 4 destination words): 5 aligned source words -> 4 x 4 funnel-shifted words -> two 8-byte stores per copy
  {
    const uint32_t* src0 = reinterpret_cast<const uint32_t*>(
        planes + (size_t)g.ref_slot * plane_bytes + (size_t)py * plane_stride + (px - xa));
    const int stride_w = plane_stride >> 2;
    constexpr int kGroups = kWinRowWords / 4;                   // 10
    // K2_COPY_UNROLL items per thread in flight: the loop is a chain of L2 round trips (the staging warps of the persistent
    // kernel have to finish a window while the search warps work through the previous group)
    const int n_items = rows * kGroups;
    for (int i0 = ptid; i0 < n_items; i0 += K2_COPY_UNROLL * pthreads) {
      uint32_t w[K2_COPY_UNROLL][5];
#pragma unroll
      for (int u = 0; u < K2_COPY_UNROLL; ++u) {
        const int i = min(i0 + u * pthreads, n_items - 1);       // clamped: a few redundant loads instead of a branch
        const int r = i / kGroups, gq = i - r * kGroups;
        const uint32_t* sp = src0 + (size_t)r * stride_w + 4 * gq;
#pragma unroll
        for (int k = 0; k < 5; ++k) w[u][k] = __ldg(sp + k);
      }
#pragma unroll
      for (int u = 0; u < K2_COPY_UNROLL; ++u) {
        const int i = i0 + u * pthreads;
        if (i < n_items) {
          const int r = i / kGroups, gq = i - r * kGroups;
#pragma unroll
          for (int sft = 0; sft < 4; ++sft) {
            // copy_words is even (M is), not a multiple of 4: 8-byte stores
            uint2* dst = reinterpret_cast<uint2*>(winw + sft * copy_words + r * kWinRowWords + 4 * gq);
            dst[0] = make_uint2(__funnelshift_r(w[u][0], w[u][1], 8 * sft), __funnelshift_r(w[u][1], w[u][2], 8 * sft));
            dst[1] = make_uint2(__funnelshift_r(w[u][2], w[u][3], 8 * sft), __funnelshift_r(w[u][3], w[u][4], 8 * sft));
          }
        }
      }
    }
  }

  // ---- 3: the current macroblock (16 rows x 4 packed words); every unit re-reads it with 16 broadcast LDS.128 so that its
  //      64 registers are free during the rate / argmin pass ----
  for (int t = ptid; t < 64; t += pthreads) {
    const int r = t >> 2, k = t & 3;
    uint32_t v;
    if (g.org_index < 0) {
      v = __ldg(reinterpret_cast<const uint32_t*>(planes + (size_t)g.cur_slot * plane_bytes + (size_t)(kMargin + 16 * g.mb_y + r) * plane_stride +
                                                  kMargin + 16 * g.mb_x + 4 * k));
    } else {
      // signed 16-bit originals (residual-prediction passes, N3) are searched as their [0,255] clip: the reference
      // samples lie in [0,255], so |org - ref| = |clip(org) - ref| + |org - clip(org)| and the second term does not
      // depend on the displacement -- same argmin, same tie-break; K3 adds the constant back to the SAD it reports
      v = pack_clip8(__ldg(reinterpret_cast<const short4*>(org_mbs + (size_t)g.org_index * 256 + r * 16 + 4 * k)));
    }
    sm.cur[r][k] = v;
  }

  // ---- 1b: per-slot job parameters (first warp) ----
  if (ptid < 32) {
#pragma unroll
    for (int q = 0; q < kSlotRegs; ++q) {
      const int sl = ptid + 32 * q;
      if (sl < NS) {
        const int jo = slot_jo[q];
        int xlo = 0, xhi = -1, ylo = 0, yhi = -1, ph = 0, pv = 0;
        uint32_t fac = 0, inv = 0;
        if (jo >= 0) {
          const jsvm_me_job& j = slot_jobs[q];
          const Window w = job_window(j, mb_w, mb_h, default_range);
          fac = j.cost_factor;
          xlo = w.xlo - g.ux0 + xa; xhi = w.xhi - g.ux0 + xa;         // relative to the staged window origin
          ylo = w.ylo - g.uy0;      yhi = w.yhi - g.uy0;
          ph = j.pred.hor; pv = j.pred.ver;
          // Rows of the union window outside this job's own window are neutralised through the rate table: their "bits" entry
          // makes (f * bits) >> 16 >= kRowOutCost, more than any valid cost of a partition of at most 128 samples
          // (SAD <= 32640, rate <= 62 bits * 2^22.5 / 65536 < 6000) and small enough for the 17-bit cost field
          // (65280 + 40001 + 6000 < 2^17) and for f * bits + f * bits_x to stay below 2^32.  Not for the 16x16 partition
          // (its valid costs reach 71000), not for f == 0 and not for factors beyond the SAD range (f > 7e6: QP 51 gives 5.5e6; an
          // SSE-style factor could make the rate itself wrap, N5): those slots keep the exact row-by-row pass.
          if (sl != 0 && fac != 0 && fac <= 7000000u) inv = (kRowOutCost * 65536u + fac - 1u) / fac;
        }
        sm.job[sl] = jo < 0 ? -1 : g.first_job + jo;
        sm.best[sl] = kKeyInvalid;
        sm.factor[sl] = fac;
        sm.xlo[sl] = (int16_t)xlo; sm.xhi[sl] = (int16_t)xhi;
        sm.ylo[sl] = (int16_t)ylo; sm.yhi[sl] = (int16_t)yhi;
        sm.ph[sl] = (int16_t)ph;   sm.pv[sl] = (int16_t)pv;
        sm.inv[sl] = inv;
      }
    }
  }

  asm volatile("bar.sync 1, %0;" :: "n"(kThreads) : "memory");    // the per-slot parameters of step 1b are visible

  // ---- 4: per staged column: bits of the horizontal difference of every slot, and the column / strip validity masks.
  //      A thread owns a column (no index arithmetic inside the slot loop); the slot parameters come from the record. ----
  for (int ci = ptid; ci < 4 * kMaxM; ci += pthreads) {
    const int sft = ci / kMaxM, m = ci - sft * kMaxM, x = 4 * m + sft;     // staged column of entry [s][m]
    const int dx4 = (g.ux0 - xa + x) << 2;                                  // quarter-pel displacement of the column
    const int y0 = ci * V, y1 = y0 + V - 1;                                 // entry ci doubles as strip index for the row masks
    Mask mk = 0, full = 0, any = 0;
#pragma unroll 1
    for (int sl = 0; sl < NS; ++sl) {
      const int xlo = sm.xlo[sl], xhi = sm.xhi[sl];
      // xGetBits with m_uiMvScaleShift = 2: table index (x << 2) - pred (N6)
      sm.bx[sl][ci] = (xhi >= xlo && x < uwx) ? (uint8_t)mv_component_bits(dx4 - sm.ph[sl]) : (uint8_t)0;
      if (x >= xlo && x <= xhi) mk |= (Mask)1 << sl;
      if (ci < nstrips && xhi >= xlo) {
        const int ylo = sm.ylo[sl], yhi = sm.yhi[sl];
        const bool some = ylo <= y1 && y0 <= yhi;
        // "full": the straight-line pass may commit all V rows of the strip -- every row is inside the job's window, or
        // the rows outside are neutralised through the rate table
        if ((ylo <= y0 && y1 <= yhi) || (some && sm.inv[sl] != 0)) full |= (Mask)1 << sl;
        if (some) any |= (Mask)1 << sl;
      }
    }
    sm.xmask[ci] = mk;
    if (ci < nstrips) { sm.yfull[ci] = full; sm.yany[ci] = any; }
  }
  // ---- 5: bits of the vertical difference, one (slot, row) entry per thread and pass; rows outside the job's window: step 1b ----
  {
    const uint32_t inv_bs = 0xffffffffu / (uint32_t)by_stride + 1u;      // ceil(2^32 / d), d >= 2: exact division of i < 2^16
    for (int i = ptid; i < NS * by_stride; i += pthreads) {
      const int sl = (int)__umulhi((uint32_t)i, inv_bs), y = i - sl * by_stride;
      uint32_t b = 0u;
      if (sm.job[sl] >= 0) b = (y >= sm.ylo[sl] && y <= sm.yhi[sl]) ? mv_component_bits(((g.uy0 + y) << 2) - sm.pv[sl]) : sm.inv[sl];
      by_tab[i] = b;
    }
  }
}

// ---- the search loop of one warp over one staged group: batches of 32 units (unit = column x strip of V displacements)
//      pulled from the stage's counter, always one batch ahead of the one being worked on ----
// QUAD: the group holds ONE job, an 8x8 block (slot 5 + quadrant): the SAD pass covers only the 8 rows x 2 words of that quadrant
// and the rate / argmin pass only that slot -- a quarter of the SAD work and a ninth of the slot work of the general loop.
// PRUNE: skip a slot's rate / key code when no SAD of the batch can beat the stage's incumbent (see live_slots).  Off for the
// 4x4-granular groups (small partitions have SADs below their rate-dominated incumbents almost everywhere: 36 % slower on C3)
// and for the small-window form (every displacement is near the winner).
template <int V, bool GRAN4, bool QUAD = false, bool PRUNE = (K2_PRUNE != 0) && !GRAN4>
__device__ __forceinline__ void k2_search_units(const K2Geo<V>& geo, const K2Stage<V, GRAN4>& st, const int lane, const int first_u0,
                                                uint32_t (&best)[K2Smem<V, GRAN4>::NBEST])
{
  using SM = K2Smem<V, GRAN4>;
  using Mask = typename SM::Mask;
  constexpr int NACC = SM::NACC;
  SM& sm = *st.sm;
  const uint32_t* winw = st.winw;
  const uint32_t* by_tab = st.by_tab;
  const int xa = geo.xa, uw = geo.uw, M = geo.M, copy_words = geo.copy_words, by_stride = geo.by_stride;

  const int units_per_strip = 4 * M;
  const int nunits = geo.nstrips * units_per_strip;
  // exact division of u < 2^16 by d < 2^16: (u * ceil(2^32 / d)) >> 32
  const uint32_t inv_ups = 0xffffffffu / (uint32_t)units_per_strip + 1u;     // = ceil(2^32 / d) for d >= 2
  const uint32_t inv_m = 0xffffffffu / (uint32_t)M + 1u;

  int u0 = first_u0, u0_next = 0;
  if (first_u0 < 0) {                                             // no static first batch: fetch it
    if (lane == 0) u0 = smem_fetch_add(&sm.next_unit, 32);
    u0 = __shfl_sync(0xffffffffu, u0, 0);
  }
  if (lane == 0) u0_next = smem_fetch_add(&sm.next_unit, 32);
  for (;;) {
    if (u0 >= nunits) break;
    const int fetched = __shfl_sync(0xffffffffu, u0_next, 0);
    if (lane == 0) u0_next = smem_fetch_add(&sm.next_unit, 32);   // prefetch the batch after the next one
    const bool live = u0 + lane < nunits;
    const int u = live ? u0 + lane : nunits - 1;                  // idle lanes redo the last unit with every slot masked off
    u0 = fetched;
    const int strip_k = (int)__umulhi((uint32_t)u, inv_ups);
    const int rem = u - strip_k * units_per_strip;
    // PRUNE: strips are visited from the middle of the union window outwards (k = 0, 1, 2, .. -> mid, mid + 1, mid - 1, ..): the
    // winners sit near the start vectors, so the bounds below are tight after the first round of batches
    const int strip = PRUNE ? ((geo.nstrips - 1) >> 1) + ((strip_k & 1) ? ((strip_k + 1) >> 1) : -(strip_k >> 1)) : strip_k;
    const int s = (int)__umulhi((uint32_t)rem, inv_m), m = rem - s * M;
    const int x = 4 * m + s - xa;                                 // column inside the union window (negative: alignment columns)
    const int ys = strip * V;
    const int cidx = s * kMaxM + m;

    if constexpr (QUAD && !GRAN4) {
      const int quad = sm.pad, sl = 5 + quad;
      const Mask mxq = live ? sm.xmask[cidx] : (Mask)0;
      const bool fullq = ((mxq & sm.yfull[strip]) >> sl) & 1, anyq = ((mxq & sm.yany[strip]) >> sl) & 1;
      if (__all_sync(0xffffffffu, !anyq)) continue;
      uint32_t accq[V];
#pragma unroll
      for (int v = 0; v < V; ++v) accq[v] = 0;
      const int r0 = (quad >> 1) * 8, k0 = (quad & 1) * 2;
      const uint32_t* rq = winw + s * copy_words + (ys + r0) * kWinRowWords + m + k0;
      {
        uint32_t ref[V][2];
#pragma unroll
        for (int r = 0; r < V - 1; ++r) { ref[r][0] = rq[r * kWinRowWords]; ref[r][1] = rq[r * kWinRowWords + 1]; }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          ref[(c + V - 1) % V][0] = rq[(c + V - 1) * kWinRowWords]; ref[(c + V - 1) % V][1] = rq[(c + V - 1) * kWinRowWords + 1];
          const uint2 cq = *reinterpret_cast<const uint2*>(&sm.cur[r0 + c][k0]);
#pragma unroll
          for (int v = 0; v < V; ++v) accq[v] = sad4(cq.x, ref[(c + v) % V][0], accq[v]);
#pragma unroll
          for (int v = 0; v < V; ++v) accq[v] = sad4(cq.y, ref[(c + v) % V][1], accq[v]);
        }
      }
      const uint32_t f = sm.factor[sl];
      const uint32_t fbx = f * (uint32_t)sm.bx[sl][cidx];
      uint32_t kv[V];
#pragma unroll
      for (int v4 = 0; v4 < V; v4 += 4) {
        const uint4 q = *reinterpret_cast<const uint4*>(&by_tab[sl * by_stride + ys + v4]);
        const uint32_t byv[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
        for (int v = 0; v < 4; ++v) kv[v4 + v] = rate_key(f, byv[v], fbx, accq[v4 + v], (uint32_t)((ys + v4 + v) * uw + x));
      }
      uint32_t b = best[0];                                         // QUAD keeps its single running minimum in best[0]
      if (fullq) {
#pragma unroll
        for (int v = 0; v < V; ++v) b = min(b, kv[v]);
      } else if (anyq) {
        const int lo = sm.ylo[sl] - ys, hi = sm.yhi[sl] - ys;
#pragma unroll
        for (int v = 0; v < V; ++v) if (v >= lo && v <= hi) b = min(b, kv[v]);
      }
      best[0] = b;
      continue;
    }

    // ---- SAD pass over the strip ----
    // Row c of the current macroblock (one broadcast LDS.128) meets the V reference rows c .. c+V-1, which sit in a sliding
    // register window (one new row = 4 LDS.32 per step): 16 x 4 x V VABSDIFF4.U8.ACC, 36 live data registers besides the
    // accumulators.  Word order 0, 2, 1, 3: the two updates of one accumulator are 2V instructions apart.
    uint32_t acc[V][NACC];
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
      for (int a = 0; a < NACC; ++a) acc[v][a] = 0;

    const uint32_t* rp = winw + s * copy_words + ys * kWinRowWords + m;
    {
      uint32_t ref[V][4];
#pragma unroll
      for (int r = 0; r < V - 1; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) ref[r][k] = rp[r * kWinRowWords + k];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
#pragma unroll
        for (int k = 0; k < 4; ++k) ref[(c + V - 1) % V][k] = rp[(c + V - 1) * kWinRowWords + k];
        const uint4 cq = *reinterpret_cast<const uint4*>(&sm.cur[c][0]);
        const uint32_t cw[4] = { cq.x, cq.y, cq.z, cq.w };
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int k = GRAN4 ? kk : ((kk & 1) * 2 + (kk >> 1));      // 8x8 granularity: 0, 2, 1, 3
          const int a = GRAN4 ? (c >> 2) * 4 + k : (c >> 3) * 2 + (k >> 1);
#pragma unroll
          for (int v = 0; v < V; ++v) acc[v][a] = sad4(cw[k], ref[(c + v) % V][k], acc[v][a]);
        }
      }
    }

    // ---- per partition: SAD + rate, packed-key min ----
    // Pass 1 is straight-line code over all partition slots (no branches, so the loads of the next slot and the 8 x NS
    // independent rate / key chains overlap): every strip row is evaluated and the result is committed only for slots
    // whose window holds this column and ALL rows of the strip.  Pass 2 (rare: the strips that straddle the top or
    // bottom edge of some job's window) re-evaluates those slots row by row.
    const Mask mx = live ? sm.xmask[cidx] : (Mask)0;
    const Mask mfull = mx & sm.yfull[strip], many = mx & sm.yany[strip];
    if (__all_sync(0xffffffffu, many == 0)) continue;               // outside every job's window (alignment columns, x >= uw)
    const Mask part = many & ~mfull;                                // slots with only SOME valid rows in this strip
    // cost = SAD + rate >= SAD, and a key replaces the incumbent only with cost <= the incumbent's: a slot whose SADs over the
    // strip all exceed the cost of the stage's incumbent (a candidate some warp of the CTA has evaluated) cannot win -- exact,
    // independent of the order in which warps see the displacements.  Lower bounds of the partition SADs
    // come from the minima of the four quadrant SADs over the strip (min(a + b) >= min a + min b); each lane sets a bit per slot
    // that can still change, one warp-wide OR makes them uniform, and pass 1 runs only for those slots -- on textured content
    // for few of them.
    uint32_t live_slots = 0xffffffffu;
    if constexpr (PRUNE && !GRAN4) {
      uint32_t q[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        q[a] = acc[0][a];
#pragma unroll
        for (int v = 1; v < V; ++v) q[a] = min(q[a], acc[v][a]);
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) q[a] <<= kIdxBits;              // in key units (SAD <= 65280: no overflow); lb <= key >> 15  <=>  lb << 15 <= key
      const uint32_t lb[9] = { q[0] + q[1] + q[2] + q[3], q[0] + q[1], q[2] + q[3], q[0] + q[2], q[1] + q[3], q[0], q[1], q[2], q[3] };
      uint32_t bits = 0;
#pragma unroll
      for (int sl = 0; sl < 9; ++sl) {
        // the stage's running minimum of the slot: all warps of the CTA fold their improvements in (below); a stale value is
        // only a looser bound
        const uint32_t incumbent = *reinterpret_cast<const volatile uint32_t*>(&sm.best[sl]);
        bits |= lb[sl] <= incumbent ? (1u << sl) : 0u;
      }
      live_slots = __reduce_or_sync(0xffffffffu, bits & (uint32_t)mfull);
#ifdef K2_PRUNE_STATS
      {
        const uint32_t inside = __reduce_or_sync(0xffffffffu, (uint32_t)mfull) & 0x1ffu;
        if (lane == 0) {
          atomicAdd(&g_prune_stats[0], 1ull);
          atomicAdd(&g_prune_stats[1], (unsigned long long)(live_slots == 0));
          atomicAdd(&g_prune_stats[2], (unsigned long long)__popc(live_slots));
          atomicAdd(&g_prune_stats[3], (unsigned long long)__popc(inside));
        }
      }
#endif
    }
    uint32_t idxv[V];
#pragma unroll
    for (int v = 0; v < V; ++v) idxv[v] = (uint32_t)((ys + v) * uw + x);

    auto slot_keys = [&](const int sl, const uint32_t (&sadv)[V], uint32_t (&kv)[V]) {
      const uint32_t f = sm.factor[sl];
      const uint32_t fbx = f * (uint32_t)sm.bx[sl][cidx];
      uint32_t byv[V];
#pragma unroll
      for (int v4 = 0; v4 < V; v4 += 4) {
        const uint4 q = *reinterpret_cast<const uint4*>(&by_tab[sl * by_stride + ys + v4]);   // warp-uniform address: broadcast LDS.128
        byv[v4] = q.x; byv[v4 + 1] = q.y; byv[v4 + 2] = q.z; byv[v4 + 3] = q.w;
      }
#pragma unroll
      for (int v = 0; v < V; ++v) kv[v] = rate_key(f, byv[v], fbx, sadv[v], idxv[v]);
    };
    auto update_full = [&](const int sl, const uint32_t (&sadv)[V]) {
      if constexpr (PRUNE && !GRAN4) { if (!((live_slots >> sl) & 1u)) return; }     // warp-uniform: see live_slots
      uint32_t kv[V];
      slot_keys(sl, sadv, kv);
      // two levels of three-input minima (VIMNMX3) instead of a serial chain through best[sl].  (A 4-instruction tree that
      // folds best[sl] into the last VIMNMX3 measured 2 % slower than this 5-instruction one.)
#pragma unroll
      for (int v = 0; v + 2 < V; v += 3) kv[v] = min(min(kv[v], kv[v + 1]), kv[v + 2]);
      uint32_t m;
      if (V == 8) m = min(min(kv[0], kv[3]), min(kv[6], kv[7]));
      else        m = min(min(kv[0], kv[3]), kv[V - 1]);
      if (!GRAN4) {
        if ((mfull >> sl) & 1) best[sl] = min(best[sl], m);
        if constexpr (PRUNE) {                                    // every warp's bound: fold the improvement into the stage now
          const uint32_t w = __reduce_min_sync(0xffffffffu, best[sl]);
          if (lane == 0 && w != kKeyInvalid) smem_min(&sm.best[sl], w);
        }
      } else {
        const uint32_t w = __reduce_min_sync(0xffffffffu, ((mfull >> sl) & 1) ? m : kKeyInvalid);
        if (lane == (sl & 31)) best[sl >> 5] = min(best[sl >> 5], w);
      }
    };
    auto update_partial = [&](const int sl, const uint32_t (&sadv)[V]) {
      uint32_t kv[V];
      slot_keys(sl, sadv, kv);
      const int lo = sm.ylo[sl] - ys, hi = sm.yhi[sl] - ys;         // valid strip rows: lo <= v <= hi
      uint32_t b = GRAN4 ? kKeyInvalid : best[GRAN4 ? 0 : sl];
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (v >= lo && v <= hi) b = min(b, kv[v]);
      if (!GRAN4) best[GRAN4 ? 0 : sl] = b;
      else {
        const uint32_t w = __reduce_min_sync(0xffffffffu, ((part >> sl) & 1) ? b : kKeyInvalid);
        if (lane == (sl & 31)) best[sl >> 5] = min(best[sl >> 5], w);
      }
    };

    // partition SADs from the accumulators (8x8 granularity: 5 adds per displacement for the 9-partition set)
    auto for_each_slot = [&](auto&& fn) {
      if (!GRAN4) {
        uint32_t t0[V], t1[V];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
          for (int v = 0; v < V; ++v) t0[v] = acc[v][q];
          fn(5 + q, t0);
        }
#pragma unroll
        for (int v = 0; v < V; ++v) { t0[v] = acc[v][0] + acc[v][1]; t1[v] = acc[v][2] + acc[v][3]; }
        fn(1, t0); fn(2, t1);                                       // 16x8 upper, lower
#pragma unroll
        for (int v = 0; v < V; ++v) t0[v] += t1[v];
        fn(0, t0);                                                  // 16x16
#pragma unroll
        for (int v = 0; v < V; ++v) { t0[v] = acc[v][0] + acc[v][2]; t1[v] = acc[v][1] + acc[v][3]; }
        fn(3, t0); fn(4, t1);                                       // 8x16 left, right
      } else {
        // 4x4 granularity: the 41 partition SADs are built hierarchically, one 8x8 quadrant at a time (25 adds per
        // displacement; few sums alive at once).  acc index of 4x4 block k of quadrant b: see slot_mask_gran4.
        uint32_t s88[4][V];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int a0 = (b >> 1) * 8 + (b & 1) * 2;                 // accumulators a0, a0+1 (upper), a0+4, a0+5 (lower)
          uint32_t t[V], up[V], lo[V];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int v = 0; v < V; ++v) t[v] = acc[v][a0 + (k >> 1) * 4 + (k & 1)];
            fn(25 + b * 4 + k, t);                                   // 4x4
          }
#pragma unroll
          for (int v = 0; v < V; ++v) { up[v] = acc[v][a0] + acc[v][a0 + 1]; lo[v] = acc[v][a0 + 4] + acc[v][a0 + 5]; }
          fn(9 + b * 2, up); fn(9 + b * 2 + 1, lo);                  // 8x4 upper, lower
#pragma unroll
          for (int v = 0; v < V; ++v) s88[b][v] = up[v] + lo[v];
          fn(5 + b, s88[b]);                                         // 8x8
#pragma unroll
          for (int v = 0; v < V; ++v) { up[v] = acc[v][a0] + acc[v][a0 + 4]; lo[v] = acc[v][a0 + 1] + acc[v][a0 + 5]; }
          fn(17 + b * 2, up); fn(17 + b * 2 + 1, lo);                // 4x8 left, right
        }
        uint32_t t0[V], t1[V];
#pragma unroll
        for (int v = 0; v < V; ++v) { t0[v] = s88[0][v] + s88[1][v]; t1[v] = s88[2][v] + s88[3][v]; }
        fn(1, t0); fn(2, t1);                                        // 16x8 upper, lower
#pragma unroll
        for (int v = 0; v < V; ++v) t0[v] += t1[v];
        fn(0, t0);                                                   // 16x16
#pragma unroll
        for (int v = 0; v < V; ++v) { t0[v] = s88[0][v] + s88[2][v]; t1[v] = s88[1][v] + s88[3][v]; }
        fn(3, t0); fn(4, t1);                                        // 8x16 left, right
      }
    };

    if (!PRUNE || live_slots != 0) for_each_slot(update_full);
    if (__any_sync(0xffffffffu, part != 0)) {
      // the keys are recomputed from an opaque copy of the column: keeping the 8 x NS keys of pass 1 alive for this rare
      // pass would spill them in the hot path
      int x2 = x;
      asm volatile("mov.b32 %0, %0;" : "+r"(x2));

#pragma unroll
      for (int v = 0; v < V; ++v) idxv[v] = (uint32_t)((ys + v) * uw + x2);
      for_each_slot([&](const int sl, const uint32_t (&sadv)[V]) {
        if (!GRAN4) { if ((part >> sl) & 1) update_partial(sl, sadv); }
        else if (__any_sync(0xffffffffu, (part >> sl) & 1)) update_partial(sl, sadv);     // warp-uniform: REDUX inside
      });
    }
  }
}

// fold a warp's minima into the stage: warp minimum of every slot (8x8 granularity) or the lane-distributed minima (4x4)
template <int V, bool GRAN4>
__device__ __forceinline__ void k2_warp_reduce(K2Smem<V, GRAN4>& sm, const int lane, const uint32_t (&best)[K2Smem<V, GRAN4>::NBEST]) {
  using SM = K2Smem<V, GRAN4>;
  if (!GRAN4) {
#pragma unroll
    for (int sl = 0; sl < SM::NBEST; ++sl) {
      const uint32_t w = __reduce_min_sync(0xffffffffu, best[sl]);
      if (lane == 0 && w != kKeyInvalid) smem_min(&sm.best[sl], w);
    }
  } else {
#pragma unroll
    for (int q = 0; q < SM::NBEST; ++q) {
      const int sl = lane + 32 * q;
      if (sl < SM::NS && best[q] != kKeyInvalid) smem_min(&sm.best[sl], best[q]);
    }
  }
}

// one key per job to HBM; K3 decodes it with the group's window geometry
template <int V, bool GRAN4>
__device__ __forceinline__ void k2_flush(const K2Smem<V, GRAN4>& sm, const int ptid, const int pthreads,
                                         uint32_t* __restrict__ keys, JobAux* __restrict__ aux) {
  for (int sl = ptid; sl < K2Smem<V, GRAN4>::NS; sl += pthreads) {
    const int job = sm.job[sl];
    if (job >= 0) {
      keys[job] = sm.best[sl];
      JobAux a; a.ux0 = sm.ux0; a.uy0 = sm.uy0; a.uw = sm.uw; a.pad = 0;
      aux[job] = a;
    }
  }
}

// ---- kernel A: one CTA per search group (small launches, 4x4-granular groups, very tall union windows) ----
// THREADS: 256 (2 CTAs per SM), or 128 with 4 CTAs per SM for small windows (search range <= ~24: a group is only ~6 batches of 32
// units, too little for 8 warps, and twice as many groups in flight hide each other's staging)
template <int V, bool GRAN4, bool QUAD = false, int THREADS = kK2Threads>
__global__ void __launch_bounds__(THREADS, 512 / THREADS)
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
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const Group& g = groups[blockIdx.x];
  if (g.n_jobs == 0) return;      // empty slot of a device-planned launch (k6_picture.cuh)
  const int tid = threadIdx.x, lane = tid & 31;
  const K2Geo<V> geo((kMargin + 16 * g.mb_x + g.ux0) & 3, g.uw, g.uh);
  const K2Stage<V, GRAN4> st(smem_raw, geo);

  // the first batch of every warp is static (u0 = 32 * warp), the counter starts behind them
  k2_build_stage<V, GRAN4, THREADS>(g, geo, st, planes, plane_bytes, plane_stride, jobs, org_mbs, mb_w, mb_h, default_range, tid, THREADS);
  __syncthreads();            // window copies and tables visible

  uint32_t best[SM::NBEST];
#pragma unroll
  for (int sl = 0; sl < SM::NBEST; ++sl) best[sl] = kKeyInvalid;
  k2_search_units<V, GRAN4, QUAD, (K2_PRUNE != 0) && !GRAN4 && THREADS == kK2Threads>(geo, st, lane, (tid >> 5) * 32, best);
  if constexpr (QUAD) {
    const uint32_t w = __reduce_min_sync(0xffffffffu, best[0]);
    if (lane == 0 && w != kKeyInvalid) smem_min(&st.sm->best[5 + st.sm->pad], w);
  } else {
    k2_warp_reduce<V, GRAN4>(*st.sm, lane, best);
  }
  __syncthreads();
  k2_flush<V, GRAN4>(*st.sm, tid, THREADS, keys, aux);
}

// ---- kernel B: persistent, one CTA per SM, staging / search warps over two shared-memory stages ----
// kK2PConsumers warps run the search loop on stage s while kK2PProducers warps stage the next group into stage s ^ 1
// (and write the finished group's keys to HBM), then join the search of the current group; stages are handed over with
// mbarriers.  No warp waits at a CTA-wide barrier, so the ALU pipe always has SAD work queued.
#ifndef K2P_CONSUMERS
#define K2P_CONSUMERS 12
#endif
#ifndef K2P_PRODUCERS
#define K2P_PRODUCERS 4
#endif
// (8x8-granular groups: 12 + 4; 4x4-granular groups stage 41 rate tables per group and search a smaller window: K2P_PRODUCERS_G4)
#ifndef K2P_PRODUCERS_G4
#define K2P_PRODUCERS_G4 5
#endif
constexpr int kK2PThreads = 32 * (K2P_CONSUMERS + K2P_PRODUCERS);       // 16 warps: 512 threads x 128 registers
template <bool GRAN4> struct K2PSplit {
  static constexpr int kProducers = GRAN4 ? K2P_PRODUCERS_G4 : K2P_PRODUCERS;
  static constexpr int kConsumers = kK2PThreads / 32 - kProducers;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

template <int V, bool GRAN4>
__global__ void __launch_bounds__(kK2PThreads, 1)
full_search_persistent_kernel(const uint8_t* __restrict__ planes, size_t plane_bytes, int plane_stride,
                              const Group* __restrict__ groups, int n_groups,
                              const jsvm_me_job* __restrict__ jobs,
                              const int16_t* __restrict__ org_mbs,
                              int mb_w, int mb_h, int default_range,
                              uint32_t* __restrict__ keys, JobAux* __restrict__ aux,
                              int stage_bytes)
{
  using SM = K2Smem<V, GRAN4>;
  constexpr int kK2PConsumers = K2PSplit<GRAN4>::kConsumers, kK2PProducers = K2PSplit<GRAN4>::kProducers;
    extern __shared__ __align__(128) uint8_t smem_raw[];
  // [0, 64): mbarriers full[2], done[2]; stages follow at 128
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* done = full + 2;
  // (stage addresses are computed from smem_raw at each use: an indexed pointer array makes the compiler fall back to
  //  generic loads instead of LDS)
#define K2P_STAGE(s_) (smem_raw + 128 + (size_t)(s_) * (size_t)stage_bytes)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_mine = ((int)blockIdx.x < n_groups) ? (n_groups - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;   // groups b, b + G, ...
  if (tid == 0) {
    mbar_init(&full[0], 32 * kK2PProducers); mbar_init(&full[1], 32 * kK2PProducers);
    mbar_init(&done[0], kK2PConsumers + kK2PProducers); mbar_init(&done[1], kK2PConsumers + kK2PProducers);   // one arrival per warp
    fence_barrier_init();
  }
  __syncthreads();

  // Every warp runs the same loop.  Iteration k: the staging warps first retire group k - 2 (wait until all sixteen warps have
  // folded their minima in, write its keys) and build the stage of group k; then EVERY warp searches -- the search warps group k,
  // the staging warps group k - 1, whose stage they finished an iteration ago and which the search warps are still working on:
  // staging takes a third of a group's search time, and for the rest of it the three staging warps pull batches from the same
  // counter instead of idling (with 13 + 3 and idle staging warps the ALU pipes of three of the four sub-cores were 78 % busy).
  const bool producer = warp >= kK2PConsumers;
  const int ptid = tid - 32 * kK2PConsumers;
  constexpr int pthreads = 32 * kK2PProducers;
  for (int k = 0; k < n_mine + 2; ++k) {
    if (producer) {
      const int s = k & 1, use = k >> 1;
      if (k >= 2) {
        // stage s was used by group k - 2: wait until every warp has folded its minima in, then write the keys out
        mbar_wait_relaxed(&done[s], (uint32_t)((use - 1) & 1));
        const Group& gp = groups[blockIdx.x + (size_t)(k - 2) * gridDim.x];
        const K2Geo<V> geop((kMargin + 16 * gp.mb_x + gp.ux0) & 3, gp.uw, gp.uh);
        const K2Stage<V, GRAN4> stp(K2P_STAGE(s), geop);
        if (ptid < 32) {                                      // the warp that re-initialises best[] in k2_build_stage
          k2_flush<V, GRAN4>(*stp.sm, ptid, 32, keys, aux);
          __syncwarp();
        }
      }
      if (k < n_mine) {
        const Group& g = groups[blockIdx.x + (size_t)k * gridDim.x];
        const K2Geo<V> geo((kMargin + 16 * g.mb_x + g.ux0) & 3, g.uw, g.uh);
        const K2Stage<V, GRAN4> st(K2P_STAGE(s), geo);
        k2_build_stage<V, GRAN4, pthreads>(g, geo, st, planes, plane_bytes, plane_stride, jobs, org_mbs, mb_w, mb_h, default_range, ptid, 0);
        mbar_arrive(&full[s]);                                // release: this thread's stage writes are visible to waiters
      }
    }
    const int kk = producer ? k - 1 : k;                      // the group this warp searches in this iteration
    if (kk >= 0 && kk < n_mine) {
      const int s = kk & 1, use = kk >> 1;
      mbar_wait(&full[s], (uint32_t)(use & 1));
      // the stage record sits behind the group's window copies: recover the geometry from the group itself
      const Group& g = groups[blockIdx.x + (size_t)kk * gridDim.x];
      const K2Geo<V> geo((kMargin + 16 * g.mb_x + g.ux0) & 3, g.uw, g.uh);
      const K2Stage<V, GRAN4> st(K2P_STAGE(s), geo);
      uint32_t best[SM::NBEST];
#pragma unroll
      for (int sl = 0; sl < SM::NBEST; ++sl) best[sl] = kKeyInvalid;
      k2_search_units<V, GRAN4>(geo, st, lane, -1, best);
      k2_warp_reduce<V, GRAN4>(*st.sm, lane, best);
      __syncwarp();
      if (lane == 0) mbar_arrive(&done[s]);                   // release: the atomicMin results are visible to the staging warps
    }
  }
#undef K2P_STAGE
}

template <int V, bool GRAN4>
inline size_t k2_smem_bytes(int max_uh) {
  using SM = K2Smem<V, GRAN4>;
  const size_t rec = (sizeof(SM) + 127) & ~(size_t)127;
  const size_t win = (((size_t)4 * (k2_rows(max_uh) * kWinRowWords + 31) * 4) + 15) & ~(size_t)15;
  const size_t by = (size_t)SM::NS * (((max_uh + V - 1) / V) * V) * 4;
  const size_t slack = (size_t)V * kWinRowBytes;            // rows the last strip reads past the staged ones (never used)
  return rec + win + by + slack;
}

// persistent kernel: barriers + two stages
template <int V, bool GRAN4>
inline size_t k2p_stage_bytes(int max_uh) { return (k2_smem_bytes<V, GRAN4>(max_uh) + 127) & ~(size_t)127; }
template <int V, bool GRAN4>
inline size_t k2p_smem_bytes(int max_uh) { return 128 + 2 * k2p_stage_bytes<V, GRAN4>(max_uh); }

}  // namespace jsvm
