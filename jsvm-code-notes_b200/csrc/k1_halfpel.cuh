// k1_halfpel.cuh -- K1: whole-frame half-pel plane (H.264 6-tap) and margin fill.
//
// Replaces (R = /root/reference/JSVM/H264Extension)
//   QuarterPelFilter::filterFrame        COM/QuarterPelFilter.cpp:61-149
//   YuvPicBuffer::xFillPlaneMargin       COM/YuvPicBuffer.cpp:1608-1641   (margin_fill_kernel)
//
// filterFrame semantics restated:
//   pass 1  tmp[y][2x]   = s[y][x] << 5
//           tmp[y][2x+1] = s[x-2] - 5 s[x-1] + 20 s[x] + 20 s[x+1] - 5 s[x+2] + s[x+3]      (unrounded)
//           for x in [-28, W+28); rows above / below the picture replicate the first / last picture row
//   pass 2  out[2y][c]   = clip255( (tmp[y][c] + 16) / 32 )
//           out[2y+1][c] = clip255( (V6(tmp)[y][c] + 512) / 1024 ),  V6 = same taps over rows y-2..y+3
//           for y in [-28, H+28), c in [-56, 2(W+28))
//   '/' truncates toward zero and gClip maps every negative value to 0, so an arithmetic shift
//   followed by the clamp gives the same bytes.
// The kernel reads the full-pel plane INCLUDING its margins; Frame::extendFrame calls fillMargin
// immediately before filterFrame (COM/Frame.cpp:308, 362), so margin rows equal the first / last
// picture row and the reference's row replication of `tmp` is reproduced by reading them.
//
// One CTA = 64 x 32 integer positions = 128 x 64 output bytes.  The (64+32) x (32+5) input tile
// with its apron is staged by ONE 3D TMA box load (UTMALDG).  The arithmetic is packed:
//   * the 2-D filter is separable and exact in integers, so the diagonal samples are computed as H6( V6( s ) ) instead of
//     the reference's V6( H6( s ) ) -- same integers, and V6 on BYTES fits 16-bit lanes;
//   * phase A: V6(s) for 72 columns x 32 rows with two samples per register (VIADD.16x2, IMAD on packed non-negative
//     lanes), stored as int16 pairs in shared memory;
//   * phase B (thread = 8 integer columns of one row = 2 x 16 output bytes): horizontal half samples with IDP.4A (u8 x s8
//     dot products) straight from the staged bytes, diagonal samples with IDP.2A (s16 x s8) from phase A's pairs, rounding
//     and clamping two samples at a time (VIADDMNMX.S16x2.RELU), bytes interleaved with PRMT, one 16-byte store per row.
#pragma once

#include <cuda.h>
#include "me_common.cuh"
#include "k2_fullsearch.cuh"   // mbarrier / TMA helpers

namespace jsvm {

constexpr int kK1TileX = 64, kK1TileY = 32;
constexpr int kK1BoxW = 96, kK1BoxH = kK1TileY + 5;     // 37 rows: y-2 .. y+34; 96 bytes: columns x0-16 .. x0+79
// (TMA tiled loads need a 16-byte aligned innermost start coordinate -- measured on B200: an unaligned one
//  raises 'illegal instruction' -- so the box starts 16 columns left of the tile instead of 2.)
constexpr int kK1Threads = 256;
constexpr int kK1VCols = 72;                             // phase A columns x0-4 .. x0+67 (taps of x0 .. x0+63 need x0-2 .. x0+66)
constexpr int kK1VPitch = kK1VCols / 2 + 2;              // 32-bit words (int16 pairs) per row of s_v: 38 (even: 8-byte aligned rows)

__device__ __forceinline__ int dp2a_lo_s16s8(uint32_t a, uint32_t b, int c) { int d; asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int dp2a_hi_s16s8(uint32_t a, uint32_t b, int c) { int d; asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

// two int16 lanes t -> two bytes clip255( (t + 16) >> 5 ), each in the HIGH byte of its lane:
// clamp t + 16 to [0, 8191] first (negative results clip to 0 anyway), then the shift is a plain logical one per lane
// (the shift is done as a multiplication by 8 on the FMA pipe: the result byte is then byte 1 of its lane and the
//  PRMT that interleaves the output bytes picks it from there)
__device__ __forceinline__ uint32_t round5_clip_x2_hi(uint32_t t) {
  return __viaddmin_s16x2_relu(t, 0x00100010u, 0x1fff1fffu) * 8u;        // lanes <= 65528: bytes 1 and 3 hold the results
}
__device__ __forceinline__ uint32_t pack_lo16(int lo, int hi) { return __byte_perm((uint32_t)lo, (uint32_t)hi, 0x5410); }

// (A persistent variant that keeps the TMA load of the next tile in flight measured 10-30 % slower: the kernel is bound by
//  instruction issue, 82 % of peak, not by load latency.)
__global__ void __launch_bounds__(kK1Threads)
halfpel_kernel(const __grid_constant__ CUtensorMap plane_map,  // [slot][H+64][W+64] u8, box {96, 37, 1}
               uint8_t* __restrict__ half_planes, size_t half_bytes, int half_stride,
               const int* __restrict__ slots, int W, int H)
{
  __shared__ __align__(128) uint8_t  s_in[kK1BoxH * kK1BoxW];
  __shared__ __align__(16)  uint32_t s_v[kK1TileY * kK1VPitch];     // V6(s) as int16 pairs (col 2i, col 2i+1), col 0 = x0 - 4
  __shared__ uint64_t s_bar;

  const int tid = threadIdx.x;
  const int slot = slots[blockIdx.z];
  const int x0 = -kMargin + kK1TileX * (int)blockIdx.x;            // first integer column of the tile (picture coordinates)
  const int y0 = -(kMargin - 4) + kK1TileY * (int)blockIdx.y;      // first integer row: region starts at -28

  if (tid == 0) { mbar_init(&s_bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&s_bar, kK1BoxH * kK1BoxW);
    tma_load_3d(s_in, &plane_map, &s_bar, x0 - 16 + kMargin, y0 - 2 + kMargin, slot);
  }
  mbar_wait(&s_bar, 0);

  // ---- phase A: vertical 6-tap on bytes.  item = 4 columns x 4 rows; tile byte 12 = column x0 - 4 ----
  if (tid < (kK1VCols / 4) * (kK1TileY / 4)) {
    const int cg = tid % (kK1VCols / 4), rb = tid / (kK1VCols / 4);                 // column group 0..17, row block 0..7
    const uint32_t* src = reinterpret_cast<const uint32_t*>(s_in + (4 * rb) * kK1BoxW + 12 + 4 * cg);   // staged row r = picture row y0 - 2 + r
    uint32_t ev[9], od[9];                                                          // bytes (0, 2) and (1, 3) of each input row as 16-bit lanes
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const uint32_t w = src[r * (kK1BoxW / 4)];
      ev[r] = __byte_perm(w, 0, 0x4240);
      od[r] = __byte_perm(w, 0, 0x4341);
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {                                                   // output row 4 rb + o: taps at staged rows o .. o+5
      // t = 20 (s2 + s3) - 5 (s1 + s4) + (s0 + s5): the products of the non-negative packed sums stay inside their lanes
      const uint32_t te = __vsub2(__vadd2(__vadd2(ev[o + 2], ev[o + 3]) * 20u, __vadd2(ev[o], ev[o + 5])), __vadd2(ev[o + 1], ev[o + 4]) * 5u);
      const uint32_t to = __vsub2(__vadd2(__vadd2(od[o + 2], od[o + 3]) * 20u, __vadd2(od[o], od[o + 5])), __vadd2(od[o + 1], od[o + 4]) * 5u);
      // (c0, c2), (c1, c3) -> (c0, c1), (c2, c3)
      uint2 pr;
      pr.x = __byte_perm(te, to, 0x5410);
      pr.y = __byte_perm(te, to, 0x7632);
      *reinterpret_cast<uint2*>(&s_v[(4 * rb + o) * kK1VPitch + 2 * cg]) = pr;
    }
  }
  __syncthreads();

  // ---- phase B: thread = row y0 + ty, integer columns x0 + 8 tx .. + 7 ----
  const int tx = tid & 7, ty = tid >> 3;
  const int y = y0 + ty;
  if (y < H + kMargin - 4) {
  const int X = 8 * tx;
  // staged bytes of row y: columns X-4 .. X+11 (tile byte 16 = column x0)
  const uint32_t* rw = reinterpret_cast<const uint32_t*>(s_in + (ty + 2) * kK1BoxW + 12 + X);
  const uint32_t w0 = rw[0], w1 = rw[1], w2 = rw[2], w3 = rw[3];

  // horizontal half samples: H(x) = s[x-2] - 5 s[x-1] + 20 s[x] + 20 s[x+1] - 5 s[x+2] + s[x+3]
  constexpr uint32_t kCoefA = 0x1414fb01u;     // bytes ( 1, -5, 20, 20 )
  constexpr uint32_t kCoefB = 0x000001fbu;     // bytes ( -5, 1, 0, 0 )
  int hh[8];
  {
    // column X+k reads bytes (k+2 .. k+7) of the 16 staged bytes
    const uint32_t a0 = __funnelshift_r(w0, w1, 16), b0 = __funnelshift_r(w1, w2, 16);
    const uint32_t a1 = __funnelshift_r(w0, w1, 24), b1 = __funnelshift_r(w1, w2, 24);
    const uint32_t a3 = __funnelshift_r(w1, w2, 8),  b3 = __funnelshift_r(w2, w3, 8);
    const uint32_t a4 = b0,                          b4 = __funnelshift_r(w2, w3, 16);
    const uint32_t a5 = b1,                          b5 = __funnelshift_r(w2, w3, 24);
    const uint32_t a7 = b3,                          b7 = w3 >> 8;
    hh[0] = dp4a_u8s8(b0, kCoefB, dp4a_u8s8(a0, kCoefA, 0));
    hh[1] = dp4a_u8s8(b1, kCoefB, dp4a_u8s8(a1, kCoefA, 0));
    hh[2] = dp4a_u8s8(w2, kCoefB, dp4a_u8s8(w1, kCoefA, 0));
    hh[3] = dp4a_u8s8(b3, kCoefB, dp4a_u8s8(a3, kCoefA, 0));
    hh[4] = dp4a_u8s8(b4, kCoefB, dp4a_u8s8(a4, kCoefA, 0));
    hh[5] = dp4a_u8s8(b5, kCoefB, dp4a_u8s8(a5, kCoefA, 0));
    hh[6] = dp4a_u8s8(w3, kCoefB, dp4a_u8s8(w2, kCoefA, 0));
    hh[7] = dp4a_u8s8(b7, kCoefB, dp4a_u8s8(a7, kCoefA, 0));
  }
  // even output row: ( s[x], clip255( (H(x) + 16) >> 5 ) ) interleaved
  uint32_t row0[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t hq = round5_clip_x2_hi(pack_lo16(hh[2 * j], hh[2 * j + 1]));               // bytes 1 and 3
    const uint32_t sw = (j < 2) ? w1 : w2;                                                    // s[X .. X+3] / s[X+4 .. X+7]
    row0[j] = __byte_perm(sw, hq, (j & 1) ? 0x7352 : 0x7150);
  }

  // vertical and diagonal half samples from V6(s) pairs: columns X-2 .. X+11 are s_v pairs (X+2)/2 .. (X+2)/2 + 6
  const uint32_t* vp = &s_v[ty * kK1VPitch + (X + 2) / 2];
  uint32_t E[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) E[i] = vp[i];
  uint32_t O[6];                                                   // ( col 2i+1, col 2i+2 )
#pragma unroll
  for (int i = 0; i < 6; ++i) O[i] = __byte_perm(E[i], E[i + 1], 0x5432);
  constexpr uint32_t kCoefC = 0x1414fb01u;     // lo: ( 1, -5 )  hi: ( 20, 20 )
  constexpr uint32_t kCoefD = 0x000001fbu;     // lo: ( -5, 1 )
  int hv[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    // HV(x) = H6 over V6(s)[x-2 .. x+3], + 512 for the rounding of ( . + 512 ) >> 10
    hv[2 * i]     = dp2a_lo_s16s8(E[i + 2], kCoefD, dp2a_hi_s16s8(E[i + 1], kCoefC, dp2a_lo_s16s8(E[i], kCoefC, 512)));
    hv[2 * i + 1] = dp2a_lo_s16s8(O[i + 2], kCoefD, dp2a_hi_s16s8(O[i + 1], kCoefC, dp2a_lo_s16s8(O[i], kCoefC, 512)));
  }
  uint32_t row1[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t vq = round5_clip_x2_hi(E[j + 1]);                                  // clip255( (V6(s) + 16) >> 5 ), columns X+2j, X+2j+1: bytes 1, 3
    // ( hv >> 10 ) clipped: bits 8..23 of each sum as int16 lanes (|hv| < 2^23), clamp to [0, 1023], x 64 -> bytes 1, 3
    const uint32_t dq = __viaddmin_s16x2_relu(__byte_perm((uint32_t)hv[2 * j], (uint32_t)hv[2 * j + 1], 0x6521), 0u, 0x03ff03ffu) * 64u;
    row1[j] = __byte_perm(vq, dq, 0x7351);
  }

  // half-pel coordinates of this thread's 16 bytes: hx = 2*x0 + 16 tx .. +15, rows 2y and 2y+1
  const int c0 = 16 * tx;
  const int hx = 2 * x0 + c0;
  const int hx_lo = -2 * (kMargin - 4), hx_hi = 2 * (W + kMargin - 4);          // written region [-56, 2(W+28))
  uint8_t* d0 = half_planes + (size_t)slot * half_bytes + (size_t)(kHpOrigin + 2 * y) * half_stride + kHpOrigin + hx;
  uint8_t* d1 = d0 + half_stride;
  if (hx >= hx_lo && hx + 16 <= hx_hi) {
    *reinterpret_cast<uint4*>(d0) = make_uint4(row0[0], row0[1], row0[2], row0[3]);
    *reinterpret_cast<uint4*>(d1) = make_uint4(row1[0], row1[1], row1[2], row1[3]);
  } else {
    if (hx >= hx_lo && hx + 8 <= hx_hi) {
      *reinterpret_cast<uint2*>(d0) = make_uint2(row0[0], row0[1]);
      *reinterpret_cast<uint2*>(d1) = make_uint2(row1[0], row1[1]);
    }
    if (hx + 8 >= hx_lo && hx + 16 <= hx_hi) {
      *reinterpret_cast<uint2*>(d0 + 8) = make_uint2(row0[2], row0[3]);
      *reinterpret_cast<uint2*>(d1 + 8) = make_uint2(row1[2], row1[3]);
    }
  }
  }
}

// YuvPicBuffer::xFillPlaneMargin: left / right replicate per picture row, then the padded first /
// last rows are copied up / down.  One thread per margin byte group; dense picture rows are
// written by the upload copy itself.
__global__ void margin_fill_kernel(uint8_t* __restrict__ planes, size_t plane_bytes, int plane_stride,
                                   const int* __restrict__ slots, int W, int H)
{
  uint8_t* p = planes + (size_t)slots[blockIdx.z] * plane_bytes;
  const int S = plane_stride, PH = H + 2 * kMargin;
  const int row = blockIdx.y;                                   // 0 .. PH-1 (padded row)
  if (row >= PH) return;
  const int src_row = min(max(row - kMargin, 0), H - 1) + kMargin;
  const uint8_t* sr = p + (size_t)src_row * S;
  uint8_t* dr = p + (size_t)row * S;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < S; x += gridDim.x * blockDim.x) {
    const bool inside = (row == src_row) && x >= kMargin && x < kMargin + W;
    if (!inside) {
      const int sx = min(max(x - kMargin, 0), W - 1) + kMargin;
      dr[x] = sr[sx];
    }
  }
}

// int16 (XPel) plane with margins -> byte plane, and u8 dense picture -> interior of the byte plane.
__global__ void xpel_to_u8_kernel(const int16_t* __restrict__ src, int src_stride, uint8_t* __restrict__ dst, int dst_stride, int cols, int rows, int* __restrict__ bad)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= cols || y >= rows) return;
  const int v = src[(size_t)y * src_stride + x];
  if (v < 0 || v > 255) atomicAdd(bad, 1);
  dst[(size_t)y * dst_stride + x] = (uint8_t)v;
}

__global__ void u8_to_xpel_kernel(const uint8_t* __restrict__ src, int src_stride, int16_t* __restrict__ dst, int dst_stride, int cols, int rows)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= cols || y >= rows) return;
  dst[(size_t)y * dst_stride + x] = src[(size_t)y * src_stride + x];
}

}  // namespace jsvm
