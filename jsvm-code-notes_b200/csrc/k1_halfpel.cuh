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
// with its apron is staged by ONE 3D TMA box load (UTMALDG), pass 1 goes to shared memory as
// int16, pass 2 writes 16-byte vectors (8-byte at the two region edges, which sit at 8 mod 16).
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
constexpr int kK1TmpStride = 2 * kK1TileX + 8;          // int16 elements per tmp row (+8: 16-byte aligned rows, spreads banks)

__global__ void __launch_bounds__(kK1Threads)
halfpel_kernel(const __grid_constant__ CUtensorMap plane_map,  // [slot][H+64][W+64] u8, box {80, 37, 1}
               uint8_t* __restrict__ half_planes, size_t half_bytes, int half_stride,
               const int* __restrict__ slots, int W, int H)
{
  __shared__ __align__(128) uint8_t  s_in[kK1BoxH * kK1BoxW];
  __shared__ __align__(16)  int16_t  s_tmp[kK1BoxH * kK1TmpStride];
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

  // ---- pass 1: horizontal 6-tap on the 37 staged rows ----
  for (int i = tid; i < kK1BoxH * (kK1TileX / 4); i += kK1Threads) {
    const int r = i / (kK1TileX / 4), xc = (i - r * (kK1TileX / 4)) * 4;     // 4 integer columns xc .. xc+3
    // tile byte 0 = column x0-16, so s[x0+xc-2] sits at byte xc+14; load the aligned words from byte xc+12
    const uint32_t* rw = reinterpret_cast<const uint32_t*>(s_in + r * kK1BoxW + xc + 12);
    const uint32_t w0 = rw[0], w1 = rw[1], w2 = rw[2];
    int b[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) { b[k] = (w0 >> (8 * k)) & 0xff; b[4 + k] = (w1 >> (8 * k)) & 0xff; b[8 + k] = (w2 >> (8 * k)) & 0xff; }
    int16_t o[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // s[x] = b[k+4]
      int t = b[k + 4] + b[k + 5];
      t <<= 2;
      t -= b[k + 3] + b[k + 6];
      t += t << 2;
      t += b[k + 2] + b[k + 7];
      o[2 * k] = (int16_t)(b[k + 4] << 5);
      o[2 * k + 1] = (int16_t)t;
    }
    *reinterpret_cast<uint4*>(&s_tmp[r * kK1TmpStride + 2 * xc]) = *reinterpret_cast<const uint4*>(o);
  }
  __syncthreads();

  // ---- pass 2: vertical 6-tap, rounding, clamp, 16 output bytes per row per thread ----
  const int tx = tid & 7, ty = tid >> 3;                      // 8 x 32 threads; thread = 16 half-pel columns x 1 integer row
  const int y = y0 + ty;
  if (y >= H + kMargin - 4) return;
  const int c0 = 16 * tx;                                      // first half-pel column inside the tile
  const int16_t* tp = &s_tmp[(ty + 2) * kK1TmpStride + c0];
  uint32_t row0[4], row1[4];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    // 6 rows x 8 int16 columns as 128-bit shared loads
    uint4 rv[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) rv[r] = *reinterpret_cast<const uint4*>(tp + 8 * half + (r - 2) * kK1TmpStride);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint32_t p0 = 0, p1 = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = 4 * q + k;                       // 0..7 inside this half
        int e[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          const uint32_t wv = (c >> 1) == 0 ? rv[r].x : (c >> 1) == 1 ? rv[r].y : (c >> 1) == 2 ? rv[r].z : rv[r].w;
          e[r] = (c & 1) ? ((int)wv >> 16) : (int)(int16_t)(wv & 0xffff);
        }
        int t = e[2] + e[3];
        t <<= 2;
        t -= e[1] + e[4];
        t += t << 2;
        t += e[0] + e[5];
        int v0 = (e[2] + 16) >> 5, v1 = (t + 512) >> 10;
        v0 = min(max(v0, 0), 255); v1 = min(max(v1, 0), 255);
        p0 |= (uint32_t)v0 << (8 * k);
        p1 |= (uint32_t)v1 << (8 * k);
      }
      row0[2 * half + q] = p0; row1[2 * half + q] = p1;
    }
  }
  // half-pel coordinates of this thread's 16 bytes: hx = 2*x0 + c0 .. +15, rows 2y and 2y+1
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
