// k7_interlayer.cuh -- K7: inter-layer resampling that feeds the enhancement-layer motion search.
//
// Residual upsampling (the signal subtracted from the original in the residual-prediction passes, ENC/MbEncoder.cpp:606-608, N3):
//   DownConvert::residualUpsampling -> xResidualUpsampling -> xCompResidualUpsampling -> xDetermineTransBlkIdcs +
//   xBasicResidualUpsampling          COM/DownConvert.cpp:618-631, 2476-2602, 2604-2731, 2733-2817, 2867-2958
//   (called through Frame::residualUpsampling, COM/Frame.cpp:398-406, from ENC/GOPEncoder.cpp:2812)
// for progressive layers (frame_mbs_only in both layers): a separable bilinear interpolation on a 1/16-sample grid that never
// interpolates ACROSS a transform-block boundary of the base layer (4x4 or 8x8 luma blocks per macroblock, 4x4 chroma blocks):
// there the nearer sample is taken.  Horizontal pass at base rows, then vertical pass; samples outside the scaled reference window
// are zero.  One thread per output sample; the two horizontally interpolated values a sample needs are recomputed, not staged
// (4 base samples per output sample, all L1 / L2 hits) -- the kernel is bound by the write of the output plane.
#pragma once

#include "me_common.cuh"

namespace jsvm {

struct ResamplePlane {                 // the arguments of xBasicResidualUpsampling for one plane
  int base_w, base_h, cur_w, cur_h;
  int l_off, t_off, r_off, b_off;
  int shift_x_m4, shift_y_m4;
  int scale_x, scale_y, offset_x, offset_y, add_x, add_y, delta_x, delta_y;
  int mb_size, mbs_per_row;            // 16 (luma) or 8 (chroma); base macroblocks per row
  int chroma;
};

__device__ __forceinline__ int clip_int(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// m_paiTransBlkIdc (xDetermineTransBlkIdcs, progressive branch): an id that is equal for two samples iff they lie in the same
// transform block of the base layer
__device__ __forceinline__ int trans_blk_idc(const ResamplePlane& P, const uint8_t* __restrict__ t8x8, int x, int y) {
  const int sh = P.chroma ? 3 : 4;                           // macroblock size 8 (chroma) or 16
  const int mb = (y >> sh) * P.mbs_per_row + (x >> sh);
  const int ix = x & (P.mb_size - 1), iy = y & (P.mb_size - 1);
  if (!P.chroma && t8x8[mb]) return ((((iy >> 3) << 1) + (ix >> 3) + (mb << 2)) << 1) + 1;
  return (((iy >> 2) << 2) + (ix >> 2) + (mb << 4)) << 1;
}

// horizontally interpolated value (scaled by 16) at output column position (x0, x1, phase) of base row y
__device__ __forceinline__ int resample_h(const ResamplePlane& P, const int16_t* __restrict__ src, const uint8_t* __restrict__ t8x8,
                                          int y, int x0, int x1, int phase) {
  const int a = src[(size_t)y * P.base_w + x0], b = src[(size_t)y * P.base_w + x1];
  if (trans_blk_idc(P, t8x8, x0, y) == trans_blk_idc(P, t8x8, x1, y)) return (16 - phase) * a + phase * b;
  return (phase < 8 ? a : b) << 4;
}

__global__ void __launch_bounds__(256)
residual_upsample_kernel(ResamplePlane P, const int16_t* __restrict__ src, const uint8_t* __restrict__ t8x8, int16_t* __restrict__ dst)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= P.cur_w || j >= P.cur_h) return;
  int out = 0;
  if (i >= P.l_off && i < P.cur_w - P.r_off && j >= P.t_off && j < P.cur_h - P.b_off) {
    // horizontal position (identical for both passes, COM/DownConvert.cpp:2889-2892, 2916-2918)
    const int pos_x16 = (int)((uint32_t)((i - P.offset_x) * P.scale_x + P.add_x) >> P.shift_x_m4) - P.delta_x;
    const int ph_x = pos_x16 & 15, px = pos_x16 >> 4;
    const int x0 = clip_int(px, 0, P.base_w - 1), x1 = clip_int(px + 1, 0, P.base_w - 1);
    const int xr = clip_int(px + (ph_x >> 3), 0, P.base_w - 1);               // rounded column: where the vertical pass looks up the block ids
    // vertical position (:2929-2935)
    const int pre = (j - P.offset_y) * P.scale_y + P.add_y;
    const int post = j >= P.offset_y ? (int)((uint32_t)pre >> P.shift_y_m4) : (pre >> P.shift_y_m4);
    const int pos_y16 = post - P.delta_y;
    const int ph_y = pos_y16 & 15, py = pos_y16 >> 4;
    const int y0 = clip_int(py, 0, P.base_h - 1), y1 = clip_int(py + 1, 0, P.base_h - 1);
    const int h0 = resample_h(P, src, t8x8, y0, x0, x1, ph_x), h1 = resample_h(P, src, t8x8, y1, x0, x1, ph_x);
    if (trans_blk_idc(P, t8x8, xr, y0) == trans_blk_idc(P, t8x8, xr, y1)) out = ((16 - ph_y) * h0 + ph_y * h1 + 128) >> 8;
    else out = ((ph_y < 8 ? h0 : h1) + 8) >> 4;
  }
  dst[(size_t)j * P.cur_w + i] = (int16_t)out;
}

}  // namespace jsvm

// ---------------------------------------------------------------------------------------------------------------------------
// Inter-layer motion prediction: MbDataCtrl::upsampleMotion -> MotionUpsampling::resample (COM/MbDataCtrl.cpp:85-1140; G.6.1,
// G.6.2, G.8.6.1).  Every enhancement-layer macroblock is independent: one thread per macroblock, all state in registers / local
// memory (16 partition ids, 2 x 16 motion vectors).  Progressive layers, frame macroblocks.
// ---------------------------------------------------------------------------------------------------------------------------
namespace jsvm {

struct MotionUpsample {                // PosCalcParam / MvScaleParam (COM/MbDataCtrl.cpp:14-84) and the slice-level switches
  int frame_mbw, frame_mbh, ref_w, ref_h;
  int crop_x0, crop_y0, crop_x1, crop_y1;
  int shift_x, shift_y, scale_x, scale_y, add_x, add_y;
  int mv_scale_x, mv_scale_y;
  int rs_change, max_list, slice_type, ref_layer_dq_id, direct_8x8_inference, check_residual_pred, mv_threshold;
};

__device__ __forceinline__ int mu_part_idc(const MotionUpsample& M, const jsvm_mb_motion* __restrict__ base, int mbx, int mby, int x_in, int y_in) {
  const int cx = (mbx << 4) + x_in, cy = (mby << 4) + y_in;
  int bx = (int)((uint32_t)(cx * M.scale_x + M.add_x) >> M.shift_x), by = (int)((uint32_t)(cy * M.scale_y + M.add_y) >> M.shift_y);
  bx = min(bx, M.ref_w - 1); by = min(by, M.ref_h - 1);
  const int mb_idx = (by >> 4) * (M.ref_w >> 4) + (bx >> 4), xi = bx & 15, yi = by & 15;
  const jsvm_mb_motion& b = base[mb_idx];
  if (b.mb_mode >= 6) return -1;                                           // MbData::isIntra
  const int b8 = ((yi >> 3) << 1) + (xi >> 3), b4 = (((yi & 7) >> 2) << 1) + ((xi & 7) >> 2);
  // aauiNxNPartIdx (:321-329) as nibbles: first block of the partition that holds block i, per (sub-)macroblock mode
  constexpr uint32_t nxn[6] = { 0x0000, 0x0000, 0x2200, 0x1010, 0x3210, 0x3210 };
  int part8 = 0, part4 = 0;
  if (b.mb_mode == 0 && b.slice_type == 1) {                                // B_Skip / B_Direct_16x16
    if (M.ref_layer_dq_id == 0) { part8 = b8; part4 = b4; }
  } else {
    part8 = (nxn[b.mb_mode] >> (4 * b8)) & 15;
    if (b.mb_mode == 4 || b.mb_mode == 5) {
      const int bm = b.blk_mode[b8];
      if (bm == 0) { if (M.ref_layer_dq_id == 0) part4 = b4; }
      else part4 = (nxn[bm - 8 + 1] >> (4 * b4)) & 15;
    }
  }
  const int x4 = ((part8 & 1) << 1) + (part4 & 1), y4 = ((part8 >> 1) << 1) + (part4 >> 1);
  return (mb_idx << 4) + (y4 << 2) + x4;
}

__device__ __forceinline__ int mu_absdiff(jsvm_mv a, jsvm_mv b) { return abs(a.hor - b.hor) + abs(a.ver - b.ver); }
__device__ __forceinline__ jsvm_mv mu_avg2(jsvm_mv a, jsvm_mv b) {
  jsvm_mv r; r.hor = (int16_t)((int16_t)((int16_t)(a.hor + b.hor) + 1) >> 1); r.ver = (int16_t)((int16_t)((int16_t)(a.ver + b.ver) + 1) >> 1); return r;
}
__device__ __forceinline__ jsvm_mv mu_avg4(jsvm_mv a, jsvm_mv b, jsvm_mv c, jsvm_mv d) {
  jsvm_mv r;
  r.hor = (int16_t)((int16_t)((int16_t)((int16_t)((int16_t)(a.hor + b.hor) + c.hor) + d.hor) + 2) >> 2);
  r.ver = (int16_t)((int16_t)((int16_t)((int16_t)((int16_t)(a.ver + b.ver) + c.ver) + d.ver) + 2) >> 2);
  return r;
}

__global__ void __launch_bounds__(64)
motion_upsample_kernel(MotionUpsample M, const jsvm_mb_motion* __restrict__ base, jsvm_mb_motion* __restrict__ out)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M.frame_mbw * M.frame_mbh) return;
  const int mbx = t % M.frame_mbw, mby = t / M.frame_mbw;
  jsvm_mb_motion o;
  memset(&o, 0, sizeof(o));
  o.blk_mode[0] = o.blk_mode[1] = o.blk_mode[2] = o.blk_mode[3] = 8;
  for (int l = 0; l < 2; ++l) for (int k = 0; k < 4; ++k) o.ref_idx[l][k] = -1;
  const bool in_crop = mbx >= M.crop_x0 && mbx < M.crop_x1 && mby >= M.crop_y0 && mby < M.crop_y1;     // xIsInCropWindow
  o.in_crop_window = in_crop; o.res_pred_safe = in_crop;
  if (!in_crop) { out[t] = o; return; }

  int idc[16];                                               // [y * 4 + x]
  bool intra_bl = true;
  for (int k = 0; k < 16; ++k) { idc[k] = mu_part_idc(M, base, mbx, mby, ((k & 3) << 2) + 1, ((k >> 2) << 2) + 1); intra_bl = intra_bl && idc[k] == -1; }
  if (intra_bl) o.mb_mode = 36;                              // INTRA_BL
  else {
    if (!M.rs_change) {
      // G.8.6.1.1: intra 4x4 blocks take the partition of an inter neighbour inside their 8x8 block (horizontal, vertical, diagonal;
      // blocks replaced earlier in the scan do not donate), then intra 8x8 blocks take their neighbours' in the same order
      for (int p8 = 0; p8 < 4; ++p8) {
        const int xp = p8 & 1, yp = p8 >> 1;
        uint32_t proc = 0;
        for (int s = 0; s < 4; ++s) {
          const int xs = s & 1, ys = s >> 1, xc = (xp << 1) + xs, yc = (yp << 1) + ys;
          if (idc[yc * 4 + xc] != -1) continue;
          const int xi = 1 - xs, yi = 1 - ys, xci = (xp << 1) + xi, yci = (yp << 1) + yi;
          proc |= 1u << (ys * 2 + xs);
          if (!(proc >> (ys * 2 + xi) & 1) && idc[yc * 4 + xci] != -1) idc[yc * 4 + xc] = idc[yc * 4 + xci];
          else if (!(proc >> (yi * 2 + xs) & 1) && idc[yci * 4 + xc] != -1) idc[yc * 4 + xc] = idc[yci * 4 + xc];
          else if (!(proc >> (yi * 2 + xi) & 1) && idc[yci * 4 + xci] != -1) idc[yc * 4 + xc] = idc[yci * 4 + xci];
        }
      }
      uint32_t proc8 = 0;
      for (int p8 = 0; p8 < 4; ++p8) {
        const int xp = p8 & 1, yp = p8 >> 1, xpi = 1 - xp, ypi = 1 - yp, xo = xp << 1, yo = yp << 1, xoi = 2 - xp, yoi = 2 - yp;
        if (idc[yo * 4 + xo] != -1) continue;
        proc8 |= 1u << (yp * 2 + xp);
        if (!(proc8 >> (yp * 2 + xpi) & 1) && idc[yo * 4 + xoi] != -1) {
          idc[yo * 4 + xo] = idc[yo * 4 + xo + 1] = idc[yo * 4 + xoi];
          idc[(yo + 1) * 4 + xo] = idc[(yo + 1) * 4 + xo + 1] = idc[(yo + 1) * 4 + xoi];
        } else if (!(proc8 >> (ypi * 2 + xp) & 1) && idc[yoi * 4 + xo] != -1) {
          idc[yo * 4 + xo] = idc[(yo + 1) * 4 + xo] = idc[yoi * 4 + xo];
          idc[yo * 4 + xo + 1] = idc[(yo + 1) * 4 + xo + 1] = idc[yoi * 4 + xo + 1];
        } else if (!(proc8 >> (ypi * 2 + xpi) & 1) && idc[yoi * 4 + xoi] != -1) {
          idc[yo * 4 + xo] = idc[yo * 4 + xo + 1] = idc[(yo + 1) * 4 + xo] = idc[(yo + 1) * 4 + xo + 1] = idc[yoi * 4 + xoi];
        }
      }
    }
    int ref8[2][4];                                          // [list][8x8 raster]
    jsvm_mv mv[2][16];                                       // [list][4x4 raster]
    for (int l = 0; l < 2; ++l) { for (int k = 0; k < 4; ++k) ref8[l][k] = -1; for (int k = 0; k < 16; ++k) mv[l][k] = jsvm_mv{0, 0}; }
    for (int l = 0; l < M.max_list; ++l) {
      int rtmp[16];
      for (int k = 0; k < 16; ++k) {                         // xGetInitialBaseRefIdxAndMv: scaled base MV (no cropping change)
        const jsvm_mb_motion& b = base[idc[k] >> 4];
        const int b4 = idc[k] & 15, rb = b.ref_idx[l][((b4 >> 3) << 1) | ((b4 & 3) >> 1)];
        rtmp[k] = rb < 1 ? -1 : rb;
        if (rb >= 1) {
          mv[l][k].hor = (int16_t)((b.mv[l][b4].hor * M.mv_scale_x + 32768) >> 16);
          mv[l][k].ver = (int16_t)((b.mv[l][b4].ver * M.mv_scale_y + 32768) >> 16);
        }
      }
      for (int b8 = 0; b8 < 4; ++b8) ref8[l][b8] = rtmp[((b8 >> 1) << 3) + ((b8 & 1) << 1)];
      if (M.rs_change) continue;
      for (int b8 = 0; b8 < 4; ++b8) {                       // smallest positive reference index of the 8x8; MVs follow a block that has it
        const int x8 = b8 & 1, y8 = b8 >> 1;
        int r = ref8[l][b8];
        for (int s = 0; s < 4; ++s) { const int v = rtmp[((y8 << 1) + (s >> 1)) * 4 + (x8 << 1) + (s & 1)]; r = r < 1 ? v : (v < 1 ? r : min(r, v)); }
        ref8[l][b8] = r;
        for (int s = 0; s < 4; ++s) {
          const int xs = s & 1, ys = s >> 1, x = (x8 << 1) + xs, y = (y8 << 1) + ys;
          if (r == rtmp[y * 4 + x]) continue;
          const int xi = (x8 << 1) + 1 - xs, yi = (y8 << 1) + 1 - ys;
          if (r == rtmp[y * 4 + xi]) mv[l][y * 4 + x] = mv[l][y * 4 + xi];
          else if (r == rtmp[yi * 4 + x]) mv[l][y * 4 + x] = mv[l][yi * 4 + x];
          else mv[l][y * 4 + x] = mv[l][yi * 4 + xi];
        }
      }
    }
    for (int b8 = 0; b8 < 4; ++b8) {                         // xDeriveBlockModeAndUpdateMv
      const int thr = M.rs_change ? 0 : 1, xo = (b8 & 1) << 1, yo = (b8 >> 1) << 1;
      const int i00 = yo * 4 + xo, i10 = i00 + 1, i01 = i00 + 4, i11 = i00 + 5;
      if (M.direct_8x8_inference && !M.rs_change && M.slice_type == 1) {
        const int ic = ((yo >> 1) * 3) * 4 + (xo >> 1) * 3;  // the corner 4x4 of the macroblock
        for (int l = 0; l < M.max_list; ++l) { const jsvm_mv tm = mv[l][ic]; mv[l][i00] = mv[l][i10] = mv[l][i01] = mv[l][i11] = tm; }
      }
      bool hor = true, ver = true, all = true;
      for (int l = 0; l < M.max_list; ++l) {
        const bool h1 = mu_absdiff(mv[l][i00], mv[l][i10]) <= thr, h2 = mu_absdiff(mv[l][i01], mv[l][i11]) <= thr;
        const bool v1 = mu_absdiff(mv[l][i00], mv[l][i01]) <= thr, v2 = mu_absdiff(mv[l][i10], mv[l][i11]) <= thr;
        const bool dg = mu_absdiff(mv[l][i00], mv[l][i11]) <= thr;
        all = all && h1 && v1 && dg; hor = hor && h1 && h2; ver = ver && v1 && v2;
      }
      const int bm = all ? 8 : hor ? 9 : ver ? 10 : 11;
      o.blk_mode[b8] = (uint8_t)bm;
      if (M.rs_change || bm == 11) continue;
      for (int l = 0; l < M.max_list; ++l) {
        if (bm == 8) mv[l][i00] = mv[l][i10] = mv[l][i01] = mv[l][i11] = mu_avg4(mv[l][i00], mv[l][i10], mv[l][i01], mv[l][i11]);
        else if (bm == 9) { const jsvm_mv a = mu_avg2(mv[l][i00], mv[l][i10]), b = mu_avg2(mv[l][i01], mv[l][i11]); mv[l][i00] = mv[l][i10] = a; mv[l][i01] = mv[l][i11] = b; }
        else { const jsvm_mv a = mu_avg2(mv[l][i00], mv[l][i01]), b = mu_avg2(mv[l][i10], mv[l][i11]); mv[l][i00] = mv[l][i01] = a; mv[l][i10] = mv[l][i11] = b; }
      }
    }
    int mode = 4;                                            // xDeriveMbMode
    if (o.blk_mode[0] == 8 && o.blk_mode[1] == 8 && o.blk_mode[2] == 8 && o.blk_mode[3] == 8) {
      auto same = [&](int l, int a, int b) {
        const jsvm_mv ma = mv[l][((a >> 1) << 3) + ((a & 1) << 1)], mb = mv[l][((b >> 1) << 3) + ((b & 1) << 1)];
        return ref8[l][a] == ref8[l][b] && ma.hor == mb.hor && ma.ver == mb.ver;
      };
      bool hm = true, vm = true;
      for (int l = 0; l < M.max_list; ++l) { hm = hm && same(l, 0, 1) && same(l, 2, 3); vm = vm && same(l, 0, 2) && same(l, 1, 3); }
      mode = vm ? (hm ? 1 : 3) : (hm ? 2 : 4);
    }
    o.mb_mode = (uint8_t)mode;
    uint32_t fb = 0;                                         // xDeriveFwdBwd
    for (int n = 3; n >= 0; --n) fb = (fb << 4) + (ref8[0][n] > 0 ? 1 : 0) + (ref8[1][n] > 0 ? 2 : 0);
    o.fwd_bwd = (uint16_t)fb;
    if (!M.rs_change)                                        // xSetInterIntraIdc
      for (int k = 0; k < 4; ++k) if (mu_part_idc(M, base, mbx, mby, (k & 1) * 15, (k >> 1) * 15) == -1) o.base_intra |= (uint8_t)(1 << k);
    for (int l = 0; l < M.max_list; ++l) {
      for (int k = 0; k < 4; ++k) o.ref_idx[l][k] = (int8_t)ref8[l][k];
      for (int k = 0; k < 16; ++k) o.mv[l][k] = mv[l][k];
    }
  }
  if (M.check_residual_pred && !M.rs_change) {               // xSetResPredSafeFlag
    bool safe = true;
    for (int k16 = 0; k16 < 16 && safe; ++k16) {
      const int x = k16 & 3, y = k16 >> 2;
      int pi[4]; bool samep = true, nointer = false;
      for (int k = 0; k < 4; ++k) { pi[k] = mu_part_idc(M, base, mbx, mby, (x << 2) + 3 * (k & 1), (y << 2) + 3 * (k >> 1)); samep = samep && pi[k] == pi[0]; nointer = nointer || pi[k] < 0; }
      if (samep || nointer) { safe = nointer; continue; }
      for (int l = 0; l < 2 && safe; ++l) {
        int rf[4]; jsvm_mv v[4];
        for (int k = 0; k < 4; ++k) {
          const jsvm_mb_motion& b = base[pi[k] >> 4]; const int b4 = pi[k] & 15;
          rf[k] = b.ref_idx[l][((b4 >> 3) << 1) | ((b4 & 3) >> 1)]; v[k] = b.mv[l][b4];
        }
        for (int k = 0; k < 4 && safe; ++k) safe = rf[k] == rf[0];
        if (safe) { const jsvm_mv avg = mu_avg4(v[0], v[1], v[2], v[3]); for (int k = 0; k < 4 && safe; ++k) safe = mu_absdiff(v[k], avg) <= M.mv_threshold; }
      }
    }
    o.res_pred_safe = safe;
  }
  out[t] = o;
}

}  // namespace jsvm
