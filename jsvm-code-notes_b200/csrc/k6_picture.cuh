// k6_picture.cuh -- K6: motion-vector prediction and per-macroblock scheduling of a whole picture's motion search on the device
// (SURVEY 8f rank 1).
//
// What the encoder does one call at a time from the host (R = /root/reference/JSVM/H264Extension):
//   MbDataAccess::getMvPredictor / xGetMvPredictor      INCC/MbDataAccess.h:289-296, 485-513; COM/MbDataAccess.cpp:137-187
//   neighbour lookup xGetBlockLeft / Above / AboveRight / AboveLeft                      INCC/MbDataAccess.h:1193-1348
//   the list-0 loops of MbEncoder::xEstimateMb16x16 / 16x8 / 8x16 / 8x8 + xEstimateSubMb8x8
//                                                        ENC/MbEncoder.cpp:4449-4560, 4830-4935, 5234-5330, 5637-5700, 5950-6040
//   (mode / reference-index bits: uiMbBits, uiBlkBits, getRefIdxBits :4372-4379, g_aucFrameBits :25-40)
// is done here for every macroblock of P pictures at once: the MV predictor of a partition comes from the FINAL motion of the
// left / above / above-right (above-left) macroblocks and from the partitions of the same macroblock mode estimated before it, so
// macroblocks are processed in wavefronts ( x + 2 y = const ), the waves of all pictures of the batch side by side, and inside a
// wave in four steps:
//   step 1   16x16, 16x8 upper, 8x16 left, 8x8 #0        (depend on neighbouring macroblocks only: ONE SAD pass per reference)
//   step 2   16x8 lower, 8x16 right, 8x8 #1              (depend on step 1 of their own mode)
//   step 3   8x8 #2                                      step 4   8x8 #3
// Each step = this file's prepare kernel (MVPs -> jobs + search groups, the planner's work, on the device) + K2 + K3.
// The macroblock mode is then the one with the smallest sum of estimateBlockWithStart costs (strict '<' in the order 16x16, 16x8,
// 8x16, 8x8, as xCheckBestEstimation compares): a motion-cost-only decision -- the rate-distortion decision of xSetRdCostInterMb
// (transform, quantisation, entropy coding) is outside the scope table.
#pragma once

#include "me_common.cuh"

namespace jsvm {

__host__ __device__ constexpr int pic_step_jobs(int step) { return step == 0 ? 4 : step == 1 ? 3 : 1; }
// partition numbering of jsvm_pic_mb: 0 16x16 | 1, 2 16x8 | 3, 4 8x16 | 5..8 8x8
__constant__ int8_t c_pic_step_part[4][4] = { {0, 1, 3, 5}, {2, 4, 6, -1}, {7, -1, -1, -1}, {8, -1, -1, -1} };
__constant__ uint8_t c_pic_part_mode[9] = { JSVM_MODE_16x16, JSVM_MODE_16x8, JSVM_MODE_16x8, JSVM_MODE_8x16, JSVM_MODE_8x16,
                                            JSVM_BLK_8x8, JSVM_BLK_8x8, JSVM_BLK_8x8, JSVM_BLK_8x8 };
__constant__ uint8_t c_pic_part_blk[9] = { 0, 0, 8, 0, 2, 0, 2, 8, 10 };
__constant__ uint8_t c_pic_part_bits[9] = { 1, 3, 3, 3, 3, 1 + 5, 1, 1, 1 };     // P slice: uiMbBits[0] / uiBlkBits[0] (+ 5 mode bits on 8x8 #0)

struct PicTemp {                 // per (picture, macroblock): winners of the nine partitions so far
  jsvm_mv mv[9];
  int8_t ref[9];
  int8_t pad[3];
  uint32_t cost[9];
};

struct MvRef { int hor, ver, ref; };     // Mv3D: ref 0 = BLOCK_NOT_AVAILABLE (INCC/Mv.h:131-135)

// Mv3D of 4x4 block (bx, by) of macroblock (mx, my) of the FINAL motion field; outside the picture: m_rcMbUnavailable
__device__ __forceinline__ MvRef pic_neighbour(const jsvm_pic_mb* __restrict__ field, int mb_w, int mb_h, int mx, int my, int b4) {
  MvRef r{0, 0, 0};
  if (mx < 0 || my < 0 || mx >= mb_w || my >= mb_h) return r;
  const jsvm_pic_mb& m = field[my * mb_w + mx];
  r.hor = m.mv[b4].hor; r.ver = m.mv[b4].ver;
  r.ref = m.ref_idx[((b4 >> 3) << 1) | ((b4 & 3) >> 1)];      // m_ascRefIdx[ m_auiBlk2Part[ b4x4 ] ]
  return r;
}

__device__ __forceinline__ int median3(int a, int b, int c) { return a > b ? (a > c ? (b > c ? b : c) : a) : (b > c ? (a > c ? a : c) : b); }

enum { PIC_MEDIAN = 0, PIC_PRED_A, PIC_PRED_B, PIC_PRED_C };

// MbDataAccess::xGetMvPredictor (COM/MbDataAccess.cpp:160-186) once A, B, C ( C already replaced by D when unavailable ) are known
__device__ __forceinline__ jsvm_mv pic_mvp(const MvRef& A, const MvRef& B, const MvRef& C, int ref, int type) {
  const MvRef* pick = nullptr;
  if ((type == PIC_PRED_A && A.ref == ref) || (A.ref == ref && B.ref != ref && C.ref != ref) || (B.ref == 0 && C.ref == 0)) pick = &A;
  else if ((type == PIC_PRED_B && B.ref == ref) || (A.ref != ref && B.ref == ref && C.ref != ref)) pick = &B;
  else if ((type == PIC_PRED_C && C.ref == ref) || (A.ref != ref && B.ref != ref && C.ref == ref)) pick = &C;
  jsvm_mv m;
  if (pick) { m.hor = (int16_t)pick->hor; m.ver = (int16_t)pick->ver; }
  else { m.hor = (int16_t)median3(A.hor, B.hor, C.hor); m.ver = (int16_t)median3(A.ver, B.ver, C.ver); }
  return m;
}

__device__ __forceinline__ uint32_t pic_ref_bits(int ref, int n_refs) {            // getRefIdxBits
  if (n_refs == 1) return 0;
  if (n_refs == 2) return 1;
  return 1u + 2u * (31u - (uint32_t)__clz(ref));                                   // g_aucFrameBits: 1, 3, 3, 5, 5, 5, 5, 7, ...
}

// One thread per (picture, macroblock of the wave).  step 0..3: collect the results of step - 1, then emit the jobs and search
// groups of this step; step 4: collect step 3, decide the macroblock mode, write the final motion.
__global__ void __launch_bounds__(128)
picture_step_kernel(const jsvm_pic_task* __restrict__ tasks, int n_pics, int mb_w, int mb_h,
                    const int2* __restrict__ wave_mbs, int n_wave, int step, int max_refs, int default_range, int cap_uh,
                    jsvm_pic_mb* __restrict__ field,               // [n_pics][mb_h * mb_w]
                    PicTemp* __restrict__ temp,                    // [n_pics][mb_h * mb_w]
                    const jsvm_me_result* __restrict__ prev_results,   // results of the previous step of this wave
                    jsvm_me_job* __restrict__ jobs, Group* __restrict__ groups)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pics * n_wave) return;
  const int pic = t / n_wave, i = t - pic * n_wave;
  const int mx = wave_mbs[i].x, my = wave_mbs[i].y;
  const jsvm_pic_task task = tasks[pic];
  const int R = task.n_refs;
  const size_t mb_idx = (size_t)pic * mb_w * mb_h + (size_t)my * mb_w + mx;
  jsvm_pic_mb* pf = field + (size_t)pic * mb_w * mb_h;
  PicTemp& tp = temp[mb_idx];

  // ---- collect the previous step: per partition the reference with the smallest cost ( uiCostTest < uiCost[0], first wins ) ----
  if (step > 0) {
    const int pj = pic_step_jobs(step - 1);
    for (int q = 0; q < pj; ++q) {
      const int part = c_pic_step_part[step - 1][q];
      uint32_t best = 0xffffffffu; int best_r = 0; jsvm_mv best_mv{0, 0};
      for (int r = 1; r <= R; ++r) {
        const jsvm_me_result& res = prev_results[((size_t)t * max_refs + (r - 1)) * pj + q];
        if (res.cost < best) { best = res.cost; best_r = r; best_mv = res.mv; }
      }
      tp.cost[part] = best; tp.ref[part] = (int8_t)best_r; tp.mv[part] = best_mv;
    }
  }

  if (step == 4) {
    // ---- macroblock mode by motion cost ----
    const uint32_t c16 = tp.cost[0], c168 = tp.cost[1] + tp.cost[2], c816 = tp.cost[3] + tp.cost[4],
                   c88 = tp.cost[5] + tp.cost[6] + tp.cost[7] + tp.cost[8];
    uint32_t best = c16; int mode = 1;
    if (c168 < best) { best = c168; mode = 2; }
    if (c816 < best) { best = c816; mode = 3; }
    if (c88 < best) { best = c88; mode = 4; }
    jsvm_pic_mb out;
    out.mode = (uint8_t)mode; out.pad[0] = out.pad[1] = out.pad[2] = 0;
    out.cost = best;
    for (int b8 = 0; b8 < 4; ++b8) {
      const int part = mode == 1 ? 0 : mode == 2 ? 1 + (b8 >> 1) : mode == 3 ? 3 + (b8 & 1) : 5 + b8;
      out.ref_idx[b8] = tp.ref[part];
      const int x0 = (b8 & 1) * 2, y0 = (b8 >> 1) * 2;
      for (int k = 0; k < 4; ++k) out.mv[(y0 + (k >> 1)) * 4 + x0 + (k & 1)] = tp.mv[part];
    }
    for (int p = 0; p < 9; ++p) { out.part_mv[p] = tp.mv[p]; out.part_ref[p] = tp.ref[p]; out.part_cost[p] = tp.cost[p]; }
    out.pad2[0] = out.pad2[1] = out.pad2[2] = 0;
    pf[my * mb_w + mx] = out;
    return;
  }

  // ---- neighbours of this step's partitions ----
  const int nj = pic_step_jobs(step);
  auto cur = [&](int part) { MvRef r{tp.mv[part].hor, tp.mv[part].ver, tp.ref[part]}; return r; };
  for (int r = 1; r <= R; ++r) {
    jsvm_me_job js[4];
    for (int q = 0; q < nj; ++q) {
      const int part = c_pic_step_part[step][q];
      MvRef A, B, C;
      int type = PIC_MEDIAN;
      switch (part) {
        case 0: case 1: case 3: case 5: {                 // partitions whose top-left 4x4 is block 0
          A = pic_neighbour(pf, mb_w, mb_h, mx - 1, my, 3);
          B = pic_neighbour(pf, mb_w, mb_h, mx, my - 1, 12);
          if (part <= 1) C = pic_neighbour(pf, mb_w, mb_h, mx + 1, my - 1, 12);       // 16 wide: above-right macroblock
          else C = pic_neighbour(pf, mb_w, mb_h, mx, my - 1, 14);                     // 8 wide: block above-right of 4x4 #1
          if (C.ref == 0) C = pic_neighbour(pf, mb_w, mb_h, mx - 1, my - 1, 15);
          type = part == 1 ? PIC_PRED_B : part == 3 ? PIC_PRED_A : PIC_MEDIAN;
        } break;
        case 2:                                            // 16x8 lower: C (right of the macroblock) never available -> D
          A = pic_neighbour(pf, mb_w, mb_h, mx - 1, my, 11); B = cur(1); C = pic_neighbour(pf, mb_w, mb_h, mx - 1, my, 7);
          type = PIC_PRED_A; break;
        case 4: case 6:                                    // 8x16 right / 8x8 #1
          A = cur(part == 4 ? 3 : 5);
          B = pic_neighbour(pf, mb_w, mb_h, mx, my - 1, 14);
          C = pic_neighbour(pf, mb_w, mb_h, mx + 1, my - 1, 12);
          if (C.ref == 0) C = pic_neighbour(pf, mb_w, mb_h, mx, my - 1, 13);
          type = part == 4 ? PIC_PRED_C : PIC_MEDIAN; break;
        case 7:                                            // 8x8 #2
          A = pic_neighbour(pf, mb_w, mb_h, mx - 1, my, 11); B = cur(5); C = cur(6); break;
        default:                                           // 8x8 #3: C not yet coded -> D = 8x8 #0
          A = cur(7); B = cur(6); C = cur(5); break;
      }
      const jsvm_mv mvp = pic_mvp(A, B, C, r, type);
      jsvm_me_job& j = js[q];
      j.cur_slot = task.cur_slot; j.ref_slot = task.ref_slot[r - 1];
      j.mb_x = (int16_t)mx; j.mb_y = (int16_t)my;
      j.blk = c_pic_part_blk[part]; j.mode = c_pic_part_mode[part];
      j.search_range = 0;
      j.start = mvp; j.pred = mvp;
      j.cost_factor = task.cost_factor;
      j.bits_in = c_pic_part_bits[part] + pic_ref_bits(r, R);
      j.org_index = -1;
    }
    // ---- the planner's work for these jobs: one group if the union of the clipped windows fits one CTA, else one group per job ----
    const size_t base = ((size_t)t * max_refs + (r - 1)) * nj;
    Window u{0, 0, 0, 0};
    for (int q = 0; q < nj; ++q) {
      const Window w1 = job_window(js[q], mb_w, mb_h, default_range);
      if (q == 0) u = w1;
      else { u.xlo = min(u.xlo, w1.xlo); u.xhi = max(u.xhi, w1.xhi); u.ylo = min(u.ylo, w1.ylo); u.yhi = max(u.yhi, w1.yhi); }
    }
    const int uw = u.xhi - u.xlo + 1, uh = u.yhi - u.ylo + 1;
    const bool together = uw <= 145 && uh <= cap_uh && uw * uh <= (1 << kIdxBits);
    for (int q = 0; q < nj; ++q) {
      jobs[base + q] = js[q];
      Group g;
      g.first_job = (int32_t)(base + (together ? 0 : q));
      g.gran4 = 0; g.cur_slot = js[q].cur_slot; g.ref_slot = js[q].ref_slot; g.mb_x = js[q].mb_x; g.mb_y = js[q].mb_y; g.org_index = -1;
      for (int s = 0; s < kSlots; ++s) g.slot_job[s] = -1;
      g.pad[0] = g.pad[1] = g.pad[2] = 0;
      if (together) {
        g.n_jobs = (int16_t)(q == 0 ? nj : 0);                                     // slots behind the first stay empty (K2 skips them)
        g.ux0 = (int16_t)u.xlo; g.uy0 = (int16_t)u.ylo; g.uw = (int16_t)uw; g.uh = (int16_t)uh;
        if (q == 0) for (int k = 0; k < nj; ++k) g.slot_job[partition_slot(js[k].mode, js[k].blk)] = (int8_t)k;
      } else {
        const Window w1 = job_window(js[q], mb_w, mb_h, default_range);
        g.n_jobs = 1;
        g.ux0 = (int16_t)w1.xlo; g.uy0 = (int16_t)w1.ylo; g.uw = (int16_t)(w1.xhi - w1.xlo + 1); g.uh = (int16_t)(w1.yhi - w1.ylo + 1);
        g.slot_job[partition_slot(js[q].mode, js[q].blk)] = 0;
      }
      groups[base + q] = g;
    }
  }
  // reference slots this picture does not use: empty groups
  for (int r = R + 1; r <= max_refs; ++r)
    for (int q = 0; q < nj; ++q) {
      Group g; memset(&g, 0, sizeof(g));
      groups[((size_t)t * max_refs + (r - 1)) * nj + q] = g;
    }
}

}  // namespace jsvm
