/*
 * oracle/me_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the JSVM motion-estimation hot path (see me_oracle.c for the
 * reference file:line each function follows).  Parity status: PINNED -- every function is
 * checked against the compiled, unmodified reference (oracle/_ref/libjsvm_ref.so, built from
 * /root/reference by oracle/Makefile) in tests/test_oracle_vs_reference.py, and against the
 * golden fixtures under tests/golden/ that were generated from that reference build by
 * tests/golden/make_golden.py.  The reference ships no golden vectors of its own.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product never links, imports or executes anything in oracle/.
 *
 * Same entry points as ref_harness.h with the prefix jsvm_ora_.
 */
#ifndef JSVM_ME_ORACLE_H
#define JSVM_ME_ORACLE_H

#include "../include/jsvm_me_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jsvm_ora_ctx jsvm_ora_ctx;

jsvm_ora_ctx* jsvm_ora_create(int width, int height, int n_slots,
                              int search_mode, int full_pel_dfunc, int sub_pel_dfunc, int search_range);
void jsvm_ora_destroy(jsvm_ora_ctx* ctx);
int  jsvm_ora_set_plane_u8(jsvm_ora_ctx* ctx, int slot, const uint8_t* luma, int stride);
int  jsvm_ora_get_fullpel_u8(jsvm_ora_ctx* ctx, int slot, uint8_t* dst, int dst_stride);
int  jsvm_ora_get_halfpel_u8(jsvm_ora_ctx* ctx, int slot, uint8_t* dst, int dst_stride);
int  jsvm_ora_search(jsvm_ora_ctx* ctx, const jsvm_me_job* jobs, int n_jobs,
                     const int16_t* org_mbs, int n_org_mbs,
                     jsvm_me_result* results, uint8_t* preds);
uint32_t jsvm_ora_cost_factor(double lambda, int sad);
uint32_t jsvm_ora_distortion(const int16_t* org16x16, const int16_t* ref, int ref_stride, int mode, int dfunc, int blk);
/* signed Exp-Golomb code length of one quarter-pel MV difference component */
uint32_t jsvm_ora_component_bits(int d);
/* number of integer search positions of one job (clipped window size), for ALG_OPS accounting */
int jsvm_ora_window_positions(jsvm_ora_ctx* ctx, const jsvm_me_job* job);

/* chroma planes, chroma originals, cost-factor pairs, motion-compensated prediction: see ref_harness.h */
int jsvm_ora_set_chroma_u8(jsvm_ora_ctx* ctx, int slot, const uint8_t* cb, const uint8_t* cr, int stride);
int jsvm_ora_set_org_chroma(jsvm_ora_ctx* ctx, const int16_t* org_chroma_mbs, int n_org_mbs);
int jsvm_ora_map_cost_factor(jsvm_ora_ctx* ctx, uint32_t factor_full_pel, uint32_t factor_sub_pel);
int jsvm_ora_compensate(jsvm_ora_ctx* ctx, const jsvm_mc_job* jobs, int n_jobs, uint8_t* pred_y, uint8_t* pred_cb, uint8_t* pred_cr);

/* ELSearchRange / FastBiSearch of MotionVectorSearchParams; the three neighbour MVs of every job (3 per job, host pointer kept) */
int jsvm_ora_set_fast_search_params(jsvm_ora_ctx* ctx, int el_search_range, int fast_bi_search);
int jsvm_ora_set_mv_candidates(jsvm_ora_ctx* ctx, const jsvm_mv* abc, int n_jobs);

/* picture-level motion search (MV prediction + per-macroblock loops), see jsvm_me_estimate_pictures */
int jsvm_ora_estimate_pictures(jsvm_ora_ctx* ctx, const jsvm_pic_task* tasks, int n_pics, jsvm_pic_mb* out);

/* residual upsampling of progressive layers (DownConvert::xResidualUpsampling), see jsvm_me_upsample_residual */
int jsvm_ora_upsample_residual(const jsvm_resize_params* params, const int16_t* base_y, const int16_t* base_cb, const int16_t* base_cr,
                               const uint8_t* base_transform8x8, int16_t* out_y, int16_t* out_cb, int16_t* out_cr);

/* inter-layer motion prediction of progressive layers (MbDataCtrl::upsampleMotion), see jsvm_me_upsample_motion */
int jsvm_ora_upsample_motion(const jsvm_motion_upsample_params* params, const jsvm_mb_motion* base, jsvm_mb_motion* out);

#ifdef __cplusplus
}
#endif
#endif
