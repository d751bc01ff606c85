/*
 * oracle/ref_harness.h -- TEST INFRASTRUCTURE ONLY.
 *
 * C entry points of oracle/_ref/libjsvm_ref.so: the UNMODIFIED JSVM reference classes
 * (compiled from /root/reference by oracle/Makefile) driven on caller-supplied planes and jobs.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (jsvm-code-notes_b200/) never does.
 *
 * Job / result records are the ones of the product ABI (include/jsvm_me_b200.h) so that the
 * reference, the C restatement (me_oracle.c) and the CUDA path can be compared field by field.
 */
#ifndef JSVM_REF_HARNESS_H
#define JSVM_REF_HARNESS_H

#include "../include/jsvm_me_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jsvm_ref_ctx jsvm_ref_ctx;

/* Builds the object graph of CreaterH264AVCEncoder::init (ENC/CreaterH264AVCEncoder.cpp:263-316) /
 * MCTFPreProcessor::init (TEST/MCTFPreProcessor/MCTFPreProcessor.cpp:83-142) for one picture size. */
jsvm_ref_ctx* jsvm_ref_create(int width, int height, int n_slots,
                              int search_mode, int full_pel_dfunc, int sub_pel_dfunc, int search_range);
void jsvm_ref_destroy(jsvm_ref_ctx* ctx);

/* Loads an 8-bit luma picture (chroma = 128) into slot's Frame and runs
 * Frame::extendFrame( QuarterPelFilter* ) = fillMargin + filterFrame (COM/Frame.cpp:296-373). */
int jsvm_ref_set_plane_u8(jsvm_ref_ctx* ctx, int slot, const uint8_t* luma, int stride);

/* Full-pel plane including the 32-sample margins: (H+64) rows x (W+64) bytes. */
int jsvm_ref_get_fullpel_u8(jsvm_ref_ctx* ctx, int slot, uint8_t* dst, int dst_stride);
/* Region of the half-pel plane that filterFrame writes: 2(H+56) rows x 2(W+56) bytes.
 * Returns the number of samples that were outside [0,255] (must be 0). */
int jsvm_ref_get_halfpel_u8(jsvm_ref_ctx* ctx, int slot, uint8_t* dst, int dst_stride);

/* estimateBlockWithStart (+ compensateBlock when preds != NULL) for every job, in array order. */
int jsvm_ref_search(jsvm_ref_ctx* ctx, const jsvm_me_job* jobs, int n_jobs,
                    const int16_t* org_mbs, int n_org_mbs,
                    jsvm_me_result* results, uint8_t* preds);

/* RateDistortion::setMbQpLambda -> getMotionCostShift( bSad ) (ENC/RateDistortion.cpp:59-70). */
uint32_t jsvm_ref_cost_factor(double lambda, int sad);

/* XDistortion function table on raw blocks (SAD / SSD / HADAMARD), for unit parity of a3/a4:
 * org: 16x16 int16 dense; ref: pointer + stride. */
uint32_t jsvm_ref_distortion(const int16_t* org16x16, const int16_t* ref, int ref_stride, int mode, int dfunc, int blk);

/* Chroma planes of a slot (dense (W/2) x (H/2) bytes); jsvm_ref_set_plane_u8 sets them to 128.  Margins are refilled. */
int jsvm_ref_set_chroma_u8(jsvm_ref_ctx* ctx, int slot, const uint8_t* cb, const uint8_t* cr, int stride);
/* Chroma originals (Cb 8x8, Cr 8x8 int16 per macroblock) of the caller-prepared originals, for SearchFuncFullPel = YUV_SAD.
 * Without them a job with org_index >= 0 is searched with the luma SAD (ENC/MotionEstimation.cpp:250-253). */
int jsvm_ref_set_org_chroma(jsvm_ref_ctx* ctx, const int16_t* org_chroma_mbs, int n_org_mbs);
/* The cost factor of the other class for jobs carrying `factor_full_pel` (see jsvm_me_map_cost_factor). */
int jsvm_ref_map_cost_factor(jsvm_ref_ctx* ctx, uint32_t factor_full_pel, uint32_t factor_sub_pel);
/* MotionCompensation::xPredLuma / xPredChroma (predBlk, xPredChromaPel, SampleWeighting) for every job. */
int jsvm_ref_compensate(jsvm_ref_ctx* ctx, const jsvm_mc_job* jobs, int n_jobs, uint8_t* pred_y, uint8_t* pred_cb, uint8_t* pred_cr);

/* MotionVectorSearchParams ELSearchRange (0 = SearchRange) / FastBiSearch; the neighbours' MVs (A, B, C per job) for FAST / TZ. */
int jsvm_ref_set_fast_search_params(jsvm_ref_ctx* ctx, int el_search_range, int fast_bi_search);
int jsvm_ref_set_mv_candidates(jsvm_ref_ctx* ctx, const jsvm_mv* abc, int n_jobs);

/* Picture-level motion search with the reference's MV prediction (see jsvm_me_estimate_pictures). */
int jsvm_ref_estimate_pictures(jsvm_ref_ctx* ctx, const jsvm_pic_task* tasks, int n_pics, jsvm_pic_mb* out);

/* Frame::residualUpsampling (DownConvert::xResidualUpsampling) on dense int16 planes; see jsvm_me_upsample_residual. */
int jsvm_ref_upsample_residual(const jsvm_resize_params* params, const int16_t* base_y, const int16_t* base_cb, const int16_t* base_cr,
                               const uint8_t* base_transform8x8, int16_t* out_y, int16_t* out_cb, int16_t* out_cr);

/* MbDataCtrl::upsampleMotion on caller-supplied base-layer motion; see jsvm_me_upsample_motion. */
int jsvm_ref_upsample_motion(const jsvm_motion_upsample_params* params, const jsvm_mb_motion* base, jsvm_mb_motion* out);

#ifdef __cplusplus
}
#endif
#endif
