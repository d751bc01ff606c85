/*
 * jsvm_me_b200.h -- C ABI of the B200-native JSVM motion-estimation hot path.
 *
 * This is the drop-in boundary: plain C, opaque handle, int status, caller-owned buffers,
 * no torch / C++ types.  Every entry point names the reference interface it replaces
 * (paths relative to /root/reference/JSVM/H264Extension):
 *
 *   ENC = src/lib/H264AVCEncoderLib, COM = src/lib/H264AVCCommonLib, INCC = include/H264AVCCommonLib
 *
 * The reference-side binding (replacement bodies for MotionEstimation.cpp /
 * MotionEstimationQuarterPel.cpp / QuarterPelFilter::filterFrame that forward to these
 * functions while the class headers stay byte-identical) is shown in INTEGRATION.md and
 * implemented in jsvm-code-notes_b200/dropin/.
 *
 * Status convention = the reference's ErrVal (include/H264AVCCommonIf.h:26-38, Macros.h:23-43):
 * 0 (Err::m_nOK) on success, -1 (Err::m_nERR) on failure; jsvm_me_last_error() gives the text.
 * There is NO CPU fallback: every compute entry point fails if no sm_100 device is usable.
 */
#ifndef JSVM_ME_B200_H
#define JSVM_ME_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JSVM_ME_OK   0
#define JSVM_ME_ERR (-1)

/* INCC/CommonDefs.h:169-190 (MbMode / BlkMode values accepted by estimateBlockWithStart) */
#define JSVM_MODE_16x16 1
#define JSVM_MODE_16x8  2
#define JSVM_MODE_8x16  3
#define JSVM_BLK_8x8    8
#define JSVM_BLK_8x4    9
#define JSVM_BLK_4x8   10
#define JSVM_BLK_4x4   11

/* INCC/CommonDefs.h:208-223 */
#define JSVM_DF_SAD      0
#define JSVM_DF_SSD      1
#define JSVM_DF_HADAMARD 2
#define JSVM_DF_YUV_SAD  3
#define JSVM_SEARCH_BLOCK  0   /* SearchMode BLOCK_SEARCH = full search (xPelBlockSearch) */
#define JSVM_SEARCH_SPIRAL 1   /* xPelSpiralSearch */
#define JSVM_SEARCH_LOG    2   /* xPelLogSearch */
#define JSVM_SEARCH_FAST   3   /* xPelLogSearch seeded with the MV predictor and the neighbours' MVs */
#define JSVM_SEARCH_TZ     4   /* xTZSearch (the default of the shipped encoder.cfg) */

/* Per-call state of estimateBlockWithStart that the fast searches read, carried in the two top bits of jsvm_me_job.mode:
 * pcBSP != NULL (bi-predictive search: halves the first step of LOG / FAST, ENC/MotionEstimation.cpp:289, 296) and
 * m_bELWithBLMv (setEL( true ): TZ uses ELSearchRange, :1112-1117). */
#define JSVM_JOB_MODE_MASK 0x3f
#define JSVM_JOB_EL        0x40
#define JSVM_JOB_BIPRED    0x80

/* Luma plane geometry, INCC/CommonDefs.h:246-247 + COM/YuvBufferCtrl.cpp:134-157: stride = width + 2*YUV_X_MARGIN (32).
 * The reference ALLOCATES YUV_Y_MARGIN = 64 rows above and below the picture (m_uiLumBaseOffset = stride*64 + 32) but, for
 * frame pictures, fills and reads only m_uiYMargin = 64/2 = 32 of them (YuvPicBuffer::fillMargin, filterFrame uses getLXMargin for
 * both axes).  The device plane therefore keeps 32 margin samples on every side; jsvm_me_upload_plane_xpel addresses the source
 * relative to `origin` with the caller's stride, so a real YuvPicBuffer (64 allocated rows) is read correctly: rows -32..H+31. */
#define JSVM_ME_MARGIN 32

typedef struct jsvm_me_ctx jsvm_me_ctx;

/* INCC/Mv.h: two Shorts, quarter-pel units, arithmetic >> on negatives. */
typedef struct jsvm_mv { int16_t hor, ver; } jsvm_mv;

/*
 * One job == one call of
 *   MotionEstimation::estimateBlockWithStart( rcMbDataAccess, rcRefFrame, rcMv, rcMvPred,
 *                                             ruiBits, ruiCost, uiBlk, uiMode, uiSearchRange, pcPW, pcBSP )
 * (ENC/MotionEstimation.cpp:129-340) for the un-weighted / fWeight==1 path, plus the MB state
 * the reference keeps in members: the MB position that MotionCompensation::initMb
 * (COM/MotionCompensation.cpp:152-167) turned into m_cMin/m_cMax, and the cost factor that
 * xSetMEPars fetched from RateDistortionIf::getMotionCostShift (ENC/MotionEstimationCost.cpp:58-65).
 * Jobs that share (cur_slot, ref_slot, mb_x, mb_y, org_index) and sit next to each other in the
 * array are searched together from one SAD pass over the window (SAD reuse across partitions).
 */
typedef struct jsvm_me_job {
  int16_t  cur_slot;       /* plane slot of the original picture (XDistortion::loadOrgMbPelData source) */
  int16_t  ref_slot;       /* plane slot of rcRefFrame (full-pel + half-pel buffers) */
  int16_t  mb_x, mb_y;     /* initMb( uiMbY, uiMbX ) */
  uint8_t  blk;            /* uiBlk: 4x4 raster index 0..15 (B4x4Idx, INCC/CommonTypes.h:97-109) */
  uint8_t  mode;           /* uiMode: JSVM_MODE_* / JSVM_BLK_* */
  int16_t  search_range;   /* uiSearchRange; 0 = the value given to jsvm_me_set_params */
  jsvm_mv  start;          /* rcMv on entry (search start, quarter-pel) */
  jsvm_mv  pred;           /* rcMvPred (quarter-pel) */
  uint32_t cost_factor;    /* m_uiCostFactor for SAD/HADAMARD = floor(65536*sqrt(lambda)), ENC/RateDistortion.cpp:59-70;
                              computed on the host with the reference's own code (double, -ffloat-store) */
  uint32_t bits_in;        /* ruiBits on entry */
  int32_t  org_index;      /* -1: original block comes from cur_slot; >=0: index of a 16x16 int16 MB buffer in
                              org_mbs (weighted / bi-pred "2*org - pred" originals prepared by the caller,
                              ENC/MotionEstimation.cpp:163-247) */
} jsvm_me_job;

/* What one estimateBlockWithStart call returns (plus intermediate values for parity checks). */
typedef struct jsvm_me_result {
  jsvm_mv  mv;             /* rcMv on exit (quarter-pel) */
  jsvm_mv  mv_fullpel;     /* result of xPelBlockSearch in integer pels (before cMv <<= 2) */
  uint32_t sad_fullpel;    /* ruiSAD of xPelBlockSearch (rate already subtracted, :405) */
  uint32_t min_sad;        /* uiMinSAD after xSubPelSearch: distortion INCLUDING the rate term (N9) */
  uint32_t mv_bits;        /* uiMvBits = xGetBits( mv ) with the sub-pel table (shift 0) */
  uint32_t bits;           /* ruiBits on exit = bits_in + mv_bits */
  uint32_t cost;           /* ruiCost for fWeight == 1.0: (min_sad - cost(mv_bits)) + cost(bits)  (:335) */
  uint32_t best_mode;      /* m_uiBestMode: n (half-pel stage winner) or n<<4 (quarter-pel winner) */
} jsvm_me_result;

/* ---- life cycle: MotionEstimationQuarterPel::create / destroy (ENC/MotionEstimationQuarterPel.cpp:54-66,
 *      ENC/MotionEstimation.cpp:31-35) ---- */
int jsvm_me_create (jsvm_me_ctx** ctx, int device);
int jsvm_me_destroy(jsvm_me_ctx* ctx);
const char* jsvm_me_last_error(void);

/* Picture geometry and frame store: YuvBufferCtrl::initSPS( height, width, YUV_Y_MARGIN, YUV_X_MARGIN [,1] )
 * (ENC/ControlMngH264AVCEncoder.cpp:56-57) + allocation of n_slots Frames (full-pel and half-pel buffers).
 * width and height must be multiples of 16.  One ctx serves one resolution (one spatial layer). */
int jsvm_me_configure(jsvm_me_ctx* ctx, int width, int height, int n_slots);

/* MotionEstimation::init reading MotionVectorSearchParams (ENC/MotionEstimation.cpp:38-66,
 * include/CodingParameter.h:41-89; value ranges checked as MotionVectorSearchParams::check, ENC/CodingParameter.cpp:17-36).
 * SearchMode: BLOCK_SEARCH (0, the measured hot path), SPIRAL (1), LOG (2), FAST (3), TZ (4; one warp per job, see
 * jsvm_me_set_fast_search_params / jsvm_me_set_mv_candidates; LOG and FAST refuse YUV_SAD as MotionVectorSearchParams::check does).
 * SearchRange: 1..4096; the packed-byte full-search kernels stage windows of up to 129 x 129 positions, so BLOCK_SEARCH with a range
 * above 64 runs unstaged on the one-warp-per-job kernel (correct, not tuned; the reference has no limit either: its bit table is sized
 * from the range, ENC/MotionEstimationCost.cpp:27-45).
 * SearchFuncFullPel: SAD (0), SSD (1), HADAMARD (2), YUV_SAD (3); SearchFuncSubPel: SAD (0),
 * SSD (1), HADAMARD (2) -- every combination.  Full-pel SAD is the measured hot path (packed-byte kernels); the other full-pel
 * functions run a plain one-CTA-per-job kernel.  YUV_SAD needs the chroma planes of the slots (jsvm_me_upload_chroma_*).
 * A job carries ONE cost factor, the one of the FULL-PEL stage (ENC/MotionEstimation.cpp:255: the SAD factor unless
 * SearchFuncFullPel is SSD).  When the sub-pel stage uses the other factor class (:325: the SSE factor iff SearchFuncSubPel is
 * SSD) the pairs must be registered with jsvm_me_map_cost_factor. */
int jsvm_me_set_params(jsvm_me_ctx* ctx, int search_mode, int full_pel_dfunc, int sub_pel_dfunc, int search_range);

/* RateDistortionIf::getMotionCostShift( bSad ) for both classes (ENC/RateDistortion.cpp:59-70, RateDistortion.h:30): registers that
 * jobs whose cost_factor (full-pel stage) is `factor_full_pel` use `factor_sub_pel` in the sub-pel stage and in the final cost
 * (xSetMEPars( 0, ... ), ENC/MotionEstimation.cpp:325-335).  Both values come from the host (double arithmetic, N11).  Only
 * consulted when the two stages use different factor classes; up to 256 pairs per context, re-registering a factor replaces it. */
int jsvm_me_map_cost_factor(jsvm_me_ctx* ctx, uint32_t factor_full_pel, uint32_t factor_sub_pel);

/* The remaining MotionVectorSearchParams the fast searches read (include/CodingParameter.h:41-89): ELSearchRange (0 = SearchRange,
 * ENC/CodingParameter.cpp:25-28; used by TZ_SEARCH for jobs flagged JSVM_JOB_EL) and FastBiSearch (jobs with a per-call search
 * range -- the refinement searches of the bi-prediction iterations, ENC/MotionEstimation.cpp:264-274 -- run xTZSearch with an empty
 * candidate list instead of xPelBlockSearch). */
int jsvm_me_set_fast_search_params(jsvm_me_ctx* ctx, int el_search_range, int fast_bi_search);

/* MbDataAccess::addMvPredictors (COM/MbDataAccess.cpp:423-428): the MVs m_cMv3D_A / B / C of the left, above and above-right (or
 * above-left) neighbours that getMvPredictor left behind, three per job, quarter-pel.  FAST_SEARCH and TZ_SEARCH test them next
 * to rcMvPred (ENC/MotionEstimation.cpp:293-305).  HOST memory read by the next jsvm_me_search calls (must stay valid; job i uses
 * abc[3i .. 3i+2]); NULL = all zero.  The _device form takes device memory and serves jsvm_me_search_device. */
int jsvm_me_set_mv_candidates(jsvm_me_ctx* ctx, const jsvm_mv* abc, int n_jobs);
int jsvm_me_set_mv_candidates_device(jsvm_me_ctx* ctx, const jsvm_mv* d_abc);

/* Run every later call on this CUDA stream (a cudaStream_t / CUstream handle; NULL = default). */
int jsvm_me_set_stream(jsvm_me_ctx* ctx, void* cuda_stream);

/* Frame planes.  A slot holds the full-pel luma plane (with its 32-sample margin) and, after
 * jsvm_me_interp_halfpel, the half-pel plane (YuvBufferCtrl with uiResolution = 1).
 *
 * upload_plane_xpel: `origin` points at sample (0,0) of a reference YuvPicBuffer luma plane
 *   (XPel = int16, YuvPicBuffer::getMbLumAddr() after initMb(0,0)); the 32-sample margins around it
 *   must already be filled (YuvPicBuffer::fillMargin, COM/YuvPicBuffer.cpp:1595-1641) and are copied
 *   as they are.  Samples must be in [0,255] (reconstructed pictures are).
 * upload_plane_u8: dense 8-bit picture without margins; the margin is generated on the device exactly
 *   as fillMargin does (edge replication). `host` may be pinned memory. */
int jsvm_me_upload_plane_xpel(jsvm_me_ctx* ctx, int slot, const int16_t* origin, int width, int height, int stride);
int jsvm_me_upload_plane_u8  (jsvm_me_ctx* ctx, int slot, const uint8_t* host, int width, int height, int stride);
/* Same as upload_plane_u8 but the source is already in device memory (bench "resident in HBM" leg). */
int jsvm_me_set_plane_device_u8(jsvm_me_ctx* ctx, int slot, const uint8_t* dev, int width, int height, int stride);

/* Chroma planes of a slot (4:2:0; YuvPicBuffer Cb / Cr with stride (W+64)/2 and 16 margin samples, COM/YuvBufferCtrl.cpp:134-157,
 * INCC/YuvPicBuffer.h:32,68-69).  Read by the YUV_SAD full-pel search (xGetYuvSAD*, ENC/Distortion.cpp:435-505, 593-647, 718-764;
 * chroma displacement ( y>>1, x>>1 ), ENC/MotionEstimation.cpp:386-387) and by jsvm_me_compensate.
 * upload_chroma_u8  : dense (W/2) x (H/2) planes; margins are generated on the device as fillMargin does.
 * upload_chroma_xpel: `cb_origin` / `cr_origin` point at sample (0,0) of the YuvPicBuffer planes (getMbCbAddr / getMbCrAddr after
 *                     initMb(0,0)) whose 16-sample margins are already filled. */
int jsvm_me_upload_chroma_u8  (jsvm_me_ctx* ctx, int slot, const uint8_t* cb, const uint8_t* cr, int stride);
int jsvm_me_upload_chroma_xpel(jsvm_me_ctx* ctx, int slot, const int16_t* cb_origin, const int16_t* cr_origin, int stride);
/* Chroma originals of the caller-prepared macroblocks (pUOrg / pVOrg of a weighted original, ENC/MotionEstimation.cpp:259-260):
 * n_org_mbs x (Cb 8x8, Cr 8x8) int16, HOST memory, for jobs with org_index >= 0.  Only read by jsvm_me_search when
 * SearchFuncFullPel is YUV_SAD; the pointer must stay valid until that call returns.  NULL clears it. */
int jsvm_me_set_org_chroma(jsvm_me_ctx* ctx, const int16_t* org_chroma_mbs, int n_org_mbs);

/* QuarterPelFilter::filterFrame( full, half ) (COM/QuarterPelFilter.cpp:61-149): builds the half-pel
 * plane of the slot; it stays on the device (its only consumer is ME). */
int jsvm_me_interp_halfpel(jsvm_me_ctx* ctx, int slot);
/* The same for several slots in ONE kernel launch (grid.z = slot), e.g. all reference pictures of a temporal level. */
int jsvm_me_interp_halfpel_batch(jsvm_me_ctx* ctx, const int* slots, int n_slots);
/* Copy the half-pel plane back in YuvPicBuffer layout: `origin` = sample (0,0) of the half-pel buffer
 * (stride 2*(width+64)); the region the reference writes, x in [-56, 2(W+28)), y in [-56, 2(H+28)), is stored. */
int jsvm_me_download_halfpel_xpel(jsvm_me_ctx* ctx, int slot, int16_t* origin, int stride);
/* Dense copy of that same region as bytes: rows 2*(H+56), cols 2*(W+56). */
int jsvm_me_download_halfpel_u8(jsvm_me_ctx* ctx, int slot, uint8_t* dst, int dst_stride);

/* estimateBlockWithStart for n jobs (host buffers; H2D of jobs, kernels, D2H of results, then sync).
 * org_mbs: n_org_mbs dense 16x16 int16 blocks (or NULL).  preds (optional, may be NULL): n*256 bytes,
 * the winning prediction block of every job as compensateBlock would return it
 * (ENC/MotionEstimationQuarterPel.cpp:266-328), stride 16, top-left aligned. */
int jsvm_me_search(jsvm_me_ctx* ctx, const jsvm_me_job* jobs, int n_jobs,
                   const int16_t* org_mbs, int n_org_mbs,
                   jsvm_me_result* results, uint8_t* preds);

/* Same, everything already in device memory, asynchronous on the ctx stream (no sync).
 * The search groups come from the host-side planner jsvm_me_plan: it fuses neighbouring jobs of one
 * macroblock (same cur/ref slot, MB position and original) into one group that is evaluated from a
 * single SAD pass; the caller copies the records to the device. */
typedef struct jsvm_me_plan_info {
  int32_t n_groups;          /* records written to groups_out */
  int32_t n_groups_8x8;      /* the first n_groups_8x8 records need only 8x8-granular SADs (9-partition set) */
  int32_t max_uh_8x8;        /* tallest union window among those / among the 4x4-granular groups (sizes shared memory) */
  int32_t max_uh_4x4;
  int64_t n_positions;       /* sum over jobs of the clipped window size (ALG_OPS accounting) */
  int64_t n_group_positions; /* sum over groups of the union window size (positions actually evaluated) */
} jsvm_me_plan_info;
size_t jsvm_me_group_size(void);    /* bytes per group record */
int jsvm_me_plan(jsvm_me_ctx* ctx, const jsvm_me_job* jobs, int n_jobs, void* groups_out, jsvm_me_plan_info* info);
/* The same planner without a device context (pure host logic; groups_out must hold n_jobs records). */
int jsvm_me_plan_host(int width, int height, int n_slots, int search_range, const jsvm_me_job* jobs, int n_jobs,
                      void* groups_out, jsvm_me_plan_info* info);
int jsvm_me_search_device(jsvm_me_ctx* ctx, const jsvm_me_job* d_jobs, int n_jobs,
                          const void* d_groups, const jsvm_me_plan_info* info,
                          const int16_t* d_org_mbs, int n_org_mbs,
                          jsvm_me_result* d_results, uint8_t* d_preds);

/* MotionEstimationQuarterPel::compensateBlock for job `job_idx` of the LAST jsvm_me_search call
 * (the reference's stateful contract): writes w x h samples to dst (XPel, dst_stride). */
int jsvm_me_fetch_pred(jsvm_me_ctx* ctx, int job_idx, int16_t* dst, int dst_stride);

/* ---- motion-compensated prediction ----
 * One job == the prediction of one luma block and its two chroma blocks as MotionCompensation::compensateMb / compensateSubMb
 * produce it (COM/MotionCompensation.cpp:1200-1405): MV limited to the macroblock's range (xGetMbPredData -> limitComponents,
 * :1390), luma by QuarterPelFilter::predBlk (COM/QuarterPelFilter.cpp:266-619: 6-tap half samples, averaged quarter samples),
 * chroma by xPredChroma / xPredChromaPel (:171-260: 1/8-sample bilinear), then SampleWeighting::weightLumaSamples /
 * weightChromaSamples (COM/SampleWeighting.cpp:159-265, 332-372: xMixB, xMixBWeight, xWeight).
 * Luma comes from the slot's half-pel plane (the same integers predBlk computes from the full-pel plane: the plane IS
 * filterFrame's output), so the reference slots need jsvm_me_interp_halfpel; chroma needs jsvm_me_upload_chroma_*. */
/* SampleWeighting::initSlice state (COM/SampleWeighting.cpp:77-138): which prediction types are weighted in this slice */
#define JSVM_MC_DEFAULT    0   /* one list: plain prediction; two lists: ( a + b + 1 ) >> 1 (xMixB) */
#define JSVM_MC_WEIGHT_UNI 1   /* !m_bWeightedPredDisableP: one-list blocks go through xWeight when the list's weight_flag is set */
#define JSVM_MC_WEIGHT_BI  2   /* !m_bWeightedPredDisableB: two-list blocks go through xMixBWeight (explicit or implicit weights) */
typedef struct jsvm_mc_job {
  int16_t ref_slot[2];     /* LIST_0 / LIST_1 reference slot, -1 = list not used */
  int16_t mb_x, mb_y;      /* initMb( uiMbY, uiMbX ): MV clamp m_cMin / m_cMax and block position */
  uint8_t blk;             /* cIdx: 4x4 raster index of the block's top-left 4x4 */
  uint8_t size_x, size_y;  /* luma block size: 16, 8 or 4 (chroma: half) */
  uint8_t weighting;       /* JSVM_MC_DEFAULT or JSVM_MC_WEIGHT_UNI | JSVM_MC_WEIGHT_BI */
  jsvm_mv mv[2];           /* quarter-pel MV per list (before the clamp) */
  int16_t weight[2][3];    /* PredWeight::getLumaWeight / getChromaCbWeight / getChromaCrWeight per list */
  int16_t offset[2][3];    /* ... Offset */
  uint8_t log_denom[2];    /* m_uiLumaLogWeightDenom, m_uiChromaLogWeightDenom (5 / 5 for implicit weights) */
  uint8_t weight_flag[2];  /* one-list case: getLumaWeightFlag / getChromaWeightFlag of the list's PredWeight */
} jsvm_mc_job;
/* pred_y: n x 256 bytes (stride 16, top-left aligned, zero padded); pred_cb / pred_cr: n x 64 bytes (stride 8) or NULL. */
int jsvm_me_compensate(jsvm_me_ctx* ctx, const jsvm_mc_job* jobs, int n_jobs, uint8_t* pred_y, uint8_t* pred_cb, uint8_t* pred_cr);
/* The same with every buffer in device memory, asynchronous on the ctx stream. */
int jsvm_me_compensate_device(jsvm_me_ctx* ctx, const jsvm_mc_job* d_jobs, int n_jobs, uint8_t* d_pred_y, uint8_t* d_pred_cb, uint8_t* d_pred_cr);

/* ---- picture-level motion search: MV prediction and per-macroblock scheduling on the device ----
 * For every macroblock of a P picture, in the dependency order of the encoder (left, above, above-right / above-left neighbours
 * first), the list-0 loops of MbEncoder::xEstimateMb16x16 / 16x8 / 8x16 / 8x8 (ENC/MbEncoder.cpp:4449-4560, 4830-4935, 5234-5330,
 * 5637-5700, 5950-6040): for every partition and every reference frame the MV predictor of MbDataAccess::getMvPredictor
 * (INCC/MbDataAccess.h:289-296, 485-513, COM/MbDataAccess.cpp:137-187), estimateBlockWithStart with start = predictor and the mode /
 * reference-index bits of those loops, the reference with the smallest cost per partition, and -- instead of the rate-distortion
 * decision of xSetRdCostInterMb, which is outside this library -- the macroblock mode with the smallest SUM of motion costs
 * (strict '<' in the order 16x16, 16x8, 8x16, 8x8).  Needs SearchMode BLOCK_SEARCH, SearchFuncFullPel SAD, SearchRange <= 64.
 * All pictures of one call are independent of each other (a temporal level of hierarchical B pictures) and advance
 * together, wavefront by wavefront; they must use the same number of reference frames. */
#define JSVM_PIC_MAX_REFS 8
typedef struct jsvm_pic_task {
  int16_t  cur_slot;                       /* original picture */
  int16_t  n_refs;                         /* active entries of reference list 0 (1..8) */
  int16_t  ref_slot[JSVM_PIC_MAX_REFS];    /* RefFrameList0[ 1 .. n_refs ] */
  uint32_t cost_factor;                    /* SAD-class motion cost factor of the picture (see jsvm_me_job) */
} jsvm_pic_task;
typedef struct jsvm_pic_mb {
  uint8_t  mode;                           /* MbMode: 1 16x16, 2 16x8, 3 8x16, 4 8x8 (four BLK_8x8) */
  uint8_t  pad[3];
  int8_t   ref_idx[4];                     /* per 8x8 block (MbMotionData::m_ascRefIdx), 1-based */
  jsvm_mv  mv[16];                         /* per 4x4 block, raster order (MbMvData::m_acMv) */
  uint32_t cost;                           /* sum of the chosen mode's partition costs */
  jsvm_mv  part_mv[9];                     /* every partition that was estimated: 16x16 | 16x8 upper, lower | 8x16 left, right | 8x8 #0..#3 */
  int8_t   part_ref[9];
  int8_t   pad2[3];
  uint32_t part_cost[9];                   /* ruiCost of the winning reference (bits: mode + reference index + MV) */
} jsvm_pic_mb;
/* out: n_pics x (height/16 * width/16) records in raster order, HOST memory; NULL keeps them on the device (jsvm_me_download_pictures). */
int jsvm_me_estimate_pictures(jsvm_me_ctx* ctx, const jsvm_pic_task* tasks, int n_pics, jsvm_pic_mb* out);
int jsvm_me_download_pictures(jsvm_me_ctx* ctx, jsvm_pic_mb* out, int n_pics);

/* ---- inter-layer resampling for the enhancement-layer motion search ----
 * The fields of ResizeParameters the progressive resampling reads (include/ResizeParameters.h:70-101; both layers frame_mbs_only,
 * no field pictures, no MBAFF -- the five configurations are progressive). */
typedef struct jsvm_resize_params {
  int32_t level_idc;                             /* m_iLevelIdc: <= 30 selects 16-bit position arithmetic */
  int32_t frame_width, frame_height;             /* current (enhancement) layer: m_iFrameWidth / m_iFrameHeight */
  int32_t ref_frame_width, ref_frame_height;     /* reference (base) layer: m_iRefLayerFrmWidth / m_iRefLayerFrmHeight */
  int32_t scaled_ref_width, scaled_ref_height;   /* m_iScaledRefFrmWidth / m_iScaledRefFrmHeight */
  int32_t left_offset, top_offset;               /* m_iLeftFrmOffset / m_iTopFrmOffset */
  int32_t chroma_phase_x, chroma_phase_y;        /* m_iChromaPhaseX / Y */
  int32_t ref_chroma_phase_x, ref_chroma_phase_y;/* m_iRefLayerChromaPhaseX / Y */
} jsvm_resize_params;
/* Frame::residualUpsampling -> DownConvert::xResidualUpsampling (COM/Frame.cpp:398-406, COM/DownConvert.cpp:2476-2958): the
 * base-layer residual, upsampled to the enhancement-layer grid, that the residual-prediction passes subtract from the original
 * before the motion search (N3).  Dense int16 planes in HOST memory: base luma ref_w x ref_h, base chroma (ref_w/2) x (ref_h/2);
 * outputs frame_w x frame_h and (frame_w/2) x (frame_h/2).  base_transform8x8: one byte per base-layer macroblock, raster order
 * (MbData::isTransformSize8x8).  Chroma pointers may all be NULL (luma only). */
int jsvm_me_upsample_residual(jsvm_me_ctx* ctx, const jsvm_resize_params* params,
                              const int16_t* base_y, const int16_t* base_cb, const int16_t* base_cr, const uint8_t* base_transform8x8,
                              int16_t* out_y, int16_t* out_cb, int16_t* out_cr);

/* MbDataCtrl::upsampleMotion / MotionUpsampling::resample (COM/MbDataCtrl.cpp:85-1110, 1115-1140; subclauses G.6.1, G.6.2, G.8.6.1 of
 * the SVC specification): the inter-layer motion prediction -- for every enhancement-layer macroblock the partitioning, reference
 * indices and motion vectors derived from the base layer, which the enhancement-layer passes use as search start and MV
 * predictor (cBLMvPred / iBLRefIdx, ENC/MbEncoder.cpp:4483-4497, 4516-4560; estimateBlockWithStart with JSVM_JOB_EL).
 * Progressive layers, no per-picture cropping change (ESS_PICT), no coefficient prediction (SCoeff / TCoeff).
 * One record per macroblock, raster order; the same record describes the base layer (input) and the prediction (output). */
typedef struct jsvm_mb_motion {
  uint8_t  mb_mode;          /* MbMode (INCC/CommonDefs.h:169-180): 0 SKIP, 1 16x16, 2 16x8, 3 8x16, 4 8x8, 5 8x8ref0, >= 6 intra, 36 INTRA_BL */
  uint8_t  slice_type;       /* input: SliceType of the macroblock's slice (0 P, 1 B, 2 I); output: 0 */
  uint8_t  blk_mode[4];      /* BlkMode per 8x8: 8 8x8, 9 8x4, 10 4x8, 11 4x4, 0 BLK_SKIP */
  uint8_t  in_crop_window;   /* output: MbData::getInCropWindowFlag */
  uint8_t  res_pred_safe;    /* output: MbData::getSafeResPred */
  uint16_t fwd_bwd;          /* output: MbData::getFwdBwd */
  uint8_t  base_intra;       /* output: bit ( x + 2 y ) = m_aabBaseIntra[x][y] */
  uint8_t  pad;
  int8_t   ref_idx[2][4];    /* per list and 8x8 block; -1 = BLOCK_NOT_PREDICTED, 0 = BLOCK_NOT_AVAILABLE */
  jsvm_mv  mv[2][16];        /* per list and 4x4 block (raster) */
} jsvm_mb_motion;
typedef struct jsvm_motion_upsample_params {
  jsvm_resize_params resize;
  int32_t slice_type;            /* of the enhancement-layer slice: 0 P (list 0 only), 1 B, 2 I (no motion) */
  int32_t ref_layer_dq_id;       /* SliceHeader::getRefLayerDQId (B_Skip / direct partitions are only split when it is 0) */
  int32_t direct_8x8_inference;  /* SPS direct_8x8_inference_flag */
  int32_t check_residual_pred;   /* bCheckResidualPred (m_uiEssRPChkEnable) and its MV threshold (m_uiMVThres) */
  int32_t mv_threshold;
} jsvm_motion_upsample_params;
/* base: (ref_h/16) x (ref_w/16) records, out: (frame_h/16) x (frame_w/16) records, HOST memory. */
int jsvm_me_upsample_motion(jsvm_me_ctx* ctx, const jsvm_motion_upsample_params* params, const jsvm_mb_motion* base, jsvm_mb_motion* out);

/* ---- multi-GPU: one process per GPU, whole (current, reference) pairs per rank, ONE exchange: the gather of the result field to
 * the rank that owns mode decision.  NCCL over NVLink / NVSwitch, issued on the ctx stream; the library is resolved at run
 * time (libnccl.so.2), single-GPU use never touches it.  The 128-byte unique id travels by whatever the launcher offers
 * (bench.py: a torch.distributed broadcast). */
int jsvm_me_comm_unique_id(uint8_t id[128]);                                        /* ncclGetUniqueId, on the root rank */
int jsvm_me_comm_init(jsvm_me_ctx* ctx, int rank, int world, const uint8_t id[128]);
/* Every rank sends n_local records; the root receives rank r's records at d_all + r * n_max (n_max >= every rank's n_local; ranks
 * may hold different numbers of pairs).  d_all is only read on the root.  Asynchronous on the ctx stream. */
int jsvm_me_gather_results_device(jsvm_me_ctx* ctx, const jsvm_me_result* d_local, int n_local, const int* n_per_rank,
                                  jsvm_me_result* d_all, int n_max, int root);
int jsvm_me_comm_destroy(jsvm_me_ctx* ctx);

int jsvm_me_sync(jsvm_me_ctx* ctx);

/* Measurement helpers (bench.py): kernel launches issued so far, and CUDA-event timing of the
 * search / interpolation kernels on the ctx stream. */
uint64_t jsvm_me_launch_count(jsvm_me_ctx* ctx);
/* Name of the full-search kernel the last search launched (for the 8x8-granular groups if there were any, else for the
 * 4x4-granular ones): "full_search_persistent_kernel<V,GRAN4>" (one CTA per SM over two shared-memory stages) or
 * "full_search_kernel<V,GRAN4>" (one CTA per group: small launches, union windows too tall for two stages);
 * a host-buffer call that was served by several launches names the kernel of its largest launch; "" before the first search. */
const char* jsvm_me_last_search_kernel(jsvm_me_ctx* ctx);
int jsvm_me_timing_begin(jsvm_me_ctx* ctx);                  /* start accumulating per-kernel event times */
int jsvm_me_timing_end(jsvm_me_ctx* ctx, double* ms_interp, double* ms_search, double* ms_subpel,
                       uint64_t* n_interp, uint64_t* n_search, uint64_t* n_subpel);
/* Time and launch count of the motion-compensation kernel inside the last timing_begin / timing_end bracket. */
int jsvm_me_timing_compensate(jsvm_me_ctx* ctx, double* ms_compensate, uint64_t* n_compensate);

#ifdef __cplusplus
}
#endif
#endif /* JSVM_ME_B200_H */
