// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (see ref_harness.h).
//
// Drives the UNMODIFIED reference classes.  Object graph and call order follow
//   TEST/MCTFPreProcessor/MCTFPreProcessor.cpp:83-142 (create / init wiring),
//   TEST/MCTFPreProcessor/MCTF.cpp:96-145, 248-277, 451-497 (SPS/PPS/SliceHeader, per-MB initMb loop),
//   ENC/CreaterH264AVCEncoder.cpp:263-316.
// Nothing from the reference is copied; this file only calls its public / protected interfaces.
#include "H264AVCEncoderLib.h"
#include "CodingParameter.h"
#include "H264AVCCommonLib/Transform.h"
#include "H264AVCCommonLib/YuvBufferCtrl.h"
#include "H264AVCCommonLib/QuarterPelFilter.h"
#include "H264AVCCommonLib/SampleWeighting.h"
#include "H264AVCCommonLib/Frame.h"
#include "H264AVCCommonLib/MbDataCtrl.h"
#include "H264AVCCommonLib/SliceHeader.h"
#include "H264AVCCommonLib/SequenceParameterSet.h"
#include "H264AVCCommonLib/PictureParameterSet.h"
#include "DownConvert.h"
#include "ResizeParameters.h"
#include "MotionEstimation.h"
#include "MotionEstimationQuarterPel.h"
#include "RateDistortion.h"
#include "Distortion.h"

#include <vector>
#include <cstring>
#include <cstdio>
#include <cstdlib>

#include "ref_harness.h"

using namespace h264;

namespace {

// RateDistortionIf whose motion cost factor is supplied per job (the job carries the value the
// real RateDistortion::setMbQpLambda would have produced; jsvm_ref_cost_factor pins that mapping).
class FixedRateDistortion : public RateDistortionIf {
public:
  FixedRateDistortion() : m_uiFactor(0), m_uiFactorSse(0) {}
  virtual ~FixedRateDistortion() {}
  ErrVal  setMbQpLambda(MbDataAccess&, UInt, Double) { return Err::m_nOK; }
  Double  getCost (UInt, UInt) { return 0; }
  Double  getFCost(UInt, UInt) { return 0; }
  UInt    getMotionCostShift(Bool bSad) { return bSad ? m_uiFactor : m_uiFactorSse; }
  ErrVal  fixMacroblockQP(MbDataAccess&) { return Err::m_nOK; }
  UInt m_uiFactor;       // SAD-class factor
  UInt m_uiFactorSse;    // SSE-class factor
};

// Probe subclass: records what xPelBlockSearch handed to xSubPelSearch and what came back.
class ProbeME : public MotionEstimationQuarterPel {
public:
  ProbeME() { xInitBuffer(); }
  virtual ~ProbeME() {}
  Void xSubPelSearch(YuvPicBuffer* pcPelData, Mv& rcMv, UInt& ruiSAD, UInt uiBlk, UInt uiMode) {
    m_cFullPelMv  = rcMv;     // already << 2
    m_uiFullPelSad = ruiSAD;
    MotionEstimationQuarterPel::xSubPelSearch(pcPelData, rcMv, ruiSAD, uiBlk, uiMode);
    m_uiMinSad = ruiSAD;
  }
  UInt bestMode() const { return m_uiBestMode; }
  UInt fullPelDFunc() const { return m_cParams.getFullPelDFunc(); }
  UInt subPelDFunc() const { return m_cParams.getSubPelDFunc(); }
  Void setFullPelDFunc(UInt ui) { m_cParams.setFullPelDFunc(ui); }
  Void setFastSearchParams(UInt uiEL, UInt uiFastBi) { m_cParams.setELSearchRange(uiEL); m_cParams.setFastBiSearch(uiFastBi); }
  UInt searchRange() const { return m_cParams.getSearchRange(); }
  // MotionCompensation::compensateMb's inner sequence for one block (COM/MotionCompensation.cpp:1217-1224, 1363-1405):
  // limitComponents, xPredLuma (predBlk + weightLumaSamples), xPredChroma (xPredChromaPel + weightChromaSamples)
  Void predictBlock(YuvMbBuffer* pcRec, UInt uiBlk, Int iSizeX, Int iSizeY, YuvPicBuffer* apcRef[2], const Mv acMv[2], const PredWeight* apcPW[2], Bool bChroma) {
    MC8x8 cMC(B_8x8_0);
    cMC.m_cIdx = B4x4Idx(uiBlk);
    for (Int n = 0; n < 2; n++) {
      if (!apcRef[n]) continue;
      cMC.m_aacMv[n][0].set(acMv[n], 1);
      cMC.m_aacMv[n][0].limitComponents(m_cMin, m_cMax);
      cMC.m_sChromaOffset[n] = 0;
      cMC.m_apcRefBuffer[n] = apcRef[n];
      cMC.m_apcPW[n] = apcPW[n];
    }
    xPredLuma(pcRec, iSizeX, iSizeY, cMC);
    if (bChroma) xPredChroma(pcRec, iSizeX >> 1, iSizeY >> 1, cMC);
  }
  Mv   m_cFullPelMv;
  UInt m_uiFullPelSad;
  UInt m_uiMinSad;
};

}  // namespace

struct jsvm_ref_ctx {
  int width, height, n_slots;
  UInt mbw, mbh;
  SequenceParameterSet* sps;
  PictureParameterSet*  pps;
  SliceHeader*          sh;
  MbDataCtrl*           mbctrl;
  YuvBufferCtrl*        full_ctrl;
  YuvBufferCtrl*        half_ctrl;
  QuarterPelFilter*     filter;
  Transform*            transform;
  SampleWeighting*      weighting;
  XDistortion*          dist;
  FixedRateDistortion*  rd;
  ProbeME*              me;
  CodingParameter*      cp;
  std::vector<Frame*>   frames;
  PredWeight            pw;
  std::vector<std::pair<uint32_t, uint32_t> > factor_map;   // full-pel stage factor -> sub-pel stage factor (other class)
  const int16_t*        org_chroma;
  int                   n_org_chroma;
  const jsvm_mv*        cands;              // 3 per job: the neighbours' MVs MbDataAccess::addMvPredictors hands to FAST / TZ
  int                   n_cands;
};

#define CHK(x) do { if ((x) != Err::m_nOK) { fprintf(stderr, "ref_harness: %s failed (%s:%d)\n", #x, __FILE__, __LINE__); return 0; } } while (0)
#define CHKI(x) do { if ((x) != Err::m_nOK) { fprintf(stderr, "ref_harness: %s failed (%s:%d)\n", #x, __FILE__, __LINE__); return -1; } } while (0)

extern "C" jsvm_ref_ctx* jsvm_ref_create(int width, int height, int n_slots,
                                         int search_mode, int full_pel_dfunc, int sub_pel_dfunc, int search_range) {
  if (width <= 0 || height <= 0 || (width & 15) || (height & 15) || n_slots <= 0) return 0;
  jsvm_ref_ctx* c = new jsvm_ref_ctx();
  c->width = width; c->height = height; c->n_slots = n_slots; c->org_chroma = 0; c->n_org_chroma = 0; c->cands = 0; c->n_cands = 0;
  c->mbw = width >> 4; c->mbh = height >> 4;

  CHK(SequenceParameterSet::create(c->sps));
  CHK(PictureParameterSet ::create(c->pps));
  c->sps->setFrameWidthInMbs (c->mbw);
  c->sps->setFrameHeightInMbs(c->mbh);
  c->sps->setDirect8x8InferenceFlag(true);
  c->pps->setNumRefIdxActive(LIST_0, 1);
  c->pps->setNumRefIdxActive(LIST_1, 1);
  c->pps->setConstrainedIntraPredFlag(true);
  c->pps->setNumSliceGroupsMinus1(0);

  CHK(YuvBufferCtrl::create(c->full_ctrl));
  CHK(YuvBufferCtrl::create(c->half_ctrl));
  CHK(c->full_ctrl->initSPS(height, width, YUV_Y_MARGIN, YUV_X_MARGIN));
  CHK(c->half_ctrl->initSPS(height, width, YUV_Y_MARGIN, YUV_X_MARGIN, 1));

  c->sh = new SliceHeader(*c->sps, *c->pps);
  c->sh->setNalUnitType(NAL_UNIT_CODED_SLICE_SCALABLE);
  c->sh->setFirstMbInSlice(0);
  c->sh->setLastMbInSlice(c->mbw * c->mbh - 1);
  c->sh->setNumMbsInSlice(c->mbw * c->mbh);
  c->sh->setSliceType(P_SLICE);
  c->sh->setDirectSpatialMvPredFlag(true);
  c->sh->setNumRefIdxActive(LIST_0, 1);
  c->sh->setNumRefIdxActive(LIST_1, 0);
  CHK(c->sh->getPredWeightTable(LIST_1).initDefaults(c->sh->getLumaLog2WeightDenom(), c->sh->getChromaLog2WeightDenom()));
  CHK(c->sh->getPredWeightTable(LIST_0).initDefaults(c->sh->getLumaLog2WeightDenom(), c->sh->getChromaLog2WeightDenom()));
  c->sh->setSliceGroupChangeCycle(1);
  c->sh->FMOInit();

  c->mbctrl = new MbDataCtrl();
  CHK(c->mbctrl->init(*c->sps));
  CHK(c->mbctrl->initSlice(*c->sh, PRE_PROCESS, false, NULL));

  CHK(QuarterPelFilter::create(c->filter));
  CHK(Transform       ::create(c->transform));
  CHK(SampleWeighting ::create(c->weighting));
  CHK(XDistortion     ::create(c->dist));
  c->rd = new FixedRateDistortion();
  c->me = new ProbeME();
  CHK(c->dist     ->init());
  CHK(c->weighting->init());
  CHK(c->filter   ->init());

  c->cp = new CodingParameter();
  MotionVectorSearchParams& p = c->cp->getMotionVectorSearchParams();
  p.setSearchMode  (search_mode);
  p.setFullPelDFunc(full_pel_dfunc);
  p.setSubPelDFunc (sub_pel_dfunc);
  p.setSearchRange (search_range);
  CHK(c->me->init(c->dist, c->cp, c->rd, c->filter, c->transform, c->weighting));
  CHK(c->me->initSlice(*c->sh));

  for (int s = 0; s < n_slots; ++s) {
    Frame* f = new Frame(*c->full_ctrl, *c->half_ctrl, FRAME, 0);
    CHK(f->init());
    CHK(f->initHalfPel());
    c->frames.push_back(f);
  }
  return c;
}

extern "C" void jsvm_ref_destroy(jsvm_ref_ctx* c) {
  if (!c) return;
  const bool dbg = getenv("JSVM_REF_DEBUG") != 0;
#define STEP(x) do { if (dbg) { fprintf(stderr, "destroy: %s\n", #x); fflush(stderr); } x; } while (0)
  for (size_t i = 0; i < c->frames.size(); ++i) {
    STEP(c->frames[i]->uninitHalfPel());
    STEP(c->frames[i]->uninit());
    STEP(delete c->frames[i]);
  }
  STEP(c->me->uninit());
  STEP(delete c->me);
  STEP(delete c->rd);
  STEP(delete c->cp);
  STEP(c->dist->destroy());
  STEP(c->weighting->uninit()); STEP(c->weighting->destroy());
  STEP(c->filter->uninit());    STEP(c->filter->destroy());
  STEP(c->transform->destroy());
  STEP(c->mbctrl->uninit()); STEP(delete c->mbctrl);
  STEP(delete c->sh);
  STEP(c->full_ctrl->uninit()); STEP(c->full_ctrl->destroy());
  STEP(c->half_ctrl->uninit()); STEP(c->half_ctrl->destroy());
  // SequenceParameterSet::destroy() deletes m_pcVUI and then `delete this` runs a destructor that touches it again
  // (COM/SequenceParameterSet.cpp:142-150): the parameter sets are deliberately leaked (a few hundred bytes).
  delete c;
#undef STEP
}

extern "C" int jsvm_ref_set_plane_u8(jsvm_ref_ctx* c, int slot, const uint8_t* luma, int stride) {
  if (!c || slot < 0 || slot >= c->n_slots) return -1;
  Frame* f = c->frames[slot];
  CHKI(c->full_ctrl->initMb());
  CHKI(c->half_ctrl->initMb());
  YuvPicBuffer* pb = f->getFullPelYuvBuffer();
  XPel* y = pb->getMbLumAddr();
  const int ls = pb->getLStride();
  for (int r = 0; r < c->height; ++r)
    for (int x = 0; x < c->width; ++x) y[r * ls + x] = luma[r * stride + x];
  XPel* cb = pb->getMbCbAddr();
  XPel* cr = pb->getMbCrAddr();
  const int cs = pb->getCStride();
  for (int r = 0; r < c->height / 2; ++r)
    for (int x = 0; x < c->width / 2; ++x) { cb[r * cs + x] = 128; cr[r * cs + x] = 128; }
  CHKI(f->extendFrame(c->filter));
  return 0;
}

extern "C" int jsvm_ref_get_fullpel_u8(jsvm_ref_ctx* c, int slot, uint8_t* dst, int dst_stride) {
  if (!c || slot < 0 || slot >= c->n_slots) return -1;
  CHKI(c->full_ctrl->initMb());
  YuvPicBuffer* pb = c->frames[slot]->getFullPelYuvBuffer();
  const XPel* y = pb->getMbLumAddr();
  const int ls = pb->getLStride();
  const int m = pb->getLXMargin();
  int bad = 0;
  for (int r = -m; r < c->height + m; ++r)
    for (int x = -m; x < c->width + m; ++x) {
      int v = y[r * ls + x];
      if (v < 0 || v > 255) ++bad;
      dst[(r + m) * dst_stride + (x + m)] = (uint8_t)v;
    }
  return bad;
}

extern "C" int jsvm_ref_get_halfpel_u8(jsvm_ref_ctx* c, int slot, uint8_t* dst, int dst_stride) {
  if (!c || slot < 0 || slot >= c->n_slots) return -1;
  CHKI(c->half_ctrl->initMb());
  YuvPicBuffer* pb = c->frames[slot]->getHalfPelYuvBuffer();
  const XPel* y = pb->getMbLumAddr();
  const int ls = pb->getLStride();
  const int m = 2 * (JSVM_ME_MARGIN - 4);   // filterFrame writes x in [-56, 2(W+28)), y in [-56, 2(H+28))
  int bad = 0;
  for (int r = -m; r < 2 * c->height + m; ++r)
    for (int x = -m; x < 2 * c->width + m; ++x) {
      int v = y[r * ls + x];
      if (v < 0 || v > 255) ++bad;
      dst[(r + m) * dst_stride + (x + m)] = (uint8_t)v;
    }
  return bad;
}

extern "C" int jsvm_ref_search(jsvm_ref_ctx* c, const jsvm_me_job* jobs, int n_jobs,
                               const int16_t* org_mbs, int n_org_mbs,
                               jsvm_me_result* results, uint8_t* preds) {
  if (!c) return -1;
  YuvMbBuffer rec;
  for (int i = 0; i < n_jobs; ++i) {
    const jsvm_me_job& j = jobs[i];
    if (j.cur_slot < 0 || j.cur_slot >= c->n_slots || j.ref_slot < 0 || j.ref_slot >= c->n_slots) return -1;
    if (j.mb_x < 0 || j.mb_x >= (int)c->mbw || j.mb_y < 0 || j.mb_y >= (int)c->mbh) return -1;
    if (j.org_index >= n_org_mbs) return -1;

    MbDataAccess* mba = 0;
    CHKI(c->mbctrl   ->initMb(mba, j.mb_y, j.mb_x));
    CHKI(c->full_ctrl->initMb(j.mb_y, j.mb_x, false));
    CHKI(c->half_ctrl->initMb(j.mb_y, j.mb_x, false));
    CHKI(c->me       ->initMb(j.mb_y, j.mb_x, *mba));

    YuvMbBuffer* org = 0;
    c->dist->loadOrgMbPelData(c->frames[j.cur_slot]->getFullPelYuvBuffer(), org);
    if (j.org_index >= 0) {
      const int16_t* src = org_mbs + (size_t)j.org_index * 256;
      XPel* d = org->getMbLumAddr();
      for (int r = 0; r < 16; ++r)
        for (int x = 0; x < 16; ++x) d[r * org->getLStride() + x] = src[r * 16 + x];
    }

    // the job carries the factor of the full-pel stage (:255: SSE class iff SearchFuncFullPel is SSD); the other class comes
    // from the registered pairs (identity if none: both stages of one class)
    uint32_t other = j.cost_factor;
    for (size_t k = 0; k < c->factor_map.size(); ++k) if (c->factor_map[k].first == j.cost_factor) other = c->factor_map[k].second;
    const bool full_sse = c->me->fullPelDFunc() == DF_SSD;
    c->rd->m_uiFactor    = full_sse ? other : j.cost_factor;
    c->rd->m_uiFactorSse = full_sse ? j.cost_factor : other;
    const UInt cfg_full = c->me->fullPelDFunc();
    if (cfg_full == DF_YUV_SAD && j.org_index >= 0) {
      if (c->org_chroma) {
        if (j.org_index >= c->n_org_chroma) return -1;
        const int16_t* src = c->org_chroma + (size_t)j.org_index * 128;
        XPel* u = org->getMbCbAddr(); XPel* v = org->getMbCrAddr();
        for (int r = 0; r < 8; ++r)
          for (int x = 0; x < 8; ++x) { u[r * org->getCStride() + x] = src[r * 8 + x]; v[r * org->getCStride() + x] = src[64 + r * 8 + x]; }
      } else {
        c->me->setFullPelDFunc(DF_SAD);     // what ENC/MotionEstimation.cpp:250-253 does for bi-pred / unequally weighted originals
      }
    }
    Mv   mv(j.start.hor, j.start.ver);
    Mv   pred(j.pred.hor, j.pred.ver);
    UInt bits = j.bits_in, cost = 0;
    // the neighbours' MVs that FAST_SEARCH / TZ_SEARCH read through MbDataAccess::addMvPredictors (public members, normally left
    // there by getMvPredictor)
    for (int k = 0; k < 3; ++k) {
      const jsvm_mv m = (c->cands && i < c->n_cands) ? c->cands[3 * (size_t)i + k] : jsvm_mv{0, 0};
      (k == 0 ? mba->m_cMv3D_A : k == 1 ? mba->m_cMv3D_B : mba->m_cMv3D_C).set(Mv(m.hor, m.ver), 1);
    }
    const UInt mode = j.mode & JSVM_JOB_MODE_MASK;
    if (j.mode & JSVM_JOB_EL) c->me->setEL(true);                    // one-shot m_bELWithBLMv
    if (j.mode & JSVM_JOB_BIPRED) {
      // a bi-predictive call whose prepared original equals the job's: gClip( 2 * org - alt ) with alt = org (8-bit originals)
      YuvMbBuffer alt;
      alt.loadLuma(*org);
      MotionEstimation::MEBiSearchParameters bsp;
      bsp.pcAltRefPelData = &alt; bsp.uiL1Search = 0; bsp.pcAltRefFrame = c->frames[j.ref_slot];
      bsp.apcWeight[0] = &c->pw; bsp.apcWeight[1] = &c->pw;
      CHKI(c->me->estimateBlockWithStart(*mba, *c->frames[j.ref_slot], mv, pred, bits, cost, j.blk, mode, (UInt)j.search_range, &c->pw, &bsp));
      // ruiCost was formed with fWeight = 0.5; the records carry the fWeight = 1 value (the drop-in applies the weight on the host)
      const UInt f = c->rd->getMotionCostShift(c->me->subPelDFunc() != DF_SSD);
      cost = (c->me->m_uiMinSad - ((f * (bits - j.bits_in)) >> 16)) + ((f * bits) >> 16);
    } else {
      CHKI(c->me->estimateBlockWithStart(*mba, *c->frames[j.ref_slot], mv, pred, bits, cost, j.blk, mode, (UInt)j.search_range, &c->pw, 0));
    }
    c->me->setFullPelDFunc(cfg_full);
    jsvm_me_result& r = results[i];
    r.mv.hor = mv.getHor(); r.mv.ver = mv.getVer();
    r.mv_fullpel.hor = c->me->m_cFullPelMv.getHor() >> 2;
    r.mv_fullpel.ver = c->me->m_cFullPelMv.getVer() >> 2;
    r.sad_fullpel = c->me->m_uiFullPelSad;
    r.min_sad     = c->me->m_uiMinSad;
    r.mv_bits     = bits - j.bits_in;
    r.bits        = bits;
    r.cost        = cost;
    r.best_mode   = c->me->bestMode();

    if (preds) {
      CHKI(c->me->compensateBlock(&rec, j.blk, mode, NULL));
      rec.set4x4Block(B4x4Idx(j.blk));
      const XPel* p = rec.getLumBlk();
      const int w = c->dist->getBlockWidth(mode), h = c->dist->getBlockHeight(mode);
      uint8_t* dp = preds + (size_t)i * 256;
      memset(dp, 0, 256);
      for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) dp[y * 16 + x] = (uint8_t)p[y * rec.getLStride() + x];
    }
  }
  return 0;
}

extern "C" uint32_t jsvm_ref_cost_factor(double lambda, int sad) {
  RateDistortion* rd = 0;
  if (RateDistortion::create(rd) != Err::m_nOK) return 0;
  // setMbQpLambda also stores the QP in the MbData of the access object, so give it a real one
  // (1x1-MB picture).  Inner scope: SliceHeader / MbDataCtrl must die before the parameter sets.
  SequenceParameterSet* sps = 0; PictureParameterSet* pps = 0;
  SequenceParameterSet::create(sps); PictureParameterSet::create(pps);
  sps->setFrameWidthInMbs(1); sps->setFrameHeightInMbs(1);
  pps->setNumSliceGroupsMinus1(0);
  uint32_t f = 0;
  {
    SliceHeader sh(*sps, *pps);
    sh.setFirstMbInSlice(0); sh.setLastMbInSlice(0); sh.setNumMbsInSlice(1); sh.setSliceType(P_SLICE);
    sh.setSliceGroupChangeCycle(1);
    sh.FMOInit();
    MbDataCtrl ctrl;
    ctrl.init(*sps);
    ctrl.initSlice(sh, PRE_PROCESS, false, NULL);
    MbDataAccess* mba = 0;
    ctrl.initMb(mba, 0, 0);
    rd->setMbQpLambda(*mba, 30, lambda);
    f = rd->getMotionCostShift(sad != 0);
    ctrl.uninit();
  }
  rd->destroy();   // sps / pps are leaked on purpose, see jsvm_ref_destroy
  return f;
}

extern "C" uint32_t jsvm_ref_distortion(const int16_t* org16x16, const int16_t* ref, int ref_stride, int mode, int dfunc, int blk) {
  XDistortion* d = 0;
  if (XDistortion::create(d) != Err::m_nOK) return 0xffffffffu;
  d->init();
  YuvMbBuffer* org = d->getYuvMbBuffer();
  XPel* o = org->getMbLumAddr();
  for (int r = 0; r < 16; ++r)
    for (int x = 0; x < 16; ++x) o[r * org->getLStride() + x] = org16x16[r * 16 + x];
  d->set4x4Block(B4x4Idx(blk));
  XDistSearchStruct s;
  d->getDistStruct(mode, (DFunc)dfunc, false, s);
  s.pYSearch = const_cast<XPel*>(ref);
  s.iYStride = ref_stride;
  s.pUSearch = s.pVSearch = 0; s.iCStride = 0;
  uint32_t v = s.Func(&s);
  d->destroy();
  return v;
}

extern "C" int jsvm_ref_set_chroma_u8(jsvm_ref_ctx* c, int slot, const uint8_t* cb_src, const uint8_t* cr_src, int stride) {
  if (!c || slot < 0 || slot >= c->n_slots) return -1;
  Frame* f = c->frames[slot];
  CHKI(c->full_ctrl->initMb());
  CHKI(c->half_ctrl->initMb());
  YuvPicBuffer* pb = f->getFullPelYuvBuffer();
  XPel* cb = pb->getMbCbAddr();
  XPel* cr = pb->getMbCrAddr();
  const int cs = pb->getCStride();
  for (int r = 0; r < c->height / 2; ++r)
    for (int x = 0; x < c->width / 2; ++x) { cb[r * cs + x] = cb_src[r * stride + x]; cr[r * cs + x] = cr_src[r * stride + x]; }
  CHKI(f->extendFrame(c->filter));          // fillMargin of all three planes (+ the unchanged half-pel plane again)
  return 0;
}

extern "C" int jsvm_ref_set_org_chroma(jsvm_ref_ctx* c, const int16_t* org_chroma_mbs, int n_org_mbs) {
  if (!c) return -1;
  c->org_chroma = org_chroma_mbs; c->n_org_chroma = org_chroma_mbs ? n_org_mbs : 0;
  return 0;
}

extern "C" int jsvm_ref_map_cost_factor(jsvm_ref_ctx* c, uint32_t factor_full_pel, uint32_t factor_sub_pel) {
  if (!c) return -1;
  for (size_t k = 0; k < c->factor_map.size(); ++k)
    if (c->factor_map[k].first == factor_full_pel) { c->factor_map[k].second = factor_sub_pel; return 0; }
  c->factor_map.push_back(std::make_pair(factor_full_pel, factor_sub_pel));
  return 0;
}

// MotionCompensation (luma predBlk, chroma xPredChromaPel, SampleWeighting) for every job.  The slice-level weighting state is
// produced by the reference's own SampleWeighting::initSlice from a slice header that matches the job's `weighting` flags.
extern "C" int jsvm_ref_compensate(jsvm_ref_ctx* c, const jsvm_mc_job* jobs, int n_jobs, uint8_t* pred_y, uint8_t* pred_cb, uint8_t* pred_cr) {
  if (!c) return -1;
  const bool old_flag = c->pps->getWeightedPredFlag();
  const UInt old_idc = c->pps->getWeightedBiPredIdc();
  int rc = 0;
  for (int i = 0; i < n_jobs && rc == 0; ++i) {
    const jsvm_mc_job& j = jobs[i];
    if (j.mb_x < 0 || j.mb_x >= (int)c->mbw || j.mb_y < 0 || j.mb_y >= (int)c->mbh || j.blk > 15) { rc = -1; break; }
    const bool uni = (j.weighting & JSVM_MC_WEIGHT_UNI) != 0, bi = (j.weighting & JSVM_MC_WEIGHT_BI) != 0;
    if (bi && !uni && (j.log_denom[0] != 5 || j.log_denom[1] != 5)) { rc = -1; break; }     // implicit weights: denominators are 5 / 5
    c->pps->setWeightedPredFlag(uni);
    c->pps->setWeightedBiPredIdc(bi ? (uni ? 1 : 2) : 0);
    SliceHeader sh(*c->sps, *c->pps);
    sh.setSliceType(bi ? B_SLICE : P_SLICE);
    sh.setLumaLog2WeightDenom(j.log_denom[0]);
    sh.setChromaLog2WeightDenom(j.log_denom[1]);
    if (c->weighting->initSlice(sh) != Err::m_nOK) { rc = -1; break; }

    MbDataAccess* mba = 0;
    if (c->mbctrl->initMb(mba, j.mb_y, j.mb_x) != Err::m_nOK || c->full_ctrl->initMb(j.mb_y, j.mb_x, false) != Err::m_nOK ||
        c->me->initMb(j.mb_y, j.mb_x, *mba) != Err::m_nOK) { rc = -1; break; }
    YuvPicBuffer* refs[2] = { 0, 0 };
    Mv mvs[2];
    PredWeight pw[2];
    const PredWeight* ppw[2] = { 0, 0 };
    for (int l = 0; l < 2; ++l) {
      if (j.ref_slot[l] < 0) continue;
      if (j.ref_slot[l] >= c->n_slots) { rc = -1; break; }
      refs[l] = c->frames[j.ref_slot[l]]->getFullPelYuvBuffer();
      mvs[l] = Mv(j.mv[l].hor, j.mv[l].ver);
      pw[l].setLumaWeightFlag(j.weight_flag[0] != 0); pw[l].setChromaWeightFlag(j.weight_flag[1] != 0);
      pw[l].setLumaWeight(j.weight[l][0]); pw[l].setChromaCbWeight(j.weight[l][1]); pw[l].setChromaCrWeight(j.weight[l][2]);
      pw[l].setLumaOffset(j.offset[l][0]); pw[l].setChromaCbOffset(j.offset[l][1]); pw[l].setChromaCrOffset(j.offset[l][2]);
      ppw[l] = &pw[l];
    }
    if (rc) break;
    YuvMbBuffer rec;
    c->me->predictBlock(&rec, j.blk, j.size_x, j.size_y, refs, mvs, ppw, pred_cb != 0);
    const B4x4Idx idx(j.blk);
    const XPel* y = rec.getYBlk(idx);
    uint8_t* dy = pred_y + (size_t)i * 256;
    memset(dy, 0, 256);
    for (int r = 0; r < j.size_y; ++r)
      for (int x = 0; x < j.size_x; ++x) dy[r * 16 + x] = (uint8_t)y[r * rec.getLStride() + x];
    if (pred_cb) {
      const XPel* u = rec.getUBlk(idx); const XPel* v = rec.getVBlk(idx);
      uint8_t* du = pred_cb + (size_t)i * 64; uint8_t* dv = pred_cr + (size_t)i * 64;
      memset(du, 0, 64); memset(dv, 0, 64);
      for (int r = 0; r < j.size_y / 2; ++r)
        for (int x = 0; x < j.size_x / 2; ++x) { du[r * 8 + x] = (uint8_t)u[r * rec.getCStride() + x]; dv[r * 8 + x] = (uint8_t)v[r * rec.getCStride() + x]; }
    }
  }
  c->pps->setWeightedPredFlag(old_flag);
  c->pps->setWeightedBiPredIdc(old_idc);
  c->weighting->initSlice(*c->sh);
  return rc;
}

extern "C" int jsvm_ref_set_fast_search_params(jsvm_ref_ctx* c, int el_search_range, int fast_bi_search) {
  if (!c) return -1;
  c->me->setFastSearchParams(el_search_range ? el_search_range : c->me->searchRange(), fast_bi_search);   // ENC/CodingParameter.cpp:25-28
  return 0;
}

extern "C" int jsvm_ref_set_mv_candidates(jsvm_ref_ctx* c, const jsvm_mv* abc, int n_jobs) {
  if (!c) return -1;
  c->cands = abc; c->n_cands = abc ? n_jobs : 0;
  return 0;
}

// Picture-level motion search (see jsvm_me_estimate_pictures): the reference's own MV prediction (MbDataAccess::getMvPredictor on a
// real MbDataCtrl, neighbours resolved by MbDataCtrl::initMb) and estimateBlockWithStart, composed as the list-0 loops of
// MbEncoder::xEstimateMb16x16 / 16x8 / 8x16 / 8x8 compose them (ENC/MbEncoder.cpp:4449-4560, 4830-4935, 5234-5330, 5637-5700,
// 5950-6040), with the macroblock mode chosen by the sum of the motion costs.
namespace {
UInt refIdxBits(int r, int n) { return n == 1 ? 0 : n == 2 ? 1 : (r < 2 ? 1 : r < 4 ? 3 : r < 8 ? 5 : 7); }   // getRefIdxBits / g_aucFrameBits
}

extern "C" int jsvm_ref_estimate_pictures(jsvm_ref_ctx* c, const jsvm_pic_task* tasks, int n_pics, jsvm_pic_mb* out) {
  if (!c) return -1;
  static const ParIdx16x8 k16x8[2] = { PART_16x8_0, PART_16x8_1 };
  static const ParIdx8x16 k8x16[2] = { PART_8x16_0, PART_8x16_1 };
  static const ParIdx8x8  k8x8 [4] = { PART_8x8_0, PART_8x8_1, PART_8x8_2, PART_8x8_3 };
  for (int p = 0; p < n_pics; ++p) {
    const jsvm_pic_task& t = tasks[p];
    const int R = t.n_refs;
    MbDataCtrl ctrl;
    CHKI(ctrl.init(*c->sps));
    CHKI(ctrl.initSlice(*c->sh, PRE_PROCESS, false, NULL));
    c->rd->m_uiFactor = t.cost_factor; c->rd->m_uiFactorSse = t.cost_factor;
    for (UInt mby = 0; mby < c->mbh; ++mby)
      for (UInt mbx = 0; mbx < c->mbw; ++mbx) {
        MbDataAccess* mba = 0;
        CHKI(ctrl.initMb(mba, mby, mbx));
        CHKI(c->full_ctrl->initMb(mby, mbx, false));
        CHKI(c->half_ctrl->initMb(mby, mbx, false));
        CHKI(c->me->initMb(mby, mbx, *mba));
        YuvMbBuffer* org = 0;
        c->dist->loadOrgMbPelData(c->frames[t.cur_slot]->getFullPelYuvBuffer(), org);
        MbMotionData& mot = mba->getMbMotionData(LIST_0);
        jsvm_pic_mb& o = out[(size_t)p * c->mbw * c->mbh + mby * c->mbw + mbx];
        memset(&o, 0, sizeof(o));
        // one partition: every reference, keep the cheapest ( uiCostTest < uiCost[0] )
        struct Est { Mv mv; int ref; UInt cost; };
        #define ESTIMATE(GETPRED, BLK, MODE, MODEBITS, RESULT)                                             \
          do { (RESULT).cost = MSYS_UINT_MAX; (RESULT).ref = 0;                                            \
            for (int r = 1; r <= R; ++r) {                                                                 \
              Mv pred; GETPRED;                                                                            \
              Mv mv = pred; UInt bits = (MODEBITS) + refIdxBits(r, R), cost = 0;                           \
              CHKI(c->me->estimateBlockWithStart(*mba, *c->frames[t.ref_slot[r - 1]], mv, pred, bits, cost, BLK, MODE, 0, &c->pw, 0)); \
              if (cost < (RESULT).cost) { (RESULT).cost = cost; (RESULT).ref = r; (RESULT).mv = mv; }      \
            } } while (0)
        Est e[9];
        mot.clear(BLOCK_NOT_AVAILABLE); mot.setAllMv(Mv::ZeroMv());
        ESTIMATE(mba->getMvPredictor(pred, r, LIST_0), 0, MODE_16x16, 1, e[0]);
        mot.clear(BLOCK_NOT_AVAILABLE); mot.setAllMv(Mv::ZeroMv());
        for (int k = 0; k < 2; ++k) {
          ESTIMATE(mba->getMvPredictor(pred, r, LIST_0, k16x8[k]), k16x8[k], MODE_16x8, 3, e[1 + k]);
          mot.setRefIdx(e[1 + k].ref, k16x8[k]); mot.setAllMv(e[1 + k].mv, k16x8[k]);
        }
        mot.clear(BLOCK_NOT_AVAILABLE); mot.setAllMv(Mv::ZeroMv());
        for (int k = 0; k < 2; ++k) {
          ESTIMATE(mba->getMvPredictor(pred, r, LIST_0, k8x16[k]), k8x16[k], MODE_8x16, 3, e[3 + k]);
          mot.setRefIdx(e[3 + k].ref, k8x16[k]); mot.setAllMv(e[3 + k].mv, k8x16[k]);
        }
        mot.clear(BLOCK_NOT_AVAILABLE); mot.setAllMv(Mv::ZeroMv());
        for (int k = 0; k < 4; ++k) {
          ESTIMATE(mba->getMvPredictor(pred, r, LIST_0, k8x8[k]), k8x8[k] + SPART_8x8, BLK_8x8, 1 + (k == 0 ? 5 : 0), e[5 + k]);
          mot.setRefIdx(e[5 + k].ref, k8x8[k]); mot.setAllMv(e[5 + k].mv, k8x8[k]);
        }
        #undef ESTIMATE
        const UInt cost[4] = { e[0].cost, e[1].cost + e[2].cost, e[3].cost + e[4].cost, e[5].cost + e[6].cost + e[7].cost + e[8].cost };
        int mode = 1; UInt best = cost[0];
        for (int m = 1; m < 4; ++m) if (cost[m] < best) { best = cost[m]; mode = m + 1; }
        // final motion of the macroblock: what the next macroblocks predict from
        mot.clear(BLOCK_NOT_AVAILABLE);
        for (int b8 = 0; b8 < 4; ++b8) {
          const int part = mode == 1 ? 0 : mode == 2 ? 1 + (b8 >> 1) : mode == 3 ? 3 + (b8 & 1) : 5 + b8;
          mot.setRefIdx(e[part].ref, k8x8[b8]); mot.setAllMv(e[part].mv, k8x8[b8]);
          o.ref_idx[b8] = (int8_t)e[part].ref;
        }
        o.mode = (uint8_t)mode; o.cost = best;
        for (int b = 0; b < 16; ++b) { const Mv& m = mot.getMv(B4x4Idx(b)); o.mv[b].hor = m.getHor(); o.mv[b].ver = m.getVer(); }
        for (int k = 0; k < 9; ++k) { o.part_mv[k].hor = e[k].mv.getHor(); o.part_mv[k].ver = e[k].mv.getVer(); o.part_ref[k] = (int8_t)e[k].ref; o.part_cost[k] = e[k].cost; }
      }
    ctrl.uninit();
  }
  return 0;
}

// Frame::residualUpsampling (DownConvert::xResidualUpsampling) on caller-supplied dense planes, progressive layers.
namespace {
void fillResizeParameters(ResizeParameters& rp, const jsvm_resize_params& p) {
  rp.m_iExtendedSpatialScalability = ESS_SEQ;
  rp.m_iLevelIdc = p.level_idc;
  rp.m_bFrameMbsOnlyFlag = true; rp.m_bFieldPicFlag = false; rp.m_bBotFieldFlag = false; rp.m_bIsMbAffFrame = false;
  rp.m_iFrameWidth = p.frame_width; rp.m_iFrameHeight = p.frame_height;
  rp.m_iWidthInSamples = p.frame_width; rp.m_iHeightInSamples = p.frame_height;
  rp.m_iChromaPhaseX = p.chroma_phase_x; rp.m_iChromaPhaseY = p.chroma_phase_y;
  rp.m_iScaledRefFrmWidth = p.scaled_ref_width; rp.m_iScaledRefFrmHeight = p.scaled_ref_height;
  rp.m_iLeftFrmOffset = p.left_offset; rp.m_iTopFrmOffset = p.top_offset;
  rp.m_iRefLayerChromaPhaseX = p.ref_chroma_phase_x; rp.m_iRefLayerChromaPhaseY = p.ref_chroma_phase_y;
  rp.m_bRefLayerFrameMbsOnlyFlag = true; rp.m_bRefLayerFieldPicFlag = false; rp.m_bRefLayerBotFieldFlag = false; rp.m_bRefLayerIsMbAffFrame = false;
  rp.m_iRefLayerFrmWidth = p.ref_frame_width; rp.m_iRefLayerFrmHeight = p.ref_frame_height;
  rp.m_iRefLayerWidthInSamples = p.ref_frame_width; rp.m_iRefLayerHeightInSamples = p.ref_frame_height;
}
struct LayerBuffers {          // a Frame of one layer size with its own buffer control
  YuvBufferCtrl* ctrl; Frame* frame;
  LayerBuffers() : ctrl(0), frame(0) {}
  int init(int w, int h) {
    if (YuvBufferCtrl::create(ctrl) != Err::m_nOK || ctrl->initSPS(h, w, YUV_Y_MARGIN, YUV_X_MARGIN) != Err::m_nOK) return -1;
    frame = new Frame(*ctrl, *ctrl, FRAME, 0);
    return frame->init() == Err::m_nOK ? 0 : -1;
  }
  ~LayerBuffers() { if (frame) { frame->uninit(); delete frame; } if (ctrl) { ctrl->uninit(); ctrl->destroy(); } }
};
}  // namespace

extern "C" int jsvm_ref_upsample_residual(const jsvm_resize_params* p, const int16_t* base_y, const int16_t* base_cb, const int16_t* base_cr,
                                          const uint8_t* base_transform8x8, int16_t* out_y, int16_t* out_cb, int16_t* out_cr) {
  ResizeParameters rp;
  fillResizeParameters(rp, *p);
  LayerBuffers base, cur;
  if (base.init(p->ref_frame_width, p->ref_frame_height) || cur.init(p->frame_width, p->frame_height)) return -1;
  // base-layer macroblock data: only the transform size is read (xDetermineTransBlkIdcs)
  SequenceParameterSet* sps = 0;
  CHKI(SequenceParameterSet::create(sps));
  sps->setFrameWidthInMbs(p->ref_frame_width >> 4); sps->setFrameHeightInMbs(p->ref_frame_height >> 4);
  MbDataCtrl ctrl;
  CHKI(ctrl.init(*sps));
  const int n_mb = (p->ref_frame_width >> 4) * (p->ref_frame_height >> 4);
  for (int i = 0; i < n_mb; ++i) ctrl.getMbDataByIndex(i).setTransformSize8x8(base_transform8x8[i] != 0);
  base.ctrl->initMb(); cur.ctrl->initMb();
  YuvPicBuffer* bb = base.frame->getFullPelYuvBuffer();
  const bool chroma = base_cb && base_cr && out_cb && out_cr;
  for (int y = 0; y < p->ref_frame_height; ++y)
    for (int x = 0; x < p->ref_frame_width; ++x) bb->getMbLumAddr()[y * bb->getLStride() + x] = base_y[(size_t)y * p->ref_frame_width + x];
  for (int y = 0; y < p->ref_frame_height / 2; ++y)
    for (int x = 0; x < p->ref_frame_width / 2; ++x) {
      bb->getMbCbAddr()[y * bb->getCStride() + x] = chroma ? base_cb[(size_t)y * (p->ref_frame_width / 2) + x] : 0;
      bb->getMbCrAddr()[y * bb->getCStride() + x] = chroma ? base_cr[(size_t)y * (p->ref_frame_width / 2) + x] : 0;
    }
  DownConvert dc;
  dc.init(p->frame_width > p->ref_frame_width ? p->frame_width : p->ref_frame_width, p->frame_height > p->ref_frame_height ? p->frame_height : p->ref_frame_height, 16);
  CHKI(cur.frame->residualUpsampling(base.frame, dc, &rp, &ctrl));
  cur.ctrl->initMb();
  YuvPicBuffer* cb = cur.frame->getFullPelYuvBuffer();
  for (int y = 0; y < p->frame_height; ++y)
    for (int x = 0; x < p->frame_width; ++x) out_y[(size_t)y * p->frame_width + x] = cb->getMbLumAddr()[y * cb->getLStride() + x];
  if (chroma)
    for (int y = 0; y < p->frame_height / 2; ++y)
      for (int x = 0; x < p->frame_width / 2; ++x) {
        out_cb[(size_t)y * (p->frame_width / 2) + x] = cb->getMbCbAddr()[y * cb->getCStride() + x];
        out_cr[(size_t)y * (p->frame_width / 2) + x] = cb->getMbCrAddr()[y * cb->getCStride() + x];
      }
  ctrl.uninit();
  return 0;
}

// MbDataCtrl::upsampleMotion on caller-supplied base-layer motion, progressive layers.
extern "C" int jsvm_ref_upsample_motion(const jsvm_motion_upsample_params* p, const jsvm_mb_motion* base, jsvm_mb_motion* out) {
  ResizeParameters rp;
  fillResizeParameters(rp, p->resize);
  SequenceParameterSet* spsB = 0; SequenceParameterSet* spsC = 0; PictureParameterSet* pps = 0;
  CHKI(SequenceParameterSet::create(spsB)); CHKI(SequenceParameterSet::create(spsC)); CHKI(PictureParameterSet::create(pps));
  spsB->setFrameWidthInMbs(p->resize.ref_frame_width >> 4); spsB->setFrameHeightInMbs(p->resize.ref_frame_height >> 4);
  spsC->setFrameWidthInMbs(p->resize.frame_width >> 4);     spsC->setFrameHeightInMbs(p->resize.frame_height >> 4);
  spsC->setDirect8x8InferenceFlag(p->direct_8x8_inference != 0);
  pps->setNumSliceGroupsMinus1(0);
  int rc = 0;
  {
    SliceHeader sh(*spsC, *pps);
    sh.setSliceType(SliceType(p->slice_type));
    sh.setRefLayerDQId(p->ref_layer_dq_id);
    sh.setSCoeffResidualPredFlag((ResizeParameters*)0);
    sh.setTCoeffLevelPredictionFlag(false);
    MbDataCtrl baseCtrl, curCtrl;
    CHKI(baseCtrl.init(*spsB)); CHKI(curCtrl.init(*spsC));
    const int nb = (p->resize.ref_frame_width >> 4) * (p->resize.ref_frame_height >> 4);
    const int nc = (p->resize.frame_width >> 4) * (p->resize.frame_height >> 4);
    static const ParIdx8x8 k8x8[4] = { PART_8x8_0, PART_8x8_1, PART_8x8_2, PART_8x8_3 };
    for (int i = 0; i < nb; ++i) {
      MbData& d = baseCtrl.getMbDataByIndex(i);
      const jsvm_mb_motion& b = base[i];
      d.initMbData(0, 0, 1, i, i, SliceType(b.slice_type));
      d.setMbMode(MbMode(b.mb_mode));
      for (int k = 0; k < 4; ++k) d.setBlkMode(Par8x8(k), BlkMode(b.blk_mode[k]));
      for (int l = 0; l < 2; ++l) {
        MbMotionData& mot = d.getMbMotionData(ListIdx(l));
        for (int k = 0; k < 4; ++k) mot.setRefIdx(b.ref_idx[l][k], k8x8[k]);
        for (int k = 0; k < 16; ++k) mot.setMv(Mv(b.mv[l][k].hor, b.mv[l][k].ver), B4x4Idx(k));
      }
    }
    if (curCtrl.upsampleMotion(&sh, &rp, &baseCtrl, NULL, NULL, false, p->check_residual_pred != 0, p->mv_threshold) != Err::m_nOK) rc = -1;
    for (int i = 0; i < nc && rc == 0; ++i) {
      MbData& d = curCtrl.getMbDataByIndex(i);
      jsvm_mb_motion& o = out[i];
      memset(&o, 0, sizeof(o));
      o.mb_mode = (uint8_t)d.getMbMode();
      for (int k = 0; k < 4; ++k) o.blk_mode[k] = (uint8_t)d.getBlkMode(Par8x8(k));
      o.in_crop_window = d.getInCropWindowFlag(); o.res_pred_safe = d.getSafeResPred(); o.fwd_bwd = d.getFwdBwd();
      for (int k = 0; k < 4; ++k) if (d.isBaseIntra(k & 1, k >> 1)) o.base_intra |= (uint8_t)(1 << k);
      for (int l = 0; l < 2; ++l) {
        MbMotionData& mot = d.getMbMotionData(ListIdx(l));
        for (int k = 0; k < 4; ++k) o.ref_idx[l][k] = mot.getRefIdx(k8x8[k]);
        for (int k = 0; k < 16; ++k) { const Mv& m = mot.getMv(B4x4Idx(k)); o.mv[l][k].hor = m.getHor(); o.mv[l][k].ver = m.getVer(); }
      }
    }
    baseCtrl.uninit(); curCtrl.uninit();
  }
  return rc;      // the parameter sets are leaked on purpose, see jsvm_ref_destroy
}
