// MotionEstimationB200.cpp -- the reference-side binding: replacement bodies for
//   ENC/MotionEstimation.cpp, ENC/MotionEstimationQuarterPel.cpp and QuarterPelFilter::filterFrame
// that keep the reference class headers byte-identical and forward the hot path to the C ABI
// (include/jsvm_me_b200.h).  H264AVCEncoderLib links against this object instead of its own
// MotionEstimation.o / MotionEstimationQuarterPel.o (selection is at link time).
//
// What stays on the host, with the reference's own code and flags (-ffloat-store):
//   * MotionCompensation::initMb / initSlice (base class, COM/MotionCompensation.cpp)
//   * SampleWeighting (bi-pred / weighted originals, COM/SampleWeighting.cpp:267-571)
//   * MotionEstimationCost (rate table for getRateCost and the final cost, ENC/MotionEstimationCost.cpp)
//   * RateDistortion (lambda -> cost factor, ENC/RateDistortion.cpp:59-70; N11)
// What goes to the device: the full-pel full search, the sub-pel refinement, the half-pel plane.
// There is no CPU search in this file: every failure of the device path returns Err::m_nERR.
#include "H264AVCEncoderLib.h"

#include "H264AVCCommonLib/SampleWeighting.h"
#include "H264AVCCommonLib/QuarterPelFilter.h"
#include "H264AVCCommonLib/Transform.h"
#include "H264AVCCommonLib/Frame.h"
#include "RateDistortionIf.h"
#include "CodingParameter.h"
#include "MotionEstimation.h"
#include "MotionEstimationQuarterPel.h"

#include <map>
#include <vector>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "jsvm_me_b200.h"

H264AVC_NAMESPACE_BEGIN

namespace {

// One device context per picture size (one per spatial layer); frame-store slots are keyed by the
// half-pel YuvPicBuffer object of a Frame (the object filterFrame fills and estimateBlockWithStart reads).
struct LayerStore {
  jsvm_me_ctx* ctx;
  int width, height, n_slots;
  std::map<const void*, int> slot_of;
  std::vector<unsigned long long> stamp;
  std::vector<const void*> owner;
  unsigned long long clock;
};

struct B200State {
  std::map<std::pair<int, int>, LayerStore*> layers;
  int search_range, sub_dfunc, full_dfunc, search_mode, el_search_range, fast_bi_search;
  bool params_valid;
  int device;
  // the buffers compensateBlock serves from (the reference keeps m_aXHPelSearch / m_aXQPelSearch)
  unsigned char last_pred[256];
  bool last_valid;
  unsigned long long n_calls, n_calls_16bit, n_calls_bipred, n_frames;
  unsigned long long n_wp_uni, n_wp_bi_explicit, n_wp_bi_implicit;     // calls whose original went through weightInverseLumaSamples
  B200State() : search_range(0), sub_dfunc(0), full_dfunc(0), search_mode(0), el_search_range(0), fast_bi_search(0), params_valid(false), device(0), last_valid(false), n_calls(0), n_calls_16bit(0),
                n_calls_bipred(0), n_frames(0), n_wp_uni(0), n_wp_bi_explicit(0), n_wp_bi_implicit(0) {
    const char* d = getenv("JSVM_ME_DEVICE");
    if (d) device = atoi(d);
  }
  ~B200State() {
    // JSVM_ME_STATS=1: what went through the device (evidence for the system parity runs)
    if (getenv("JSVM_ME_STATS"))
      fprintf(stderr, "jsvm_me_b200 stats: filterFrame %llu, estimateBlockWithStart %llu (signed 16-bit originals %llu, bi-pred %llu, weighted uni %llu, weighted bi explicit %llu, "
                      "weighted bi implicit %llu), layers %zu\n",
              n_frames, n_calls, n_calls_16bit, n_calls_bipred, n_wp_uni, n_wp_bi_explicit, n_wp_bi_implicit, layers.size());
  }
};

B200State& state() { static B200State s; return s; }

ErrVal reportFailure(const char* what) {
  fprintf(stderr, "jsvm_me_b200: %s failed: %s\n", what, jsvm_me_last_error());
  return Err::m_nERR;
}

LayerStore* layerFor(int width, int height) {
  B200State& st = state();
  const std::pair<int, int> key(width, height);
  std::map<std::pair<int, int>, LayerStore*>::iterator it = st.layers.find(key);
  if (it != st.layers.end()) return it->second;
  LayerStore* ls = new LayerStore();
  ls->ctx = 0; ls->width = width; ls->height = height; ls->n_slots = 48; ls->clock = 0;
  if (jsvm_me_create(&ls->ctx, st.device) != JSVM_ME_OK || jsvm_me_configure(ls->ctx, width, height, ls->n_slots) != JSVM_ME_OK) {
    reportFailure("jsvm_me_create / jsvm_me_configure");
    delete ls;
    return 0;
  }
  ls->stamp.assign(ls->n_slots, 0);
  ls->owner.assign(ls->n_slots, (const void*)0);
  st.layers[key] = ls;
  return ls;
}

int slotForFill(LayerStore* ls, const void* key) {
  std::map<const void*, int>::iterator it = ls->slot_of.find(key);
  int slot;
  if (it != ls->slot_of.end()) slot = it->second;
  else {
    slot = 0;
    for (int i = 1; i < ls->n_slots; ++i) if (ls->stamp[i] < ls->stamp[slot]) slot = i;   // free or least recently filled
    if (ls->owner[slot]) ls->slot_of.erase(ls->owner[slot]);
    ls->owner[slot] = key;
    ls->slot_of[key] = slot;
  }
  ls->stamp[slot] = ++ls->clock;
  return slot;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// QuarterPelFilter::filterFrame -- replaces COM/QuarterPelFilter.cpp:61-149.
// (The reference object keeps every other QuarterPelFilter member; the dropin Makefile weakens only
//  this symbol in the reference's QuarterPelFilter.o so that this strong definition wins.)
// The half-pel plane is built and kept on the device; the host half-pel buffer is not written
// (its only consumer is motion estimation).
// ------------------------------------------------------------------------------------------------
ErrVal QuarterPelFilter::filterFrame( YuvPicBuffer* pcPelBuffer, YuvPicBuffer* pcHalfPelBuffer )
{
  ROT( NULL == pcPelBuffer );
  ROT( NULL == pcHalfPelBuffer );
  const Int iWidth = pcPelBuffer->getLWidth(), iHeight = pcPelBuffer->getLHeight();
  LayerStore* ls = layerFor( iWidth, iHeight );
  ROF( ls );
  const int slot = slotForFill( ls, pcHalfPelBuffer );
  state().n_frames++;
  if( jsvm_me_upload_plane_xpel( ls->ctx, slot, pcPelBuffer->getMbLumAddr(), iWidth, iHeight, pcPelBuffer->getLStride() ) != JSVM_ME_OK )
    return reportFailure( "jsvm_me_upload_plane_xpel" );
  if( jsvm_me_interp_halfpel( ls->ctx, slot ) != JSVM_ME_OK )
    return reportFailure( "jsvm_me_interp_halfpel" );
  if( state().full_dfunc == JSVM_DF_YUV_SAD )              // the YUV-SAD search also reads Cb / Cr (margins were filled by extendFrame)
  {
    if( jsvm_me_upload_chroma_xpel( ls->ctx, slot, pcPelBuffer->getMbCbAddr(), pcPelBuffer->getMbCrAddr(), pcPelBuffer->getCStride() ) != JSVM_ME_OK )
      return reportFailure( "jsvm_me_upload_chroma_xpel" );
  }
  return Err::m_nOK;
}

// ------------------------------------------------------------------------------------------------
// MotionEstimation
// ------------------------------------------------------------------------------------------------
MotionEstimation::MotionEstimation() :
  m_pcQuarterPelFilter( NULL ),
  m_iMaxLogStep( 0 ),
  m_pcMvSpiralSearch( NULL ),
  m_uiSpiralSearchEntries( 0 ),
  m_pcXDistortion( NULL ),
  m_bELWithBLMv( false )
{
}

MotionEstimation::~MotionEstimation()
{
}

ErrVal MotionEstimation::destroy()
{
  delete this;
  return Err::m_nOK;
}

ErrVal MotionEstimation::init( XDistortion*      pcXDistortion,
                               CodingParameter*  pcCodingParameter,
                               RateDistortionIf* pcRateDistortionIf,
                               QuarterPelFilter* pcQuarterPelFilter,
                               Transform*        pcTransform,
                               SampleWeighting*  pcSampleWeighting )
{
  RNOK( MotionCompensation::init( pcQuarterPelFilter, pcTransform, pcSampleWeighting ) );
  ROT( NULL == pcXDistortion );
  ROT( NULL == pcCodingParameter );
  ROT( NULL == pcQuarterPelFilter );
  m_pcXDistortion      = pcXDistortion;
  m_pcQuarterPelFilter = pcQuarterPelFilter;
  m_cParams            = pcCodingParameter->getMotionVectorSearchParams();
  // the rate table serves getRateCost() and the final cost (ENC/MotionEstimation.h:84, .cpp:331-335)
  RNOK( MotionEstimationCost::xInit( ( m_cParams.getSearchRange() << 2 ), pcRateDistortionIf ) );

  B200State& st = state();
  st.params_valid = false;
  // only what the device path implements; anything else is an error, not a silent CPU search
  const UInt uiFull = m_cParams.getFullPelDFunc(), uiSub = m_cParams.getSubPelDFunc();
  if( m_cParams.getSearchRange() < 1 || m_cParams.getSearchRange() > 4096 || m_cParams.getELSearchRange() > 4096 )
  {
    fprintf( stderr, "jsvm_me_b200: unsupported MotionVectorSearchParams (SearchRange / ELSearchRange must be in 1..4096)\n" );
    return Err::m_nERR;
  }
  if( ( m_cParams.getSearchMode() == LOG_SEARCH || m_cParams.getSearchMode() == FAST_SEARCH ) && uiFull == DF_YUV_SAD )
  {
    fprintf( stderr, "jsvm_me_b200: SearchMode %u with SearchFuncFullPel 3 is undefined in the reference (xPelLogSearch leaves the chroma pointers unset)\n",
             m_cParams.getSearchMode() );
    return Err::m_nERR;
  }
  st.search_mode     = (int)m_cParams.getSearchMode();
  st.el_search_range = (int)m_cParams.getELSearchRange();
  st.fast_bi_search  = (int)m_cParams.getFastBiSearch();
  st.search_range = (int)m_cParams.getSearchRange();
  st.sub_dfunc    = (int)uiSub;
  st.full_dfunc   = (int)uiFull;
  st.params_valid = true;
  return Err::m_nOK;
}

ErrVal MotionEstimation::uninit()
{
  RNOK( MotionCompensation::uninit() );
  RNOK( MotionEstimationCost::xUninit() );
  return Err::m_nOK;
}

ErrVal MotionEstimation::initMb( UInt uiMbPosY, UInt uiMbPosX, MbDataAccess& rcMbDataAccess )
{
  RNOK( MotionCompensation::initMb( uiMbPosY, uiMbPosX, rcMbDataAccess ) );
  ROT( rcMbDataAccess.getMbData().getFieldFlag() );      // frame macroblocks only
  // m_cMin / m_cMax are a pure function of the MB position and the picture size; the device recomputes them.
  // Recover the position from the clamp the base class just stored (COM/MotionCompensation.cpp:160-161).
  return Err::m_nOK;
}

ErrVal
MotionEstimation::estimateBlockWithStart( const MbDataAccess&          rcMbDataAccess,
                                          const Frame&                 rcRefFrame,
                                          Mv&                          rcMv,
                                          const Mv&                    rcMvPred,
                                          UInt&                        ruiBits,
                                          UInt&                        ruiCost,
                                          const UInt                   uiBlk,
                                          const UInt                   uiMode,
                                          const UInt                   uiSearchRange,
                                          const PredWeight*            pcPW,
                                          const MEBiSearchParameters*  pcBSP )
{
  B200State& st = state();
  ROF( st.params_valid );
  const LumaIdx cIdx  = B4x4Idx( uiBlk );
  const Int     iXSize = m_pcXDistortion->getBlockWidth ( uiMode );
  const Int     iYSize = m_pcXDistortion->getBlockHeight( uiMode );

  // ---- which "original" is searched for, and with which weight (host, reference SampleWeighting) ----
  YuvMbBuffer  cPrepared;
  YuvMbBuffer* pcOrg   = m_pcXDistortion->getYuvMbBuffer();
  Double       fWeight = 1.0;
  Double       afCW[2] = { 1.0, 1.0 };
  const Bool   bYuvSad = ( st.full_dfunc == JSVM_DF_YUV_SAD );
  if( pcBSP )
  {
    ROF( pcBSP->pcAltRefPelData && pcBSP->pcAltRefFrame && pcBSP->apcWeight[LIST_0] && pcBSP->apcWeight[LIST_1] );
    cPrepared.set4x4Block( cIdx );
    m_pcXDistortion->set4x4Block( cIdx );
    pcBSP->pcAltRefPelData->set4x4Block( cIdx );
    const Bool bImplicit = ( rcMbDataAccess.getSH().getPPS().getWeightedBiPredIdc() == 2 );
    const Bool bExplicit = !bImplicit && ( pcBSP->apcWeight[LIST_0]->getLumaWeightFlag() || pcBSP->apcWeight[LIST_1]->getLumaWeightFlag() );
    Int iScale = 128;
    if( bImplicit )
    {
      const Frame* pcL0 = pcBSP->uiL1Search ? pcBSP->pcAltRefFrame : &rcRefFrame;
      const Frame* pcL1 = pcBSP->uiL1Search ? &rcRefFrame : pcBSP->pcAltRefFrame;
      iScale = rcMbDataAccess.getSH().getDistScaleFactorWP( pcL0, pcL1 );
    }
    if( bImplicit && iScale != 128 )
    {
      PredWeight acIPW[2];
      acIPW[1].scaleL1Weight( iScale );
      acIPW[0].scaleL0Weight( acIPW[1] );
      st.n_wp_bi_implicit++;
      m_pcSampleWeighting->weightInverseLumaSamples( &cPrepared, pcOrg, pcBSP->pcAltRefPelData,
                                                     &acIPW[pcBSP->uiL1Search], &acIPW[1 - pcBSP->uiL1Search], fWeight, iYSize, iXSize );
    }
    else if( bExplicit )
    {
      st.n_wp_bi_explicit++;
      m_pcSampleWeighting->weightInverseLumaSamples( &cPrepared, pcOrg, pcBSP->pcAltRefPelData,
                                                     pcBSP->apcWeight[pcBSP->uiL1Search], pcBSP->apcWeight[1 - pcBSP->uiL1Search], fWeight, iYSize, iXSize );
    }
    else
    {
      m_pcSampleWeighting->inverseLumaSamples( &cPrepared, pcOrg, pcBSP->pcAltRefPelData, iYSize, iXSize );   // gClip( 2*org - pred )
      fWeight = 0.5;
    }
    pcOrg = &cPrepared;
  }
  else
  {
    ROF( pcPW );
    if( pcPW->getLumaWeightFlag() || ( bYuvSad && pcPW->getChromaWeightFlag() ) )
    {
      cPrepared.set4x4Block( cIdx );
      m_pcXDistortion->set4x4Block( cIdx );
      st.n_wp_uni++;
      m_pcSampleWeighting->weightInverseLumaSamples( &cPrepared, pcOrg, pcPW, fWeight, iYSize, iXSize );
      if( bYuvSad )
        m_pcSampleWeighting->weightInverseChromaSamples( &cPrepared, pcOrg, pcPW, afCW, iYSize, iXSize );
      pcOrg = &cPrepared;
    }
  }
  // ENC/MotionEstimation.cpp:250-253: YUV-SAD only for uni-directional searches with equal component weights
  const Bool bChromaInSearch = bYuvSad && !pcBSP && fWeight == afCW[0] && fWeight == afCW[1];

  // ---- the job ----
  YuvPicBuffer* pcHalf = const_cast<Frame&>( rcRefFrame ).getHalfPelYuvBuffer();
  YuvPicBuffer* pcFull = const_cast<Frame&>( rcRefFrame ).getFullPelYuvBuffer();
  LayerStore* ls = layerFor( pcFull->getLWidth(), pcFull->getLHeight() );
  ROF( ls );
  std::map<const void*, int>::iterator it = ls->slot_of.find( pcHalf );
  if( it == ls->slot_of.end() )
  {
    fprintf( stderr, "jsvm_me_b200: reference frame has no half-pel plane on the device (extendFrame not called?)\n" );
    return Err::m_nERR;
  }
  if( st.n_calls == 0 || st.search_range != 0 )
  {
    if( jsvm_me_set_params( ls->ctx, st.search_mode, st.full_dfunc, st.sub_dfunc, st.search_range ) != JSVM_ME_OK )
      return reportFailure( "jsvm_me_set_params" );
    if( jsvm_me_set_fast_search_params( ls->ctx, st.el_search_range, st.fast_bi_search ) != JSVM_ME_OK )
      return reportFailure( "jsvm_me_set_fast_search_params" );
  }
  st.n_calls++;

  // MB position: m_cMin = ( -16*mbX - 24 ) * 4  (COM/MotionCompensation.cpp:160-161)
  const Int iMbX = ( -( m_cMin.getHor() >> 2 ) - 24 ) >> 4;
  const Int iMbY = ( -( m_cMin.getVer() >> 2 ) - 24 ) >> 4;

  // cost factors of the two stages (ENC/MotionEstimation.cpp:255, 325): SSE class iff the stage's function is SSD (N7)
  xSetMEPars( 2, st.full_dfunc != JSVM_DF_SSD );
  const UInt uiFactorFull = m_uiCostFactor;
  xSetMEPars( 0, st.sub_dfunc != JSVM_DF_SSD );            // stays set: also used below for the final cost
  if( ( st.full_dfunc == JSVM_DF_SSD ) != ( st.sub_dfunc == JSVM_DF_SSD ) )
  {
    if( jsvm_me_map_cost_factor( ls->ctx, uiFactorFull, m_uiCostFactor ) != JSVM_ME_OK )
      return reportFailure( "jsvm_me_map_cost_factor" );
  }
  jsvm_me_job cJob;
  cJob.cur_slot     = (int16_t)it->second;                 // unused: the original comes from org_mbs
  cJob.ref_slot     = (int16_t)it->second;
  cJob.mb_x         = (int16_t)iMbX;
  cJob.mb_y         = (int16_t)iMbY;
  cJob.blk          = (uint8_t)uiBlk;
  cJob.mode         = (uint8_t)( uiMode | ( pcBSP ? JSVM_JOB_BIPRED : 0 ) | ( m_bELWithBLMv ? JSVM_JOB_EL : 0 ) );
  cJob.search_range = (int16_t)uiSearchRange;
  cJob.start.hor    = rcMv.getHor();      cJob.start.ver = rcMv.getVer();
  cJob.pred.hor     = rcMvPred.getHor();  cJob.pred.ver  = rcMvPred.getVer();
  cJob.cost_factor  = uiFactorFull;
  cJob.bits_in      = ruiBits;
  cJob.org_index    = 0;

  int16_t asOrg[256];
  const XPel* pOrgMb   = pcOrg->getMbLumAddr();
  const Int   iOrgStride = pcOrg->getLStride();
  for( Int y = 0; y < 16; y++ )
    for( Int x = 0; x < 16; x++ ) asOrg[y * 16 + x] = pOrgMb[y * iOrgStride + x];
  {
    const Int bx = 4 * ( uiBlk & 3 ), by = 4 * ( uiBlk >> 2 );
    bool bWide = false;
    for( Int y = by; y < by + iYSize; y++ )
      for( Int x = bx; x < bx + iXSize; x++ ) bWide |= ( asOrg[y * 16 + x] < 0 || asOrg[y * 16 + x] > 255 );
    if( bWide ) st.n_calls_16bit++;
    if( pcBSP ) st.n_calls_bipred++;
  }
  // samples of the MB outside the searched partition may be stale in a prepared buffer: only the partition matters,
  // keep them in range so that the 8-bit check of the library looks at the partition alone
  if( pcOrg == &cPrepared )
  {
    const Int bx = 4 * ( uiBlk & 3 ), by = 4 * ( uiBlk >> 2 );
    for( Int y = 0; y < 16; y++ )
      for( Int x = 0; x < 16; x++ )
        if( x < bx || x >= bx + iXSize || y < by || y >= by + iYSize ) asOrg[y * 16 + x] = 0;
  }

  // chroma original of the macroblock for the YUV-SAD search (pUOrg / pVOrg, :259-260); without it the library searches luma only
  int16_t asOrgChroma[128];
  if( bChromaInSearch )
  {
    const XPel* pU = pcOrg->getMbCbAddr();
    const XPel* pV = pcOrg->getMbCrAddr();
    const Int   iCS = pcOrg->getCStride();
    for( Int y = 0; y < 8; y++ )
      for( Int x = 0; x < 8; x++ ) { asOrgChroma[y * 8 + x] = pU[y * iCS + x]; asOrgChroma[64 + y * 8 + x] = pV[y * iCS + x]; }
  }
  if( bYuvSad && jsvm_me_set_org_chroma( ls->ctx, bChromaInSearch ? asOrgChroma : 0, bChromaInSearch ? 1 : 0 ) != JSVM_ME_OK )
    return reportFailure( "jsvm_me_set_org_chroma" );
  // the neighbours' MVs for FAST_SEARCH / TZ_SEARCH (MbDataAccess::addMvPredictors, :293-305)
  jsvm_mv acCand[3] = { { 0, 0 }, { 0, 0 }, { 0, 0 } };
  if( !uiSearchRange && ( st.search_mode == JSVM_SEARCH_FAST || st.search_mode == JSVM_SEARCH_TZ ) )
  {
    std::vector<Mv> cList;
    rcMbDataAccess.addMvPredictors( cList );
    for( size_t n = 0; n < 3 && n < cList.size(); n++ ) { acCand[n].hor = cList[n].getHor(); acCand[n].ver = cList[n].getVer(); }
    if( jsvm_me_set_mv_candidates( ls->ctx, acCand, 1 ) != JSVM_ME_OK )
      return reportFailure( "jsvm_me_set_mv_candidates" );
  }

  jsvm_me_result cRes;
  if( jsvm_me_search( ls->ctx, &cJob, 1, asOrg, 1, &cRes, st.last_pred ) != JSVM_ME_OK )
    return reportFailure( "jsvm_me_search" );
  st.last_valid = true;

  // ---- ENC/MotionEstimation.cpp:331-337, host double arithmetic (N9, N11) ----
  const UInt uiMinSAD = cRes.min_sad;
  const UInt uiMvBits = cRes.mv_bits;
  ruiBits += uiMvBits;
  ruiCost  = (UInt)floor( fWeight * (Double)( uiMinSAD - xGetCost( uiMvBits ) ) ) + xGetCost( ruiBits );
  rcMv.set( cRes.mv.hor, cRes.mv.ver );
  m_bELWithBLMv = false;
  return Err::m_nOK;
}

// ------------------------------------------------------------------------------------------------
// MotionEstimationQuarterPel
// ------------------------------------------------------------------------------------------------
MotionEstimationQuarterPel::MotionEstimationQuarterPel() :
  m_uiBestMode( 0 )
{
}

MotionEstimationQuarterPel::~MotionEstimationQuarterPel()
{
}

ErrVal MotionEstimationQuarterPel::create( MotionEstimation*& rpcMotionEstimation )
{
  MotionEstimationQuarterPel* pc = new MotionEstimationQuarterPel;
  rpcMotionEstimation = pc;
  ROT( NULL == rpcMotionEstimation );
  return Err::m_nOK;
}

// The sub-pel search runs on the device as part of jsvm_me_search; this virtual is never reached.
Void MotionEstimationQuarterPel::xSubPelSearch( YuvPicBuffer*, Mv&, UInt&, UInt, UInt )
{
  fprintf( stderr, "jsvm_me_b200: xSubPelSearch must not be called on the host\n" );
  abort();
}

// Returns the prediction of the immediately preceding estimateBlockWithStart (the reference's stateful contract).
ErrVal MotionEstimationQuarterPel::compensateBlock( YuvMbBuffer* pcRecPelData,
                                                    UInt         uiBlk,
                                                    UInt         uiMode,
                                                    YuvMbBuffer* pcRefPelData2 )
{
  B200State& st = state();
  ROF( st.last_valid );
  pcRecPelData->set4x4Block( B4x4Idx( uiBlk ) );
  const Int iStride = pcRecPelData->getLStride();
  XPel*     pDes    = pcRecPelData->getLumBlk();
  const Int iXSize  = m_pcXDistortion->getBlockWidth ( uiMode );
  const Int iYSize  = m_pcXDistortion->getBlockHeight( uiMode );
  if( pcRefPelData2 == NULL )
  {
    for( Int y = 0; y < iYSize; y++ )
      for( Int x = 0; x < iXSize; x++ ) pDes[y * iStride + x] = st.last_pred[y * 16 + x];
  }
  else
  {
    pcRefPelData2->set4x4Block( B4x4Idx( uiBlk ) );
    const Int   iStride2 = pcRefPelData2->getLStride();
    const XPel* pSrc2    = pcRefPelData2->getLumBlk();
    for( Int y = 0; y < iYSize; y++ )
      for( Int x = 0; x < iXSize; x++ ) pDes[y * iStride + x] = ( st.last_pred[y * 16 + x] + pSrc2[y * iStride2 + x] + 1 ) >> 1;
  }
  return Err::m_nOK;
}

H264AVC_NAMESPACE_END
