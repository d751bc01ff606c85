/*
 * oracle/me_oracle.c -- TEST INFRASTRUCTURE ONLY (never part of the product path).
 *
 * CPU restatement, in plain C99, of the JSVM 9.19.8 motion-estimation hot path.  It is the
 * checker the CUDA path is compared with; see me_oracle.h for how it is itself pinned against
 * the compiled reference.  Scalar and single-threaded on purpose (as the reference is).
 *
 * Citations use R = /root/reference/JSVM/H264Extension, ENC = R/src/lib/H264AVCEncoderLib,
 * COM = R/src/lib/H264AVCCommonLib, INCC = R/include/H264AVCCommonLib.
 */
#include "me_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <limits.h>

#define MARGIN   JSVM_ME_MARGIN          /* INCC/CommonDefs.h:246 YUV_X_MARGIN; y margin = 64/2 (COM/YuvBufferCtrl.cpp:134-135) */
#define HP_ORG   (2 * MARGIN)            /* origin offset inside our half-pel plane (both axes) */

#define CMARGIN  (MARGIN / 2)            /* chroma margin, INCC/YuvPicBuffer.h:68-69 */

typedef struct {
  uint8_t* full;   /* (H+64) x (W+64), origin at (32,32) */
  uint8_t* half;   /* 2(H+64) x 2(W+64), origin at (64,64); only [-56, 2(W+28)) x [-56, 2(H+28)) is written */
  uint8_t* chroma[2];   /* Cb, Cr: (H/2+32) x (W/2+32), origin at (16,16); 128 until jsvm_ora_set_chroma_u8 */
  int      valid;
} ora_plane;

struct jsvm_ora_ctx {
  int width, height, n_slots;
  int search_mode, full_dfunc, sub_dfunc, search_range;
  ora_plane* planes;
  uint32_t factor_map[256][2]; int n_factor_map;   /* full-pel stage factor -> sub-pel stage factor */
  const int16_t* org_chroma; int n_org_chroma;
  int el_search_range, fast_bi_search;              /* MotionVectorSearchParams ELSearchRange / FastBiSearch */
  const jsvm_mv* cands; int n_cands;                /* 3 per job: m_cMv3D_A / B / C (MbDataAccess::addMvPredictors) */
};

static inline int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }   /* gClip, INCC/GlobalFunctions.h:37-44 */

/* ---------------------------------------------------------------------------------------------
 * a6  MotionEstimationCost::xGetComponentBits (ENC/MotionEstimationCost.cpp:75-87):
 *     length of the signed Exp-Golomb code of one quarter-pel difference.
 * ------------------------------------------------------------------------------------------- */
uint32_t jsvm_ora_component_bits(int d) {
  uint32_t len = 1;
  uint32_t a = (uint32_t)(d < 0 ? -d : d);
  uint32_t t = (a << 1) + 1;
  while (t != 1) { t >>= 1; len += 2; }
  return len;
}

/* xGetCost( bits ) = ( m_uiCostFactor * bits ) >> 16 in 32-bit unsigned arithmetic (ENC/MotionEstimationCost.h:27, N5). */
static inline uint32_t rate_cost(uint32_t factor, uint32_t bits) { return (uint32_t)(factor * bits) >> 16; }

/* a7  RateDistortion::setMbQpLambda (ENC/RateDistortion.cpp:59-70). The reference is built with
 * -ffloat-store; on x86-64/SSE2 doubles are already IEEE binary64 so plain C gives the same value. */
uint32_t jsvm_ora_cost_factor(double lambda, int sad) {
  volatile double s = sqrt(lambda);
  volatile double v = sad ? 65536.0 * s : 65536.0 * lambda;
  return (uint32_t)floor(v);
}

/* ---------------------------------------------------------------------------------------------
 * Block geometry: XDistortion::init tables m_aiRows / m_aiCols (ENC/Distortion.cpp:50-74).
 * ------------------------------------------------------------------------------------------- */
static int mode_size(int mode, int* w, int* h) {
  switch (mode & JSVM_JOB_MODE_MASK) {
    case JSVM_MODE_16x16: *w = 16; *h = 16; return 0;
    case JSVM_MODE_16x8:  *w = 16; *h = 8;  return 0;
    case JSVM_MODE_8x16:  *w = 8;  *h = 16; return 0;
    case JSVM_BLK_8x8:    *w = 8;  *h = 8;  return 0;
    case JSVM_BLK_8x4:    *w = 8;  *h = 4;  return 0;
    case JSVM_BLK_4x8:    *w = 4;  *h = 8;  return 0;
    case JSVM_BLK_4x4:    *w = 4;  *h = 4;  return 0;
    default: return -1;
  }
}

/* a4  XDistortion::xCalcHadamard4x4 (ENC/Distortion.cpp:291-340): vertical then horizontal 4-point
 * butterflies on the difference block, sum of absolute coefficients, halved PER 4x4 (N8). */
static uint32_t hadamard4x4(const int* d /* 4x4 diffs, row-major */) {
  int m[16], t[4];
  uint32_t sum = 0;
  for (int n = 0; n < 4; ++n) {        /* columns */
    t[0] = d[0 * 4 + n] + d[3 * 4 + n];
    t[1] = d[1 * 4 + n] + d[2 * 4 + n];
    t[2] = d[1 * 4 + n] - d[2 * 4 + n];
    t[3] = d[0 * 4 + n] - d[3 * 4 + n];
    m[0 * 4 + n] = t[0] + t[1];
    m[2 * 4 + n] = t[0] - t[1];
    m[1 * 4 + n] = t[3] + t[2];
    m[3 * 4 + n] = t[3] - t[2];
  }
  for (int n = 0; n < 4; ++n) {        /* rows */
    int a = m[n * 4 + 0] + m[n * 4 + 3];
    int b = m[n * 4 + 1] + m[n * 4 + 2];
    sum += (uint32_t)abs(a + b);
    sum += (uint32_t)abs(a - b);
    a = m[n * 4 + 1] - m[n * 4 + 2];
    b = m[n * 4 + 0] - m[n * 4 + 3];
    sum += (uint32_t)abs(a + b);
    sum += (uint32_t)abs(a - b);
  }
  return sum / 2;
}

/* Generic "sample fetch" distortion so that the same code serves the full-pel plane, the
 * half-pel candidates and the quarter-pel candidates.  get(user, x, y) returns the candidate
 * prediction sample for block pixel (x, y).
 * a3  xGetSAD16x/8x/4x (ENC/Distortion.cpp:397-429, 563-587, 691-711);
 * a5  xGetSSE*       (:512-560, 653-686, 769-794);
 * a4  xGetHAD16x/8x/4x (:798-855). */
typedef int (*sample_fn)(const void* user, int x, int y);

static uint32_t distortion(const int16_t* org /* stride 16 */, sample_fn get, const void* user, int w, int h, int dfunc) {
  uint32_t sum = 0;
  if (dfunc == JSVM_DF_HADAMARD) {
    for (int by = 0; by < h; by += 4)
      for (int bx = 0; bx < w; bx += 4) {
        int d[16];
        for (int y = 0; y < 4; ++y)
          for (int x = 0; x < 4; ++x) d[y * 4 + x] = org[(by + y) * 16 + bx + x] - get(user, bx + x, by + y);
        sum += hadamard4x4(d);
      }
    return sum;
  }
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int d = org[y * 16 + x] - get(user, x, y);
      sum += (dfunc == JSVM_DF_SSD) ? (uint32_t)(d * d) : (uint32_t)abs(d);
    }
  return sum;
}

typedef struct { const int16_t* p; int stride; } raw16_src;
static int get_raw16(const void* u, int x, int y) { const raw16_src* s = (const raw16_src*)u; return s->p[y * s->stride + x]; }

uint32_t jsvm_ora_distortion(const int16_t* org16x16, const int16_t* ref, int ref_stride, int mode, int dfunc, int blk) {
  int w, h;
  if (mode_size(mode, &w, &h)) return 0xffffffffu;
  raw16_src s = { ref, ref_stride };
  return distortion(org16x16 + ((blk >> 2) * 4) * 16 + (blk & 3) * 4, get_raw16, &s, w, h, dfunc);
}

/* ---------------------------------------------------------------------------------------------
 * Planes
 * ------------------------------------------------------------------------------------------- */
jsvm_ora_ctx* jsvm_ora_create(int width, int height, int n_slots,
                              int search_mode, int full_pel_dfunc, int sub_pel_dfunc, int search_range) {
  if (width <= 0 || height <= 0 || (width & 15) || (height & 15) || n_slots <= 0) return NULL;
  jsvm_ora_ctx* c = (jsvm_ora_ctx*)calloc(1, sizeof(*c));
  c->width = width; c->height = height; c->n_slots = n_slots;
  c->search_mode = search_mode; c->el_search_range = search_range; c->full_dfunc = full_pel_dfunc; c->sub_dfunc = sub_pel_dfunc; c->search_range = search_range;
  c->planes = (ora_plane*)calloc((size_t)n_slots, sizeof(ora_plane));
  for (int s = 0; s < n_slots; ++s) {
    c->planes[s].full = (uint8_t*)calloc((size_t)(width + 2 * MARGIN) * (height + 2 * MARGIN), 1);
    c->planes[s].half = (uint8_t*)calloc((size_t)4 * (width + 2 * MARGIN) * (height + 2 * MARGIN), 1);
    for (int k = 0; k < 2; ++k) {
      c->planes[s].chroma[k] = (uint8_t*)malloc((size_t)(width / 2 + 2 * CMARGIN) * (height / 2 + 2 * CMARGIN));
      memset(c->planes[s].chroma[k], 128, (size_t)(width / 2 + 2 * CMARGIN) * (height / 2 + 2 * CMARGIN));
    }
  }
  return c;
}

void jsvm_ora_destroy(jsvm_ora_ctx* c) {
  if (!c) return;
  for (int s = 0; s < c->n_slots; ++s) { free(c->planes[s].full); free(c->planes[s].half); free(c->planes[s].chroma[0]); free(c->planes[s].chroma[1]); }
  free(c->planes);
  free(c);
}

/* YuvPicBuffer::xFillPlaneMargin (COM/YuvPicBuffer.cpp:1608-1641): left/right replicate per row,
 * then whole padded rows are copied down / up. */
static void fill_margin_m(uint8_t* full, int W, int H, int M) {
  const int S = W + 2 * M;
  uint8_t* org = full + M * S + M;
  for (int y = 0; y < H; ++y) {
    uint8_t* r = org + y * S;
    for (int m = -M; m < 0; ++m) r[m] = r[0];
    for (int m = W; m < W + M; ++m) r[m] = r[W - 1];
  }
  for (int n = 0; n < M; ++n) {
    memcpy(org - M + (H + n) * S, org - M + (H - 1) * S, (size_t)S);
    memcpy(org - M - (n + 1) * S, org - M, (size_t)S);
  }
}
static void fill_margin(uint8_t* full, int W, int H) { fill_margin_m(full, W, H, MARGIN); }

/* a12 QuarterPelFilter::filterFrame (COM/QuarterPelFilter.cpp:61-149).
 * Pass 1 (:83-102): per picture row, x in [-28, W+28): tmp[2x] = s[x] << 5,
 *   tmp[2x+1] = s[x-2] - 5 s[x-1] + 20 s[x] + 20 s[x+1] - 5 s[x+2] + s[x+3]  (unrounded int).
 *   The temp rows are zero outside that x range (memset :79) and replicated 32x below / above
 *   from the LAST / FIRST picture row (:104-116) -- not from the plane's vertical margin.
 * Pass 2 (:125-144): y in [-28, H+28), x in [-56, 2(W+28)):
 *   out[2y][x]   = gClip( (tmp[y][x] + 16) / 32 )
 *   out[2y+1][x] = gClip( (V6(tmp)[y][x] + 512) / 1024 ), V6 = the same taps over rows y-2..y+3.
 *   '/' truncates toward zero; gClip then maps every negative quotient to 0. */
static void filter_frame(const uint8_t* full, uint8_t* half, int W, int H) {
  const int S = W + 2 * MARGIN;
  const int TS = 2 * S;                 /* iTmpXSize */
  const int TR = H + 2 * MARGIN;        /* iTmpYSize */
  const int mn = MARGIN - 4;            /* iMarginNew = 28 */
  int32_t* tmp = (int32_t*)calloc((size_t)TS * TR, sizeof(int32_t));
  const uint8_t* src = full + MARGIN * S + MARGIN;
  for (int y = 0; y < H; ++y) {
    int32_t* ps = tmp + (size_t)(MARGIN + y) * TS + 2 * MARGIN;
    const uint8_t* s = src + y * S;
    for (int x = -mn; x < W + mn; ++x) {
      int t = s[x] + s[x + 1];
      t <<= 2;
      t -= s[x - 1];
      t -= s[x + 2];
      t += t << 2;
      t += s[x - 2];
      t += s[x + 3];
      ps[2 * x] = s[x] << 5;
      ps[2 * x + 1] = t;
    }
  }
  for (int n = 0; n < MARGIN; ++n) {
    memcpy(tmp + (size_t)(MARGIN + H + n) * TS, tmp + (size_t)(MARGIN + H - 1) * TS, (size_t)TS * sizeof(int32_t));
    memcpy(tmp + (size_t)(MARGIN - 1 - n) * TS, tmp + (size_t)MARGIN * TS, (size_t)TS * sizeof(int32_t));
  }
  const int HS = 2 * S;                 /* half-pel stride */
  for (int y = -mn; y < H + mn; ++y) {
    const int32_t* ps = tmp + (size_t)(MARGIN + y) * TS + 2 * MARGIN;
    uint8_t* d0 = half + (size_t)(HP_ORG + 2 * y) * HS + HP_ORG;
    uint8_t* d1 = d0 + HS;
    for (int x = -2 * mn; x < 2 * (W + mn); ++x) {
      int t = ps[x] + ps[x + TS];
      t <<= 2;
      t -= ps[x - TS];
      t -= ps[x + 2 * TS];
      t += t << 2;
      t += ps[x - 2 * TS];
      t += ps[x + 3 * TS];
      d0[x] = (uint8_t)clip255((ps[x] + 16) / 32);
      d1[x] = (uint8_t)clip255((t + 512) / 1024);
    }
  }
  free(tmp);
}

int jsvm_ora_set_plane_u8(jsvm_ora_ctx* c, int slot, const uint8_t* luma, int stride) {
  if (!c || slot < 0 || slot >= c->n_slots) return -1;
  const int W = c->width, H = c->height, S = W + 2 * MARGIN;
  ora_plane* p = &c->planes[slot];
  for (int y = 0; y < H; ++y) memcpy(p->full + (size_t)(MARGIN + y) * S + MARGIN, luma + (size_t)y * stride, (size_t)W);
  fill_margin(p->full, W, H);
  memset(p->half, 0, (size_t)4 * S * (H + 2 * MARGIN));
  filter_frame(p->full, p->half, W, H);
  p->valid = 1;
  return 0;
}

/* Chroma planes of a slot and their margins (YuvPicBuffer::fillMargin, COM/YuvPicBuffer.cpp:1595-1606: the same replication
 * with the chroma margin). */
int jsvm_ora_set_chroma_u8(jsvm_ora_ctx* c, int slot, const uint8_t* cb, const uint8_t* cr, int stride) {
  if (!c || slot < 0 || slot >= c->n_slots) return -1;
  const int CW = c->width / 2, CH = c->height / 2, CS = CW + 2 * CMARGIN;
  for (int k = 0; k < 2; ++k) {
    uint8_t* p = c->planes[slot].chroma[k];
    for (int y = 0; y < CH; ++y) memcpy(p + (size_t)(CMARGIN + y) * CS + CMARGIN, (k ? cr : cb) + (size_t)y * stride, (size_t)CW);
    fill_margin_m(p, CW, CH, CMARGIN);
  }
  return 0;
}

int jsvm_ora_set_org_chroma(jsvm_ora_ctx* c, const int16_t* org_chroma_mbs, int n_org_mbs) {
  if (!c) return -1;
  c->org_chroma = org_chroma_mbs; c->n_org_chroma = org_chroma_mbs ? n_org_mbs : 0;
  return 0;
}

int jsvm_ora_map_cost_factor(jsvm_ora_ctx* c, uint32_t factor_full_pel, uint32_t factor_sub_pel) {
  if (!c) return -1;
  for (int k = 0; k < c->n_factor_map; ++k)
    if (c->factor_map[k][0] == factor_full_pel) { c->factor_map[k][1] = factor_sub_pel; return 0; }
  if (c->n_factor_map >= 256) return -1;
  c->factor_map[c->n_factor_map][0] = factor_full_pel; c->factor_map[c->n_factor_map][1] = factor_sub_pel;
  c->n_factor_map++;
  return 0;
}

int jsvm_ora_get_fullpel_u8(jsvm_ora_ctx* c, int slot, uint8_t* dst, int dst_stride) {
  if (!c || slot < 0 || slot >= c->n_slots) return -1;
  const int S = c->width + 2 * MARGIN;
  for (int y = 0; y < c->height + 2 * MARGIN; ++y) memcpy(dst + (size_t)y * dst_stride, c->planes[slot].full + (size_t)y * S, (size_t)S);
  return 0;
}

int jsvm_ora_get_halfpel_u8(jsvm_ora_ctx* c, int slot, uint8_t* dst, int dst_stride) {
  if (!c || slot < 0 || slot >= c->n_slots) return -1;
  const int HS = 2 * (c->width + 2 * MARGIN);
  const int m = 2 * (MARGIN - 4);
  const int rows = 2 * c->height + 2 * m, cols = 2 * c->width + 2 * m;
  for (int y = 0; y < rows; ++y)
    memcpy(dst + (size_t)y * dst_stride, c->planes[slot].half + (size_t)(HP_ORG - m + y) * HS + (HP_ORG - m), (size_t)cols);
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Search
 * ------------------------------------------------------------------------------------------- */
static inline int16_t sat16(int v) { return (int16_t)(v < SHRT_MIN ? SHRT_MIN : (v > SHRT_MAX ? SHRT_MAX : v)); }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

typedef struct { int neg_ver, pos_ver, neg_hor, pos_hor; } search_rect;

/* a8 MotionCompensation::initMb (COM/MotionCompensation.cpp:152-167) -> m_cMin / m_cMax (quarter-pel, per MB);
 * a2 limitComponents + '>>= 2' (ENC/MotionEstimation.cpp:373-374, INCC/Mv.h:80-85, 118-122);
 *    SearchRect::init (ENC/MotionEstimation.h:44-56) with its Short casts (N2). */
static void search_window(const jsvm_ora_ctx* c, const jsvm_me_job* j, int* cx, int* cy, search_rect* r) {
  const int mbw = c->width >> 4, mbh = c->height >> 4;
  int16_t min_h = sat16(((-(int)j->mb_x << 4) - 24) << 2), min_v = sat16(((-(int)j->mb_y << 4) - 24) << 2);
  int16_t max_h = sat16((((mbw - j->mb_x) << 4) + 8) << 2), max_v = sat16((((mbh - j->mb_y) << 4) + 8) << 2);
  int16_t h = (int16_t)imin(max_h, imax(min_h, j->start.hor));
  int16_t v = (int16_t)imin(max_v, imax(min_v, j->start.ver));
  h >>= 2; v >>= 2;
  int16_t mnh = min_h >> 2, mnv = min_v >> 2, mxh = max_h >> 2, mxv = max_v >> 2;
  int16_t lim = (int16_t)(j->search_range ? j->search_range : c->search_range);
  int16_t pos_v = (int16_t)(mxv - v), neg_v = (int16_t)(v - mnv);
  int16_t pos_h = (int16_t)(mxh - h), neg_h = (int16_t)(h - mnh);
  r->neg_ver = imin(lim, neg_v) - v;
  r->pos_ver = imin(lim, pos_v) + v;
  r->neg_hor = imin(lim, neg_h) - h;
  r->pos_hor = imin(lim, pos_h) + h;
  *cx = h; *cy = v;
}

int jsvm_ora_window_positions(jsvm_ora_ctx* c, const jsvm_me_job* job) {
  int cx, cy; search_rect r;
  search_window(c, job, &cx, &cy, &r);
  int nx = r.pos_hor + r.neg_hor + 1, ny = r.pos_ver + r.neg_ver + 1;
  return (nx > 0 && ny > 0) ? nx * ny : 0;
}

typedef struct { const uint8_t* p; int stride; } u8_src;                 /* full-pel plane candidate */
static int get_u8(const void* u, int x, int y) { const u8_src* s = (const u8_src*)u; return s->p[y * s->stride + x]; }

typedef struct { const uint8_t* p; int stride; } hp_src;                 /* half-pel candidate: every 2nd sample */
static int get_hp(const void* u, int x, int y) { const hp_src* s = (const hp_src*)u; return s->p[2 * y * s->stride + 2 * x]; }

/* a11 QuarterPelFilter::xXFilter1..4 (COM/QuarterPelFilter.cpp:152-262) for one of the 8 neighbours
 * (dx,dy) of the half-pel winner at c: filters 0 and 3 (winner on an integer or on the diagonal
 * half sample) average the two AXIS neighbours for the diagonal quarter positions; filters 1 and 2
 * average the centre with the neighbour itself. Rounding (p+q+1)>>1. */
typedef struct { const uint8_t* c; int stride; int dx, dy; int axis_diag; } qp_src;
static int get_qp(const void* u, int x, int y) {
  const qp_src* s = (const qp_src*)u;
  const uint8_t* o = s->c + 2 * y * s->stride + 2 * x;
  if (s->axis_diag && s->dx && s->dy) return (o[s->dx] + o[s->dy * s->stride] + 1) >> 1;
  return (o[0] + o[s->dy * s->stride + s->dx] + 1) >> 1;
}

/* candidate tables, ENC/MotionEstimationQuarterPel.cpp:11-41 */
static const uint8_t kFilterOfHalf[9] = { 0, 2, 2, 1, 1, 3, 3, 3, 3 };
static const int8_t kHalfMv[9][2] = { {0,0}, {0,-2}, {0,2}, {-2,0}, {2,0}, {-2,-2}, {2,-2}, {-2,2}, {2,2} };
static const int8_t kQuarterMv[9][2] = { {0,0}, {0,-1}, {0,1}, {-1,-1}, {1,-1}, {-1,0}, {1,0}, {-1,1}, {1,1} };


/* ---------------------------------------------------------------------------------------------
 * Full-pel stage.  fp_cost() is m_cXDSS.Func + xGetCost( x, y ) at one integer displacement; the strategies below restate
 *   xPelBlockSearch  ENC/MotionEstimation.cpp:351-412      xPelSpiralSearch :415-468 (spiral table built in init, :69-113)
 *   xPelLogSearch    :472-709 (LOG_SEARCH and, with the candidate list, FAST_SEARCH)
 *   xTZSearch        :1106-1292 with xTZCheckPoint :733-757, xTZ2PointSearch :760-880, xTZ8PointDiamondSearch :929-1103 and
 *                    the TZ_SEARCH_CONFIGURATION constants :713-730
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const jsvm_ora_ctx* c; const jsvm_me_job* j;
  int w, h, dfunc, S, CS;
  uint32_t f;
  const int16_t* org; const int16_t* corg;       /* org: stride 16; corg: [2][64], stride 8 */
  const uint8_t* ref0; const uint8_t* cref0[2];
} fp_eval;

static uint32_t fp_cost(const fp_eval* e, int x, int y) {
  uint32_t sad;
  if (e->dfunc == JSVM_DF_SAD || e->dfunc == JSVM_DF_YUV_SAD) {   /* same sums as distortion(); direct loops keep the checker fast */
    const uint8_t* rp = e->ref0 + y * e->S + x;
    sad = 0;
    for (int yy = 0; yy < e->h; ++yy)
      for (int xx = 0; xx < e->w; ++xx) sad += (uint32_t)abs(e->org[yy * 16 + xx] - rp[yy * e->S + xx]);
    if (e->dfunc == JSVM_DF_YUV_SAD)
      for (int k = 0; k < 2; ++k) {
        const uint8_t* cp = e->cref0[k] + (y >> 1) * e->CS + (x >> 1);
        for (int yy = 0; yy < e->h / 2; ++yy)
          for (int xx = 0; xx < e->w / 2; ++xx) sad += (uint32_t)abs(e->corg[k * 64 + yy * 8 + xx] - cp[yy * e->CS + xx]);
      }
  } else {
    u8_src s = { e->ref0 + y * e->S + x, e->S };
    sad = distortion(e->org, get_u8, &s, e->w, e->h, e->dfunc);
  }
  /* xGetCost( x, y ) with m_uiMvScaleShift = 2 (N6): table index (x<<2) - pred */
  return sad + rate_cost(e->f, jsvm_ora_component_bits((x << 2) - e->j->pred.hor) + jsvm_ora_component_bits((y << 2) - e->j->pred.ver));
}
static uint32_t fp_rate(const fp_eval* e, int x, int y) {
  return rate_cost(e->f, jsvm_ora_component_bits((x << 2) - e->j->pred.hor) + jsvm_ora_component_bits((y << 2) - e->j->pred.ver));
}

static void mb_limits(const jsvm_ora_ctx* c, const jsvm_me_job* j, int16_t* min_h, int16_t* min_v, int16_t* max_h, int16_t* max_v) {
  const int mbw = c->width >> 4, mbh = c->height >> 4;
  *min_h = sat16(((-(int)j->mb_x << 4) - 24) << 2); *min_v = sat16(((-(int)j->mb_y << 4) - 24) << 2);
  *max_h = sat16((((mbw - j->mb_x) << 4) + 8) << 2); *max_v = sat16((((mbh - j->mb_y) << 4) + 8) << 2);
}
/* Mv::limitComponents + '>>= 2' of one vector */
static void limit_pel(const jsvm_ora_ctx* c, const jsvm_me_job* j, jsvm_mv mv, int* x, int* y) {
  int16_t a, b, cc, d; mb_limits(c, j, &a, &b, &cc, &d);
  *x = (int16_t)imin(cc, imax(a, mv.hor)) >> 2; *y = (int16_t)imin(d, imax(b, mv.ver)) >> 2;
}
/* SearchRect::init around the clipped start with range `lim` */
static void rect_init(const jsvm_ora_ctx* c, const jsvm_me_job* j, int lim_in, int* cx, int* cy, search_rect* r) {
  int16_t min_h, min_v, max_h, max_v; mb_limits(c, j, &min_h, &min_v, &max_h, &max_v);
  limit_pel(c, j, j->start, cx, cy);
  const int16_t h = (int16_t)*cx, v = (int16_t)*cy, lim = (int16_t)lim_in;
  const int16_t mnh = min_h >> 2, mnv = min_v >> 2, mxh = max_h >> 2, mxv = max_v >> 2;
  const int16_t pos_v = (int16_t)(mxv - v), neg_v = (int16_t)(v - mnv), pos_h = (int16_t)(mxh - h), neg_h = (int16_t)(h - mnh);
  r->neg_ver = imin(lim, neg_v) - v; r->pos_ver = imin(lim, pos_v) + v;
  r->neg_hor = imin(lim, neg_h) - h; r->pos_hor = imin(lim, pos_h) + h;
}

static const jsvm_mv kZeroCands[3] = { {0, 0}, {0, 0}, {0, 0} };

static void search_block(const fp_eval* e, int* bx, int* by, uint32_t* best_io) {
  int cx, cy; search_rect r;
  search_window(e->c, e->j, &cx, &cy, &r);
  uint32_t best = 0xffffffffu;
  int best_x = cx, best_y = cy;                             /* rcMv keeps the clipped start if the window is empty */
  for (int y = -r.neg_ver; y <= r.pos_ver; ++y)
    for (int x = -r.neg_hor; x <= r.pos_hor; ++x) {
      const uint32_t sad = fp_cost(e, x, y);
      if (best > sad) { best = sad; best_x = x; best_y = y; }           /* first strict minimum in raster order (N1) */
    }
  *best_io = best - fp_rate(e, best_x, best_y); *bx = best_x; *by = best_y;
}

/* xPelSpiralSearch: the start is limited to [ m_cMin + 4 SR, m_cMax - 4 SR ], then all (2 SR + 1)^2 offsets of the spiral table
 * are tried in table order; SR is always the configured range.  Returns -1 where the reference would read outside its planes
 * (picture too small for the range: the limited interval is empty). */
static int search_spiral(const fp_eval* e, int* bx, int* by, uint32_t* best_io) {
  const int sr = e->c->search_range;
  int16_t min_h, min_v, max_h, max_v; mb_limits(e->c, e->j, &min_h, &min_v, &max_h, &max_v);
  const int16_t lo_h = (int16_t)(min_h + 4 * sr), lo_v = (int16_t)(min_v + 4 * sr), hi_h = (int16_t)(max_h - 4 * sr), hi_v = (int16_t)(max_v - 4 * sr);
  if (lo_h > hi_h || lo_v > hi_v) return -1;
  const int cx = (int16_t)imin(hi_h, imax(lo_h, e->j->start.hor)) >> 2, cy = (int16_t)imin(hi_v, imax(lo_v, e->j->start.ver)) >> 2;
  uint32_t best = fp_cost(e, cx, cy);                        /* entry 0 = (0,0) */
  int ox = 0, oy = 0;
#define SPIRAL_TRY(dx, dy) do { const uint32_t sad_ = fp_cost(e, cx + (dx), cy + (dy)); if (best > sad_) { best = sad_; ox = (dx); oy = (dy); } } while (0)
  for (int off = 1; off <= sr; ++off) {
    for (int l = -off + 1; l < off; ++l) { SPIRAL_TRY(l, -off); SPIRAL_TRY(l, off); }
    for (int l = -off; l <= off; ++l)    { SPIRAL_TRY(-off, l); SPIRAL_TRY(off, l); }
  }
#undef SPIRAL_TRY
  *bx = cx + ox; *by = cy + oy; *best_io = best - fp_rate(e, *bx, *by);
  return 0;
}

/* MotionEstimation::init, :50-58 */
static int max_log_step(int search_range) {
  unsigned lim = (unsigned)search_range >> 4; int step = 1;
  while (lim) { lim >>= 1; step <<= 1; }
  return step;
}

/* xPelLogSearch( ..., bFme = (cands != NULL), uiStep ) */
static void search_log(const fp_eval* e, const jsvm_mv* cands, int step_start, int* bx, int* by, uint32_t* best_io) {
  int x, y; search_rect r;
  rect_init(e->c, e->j, e->c->search_range, &x, &y, &r);
  uint32_t best = fp_cost(e, x, y);
  if (cands) {
    jsvm_mv list[4] = { e->j->pred, cands[0], cands[1], cands[2] };       /* :283-284: rcMvPred, then A, B, C */
    for (int n = 0; n < 4; ++n) {
      int px, py; limit_pel(e->c, e->j, list[n], &px, &py);
      const uint32_t sad = fp_cost(e, px, py);
      if (best > sad) { best = sad; x = px; y = py; }
    }
  }
  int dx_old = 0, dy_old = 0;
  for (int step = step_start; step != 0; step >>= 1) {
    int go = 1;
    dx_old = 0; dy_old = 0;
    while (go) {
      int dx = 0, dy = 0;
      uint32_t sad;
      if (dx_old <= 0 && -x <= r.neg_hor - step) { sad = fp_cost(e, x - step, y); if (best > sad) { best = sad; dx = -step; } }
      if (dx_old >= 0 && x < r.pos_hor - step)   { sad = fp_cost(e, x + step, y); if (best > sad) { best = sad; dx = step; } }
      if (dy_old <= 0 && -y <= r.neg_ver - step) { sad = fp_cost(e, x, y - step); if (best > sad) { best = sad; dx = 0; dy = -step; } }
      if (dy_old >= 0 && y < r.pos_ver - step)   { sad = fp_cost(e, x, y + step); if (best > sad) { best = sad; dx = 0; dy = step; } }
      go = (dx != 0 || dy != 0);
      if (go) { dx_old = dx; dy_old = dy; x += dx; y += dy; }
    }
  }
  {                                                         /* the four diagonal neighbours, once */
    int dx = 0, dy = 0;
    uint32_t sad;
    if (dx_old != 1 && -x <= r.neg_hor - 1) {
      if (dy_old != 1 && -y <= r.neg_ver - 1) { sad = fp_cost(e, x - 1, y - 1); if (best > sad) { best = sad; dx = -1; dy = -1; } }
      if (dy_old != -1 && y < r.pos_ver - 1)  { sad = fp_cost(e, x - 1, y + 1); if (best > sad) { best = sad; dx = -1; dy = 1; } }
    }
    if (dx_old != -1 && x < r.pos_hor - 1) {
      if (dy_old != 1 && -y <= r.neg_ver - 1) { sad = fp_cost(e, x + 1, y - 1); if (best > sad) { best = sad; dx = 1; dy = -1; } }
      if (dy_old != -1 && y < r.pos_ver - 1)  { sad = fp_cost(e, x + 1, y + 1); if (best > sad) { best = sad; dx = 1; dy = 1; } }
    }
    x += dx; y += dy;
  }
  *bx = x; *by = y; *best_io = best - fp_rate(e, x, y);
}

/* ---- xTZSearch ---- */
typedef struct { const fp_eval* e; search_rect r; uint32_t best; int bx, by; unsigned dist; int nr; } tz_state;

static void tz_check(tz_state* t, int x, int y, int nr, unsigned dist) {                         /* xTZCheckPoint */
  const uint32_t sad = fp_cost(t->e, x, y);
  if (t->best > sad) { t->best = sad; t->bx = x; t->by = y; t->dist = dist; t->nr = nr; }
}

static void tz_two_points(tz_state* t) {                                                          /* xTZ2PointSearch */
  const int x = t->bx, y = t->by;
  const int l = (x - 1) >= -t->r.neg_hor, rr = (x + 1) <= t->r.pos_hor, u = (y - 1) >= -t->r.neg_ver, d = (y + 1) <= t->r.pos_ver;
  switch (t->nr) {
    case 1: if (l) tz_check(t, x - 1, y, 0, 2); if (u) tz_check(t, x, y - 1, 0, 2); break;
    case 2: if (u) { if (l) tz_check(t, x - 1, y - 1, 0, 2); if (rr) tz_check(t, x + 1, y - 1, 0, 2); } break;
    case 3: if (u) tz_check(t, x, y - 1, 0, 2); if (rr) tz_check(t, x + 1, y, 0, 2); break;
    case 4: if (l) { if (d) tz_check(t, x - 1, y + 1, 0, 2); if (u) tz_check(t, x - 1, y - 1, 0, 2); } break;
    case 5: if (rr) { if (u) tz_check(t, x + 1, y - 1, 0, 2); if (d) tz_check(t, x + 1, y + 1, 0, 2); } break;
    case 6: if (l) tz_check(t, x - 1, y, 0, 2); if (d) tz_check(t, x, y + 1, 0, 2); break;
    case 7: if (d) { if (l) tz_check(t, x - 1, y + 1, 0, 2); if (rr) tz_check(t, x + 1, y + 1, 0, 2); } break;
    case 8: if (rr) tz_check(t, x + 1, y, 0, 2); if (d) tz_check(t, x, y + 1, 0, 2); break;
    default: break;                                          /* AF() in the reference: not reachable (see search_tz) */
  }
}

static void tz_diamond(tz_state* t, int sx, int sy, int dist) {                                   /* xTZ8PointDiamondSearch */
  const search_rect* r = &t->r;
  const int top = sy - dist, bot = sy + dist, left = sx - dist, right = sx + dist;
  const int t_ok = top >= -r->neg_ver, b_ok = bot <= r->pos_ver, l_ok = left >= -r->neg_hor, r_ok = right <= r->pos_hor;
  if (dist == 1) {
    if (t_ok) tz_check(t, sx, top, 2, 1);
    if (l_ok) tz_check(t, left, sy, 4, 1);
    if (r_ok) tz_check(t, right, sy, 5, 1);
    if (b_ok) tz_check(t, sx, bot, 7, 1);
  } else if (dist <= 8) {
    const int h = dist >> 1, top2 = sy - h, bot2 = sy + h, left2 = sx - h, right2 = sx + h;
    if (t_ok && l_ok && r_ok && b_ok) {
      tz_check(t, sx, top, 2, dist); tz_check(t, left2, top2, 1, h); tz_check(t, right2, top2, 3, h); tz_check(t, left, sy, 4, dist);
      tz_check(t, right, sy, 5, dist); tz_check(t, left2, bot2, 6, h); tz_check(t, right2, bot2, 8, h); tz_check(t, sx, bot, 7, dist);
    } else {
      if (t_ok) tz_check(t, sx, top, 2, dist);
      if (top2 >= -r->neg_ver) {
        if (left2 >= -r->neg_hor) tz_check(t, left2, top2, 1, h);
        if (right2 <= r->pos_hor) tz_check(t, right2, top2, 3, h);
      }
      if (l_ok) tz_check(t, left, sy, 4, dist);
      if (r_ok) tz_check(t, right, sy, 5, dist);
      if (bot2 <= r->pos_ver) {
        if (left2 >= -r->neg_hor) tz_check(t, left2, bot2, 6, h);
        if (right2 <= r->pos_hor) tz_check(t, right2, bot2, 8, h);
      }
      if (b_ok) tz_check(t, sx, bot, 7, dist);
    }
  } else {
    const int q = dist >> 2;
    const int all = t_ok && l_ok && r_ok && b_ok;
    if (all || t_ok) tz_check(t, sx, top, 0, dist);
    if (all || l_ok) tz_check(t, left, sy, 0, dist);
    if (all || r_ok) tz_check(t, right, sy, 0, dist);
    if (all || b_ok) tz_check(t, sx, bot, 0, dist);
    for (int k = 1; k < 4; ++k) {
      const int yt = top + q * k, yb = bot - q * k, xl = sx - q * k, xr = sx + q * k;
      if (all) { tz_check(t, xl, yt, 0, dist); tz_check(t, xr, yt, 0, dist); tz_check(t, xl, yb, 0, dist); tz_check(t, xr, yb, 0, dist); }
      else {
        if (yt >= -r->neg_ver) { if (xl >= -r->neg_hor) tz_check(t, xl, yt, 0, dist); if (xr <= r->pos_hor) tz_check(t, xr, yt, 0, dist); }
        if (yb <= r->pos_ver)  { if (xl >= -r->neg_hor) tz_check(t, xl, yb, 0, dist); if (xr <= r->pos_hor) tz_check(t, xr, yb, 0, dist); }
      }
    }
  }
}

/* cands == NULL: empty candidate list (the FastBiSearch refinement).  Configuration constants of :713-730: raster 3, other
 * predicted MVs and the zero vector tested, diamond first search without early stop, raster search when the best distance
 * exceeds 3, star refinement with diamonds. */
static void search_tz(const fp_eval* e, const jsvm_mv* cands, int range, int* bx, int* by, uint32_t* best_io) {
  tz_state t; int sx, sy;
  t.e = e; t.best = 0xffffffffu; t.bx = t.by = 0; t.dist = 0; t.nr = 0;
  rect_init(e->c, e->j, range, &sx, &sy, &t.r);
  tz_check(&t, sx, sy, 0, 0);
  if (cands) {
    jsvm_mv list[4] = { e->j->pred, cands[0], cands[1], cands[2] };
    for (int n = 0; n < 4; ++n) { int px, py; limit_pel(e->c, e->j, list[n], &px, &py); tz_check(&t, px, py, 0, 0); }
  }
  tz_check(&t, 0, 0, 0, 0);
  sx = t.bx; sy = t.by;
  for (int dist = 1; dist <= range; dist *= 2) tz_diamond(&t, sx, sy, dist);
  if (t.dist == 1) { t.dist = 0; tz_two_points(&t); }
  if (t.dist > 3) {
    t.dist = 3;
    for (int y = -t.r.neg_ver; y <= t.r.pos_ver; y += 3)
      for (int x = -t.r.neg_hor; x <= t.r.pos_hor; x += 3) tz_check(&t, x, y, 0, 3);
  }
  while (t.dist > 0) {
    sx = t.bx; sy = t.by; t.dist = 0; t.nr = 0;
    for (int dist = 1; dist < range + 1; dist *= 2) tz_diamond(&t, sx, sy, dist);
    if (t.dist == 1) { t.dist = 0; if (t.nr != 0) tz_two_points(&t); }
  }
  *bx = t.bx; *by = t.by; *best_io = t.best - fp_rate(e, t.bx, t.by);
}

int jsvm_ora_set_fast_search_params(jsvm_ora_ctx* c, int el_search_range, int fast_bi_search) {
  if (!c) return -1;
  c->el_search_range = el_search_range ? el_search_range : c->search_range;      /* ENC/CodingParameter.cpp:25-28 */
  c->fast_bi_search = fast_bi_search;
  return 0;
}

int jsvm_ora_set_mv_candidates(jsvm_ora_ctx* c, const jsvm_mv* abc, int n_jobs) {
  if (!c) return -1;
  c->cands = abc; c->n_cands = abc ? n_jobs : 0;
  return 0;
}

int jsvm_ora_search(jsvm_ora_ctx* c, const jsvm_me_job* jobs, int n_jobs,
                    const int16_t* org_mbs, int n_org_mbs,
                    jsvm_me_result* results, uint8_t* preds) {
  if (!c) return -1;
  const int W = c->width, H = c->height, S = W + 2 * MARGIN, HS = 2 * S;
  for (int i = 0; i < n_jobs; ++i) {
    const jsvm_me_job* j = &jobs[i];
    int w, h;
    if (mode_size(j->mode, &w, &h)) return -1;
    if (j->cur_slot < 0 || j->cur_slot >= c->n_slots || j->ref_slot < 0 || j->ref_slot >= c->n_slots) return -1;
    if (j->mb_x < 0 || j->mb_x >= (W >> 4) || j->mb_y < 0 || j->mb_y >= (H >> 4) || j->blk > 15) return -1;
    if (j->org_index >= n_org_mbs) return -1;
    const ora_plane* cur = &c->planes[j->cur_slot];
    const ora_plane* ref = &c->planes[j->ref_slot];
    if (!ref->valid || (j->org_index < 0 && !cur->valid)) return -1;

    /* ENC/MotionEstimation.cpp:142-152: position org / ref pointers on the 4x4 block uiBlk */
    const int bx0 = 16 * j->mb_x + 4 * (j->blk & 3), by0 = 16 * j->mb_y + 4 * (j->blk >> 2);
    int16_t org[256];                                       /* block-relative, stride 16 */
    memset(org, 0, sizeof(org));
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        if (j->org_index >= 0)
          org[y * 16 + x] = org_mbs[(size_t)j->org_index * 256 + (4 * (j->blk >> 2) + y) * 16 + 4 * (j->blk & 3) + x];
        else
          org[y * 16 + x] = cur->full[(size_t)(MARGIN + by0 + y) * S + MARGIN + bx0 + x];
      }
    /* xSetMEPars( 2, 1 != FullPelDFunc ) / xSetMEPars( 0, 1 != ( 1 & SubPelDFunc ) ) (ENC/MotionEstimation.cpp:255, 325): the job
     * carries the factor of the full-pel stage; when the sub-pel stage is of the other class its factor comes from the map */
    const uint32_t f_full = j->cost_factor;
    uint32_t f = f_full;
    if ((c->full_dfunc == JSVM_DF_SSD) != (c->sub_dfunc == JSVM_DF_SSD)) {
      int found = 0;
      for (int k = 0; k < c->n_factor_map; ++k) if (c->factor_map[k][0] == f_full) { f = c->factor_map[k][1]; found = 1; }
      if (!found) return -1;
    }
    /* a5 YUV_SAD (xGetYuvSAD16x/8x/4x, ENC/Distortion.cpp:435-505, 593-647, 718-764): luma SAD + SAD of the (w/2) x (h/2) Cb and
     * Cr blocks displaced by ( x>>1, y>>1 ) (ENC/MotionEstimation.cpp:386-387).  Falls back to SAD for originals prepared
     * without chroma (bi-pred / unequal weights, :250-253). */
    int full_dfunc = c->full_dfunc;
    int16_t corg[2][64];
    memset(corg, 0, sizeof(corg));
    if (full_dfunc == JSVM_DF_YUV_SAD) {
      if ((j->org_index >= 0 && !c->org_chroma) || (j->mode & JSVM_JOB_BIPRED)) full_dfunc = JSVM_DF_SAD;
      else {
        if (j->org_index >= 0 && j->org_index >= c->n_org_chroma) return -1;
        const int CS = W / 2 + 2 * CMARGIN;
        for (int k = 0; k < 2; ++k)
          for (int y = 0; y < h / 2; ++y)
            for (int x = 0; x < w / 2; ++x) {
              const int cy = 2 * (j->blk >> 2) + y, cx = 2 * (j->blk & 3) + x;
              corg[k][y * 8 + x] = j->org_index >= 0 ? c->org_chroma[(size_t)j->org_index * 128 + k * 64 + cy * 8 + cx]
                                                     : cur->chroma[k][(size_t)(CMARGIN + 8 * j->mb_y + cy) * CS + CMARGIN + 8 * j->mb_x + cx];
            }
      }
    }

    /* ---- full-pel stage: the search strategy of ENC/MotionEstimation.cpp:262-307 ---- */
    fp_eval ev;
    ev.c = c; ev.j = j; ev.w = w; ev.h = h; ev.org = org; ev.corg = &corg[0][0]; ev.dfunc = full_dfunc; ev.f = f_full;
    ev.S = S; ev.ref0 = ref->full + (size_t)(MARGIN + by0) * S + MARGIN + bx0;
    ev.CS = W / 2 + 2 * CMARGIN;
    for (int k = 0; k < 2; ++k) ev.cref0[k] = ref->chroma[k] + (size_t)(CMARGIN + by0 / 2) * ev.CS + CMARGIN + bx0 / 2;
    uint32_t best = 0xffffffffu;                            /* MSYS_UINT_MAX */
    int best_x = 0, best_y = 0;
    const jsvm_mv* cand = (c->cands && i < c->n_cands) ? c->cands + 3 * (size_t)i : NULL;
    const int bipred = (j->mode & JSVM_JOB_BIPRED) != 0, el = (j->mode & JSVM_JOB_EL) != 0;
    if (j->search_range) {                                  /* :264-274: the iterative bi-prediction refinement */
      if (c->fast_bi_search) search_tz(&ev, NULL, j->search_range, &best_x, &best_y, &best);
      else search_block(&ev, &best_x, &best_y, &best);
    } else switch (c->search_mode) {
      case JSVM_SEARCH_BLOCK:  search_block(&ev, &best_x, &best_y, &best); break;
      case JSVM_SEARCH_SPIRAL: if (search_spiral(&ev, &best_x, &best_y, &best)) return -1; break;
      case JSVM_SEARCH_LOG: case JSVM_SEARCH_FAST:
        if (full_dfunc == JSVM_DF_YUV_SAD) return -1;       /* xPelLogSearch leaves pUSearch / pVSearch unset: undefined in the reference */
        if (c->search_mode == JSVM_SEARCH_LOG) search_log(&ev, NULL, max_log_step(c->search_range) << (bipred ? 0 : 1), &best_x, &best_y, &best);
        else search_log(&ev, cand ? cand : kZeroCands, bipred ? 1 : 2, &best_x, &best_y, &best);
        break;
      case JSVM_SEARCH_TZ:     search_tz(&ev, cand ? cand : kZeroCands, el ? c->el_search_range : c->search_range, &best_x, &best_y, &best); break;
      default: return -1;
    }
    jsvm_me_result* out = &results[i];
    out->mv_fullpel.hor = (int16_t)best_x; out->mv_fullpel.ver = (int16_t)best_y;
    out->sad_fullpel = best;

    /* ---- a9 xSubPelSearch (ENC/MotionEstimationQuarterPel.cpp:94-167) ---- */
    int16_t mvh = (int16_t)(best_x << 2), mvv = (int16_t)(best_y << 2);           /* cMv <<= 2 (:313) */
    /* a10 xCompensateBlocksHalf (:226-263): base = half-pel block origin + (mv>>1) */
    const uint8_t* hp0 = ref->half + (size_t)(HP_ORG + 2 * by0) * HS + HP_ORG + 2 * bx0;
    const uint8_t* hbase = hp0 + (mvh >> 1) + (mvv >> 1) * HS;
    uint32_t best_sad = 0xffffffffu;
    uint32_t best_mode = 0;
    int16_t bh = mvh, bv = mvv;
    for (int n = 0; n < 9; ++n) {
      /* the 9 buffers of xInitBuffer (:69-79) are the half-pel plane sampled at offset (dx/2, dy/2) */
      hp_src s = { hbase + (kHalfMv[n][0] / 2) + (kHalfMv[n][1] / 2) * HS, HS };
      uint32_t sad = distortion(org, get_hp, &s, w, h, c->sub_dfunc);
      int16_t th = (int16_t)(mvh + kHalfMv[n][0]), tv = (int16_t)(mvv + kHalfMv[n][1]);
      sad += rate_cost(f, jsvm_ora_component_bits(th - j->pred.hor) + jsvm_ora_component_bits(tv - j->pred.ver));
      if (sad < best_sad) { best_mode = (uint32_t)n; best_sad = sad; bh = th; bv = tv; }
    }
    mvh = bh; mvv = bv;
    const uint8_t* qc = hp0 + (mvh >> 1) + (mvv >> 1) * HS;       /* :137-139 */
    const int filt = kFilterOfHalf[best_mode];
    const int axis_diag = (filt == 0 || filt == 3);
    int q_winner = 0;
    for (int n = 1; n < 9; ++n) {
      qp_src s = { qc, HS, kQuarterMv[n][0], kQuarterMv[n][1], axis_diag };
      uint32_t sad = distortion(org, get_qp, &s, w, h, c->sub_dfunc);
      int16_t th = (int16_t)(mvh + kQuarterMv[n][0]), tv = (int16_t)(mvv + kQuarterMv[n][1]);
      sad += rate_cost(f, jsvm_ora_component_bits(th - j->pred.hor) + jsvm_ora_component_bits(tv - j->pred.ver));
      if (sad < best_sad) { best_mode = (uint32_t)n << 4; best_sad = sad; bh = th; bv = tv; q_winner = n; }
    }
    /* ---- ENC/MotionEstimation.cpp:331-337 (fWeight = 1.0) ---- */
    uint32_t mv_bits = jsvm_ora_component_bits(bh - j->pred.hor) + jsvm_ora_component_bits(bv - j->pred.ver);
    out->mv.hor = bh; out->mv.ver = bv;
    out->min_sad = best_sad;
    out->mv_bits = mv_bits;
    out->bits = j->bits_in + mv_bits;
    out->cost = (uint32_t)floor(1.0 * (double)(uint32_t)(best_sad - rate_cost(f, mv_bits))) + rate_cost(f, out->bits);
    out->best_mode = best_mode;

    /* ---- a13 compensateBlock (ENC/MotionEstimationQuarterPel.cpp:266-328), single-buffer branch ---- */
    if (preds) {
      uint8_t* dp = preds + (size_t)i * 256;
      memset(dp, 0, 256);
      if (best_mode < 0x10) {
        hp_src s = { hbase + (kHalfMv[best_mode][0] / 2) + (kHalfMv[best_mode][1] / 2) * HS, HS };
        for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) dp[y * 16 + x] = (uint8_t)get_hp(&s, x, y);
      } else {
        qp_src s = { qc, HS, kQuarterMv[q_winner][0], kQuarterMv[q_winner][1], axis_diag };
        for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) dp[y * 16 + x] = (uint8_t)get_qp(&s, x, y);
      }
    }
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * f3  Motion-compensated prediction: MotionCompensation::xGetMbPredData (limitComponents, COM/MotionCompensation.cpp:1390),
 *     QuarterPelFilter::predBlk and its helpers (COM/QuarterPelFilter.cpp:266-619), xPredChromaPel (COM/MotionCompensation.cpp:206-276),
 *     SampleWeighting::weightLumaSamples / weightChromaSamples (COM/SampleWeighting.cpp:159-265) with xMixB / xMixBWeight / xWeight
 *     (:382-461).  Luma is filtered from the FULL-pel plane as predBlk does (the product gathers from the half-pel plane).
 * ------------------------------------------------------------------------------------------- */
static int tap6(int a, int b, int c2, int d, int e, int f) { return a - 5 * b + 20 * c2 + 20 * d - 5 * e + f; }
static int h6(const uint8_t* p) { return tap6(p[-2], p[-1], p[0], p[1], p[2], p[3]); }
static int v6(const uint8_t* p, int S) { return tap6(p[-2 * S], p[-S], p[0], p[S], p[2 * S], p[3 * S]); }

static int pred_luma_sample(const uint8_t* p /* integer sample at floor(mv) */, int S, int dx, int dy) {
  if (dy == 0) {
    if (dx == 0) return p[0];
    const int b = clip255((h6(p) + 16) / 32);
    return dx == 2 ? b : (b + p[dx >> 1] + 1) / 2;                       /* xPredDy0Dx2 / xPredDy0Dx13 */
  }
  if (dx == 0) {
    const int hh = clip255((v6(p, S) + 16) / 32);
    return dy == 2 ? hh : (hh + p[(dy >> 1) * S] + 1) / 2;               /* xPredDx0Dy2 / xPredDx0Dy13 */
  }
  if (dx == 2) {                                                       /* xPredDx2Dy2 / xPredDx2Dy13: rows of unrounded H-6-taps */
    int t[6];
    for (int n = 0; n < 6; ++n) t[n] = h6(p + (n - 2) * S);
    const int jj = clip255((tap6(t[0], t[1], t[2], t[3], t[4], t[5]) + 512) / 1024);
    if (dy == 2) return jj;
    return (jj + clip255((t[dy == 1 ? 2 : 3] + 16) / 32) + 1) / 2;
  }
  if (dy == 2) {                                                       /* xPredDy2Dx13: columns of unrounded V-6-taps */
    int t[6];
    for (int n = 0; n < 6; ++n) t[n] = v6(p + n - 2, S);
    const int jj = clip255((tap6(t[0], t[1], t[2], t[3], t[4], t[5]) + 512) / 1024);
    return (jj + clip255((t[dx == 1 ? 2 : 3] + 16) / 32) + 1) / 2;
  }
  /* xPredElse: H-half sample of row (dy==1 ? 0 : 1), V-half sample of column (dx==1 ? 0 : 1) */
  const int bx = clip255((h6(p + (dy == 1 ? 0 : S)) + 16) / 32);
  const int hy = clip255((v6(p + (dx == 1 ? 0 : 1), S) + 16) / 32);
  return (bx + hy + 1) >> 1;
}

static int pred_chroma_sample(const uint8_t* p, int S, int xF1, int yF1) {        /* xPredChromaPel, all four branches */
  if (xF1 == 0 && yF1 == 0) return p[0];
  if (xF1 == 0) return ((8 - yF1) * p[0] + yF1 * p[S] + 4) >> 3;
  if (yF1 == 0) return ((8 - xF1) * p[0] + xF1 * p[1] + 4) >> 3;
  const int xF0 = 8 - xF1, yF0 = 8 - yF1;
  return (xF0 * (yF0 * p[0] + yF1 * p[S]) + xF1 * (yF0 * p[1] + yF1 * p[S + 1]) + 0x20) >> 6;
}

static int weight_samples(const jsvm_mc_job* j, int comp, int a, int b, int has0, int has1) {
  const int d = j->log_denom[comp ? 1 : 0];
  if (has0 && has1) {
    if (!(j->weighting & JSVM_MC_WEIGHT_BI)) return (a + b + 1) >> 1;                              /* xMixB */
    const int off = (j->offset[0][comp] + j->offset[1][comp] + 1) >> 1;                           /* xMixBWeight */
    return clip255(((j->weight[0][comp] * a + j->weight[1][comp] * b + (1 << d)) >> (d + 1)) + off);
  }
  const int l = has0 ? 0 : 1, v = has0 ? a : b;
  if ((j->weighting & JSVM_MC_WEIGHT_UNI) && j->weight_flag[comp ? 1 : 0])                         /* xWeight */
    return clip255((j->weight[l][comp] * v + (((1 + j->offset[l][comp] * 2) << d) >> 1)) >> d);
  return v;
}

int jsvm_ora_compensate(jsvm_ora_ctx* c, const jsvm_mc_job* jobs, int n_jobs, uint8_t* pred_y, uint8_t* pred_cb, uint8_t* pred_cr) {
  if (!c) return -1;
  const int W = c->width, H = c->height, S = W + 2 * MARGIN, CS = W / 2 + 2 * CMARGIN;
  const int mbw = W >> 4, mbh = H >> 4;
  for (int i = 0; i < n_jobs; ++i) {
    const jsvm_mc_job* j = &jobs[i];
    if (j->mb_x < 0 || j->mb_x >= mbw || j->mb_y < 0 || j->mb_y >= mbh || j->blk > 15) return -1;
    const int16_t min_h = sat16(((-(int)j->mb_x << 4) - 24) << 2), min_v = sat16(((-(int)j->mb_y << 4) - 24) << 2);
    const int16_t max_h = sat16((((mbw - j->mb_x) << 4) + 8) << 2), max_v = sat16((((mbh - j->mb_y) << 4) + 8) << 2);
    const int bx0 = 16 * j->mb_x + 4 * (j->blk & 3), by0 = 16 * j->mb_y + 4 * (j->blk >> 2);
    int has[2], mvx[2], mvy[2];
    for (int l = 0; l < 2; ++l) {
      has[l] = j->ref_slot[l] >= 0;
      if (has[l] && (j->ref_slot[l] >= c->n_slots || !c->planes[j->ref_slot[l]].valid)) return -1;
      mvx[l] = imin(max_h, imax(min_h, j->mv[l].hor));
      mvy[l] = imin(max_v, imax(min_v, j->mv[l].ver));
    }
    if (!has[0] && !has[1]) return -1;
    uint8_t* dy = pred_y + (size_t)i * 256;
    memset(dy, 0, 256);
    for (int y = 0; y < j->size_y; ++y)
      for (int x = 0; x < j->size_x; ++x) {
        int v[2] = { 0, 0 };
        for (int l = 0; l < 2; ++l)
          if (has[l]) {
            const uint8_t* p = c->planes[j->ref_slot[l]].full + (size_t)(MARGIN + by0 + y + (mvy[l] >> 2)) * S + MARGIN + bx0 + x + (mvx[l] >> 2);
            v[l] = pred_luma_sample(p, S, mvx[l] & 3, mvy[l] & 3);
          }
        dy[y * 16 + x] = (uint8_t)weight_samples(j, 0, v[0], v[1], has[0], has[1]);
      }
    if (pred_cb && pred_cr) {
      uint8_t* d[2] = { pred_cb + (size_t)i * 64, pred_cr + (size_t)i * 64 };
      for (int k = 0; k < 2; ++k) {
        memset(d[k], 0, 64);
        for (int y = 0; y < j->size_y / 2; ++y)
          for (int x = 0; x < j->size_x / 2; ++x) {
            int v[2] = { 0, 0 };
            for (int l = 0; l < 2; ++l)
              if (has[l]) {
                const uint8_t* p = c->planes[j->ref_slot[l]].chroma[k] + (size_t)(CMARGIN + by0 / 2 + y + (mvy[l] >> 3)) * CS + CMARGIN + bx0 / 2 + x + (mvx[l] >> 3);
                v[l] = pred_chroma_sample(p, CS, mvx[l] & 7, mvy[l] & 7);
              }
            d[k][y * 8 + x] = (uint8_t)weight_samples(j, 1 + k, v[0], v[1], has[0], has[1]);
          }
      }
    }
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * f1  Picture-level motion search: MV prediction (MbDataAccess::xGetMvPredictor, COM/MbDataAccess.cpp:137-187; neighbour lookup
 *     INCC/MbDataAccess.h:1193-1348) + the list-0 loops of MbEncoder::xEstimateMb16x16 / 16x8 / 8x16 / 8x8
 *     (ENC/MbEncoder.cpp:4449-4560, 4830-4935, 5234-5330, 5637-5700, 5950-6040), macroblocks in raster order, the macroblock mode
 *     by the smallest sum of motion costs.  Sequential, as the encoder runs it.
 * ------------------------------------------------------------------------------------------- */
typedef struct { int hor, ver, ref; } mv3d;                 /* ref 0 = BLOCK_NOT_AVAILABLE (INCC/Mv.h:131-135) */

static mv3d field_block(const jsvm_pic_mb* f, int mbw, int mbh, int mx, int my, int b4) {
  mv3d r = { 0, 0, 0 };
  if (mx < 0 || my < 0 || mx >= mbw || my >= mbh) return r;          /* m_rcMbUnavailable */
  const jsvm_pic_mb* m = &f[my * mbw + mx];
  r.hor = m->mv[b4].hor; r.ver = m->mv[b4].ver; r.ref = m->ref_idx[((b4 >> 3) << 1) | ((b4 & 3) >> 1)];
  return r;
}
static int med3(int a, int b, int c) { return a > b ? (a > c ? (b > c ? b : c) : a) : (b > c ? (a > c ? a : c) : b); }

/* type: 0 MEDIAN, 1 PRED_A, 2 PRED_B, 3 PRED_C */
static jsvm_mv mv_predictor(mv3d A, mv3d B, mv3d C, mv3d D, int ref, int type) {
  jsvm_mv m;
  if (C.ref == 0) C = D;
  if ((type == 1 && A.ref == ref) || (A.ref == ref && B.ref != ref && C.ref != ref) || (B.ref == 0 && C.ref == 0)) { m.hor = (int16_t)A.hor; m.ver = (int16_t)A.ver; return m; }
  if ((type == 2 && B.ref == ref) || (A.ref != ref && B.ref == ref && C.ref != ref)) { m.hor = (int16_t)B.hor; m.ver = (int16_t)B.ver; return m; }
  if ((type == 3 && C.ref == ref) || (A.ref != ref && B.ref != ref && C.ref == ref)) { m.hor = (int16_t)C.hor; m.ver = (int16_t)C.ver; return m; }
  m.hor = (int16_t)med3(A.hor, B.hor, C.hor); m.ver = (int16_t)med3(A.ver, B.ver, C.ver);
  return m;
}

static uint32_t ref_idx_bits(int r, int n) { return n == 1 ? 0u : n == 2 ? 1u : (r < 2 ? 1u : r < 4 ? 3u : r < 8 ? 5u : 7u); }

int jsvm_ora_estimate_pictures(jsvm_ora_ctx* c, const jsvm_pic_task* tasks, int n_pics, jsvm_pic_mb* out) {
  if (!c) return -1;
  const int mbw = c->width >> 4, mbh = c->height >> 4;
  static const uint8_t part_mode[9] = { JSVM_MODE_16x16, JSVM_MODE_16x8, JSVM_MODE_16x8, JSVM_MODE_8x16, JSVM_MODE_8x16, JSVM_BLK_8x8, JSVM_BLK_8x8, JSVM_BLK_8x8, JSVM_BLK_8x8 };
  static const uint8_t part_blk[9]  = { 0, 0, 8, 0, 2, 0, 2, 8, 10 };
  static const uint8_t part_bits[9] = { 1, 3, 3, 3, 3, 6, 1, 1, 1 };
  for (int p = 0; p < n_pics; ++p) {
    const jsvm_pic_task* t = &tasks[p];
    const int R = t->n_refs;
    jsvm_pic_mb* f = out + (size_t)p * mbw * mbh;
    memset(f, 0, sizeof(jsvm_pic_mb) * (size_t)mbw * mbh);
    for (int my = 0; my < mbh; ++my)
      for (int mx = 0; mx < mbw; ++mx) {
        jsvm_pic_mb* o = &f[my * mbw + mx];
        mv3d cur[9];                                       /* partitions of this macroblock estimated so far */
        for (int part = 0; part < 9; ++part) {
          /* neighbours of the partition's corner blocks: A left of the first 4x4, B above it, C above-right of the LAST 4x4 of
           * the top row, D above-left of the first (used when C is not available) */
          mv3d A, B, C, D = { 0, 0, 0 };
          int type = 0;
          switch (part) {
            case 0: case 1: case 3: case 5:
              A = field_block(f, mbw, mbh, mx - 1, my, 3); B = field_block(f, mbw, mbh, mx, my - 1, 12);
              C = part <= 1 ? field_block(f, mbw, mbh, mx + 1, my - 1, 12) : field_block(f, mbw, mbh, mx, my - 1, 14);
              D = field_block(f, mbw, mbh, mx - 1, my - 1, 15);
              type = part == 1 ? 2 : part == 3 ? 1 : 0;
              break;
            case 2: A = field_block(f, mbw, mbh, mx - 1, my, 11); B = cur[1]; C.hor = C.ver = C.ref = 0; D = field_block(f, mbw, mbh, mx - 1, my, 7); type = 1; break;
            case 4: case 6:
              A = cur[part == 4 ? 3 : 5]; B = field_block(f, mbw, mbh, mx, my - 1, 14);
              C = field_block(f, mbw, mbh, mx + 1, my - 1, 12); D = field_block(f, mbw, mbh, mx, my - 1, 13);
              type = part == 4 ? 3 : 0;
              break;
            case 7: A = field_block(f, mbw, mbh, mx - 1, my, 11); B = cur[5]; C = cur[6]; break;
            default: A = cur[7]; B = cur[6]; C.hor = C.ver = C.ref = 0; D = cur[5]; break;
          }
          uint32_t best = 0xffffffffu; int best_r = 0; jsvm_mv best_mv = { 0, 0 };
          for (int r = 1; r <= R; ++r) {
            jsvm_me_job j; jsvm_me_result res;
            memset(&j, 0, sizeof(j));
            j.cur_slot = t->cur_slot; j.ref_slot = t->ref_slot[r - 1]; j.mb_x = (int16_t)mx; j.mb_y = (int16_t)my;
            j.blk = part_blk[part]; j.mode = part_mode[part];
            j.pred = mv_predictor(A, B, C, D, r, type); j.start = j.pred;
            j.cost_factor = t->cost_factor; j.bits_in = part_bits[part] + ref_idx_bits(r, R); j.org_index = -1;
            if (jsvm_ora_search(c, &j, 1, NULL, 0, &res, NULL)) return -1;
            if (res.cost < best) { best = res.cost; best_r = r; best_mv = res.mv; }
          }
          cur[part].hor = best_mv.hor; cur[part].ver = best_mv.ver; cur[part].ref = best_r;
          o->part_mv[part] = best_mv; o->part_ref[part] = (int8_t)best_r; o->part_cost[part] = best;
        }
        const uint32_t cost[4] = { o->part_cost[0], o->part_cost[1] + o->part_cost[2], o->part_cost[3] + o->part_cost[4],
                                   o->part_cost[5] + o->part_cost[6] + o->part_cost[7] + o->part_cost[8] };
        int mode = 1; uint32_t best = cost[0];
        for (int m = 1; m < 4; ++m) if (cost[m] < best) { best = cost[m]; mode = m + 1; }
        o->mode = (uint8_t)mode; o->cost = best;
        for (int b8 = 0; b8 < 4; ++b8) {
          const int part = mode == 1 ? 0 : mode == 2 ? 1 + (b8 >> 1) : mode == 3 ? 3 + (b8 & 1) : 5 + b8;
          o->ref_idx[b8] = o->part_ref[part];
          for (int k = 0; k < 4; ++k) o->mv[((b8 >> 1) * 2 + (k >> 1)) * 4 + (b8 & 1) * 2 + (k & 1)] = o->part_mv[part];
        }
      }
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * f4  Residual upsampling for progressive layers: DownConvert::xCompResidualUpsampling, xDetermineTransBlkIdcs,
 *     xBasicResidualUpsampling (COM/DownConvert.cpp:2733-2817, 2604-2650, 2867-2958).  Two full passes over an int image buffer,
 *     as the reference does it (the product fuses them per output sample).
 * ------------------------------------------------------------------------------------------- */
static int ceil_log2(int i) { int s = 0; --i; while (i > 0) { ++s; i >>= 1; } return s; }
static int xclip(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void upsample_residual_plane(const jsvm_resize_params* rp, int chroma, const int16_t* src, const uint8_t* t8x8, int16_t* dst) {
  const int f = chroma ? 2 : 1;
  const int ref_phase_x = chroma ? rp->ref_chroma_phase_x : 0, ref_phase_y = chroma ? rp->ref_chroma_phase_y : 0;
  const int phase_x = chroma ? rp->chroma_phase_x : 0, phase_y = chroma ? rp->chroma_phase_y : 0;
  const int base_w = rp->ref_frame_width / f, base_h = rp->ref_frame_height / f;
  const int out_w = rp->scaled_ref_width / f, out_h = rp->scaled_ref_height / f;
  const int cur_w = rp->frame_width / f, cur_h = rp->frame_height / f;
  const int l_off = rp->left_offset / f, t_off = rp->top_offset / f, r_off = cur_w - l_off - out_w, b_off = cur_h - t_off - out_h;
  const int shift_x = rp->level_idc <= 30 ? 16 : 31 - ceil_log2(base_w), shift_y = rp->level_idc <= 30 ? 16 : 31 - ceil_log2(base_h);
  const int scale_x = (int)((((unsigned)base_w << shift_x) + (unsigned)(out_w >> 1)) / (unsigned)out_w);
  const int scale_y = (int)((((unsigned)base_h << shift_y) + (unsigned)(out_h >> 1)) / (unsigned)out_h);
  const int add_x = (((base_w * (2 + phase_x)) << (shift_x - 2)) + (out_w >> 1)) / out_w + (1 << (shift_x - 5));
  const int add_y = (((base_h * (2 + phase_y)) << (shift_y - 2)) + (out_h >> 1)) / out_h + (1 << (shift_y - 5));
  const int delta_x = 4 * (2 + ref_phase_x), delta_y = 4 * (2 + ref_phase_y);
  const int mbs = chroma ? 8 : 16, mbw = rp->ref_frame_width / 16;
  const int stride = cur_w > base_w ? cur_w : base_w, rows = cur_h > base_h ? cur_h : base_h;
  int* img = (int*)calloc((size_t)stride * rows, sizeof(int));
  int* idc = (int*)calloc((size_t)stride * rows, sizeof(int));
  int* tmp = (int*)calloc((size_t)(stride > rows ? stride : rows), sizeof(int));
  for (int y = 0; y < base_h; ++y)
    for (int x = 0; x < base_w; ++x) {
      img[y * stride + x] = src[(size_t)y * base_w + x];
      const int mb = (y / mbs) * mbw + x / mbs, ix = x % mbs, iy = y % mbs;
      idc[y * stride + x] = (!chroma && t8x8[mb]) ? ((((iy >> 3) << 1) + (ix >> 3) + (mb << 2)) << 1) + 1 : (((iy >> 2) << 2) + (ix >> 2) + (mb << 4)) << 1;
    }
  for (int j = 0; j < base_h; ++j) {                        /* horizontal pass */
    int* ps = img + j * stride;
    for (int i = 0; i < cur_w; ++i) {
      if (i < l_off || i >= cur_w - r_off) { tmp[i] = 0; continue; }
      const int p16 = (int)((unsigned)((i - l_off) * scale_x + add_x) >> (shift_x - 4)) - delta_x;
      const int ph = p16 & 15, x0 = xclip(p16 >> 4, 0, base_w - 1), x1 = xclip((p16 >> 4) + 1, 0, base_w - 1);
      tmp[i] = idc[j * stride + x0] == idc[j * stride + x1] ? (16 - ph) * ps[x0] + ph * ps[x1] : ps[ph < 8 ? x0 : x1] << 4;
    }
    memcpy(ps, tmp, (size_t)cur_w * sizeof(int));
  }
  for (int i = 0; i < cur_w; ++i) {                         /* vertical pass */
    const int p16x = (int)((unsigned)((i - l_off) * scale_x + add_x) >> (shift_x - 4)) - delta_x;
    const int xr = xclip((p16x >> 4) + ((p16x & 15) >> 3), 0, base_w - 1);
    for (int j = 0; j < cur_h; ++j) {
      if (i < l_off || i >= cur_w - r_off || j < t_off || j >= cur_h - b_off) { tmp[j] = 0; continue; }
      const int pre = (j - t_off) * scale_y + add_y;
      const int post = j >= t_off ? (int)((unsigned)pre >> (shift_y - 4)) : (pre >> (shift_y - 4));
      const int p16 = post - delta_y, ph = p16 & 15, y0 = xclip(p16 >> 4, 0, base_h - 1), y1 = xclip((p16 >> 4) + 1, 0, base_h - 1);
      tmp[j] = idc[y0 * stride + xr] == idc[y1 * stride + xr] ? ((16 - ph) * img[y0 * stride + i] + ph * img[y1 * stride + i] + 128) >> 8
                                                              : (img[(ph < 8 ? y0 : y1) * stride + i] + 8) >> 4;
    }
    for (int j = 0; j < cur_h; ++j) dst[(size_t)j * cur_w + i] = (int16_t)tmp[j];
  }
  free(img); free(idc); free(tmp);
}

int jsvm_ora_upsample_residual(const jsvm_resize_params* rp, const int16_t* base_y, const int16_t* base_cb, const int16_t* base_cr,
                               const uint8_t* base_transform8x8, int16_t* out_y, int16_t* out_cb, int16_t* out_cr) {
  if (!rp || !base_y || !out_y || !base_transform8x8) return -1;
  upsample_residual_plane(rp, 0, base_y, base_transform8x8, out_y);
  if (base_cb && base_cr && out_cb && out_cr) {
    upsample_residual_plane(rp, 1, base_cb, base_transform8x8, out_cb);
    upsample_residual_plane(rp, 1, base_cr, base_transform8x8, out_cr);
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * f4  Inter-layer motion prediction for progressive layers: MbDataCtrl::upsampleMotion -> MotionUpsampling::resample
 *     (COM/MbDataCtrl.cpp:1115-1140, 122-150) with xInitMb :152-195, xIsInCropWindow :197-208, xGetRefLayerMb :211-288 (G.6.1),
 *     xGetRefLayerPartIdc :290-365 (G.6.2), xSetPartIdcArray :368-473 (G.8.6.1.1), xGetInitialBaseRefIdxAndMv :476-548,
 *     xGetRefIdxAndInitialMvPred :561-625, xDeriveBlockModeAndUpdateMv :636-723 (G.8.6.1.2), xDeriveMbMode :737-759,
 *     xDeriveFwdBwd :762-773, xSetInterIntraIdc :776-798, xSetResPredSafeFlag :801-859, xSetPredMbData :862-945;
 *     PosCalcParam / MvScaleParam :14-84.  Frame macroblocks only, no cropping change per picture, no coefficient prediction.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const jsvm_motion_upsample_params* p; const jsvm_mb_motion* base;
  int rs_change, ref_w, ref_h, shift_x, shift_y, scale_x, scale_y, add_x, add_y;       /* PosCalcParam */
  int mv_scale_x, mv_scale_y;                                                          /* MvScaleParam */
  int max_list;
  int mbx, mby;
} mu_ctx;

static int mu_spatial_change(const jsvm_resize_params* r) {          /* getSpatialResolutionChangeFlag, progressive */
  if (r->left_offset % 16 || r->top_offset % 16) return 1;
  if (r->ref_frame_width != r->scaled_ref_width || r->ref_frame_height != r->scaled_ref_height) return 1;
  if (r->ref_chroma_phase_x != r->chroma_phase_x || r->ref_chroma_phase_y != r->chroma_phase_y) return 1;
  return 0;
}
static int mu_restricted_change(const jsvm_resize_params* r) {       /* getRestrictedSpatialResolutionChangeFlag */
  if (!mu_spatial_change(r)) return 1;
  if (r->left_offset % 16 || r->top_offset % 16) return 0;
  if (!(r->ref_frame_height == r->scaled_ref_height || 2 * r->ref_frame_height == r->scaled_ref_height)) return 0;
  if (!(r->ref_frame_width == r->scaled_ref_width || 2 * r->ref_frame_width == r->scaled_ref_width)) return 0;
  return 1;
}

/* G.6.1 + G.6.2: partition indication ( base macroblock index << 4 | 4x4 block of the partition's first block ), -1 = intra */
static int mu_part_idc(const mu_ctx* m, int x_in_mb, int y_in_mb) {
  const int cx = (m->mbx << 4) + x_in_mb, cy = (m->mby << 4) + y_in_mb;
  int bx = (int)((unsigned)(cx * m->scale_x + m->add_x) >> m->shift_x), by = (int)((unsigned)(cy * m->scale_y + m->add_y) >> m->shift_y);
  bx = imin(bx, m->ref_w - 1); by = imin(by, m->ref_h - 1);
  const int mb_idx = (by >> 4) * (m->ref_w >> 4) + (bx >> 4), xi = bx & 15, yi = by & 15;
  const jsvm_mb_motion* b = &m->base[mb_idx];
  if (b->mb_mode >= 6) return -1;
  const int b8 = ((yi >> 3) << 1) + (xi >> 3), b4 = (((yi & 7) >> 2) << 1) + ((xi & 7) >> 2);
  static const uint8_t nxn[6][4] = { {0,0,0,0}, {0,0,0,0}, {0,0,2,2}, {0,1,0,1}, {0,1,2,3}, {0,1,2,3} };
  int part8 = 0, part4 = 0;
  if (b->mb_mode == 0 && b->slice_type == 1) {                          /* B_Skip / B_Direct_16x16 */
    if (m->p->ref_layer_dq_id == 0) { part8 = b8; part4 = b4; }
  } else {
    part8 = nxn[b->mb_mode][b8];
    if (b->mb_mode == 4 || b->mb_mode == 5) {
      const int bm = b->blk_mode[b8];
      if (bm == 0) { if (m->p->ref_layer_dq_id == 0) part4 = b4; }
      else part4 = nxn[bm - 8 + 1][b4];
    }
  }
  const int x4 = ((part8 & 1) << 1) + (part4 & 1), y4 = ((part8 >> 1) << 1) + (part4 >> 1);
  return (mb_idx << 4) + (y4 << 2) + x4;
}

static int mu_min_ref(int a, int b) { return a < 1 ? b : (b < 1 ? a : imin(a, b)); }
static int mu_mv_diff(jsvm_mv a, jsvm_mv b) { return abs(a.hor - b.hor) + abs(a.ver - b.ver); }
static jsvm_mv mu_avg(const jsvm_mv* v, int n) {          /* ( a + b [+ c + d] + Mv( n/2, n/2 ) ) >> log2 n on Shorts */
  int h = n / 2, w = n / 2;
  for (int i = 0; i < n; ++i) { h = (int16_t)(h + v[i].hor); w = (int16_t)(w + v[i].ver); }
  jsvm_mv r; r.hor = (int16_t)(h >> (n == 4 ? 2 : 1)); r.ver = (int16_t)(w >> (n == 4 ? 2 : 1));
  return r;
}

static void mu_resample_mb(mu_ctx* m, jsvm_mb_motion* o) {
  const jsvm_resize_params* r = &m->p->resize;
  memset(o, 0, sizeof(*o));
  for (int k = 0; k < 4; ++k) o->blk_mode[k] = 8;
  memset(o->ref_idx, -1, sizeof(o->ref_idx));
  /* xIsInCropWindow */
  const int x0 = (r->left_offset + 15) / 16, y0 = (r->top_offset + 15) / 16;
  const int x1 = (r->left_offset + r->scaled_ref_width) / 16, y1 = (r->top_offset + r->scaled_ref_height) / 16;
  const int in_crop = m->mbx >= x0 && m->mbx < x1 && m->mby >= y0 && m->mby < y1;
  o->in_crop_window = (uint8_t)in_crop;
  o->res_pred_safe = (uint8_t)in_crop;
  if (!in_crop) return;
  int idc[4][4];                                                /* [x][y] */
  int intra_bl = 1;
  for (int y = 0; y < 4; ++y)
    for (int x = 0; x < 4; ++x) { idc[x][y] = mu_part_idc(m, (x << 2) + 1, (y << 2) + 1); intra_bl = intra_bl && idc[x][y] == -1; }
  if (intra_bl) { o->mb_mode = 36; }
  else {
    if (!m->rs_change) {
      for (int yp = 0; yp < 2; ++yp)                            /* replace -1 on a 4x4 basis */
        for (int xp = 0; xp < 2; ++xp) {
          int proc[2][2] = { {0, 0}, {0, 0} };
          for (int ys = 0; ys < 2; ++ys)
            for (int xs = 0; xs < 2; ++xs) {
              const int yc = (yp << 1) + ys, xc = (xp << 1) + xs;
              if (idc[xc][yc] != -1) continue;
              const int yi = 1 - ys, xi = 1 - xs, yci = (yp << 1) + yi, xci = (xp << 1) + xi;
              proc[xs][ys] = 1;
              if (!proc[xi][ys] && idc[xci][yc] != -1) idc[xc][yc] = idc[xci][yc];
              else if (!proc[xs][yi] && idc[xc][yci] != -1) idc[xc][yc] = idc[xc][yci];
              else if (!proc[xi][yi] && idc[xci][yci] != -1) idc[xc][yc] = idc[xci][yci];
            }
        }
      int proc8[2][2] = { {0, 0}, {0, 0} };                     /* replace -1 on an 8x8 basis */
      for (int yp = 0; yp < 2; ++yp)
        for (int xp = 0; xp < 2; ++xp) {
          const int ypi = 1 - yp, xpi = 1 - xp, yo = yp << 1, xo = xp << 1, yoi = 2 - yp, xoi = 2 - xp;
          if (idc[xo][yo] != -1) continue;
          proc8[xp][yp] = 1;
          if (!proc8[xpi][yp] && idc[xoi][yo] != -1) {
            idc[xo][yo] = idc[xoi][yo]; idc[xo + 1][yo] = idc[xoi][yo]; idc[xo][yo + 1] = idc[xoi][yo + 1]; idc[xo + 1][yo + 1] = idc[xoi][yo + 1];
          } else if (!proc8[xp][ypi] && idc[xo][yoi] != -1) {
            idc[xo][yo] = idc[xo][yoi]; idc[xo + 1][yo] = idc[xo + 1][yoi]; idc[xo][yo + 1] = idc[xo][yoi]; idc[xo + 1][yo + 1] = idc[xo + 1][yoi];
          } else if (!proc8[xpi][ypi] && idc[xoi][yoi] != -1) {
            idc[xo][yo] = idc[xo + 1][yo] = idc[xo][yo + 1] = idc[xo + 1][yo + 1] = idc[xoi][yoi];
          }
        }
    }
    int ref8[2][2][2];                                          /* [list][x][y] */
    jsvm_mv mv[2][4][4];                                        /* [list][x][y] */
    memset(mv, 0, sizeof(mv));
    for (int l = 0; l < 2; ++l) for (int x = 0; x < 2; ++x) for (int y = 0; y < 2; ++y) ref8[l][x][y] = -1;
    for (int l = 0; l < m->max_list; ++l) {
      int rtmp[4][4];
      for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 4; ++x) {
          const jsvm_mb_motion* b = &m->base[idc[x][y] >> 4];
          const int b4 = idc[x][y] & 15, rb = b->ref_idx[l][((b4 >> 3) << 1) | ((b4 & 3) >> 1)];
          if (rb < 1) { rtmp[x][y] = -1; mv[l][x][y].hor = mv[l][x][y].ver = 0; continue; }
          rtmp[x][y] = rb;
          mv[l][x][y].hor = (int16_t)((b->mv[l][b4].hor * m->mv_scale_x + 32768) >> 16);
          mv[l][x][y].ver = (int16_t)((b->mv[l][b4].ver * m->mv_scale_y + 32768) >> 16);
        }
      ref8[l][0][0] = rtmp[0][0]; ref8[l][0][1] = rtmp[0][2]; ref8[l][1][0] = rtmp[2][0]; ref8[l][1][1] = rtmp[2][2];
      if (m->rs_change) continue;
      for (int y8 = 0; y8 < 2; ++y8)
        for (int x8 = 0; x8 < 2; ++x8) {
          for (int y4 = 0; y4 < 2; ++y4)
            for (int x4 = 0; x4 < 2; ++x4) ref8[l][x8][y8] = mu_min_ref(ref8[l][x8][y8], rtmp[(x8 << 1) + x4][(y8 << 1) + y4]);
          for (int ys = 0; ys < 2; ++ys)
            for (int xs = 0; xs < 2; ++xs) {
              const int y = (y8 << 1) + ys, x = (x8 << 1) + xs;
              if (ref8[l][x8][y8] == rtmp[x][y]) continue;
              const int yi = (y8 << 1) + 1 - ys, xi = (x8 << 1) + 1 - xs;
              if (ref8[l][x8][y8] == rtmp[xi][y]) mv[l][x][y] = mv[l][xi][y];
              else if (ref8[l][x8][y8] == rtmp[x][yi]) mv[l][x][y] = mv[l][x][yi];
              else mv[l][x][y] = mv[l][xi][yi];
            }
        }
    }
    for (int b8 = 0; b8 < 4; ++b8) {                            /* xDeriveBlockModeAndUpdateMv */
      const int thr = m->rs_change ? 0 : 1, xo = (b8 & 1) << 1, yo = (b8 >> 1) << 1;
      int hor = 1, ver = 1, all = 1;
      if (m->p->direct_8x8_inference && !m->rs_change && m->p->slice_type == 1) {
        const int xc = (xo >> 1) * 3, yc = (yo >> 1) * 3;
        for (int l = 0; l < m->max_list; ++l) {
          const jsvm_mv t = mv[l][xc][yc];
          mv[l][xo][yo] = mv[l][xo + 1][yo] = mv[l][xo][yo + 1] = mv[l][xo + 1][yo + 1] = t;
        }
      }
      for (int l = 0; l < m->max_list; ++l) {
        const int h1 = mu_mv_diff(mv[l][xo][yo], mv[l][xo + 1][yo]) <= thr, h2 = mu_mv_diff(mv[l][xo][yo + 1], mv[l][xo + 1][yo + 1]) <= thr;
        const int v1 = mu_mv_diff(mv[l][xo][yo], mv[l][xo][yo + 1]) <= thr, v2 = mu_mv_diff(mv[l][xo + 1][yo], mv[l][xo + 1][yo + 1]) <= thr;
        const int dg = mu_mv_diff(mv[l][xo][yo], mv[l][xo + 1][yo + 1]) <= thr;
        all = all && h1 && v1 && dg; hor = hor && h1 && h2; ver = ver && v1 && v2;
      }
      const int bm = all ? 8 : hor ? 9 : ver ? 10 : 11;
      o->blk_mode[b8] = (uint8_t)bm;
      if (m->rs_change || bm == 11) continue;
      for (int l = 0; l < m->max_list; ++l) {
        if (bm == 8) {
          jsvm_mv v[4] = { mv[l][xo][yo], mv[l][xo + 1][yo], mv[l][xo][yo + 1], mv[l][xo + 1][yo + 1] };
          mv[l][xo][yo] = mv[l][xo + 1][yo] = mv[l][xo][yo + 1] = mv[l][xo + 1][yo + 1] = mu_avg(v, 4);
        } else if (bm == 9) {
          jsvm_mv a[2] = { mv[l][xo][yo], mv[l][xo + 1][yo] }, b[2] = { mv[l][xo][yo + 1], mv[l][xo + 1][yo + 1] };
          mv[l][xo][yo] = mv[l][xo + 1][yo] = mu_avg(a, 2); mv[l][xo][yo + 1] = mv[l][xo + 1][yo + 1] = mu_avg(b, 2);
        } else {
          jsvm_mv a[2] = { mv[l][xo][yo], mv[l][xo][yo + 1] }, b[2] = { mv[l][xo + 1][yo], mv[l][xo + 1][yo + 1] };
          mv[l][xo][yo] = mv[l][xo][yo + 1] = mu_avg(a, 2); mv[l][xo + 1][yo] = mv[l][xo + 1][yo + 1] = mu_avg(b, 2);
        }
      }
    }
    int mode = 4;                                               /* xDeriveMbMode */
    if (o->blk_mode[0] == 8 && o->blk_mode[1] == 8 && o->blk_mode[2] == 8 && o->blk_mode[3] == 8) {
      int hm = 1, vm = 1;
#define SAME8(l, a, b) (ref8[l][(a) & 1][(a) >> 1] == ref8[l][(b) & 1][(b) >> 1] && \
                        mv[l][((a) & 1) << 1][((a) >> 1) << 1].hor == mv[l][((b) & 1) << 1][((b) >> 1) << 1].hor && \
                        mv[l][((a) & 1) << 1][((a) >> 1) << 1].ver == mv[l][((b) & 1) << 1][((b) >> 1) << 1].ver)
      for (int l = 0; l < m->max_list; ++l) { hm = hm && SAME8(l, 0, 1) && SAME8(l, 2, 3); vm = vm && SAME8(l, 0, 2) && SAME8(l, 1, 3); }
#undef SAME8
      static const uint8_t modes[4] = { 4, 2, 3, 1 };
      mode = modes[(vm ? 2 : 0) + (hm ? 1 : 0)];
    }
    o->mb_mode = (uint8_t)mode;
    unsigned fb = 0;                                            /* xDeriveFwdBwd */
    for (int n = 3; n >= 0; --n) { fb <<= 4; fb += ref8[0][n & 1][n >> 1] > 0 ? 1 : 0; fb += ref8[1][n & 1][n >> 1] > 0 ? 2 : 0; }
    o->fwd_bwd = (uint16_t)fb;
    if (!m->rs_change)                                          /* xSetInterIntraIdc */
      for (int y = 0; y < 2; ++y) for (int x = 0; x < 2; ++x) if (mu_part_idc(m, x * 15, y * 15) == -1) o->base_intra |= (uint8_t)(1 << (x + 2 * y));
    for (int l = 0; l < m->max_list; ++l) {                     /* xSetPredMbData */
      for (int b8 = 0; b8 < 4; ++b8) o->ref_idx[l][b8] = (int8_t)ref8[l][b8 & 1][b8 >> 1];
      for (int b4 = 0; b4 < 16; ++b4) o->mv[l][b4] = mv[l][b4 & 3][b4 >> 2];
    }
  }
  /* xSetResPredSafeFlag */
  if (m->p->check_residual_pred && !m->rs_change) {
    int safe = 1;
    for (int y = 0; y < 4 && safe; ++y)
      for (int x = 0; x < 4 && safe; ++x) {
        int pi[4], same = 1, nointer = 0;
        for (int k = 0; k < 4; ++k) { pi[k] = mu_part_idc(m, (x << 2) + 3 * (k & 1), (y << 2) + 3 * (k >> 1)); same = same && pi[k] == pi[0]; nointer = nointer || pi[k] < 0; }
        if (same || nointer) { safe = nointer; continue; }
        for (int l = 0; l < 2 && safe; ++l) {
          int rf[4]; jsvm_mv v[4];
          for (int k = 0; k < 4; ++k) {
            const jsvm_mb_motion* b = &m->base[pi[k] >> 4]; const int b4 = pi[k] & 15;
            rf[k] = b->ref_idx[l][((b4 >> 3) << 1) | ((b4 & 3) >> 1)]; v[k] = b->mv[l][b4];
          }
          for (int k = 0; k < 4 && safe; ++k) safe = rf[k] == rf[0];
          if (safe) { const jsvm_mv avg = mu_avg(v, 4); for (int k = 0; k < 4 && safe; ++k) safe = mu_mv_diff(v[k], avg) <= m->p->mv_threshold; }
        }
      }
    o->res_pred_safe = (uint8_t)safe;
  }
}

int jsvm_ora_upsample_motion(const jsvm_motion_upsample_params* p, const jsvm_mb_motion* base, jsvm_mb_motion* out) {
  if (!p || !base || !out) return -1;
  const jsvm_resize_params* r = &p->resize;
  mu_ctx m;
  m.p = p; m.base = base;
  m.rs_change = mu_restricted_change(r);
  m.ref_w = r->ref_frame_width; m.ref_h = r->ref_frame_height;
  m.shift_x = r->level_idc <= 30 ? 16 : 31 - ceil_log2(m.ref_w); m.shift_y = r->level_idc <= 30 ? 16 : 31 - ceil_log2(m.ref_h);
  m.scale_x = (int)((((unsigned)m.ref_w << m.shift_x) + (unsigned)(r->scaled_ref_width >> 1)) / (unsigned)r->scaled_ref_width);
  m.scale_y = (int)((((unsigned)m.ref_h << m.shift_y) + (unsigned)(r->scaled_ref_height >> 1)) / (unsigned)r->scaled_ref_height);
  m.add_x = (1 << (m.shift_x - 1)) - r->left_offset * m.scale_x; m.add_y = (1 << (m.shift_y - 1)) - r->top_offset * m.scale_y;
  m.mv_scale_x = (int)((((uint64_t)r->scaled_ref_width << 16) + (uint64_t)(m.ref_w >> 1)) / (uint64_t)m.ref_w);
  m.mv_scale_y = (int)((((uint64_t)r->scaled_ref_height << 16) + (uint64_t)(m.ref_h >> 1)) / (uint64_t)m.ref_h);
  m.max_list = p->slice_type == 1 ? 2 : (p->slice_type == 2 ? 0 : 1);
  const int mbw = r->frame_width >> 4, mbh = r->frame_height >> 4;
  for (m.mby = 0; m.mby < mbh; ++m.mby)
    for (m.mbx = 0; m.mbx < mbw; ++m.mbx) mu_resample_mb(&m, &out[m.mby * mbw + m.mbx]);
  return 0;
}
