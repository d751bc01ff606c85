// jsvm_me_capi.cu -- the C ABI (include/jsvm_me_b200.h) over the sm_100a kernels K1 / K2 / K3.
//
// Host-side state mirrors what the reference keeps in its objects:
//   * the frame store (full-pel + half-pel YuvPicBuffers of each Frame)      -> device plane pools
//   * MotionVectorSearchParams (search range, sub-pel DFunc)                  -> jsvm_me_set_params
//   * the buffers MotionEstimationQuarterPel keeps for compensateBlock       -> last-call preds
// There is no CPU implementation behind any compute entry point: without a usable sm_100 device
// jsvm_me_create fails and everything else refuses to run.
#include <cuda.h>
#include <dlfcn.h>
#include <nccl.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <thread>
#include <functional>
#include <mutex>
#include <condition_variable>
#include <chrono>
#include <string>

#include "../../include/jsvm_me_b200.h"
#include "me_common.cuh"
#include "k1_halfpel.cuh"
#include "k2_fullsearch.cuh"
#include "k2_generic.cuh"
#include "k3_subpel.cuh"
#include "k4_mc.cuh"
#include "k5_fastsearch.cuh"
#include "k6_picture.cuh"
#include "k7_interlayer.cuh"

using namespace jsvm;

namespace {

thread_local char g_err[512] = "";

int fail(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return JSVM_ME_ERR;
}

#define CU_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
  return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

constexpr int kV8 = 8;        // strip height of the 8x8-granular search kernel
constexpr int kV4 = 4;        // strip height of the 4x4-granular search kernel
constexpr int kMaxUH = 176;   // tallest union window one CTA handles (by-table rows, shared memory)
constexpr int kSmallUH = 56;  // union windows up to this height (search range <= ~20) use the 128-thread form of the one-CTA-per-group kernel
constexpr size_t kMaxDynSmemTotal = 228 * 1024 - 4 * 1024;   // shared memory of an SM available to resident CTAs (1 KB reserved per CTA)
constexpr int kMaxSearchRange = 64;   // staged windows of at most 129 x 129 positions per job (packed-byte kernels, k2_generic.cuh)
constexpr int kMaxWideRange = 4096;   // beyond 64 the searches run unstaged, one warp per job (K5); windows stay clipped to the padded picture

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename T>
struct DevBuf {
  T* p = nullptr; size_t cap = 0;
  int reserve(size_t n) {
    if (n <= cap) return JSVM_ME_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = std::max(n, (size_t)1024);
    cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    if (e != cudaSuccess) return fail("cudaMalloc(%zu bytes) failed: %s", want * sizeof(T), cudaGetErrorString(e));
    cap = want;
    return JSVM_ME_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct jsvm_me_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 0;
  // geometry
  int W = 0, H = 0, n_slots = 0;
  int plane_stride = 0, plane_rows = 0; size_t plane_bytes = 0;
  int half_stride = 0, half_rows = 0;   size_t half_bytes = 0;
  uint8_t* d_planes = nullptr;
  uint8_t* d_half = nullptr;
  int* d_slot_ids = nullptr;            // scratch list of slots for batched launches (n_slots entries)
  int* d_dirty_ids = nullptr;           // slots whose margins have not been filled since their last upload (n_slots entries)
  std::vector<int> margin_dirty;        // the same list on the host: uploads only queue their copy, the first reader fills the margins
  std::vector<char> slot_full, slot_half, slot_chroma;
  // chroma planes [slot][Cb, Cr][H/2 + 32][W/2 + 32], origin (16,16); allocated by the first chroma upload
  uint8_t* d_chroma = nullptr;
  int chroma_stride = 0, chroma_rows = 0; size_t chroma_bytes = 0;     // per plane
  // full-pel -> sub-pel cost-factor pairs (jsvm_me_map_cost_factor) and their device copy
  std::vector<uint2> factor_map;
  DevBuf<uint2> d_factor_map;
  bool factor_map_dirty = false;
  // chroma originals of caller-prepared macroblocks (jsvm_me_set_org_chroma)
  const int16_t* org_chroma = nullptr; int n_org_chroma = 0;
  DevBuf<int16_t> d_org_chroma;
  // jsvm_me_compensate (host-buffer form)
  DevBuf<jsvm_mc_job> d_mc_jobs;
  DevBuf<uint8_t> d_mc_pred;
  CUtensorMap map_k1{};
  EncodeTiledFn encode = nullptr;
  // params
  int search_range = 0, sub_dfunc = JSVM_DF_SAD, full_dfunc = JSVM_DF_SAD;
  int search_mode = JSVM_SEARCH_BLOCK, el_search_range = 0, fast_bi_search = 0;
  bool params_set = false;
  // neighbour MVs (A, B, C per job) for FAST_SEARCH / TZ_SEARCH: host pointer for jsvm_me_search, device pointer for jsvm_me_search_device
  const jsvm_mv* cands = nullptr; int n_cands = 0;
  const jsvm_mv* d_cands_user = nullptr;
  DevBuf<jsvm_mv> d_cands;
  // per-call device buffers (host-buffer API)
  DevBuf<jsvm_me_job> d_jobs;
  DevBuf<Group> d_groups;
  DevBuf<uint32_t> d_keys;
  DevBuf<uint32_t> d_full_costs;         // generic (SSD / Hadamard) full-pel search only
  DevBuf<JobAux> d_aux;
  DevBuf<jsvm_me_result> d_results;
  DevBuf<uint8_t> d_preds;
  DevBuf<int16_t> d_org;
  DevBuf<uint8_t> d_stage;              // staging for plane uploads
  int* d_bad = nullptr;
  std::vector<Group> h_groups;
  Group* pinned_groups[4] = {nullptr, nullptr, nullptr, nullptr};   // pinned ring: truly asynchronous group uploads
  size_t pinned_cap = 0;
  cudaEvent_t chunk_done[4] = {nullptr, nullptr, nullptr, nullptr};
  // host-buffer search: uploads and result copy-backs run on their own streams so that they overlap the kernels
  cudaStream_t up_stream = nullptr, down_stream = nullptr;
  cudaEvent_t ev_entry = nullptr, ev_kernels = nullptr;
  // small calls (the drop-in's one-job searches): one pinned, device-mapped arena that the kernels read and write directly --
  // no staging copies, two launches and one synchronisation per call
  struct SmallArena {
    jsvm_me_job job[64]; Group group[64]; jsvm_me_result result[64];
    int16_t org[64 * 256]; uint8_t pred[64 * 256];
    jsvm_mv cands[64 * 3];
  };
  SmallArena* small = nullptr;
  bool last_small = false;
  // state of the last host-API search (compensateBlock contract)
  std::vector<jsvm_me_job> last_jobs;
  bool last_has_preds = false;
  // measurement
  uint64_t launches = 0;
  const char* last_k2 = "";
  int last_k2_units = 0;                 // groups / jobs served by last_k2 in the current call (a call of several launches reports its largest one)
  bool timing = false;
  struct Timed { cudaEvent_t e0, e1; int kind; };
  std::vector<Timed> timed;
  double last_mc_ms = 0; uint64_t last_mc_n = 0;
  // picture-level search (jsvm_me_estimate_pictures)
  std::vector<int2> wave_mbs; std::vector<int> wave_off;     // macroblocks sorted by wavefront x + 2y, offsets per wave
  DevBuf<int2> d_wave_mbs;
  DevBuf<jsvm_pic_task> d_pic_tasks;
  DevBuf<jsvm_pic_mb> d_pic_field;
  DevBuf<PicTemp> d_pic_temp;
  DevBuf<jsvm_me_result> d_pic_results[2];
  int pic_n_last = 0;
  // the wave loop of jsvm_me_estimate_pictures as a CUDA graph (3302 dependent launches at 1080p), re-captured when its key changes
  cudaGraphExec_t pic_graph = nullptr; cudaStream_t pic_stream = nullptr; cudaEvent_t pic_ev[2] = {nullptr, nullptr};
  struct PicKey { int n_pics, refs, sub, sr; const void* p[8]; } pic_key{};
  const void* pic_aux = nullptr; uint64_t pic_launches = 0;
  // inter-layer resampling scratch
  DevBuf<int16_t> d_rs_src, d_rs_dst;
  DevBuf<uint8_t> d_rs_t8x8;
  DevBuf<jsvm_mb_motion> d_mu_base, d_mu_out;
  // multi-GPU gather
  ncclComm_t comm = nullptr; int comm_rank = 0, comm_world = 1;
};

namespace {

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to (function, device), not to a context: several contexts on one device (the
// drop-in keeps one per spatial layer) share it, so the configured maximum is process-wide state per device and is only ever raised.
constexpr int kMaxDevices = 64;
std::mutex g_smem_mu;
size_t g_smem_configured[kMaxDevices][5] = {};      // {group, persistent} x {8x8, 4x4}, group 8x8 with 128 threads
template <typename Fn>
int raise_dyn_smem(int device, int which, Fn* fn, size_t smem) {
  std::lock_guard<std::mutex> lk(g_smem_mu);
  size_t& configured = g_smem_configured[device % kMaxDevices][which];
  if (smem > configured) {
    CU_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  return JSVM_ME_OK;
}

int make_map(jsvm_me_ctx* c, CUtensorMap* map, void* base, int box_w, int box_h) {
  const cuuint64_t dims[3]    = { (cuuint64_t)c->plane_stride, (cuuint64_t)c->plane_rows, (cuuint64_t)c->n_slots };
  const cuuint64_t strides[2] = { (cuuint64_t)c->plane_stride, (cuuint64_t)c->plane_bytes };   // bytes, dims 1 and 2
  const cuuint32_t box[3]     = { (cuuint32_t)box_w, (cuuint32_t)box_h, 1 };
  const cuuint32_t estr[3]    = { 1, 1, 1 };
  CUresult r = c->encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return JSVM_ME_OK;
}

void begin_timed(jsvm_me_ctx* c, int kind) {
  if (!c->timing) return;
  jsvm_me_ctx::Timed t; t.kind = kind;
  cudaEventCreate(&t.e0); cudaEventCreate(&t.e1);
  cudaEventRecord(t.e0, c->stream);
  c->timed.push_back(t);
}
void end_timed(jsvm_me_ctx* c) {
  if (!c->timing) return;
  cudaEventRecord(c->timed.back().e1, c->stream);
}

int check_slot(jsvm_me_ctx* c, int slot) {
  if (!c) return fail("null context");
  if (!c->d_planes) return fail("jsvm_me_configure has not been called");
  if (slot < 0 || slot >= c->n_slots) return fail("slot %d out of range [0,%d)", slot, c->n_slots);
  return JSVM_ME_OK;
}

int launch_margin_fill(jsvm_me_ctx* c, const int* d_slots, int n) {
  dim3 grid(2, c->plane_rows, n), block(256);
  margin_fill_kernel<<<grid, block, 0, c->stream>>>(c->d_planes, c->plane_bytes, c->plane_stride, d_slots, c->W, c->H);
  c->launches++;
  CU_OK(cudaGetLastError());
  return JSVM_ME_OK;
}

// YuvPicBuffer::fillMargin of every plane uploaded since the last reader, in ONE launch: an upload only queues its copy (a copy,
// a 4-byte copy of the slot number and a kernel per plane kept the copy engine idle a third of the time: 17 planes of 1080p went
// up at 36 GB/s).  Every entry point that reads full-pel planes calls this first.
int flush_margins(jsvm_me_ctx* c) {
  if (c->margin_dirty.empty()) return JSVM_ME_OK;
  const int n = (int)c->margin_dirty.size();
  CU_OK(cudaMemcpyAsync(c->d_dirty_ids, c->margin_dirty.data(), sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));   // pageable: staged before return
  c->margin_dirty.clear();
  return launch_margin_fill(c, c->d_dirty_ids, n);
}
void mark_margin_dirty(jsvm_me_ctx* c, int slot) {
  if (std::find(c->margin_dirty.begin(), c->margin_dirty.end(), slot) == c->margin_dirty.end()) c->margin_dirty.push_back(slot);
}

// The planner: restates nothing from the reference -- it only decides which estimateBlockWithStart
// calls can share one SAD pass (same macroblock, planes and original; distinct partition slots;
// union of the clipped windows small enough for one CTA).  Pure host logic; large job lists are cut at
// macroblock boundaries and planned by several threads (the result equals the serial plan).
struct PlanChunk {
  std::vector<Group> g8, g4;
  int64_t npos = 0;
  int err_job = -1;
  char err[160] = "";
};

inline bool same_mb(const jsvm_me_job& a, const jsvm_me_job& b) {
  return a.cur_slot == b.cur_slot && a.ref_slot == b.ref_slot && a.mb_x == b.mb_x && a.mb_y == b.mb_y && a.org_index == b.org_index;
}

void plan_range(int width, int height, int n_slots, int search_range, const jsvm_me_job* jobs, int begin, int end, PlanChunk& out) {
  const int mbw = width >> 4, mbh = height >> 4;
  Group cur; bool open = false; Window cw{};
  auto close = [&]() {
    if (!open) return;
    cur.ux0 = (int16_t)cw.xlo; cur.uy0 = (int16_t)cw.ylo;
    cur.uw = (int16_t)(cw.xhi - cw.xlo + 1); cur.uh = (int16_t)(cw.yhi - cw.ylo + 1);
    (cur.gran4 ? out.g4 : out.g8).push_back(cur);
    open = false;
  };
  auto bad = [&](int i, const char* fmt, int a = 0, int b = 0, int c = 0, int d = 0) {
    out.err_job = i;
    snprintf(out.err, sizeof(out.err), fmt, a, b, c, d);
  };
  out.g8.reserve((size_t)(end - begin) / 8 + 16);
  for (int i = begin; i < end; ++i) {
    const jsvm_me_job& j = jobs[i];
    int w, h;
    if (!mode_size(j.mode, w, h)) return bad(i, "unsupported mode %d", (int)j.mode);
    if (j.blk > 15) return bad(i, "blk %d out of range", (int)j.blk);
    const int slot = partition_slot(j.mode, j.blk);
    if (slot < 0) return bad(i, "blk %d is not a legal origin for mode %d", (int)j.blk, (int)j.mode);
    if (j.cur_slot < 0 || j.cur_slot >= n_slots || j.ref_slot < 0 || j.ref_slot >= n_slots) return bad(i, "plane slot out of range");
    if (j.mb_x < 0 || j.mb_x >= mbw || j.mb_y < 0 || j.mb_y >= mbh) return bad(i, "macroblock (%d,%d) outside the picture", j.mb_x, j.mb_y);
    // the staged windows of every full-search kernel (packed-byte and generic) are sized for search ranges up to kMaxSearchRange
    if (j.search_range < 0 || j.search_range > kMaxSearchRange) return bad(i, "per-job search range %d outside [0,%d]", (int)j.search_range, kMaxSearchRange);
    const Window w1 = job_window(j, mbw, mbh, search_range);
    const int ww = w1.xhi - w1.xlo + 1, wh = w1.yhi - w1.ylo + 1;
    if (ww < 1 || wh < 1) return bad(i, "empty search window");
    if (ww > kMaxUW || wh > kMaxUH || ww * wh > (1 << kIdxBits))
      return bad(i, "search window %dx%d exceeds the %dx%d limit of this build (search range <= 64)", ww, wh, kMaxUW, kMaxUH);
    out.npos += (int64_t)ww * wh;
    bool join = open && (i - cur.first_job) < 127 && same_mb(j, jobs[cur.first_job]) && cur.slot_job[slot] < 0;
    if (join) {
      Window u = cw;
      u.xlo = std::min(u.xlo, w1.xlo); u.xhi = std::max(u.xhi, w1.xhi);
      u.ylo = std::min(u.ylo, w1.ylo); u.yhi = std::max(u.yhi, w1.yhi);
      const int uw = u.xhi - u.xlo + 1, uh = u.yhi - u.ylo + 1;
      if (uw > kMaxUW || uh > kMaxUH || uw * uh > (1 << kIdxBits)) join = false;
      else cw = u;
    }
    if (!join) {
      close();
      memset(&cur, 0, sizeof(cur));
      cur.first_job = i; cur.n_jobs = 0; cur.gran4 = 0;
      cur.cur_slot = j.cur_slot; cur.ref_slot = j.ref_slot; cur.mb_x = j.mb_x; cur.mb_y = j.mb_y; cur.org_index = j.org_index;
      memset(cur.slot_job, -1, sizeof(cur.slot_job));
      cw = w1; open = true;
    }
    cur.slot_job[slot] = (int8_t)(i - cur.first_job);
    cur.n_jobs = (int16_t)(i - cur.first_job + 1);
    if (slot >= kSlotsGran8) cur.gran4 = 1;
  }
  close();
}

int plan_jobs(int width, int height, int n_slots, int search_range, const jsvm_me_job* jobs, int n_jobs, std::vector<Group>& out, jsvm_me_plan_info* info) {
  int n_thr = 1;
  if (n_jobs >= 32768) n_thr = (int)std::min<unsigned>(std::min<unsigned>(16u, (unsigned)(n_jobs / 16384)), std::max(1u, std::thread::hardware_concurrency()));
  // chunk boundaries only where the macroblock changes: the serial planner starts a new group there anyway
  std::vector<int> cut(n_thr + 1, n_jobs);
  cut[0] = 0;
  for (int t = 1; t < n_thr; ++t) {
    int i = std::max(cut[t - 1], (int)((int64_t)n_jobs * t / n_thr));
    while (i < n_jobs && i > 0 && same_mb(jobs[i], jobs[i - 1])) ++i;
    cut[t] = i;
  }
  std::vector<PlanChunk> chunks(n_thr);
  if (n_thr == 1) plan_range(width, height, n_slots, search_range, jobs, 0, n_jobs, chunks[0]);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_thr; ++t)
      th.emplace_back(plan_range, width, height, n_slots, search_range, jobs, cut[t], cut[t + 1], std::ref(chunks[t]));
    for (auto& x : th) x.join();
  }
  size_t n8 = 0, n4 = 0;
  for (auto& c : chunks) {
    if (c.err_job >= 0) return fail("job %d: %s", c.err_job, c.err);
    n8 += c.g8.size(); n4 += c.g4.size();
  }
  out.resize(n8 + n4);
  size_t o8 = 0, o4 = n8;
  info->max_uh_8x8 = 0; info->max_uh_4x4 = 0; info->n_positions = 0; info->n_group_positions = 0;
  for (auto& c : chunks) {
    if (!c.g8.empty()) memcpy(&out[o8], c.g8.data(), c.g8.size() * sizeof(Group));
    if (!c.g4.empty()) memcpy(&out[o4], c.g4.data(), c.g4.size() * sizeof(Group));
    o8 += c.g8.size(); o4 += c.g4.size();
    info->n_positions += c.npos;
  }
  for (size_t i = 0; i < out.size(); ++i) {
    const Group& g = out[i];
    int32_t& mx = g.gran4 ? info->max_uh_4x4 : info->max_uh_8x8;
    mx = std::max<int32_t>(mx, g.uh);
    info->n_group_positions += (int64_t)g.uw * g.uh;
  }
  info->n_groups = (int32_t)out.size();
  info->n_groups_8x8 = (int32_t)n8;
  return JSVM_ME_OK;
}

template <int V, bool GRAN4>
int launch_search(jsvm_me_ctx* c, const Group* d_groups, int first_group, int n_groups, int max_uh,
                  const jsvm_me_job* d_jobs, const int16_t* d_org, uint32_t* d_keys, JobAux* d_aux) {
  if (n_groups <= 0) return JSVM_ME_OK;
  const size_t smem = k2_smem_bytes<V, GRAN4>(max_uh);
  // small windows (8x8-granular groups, union windows of at most kSmallUH rows): 128-thread CTAs, four per SM
  static const int small_uh = getenv("JSVM_ME_SMALL_UH") ? atoi(getenv("JSVM_ME_SMALL_UH")) : kSmallUH;
  const bool small = !GRAN4 && max_uh <= small_uh && 4 * smem <= kMaxDynSmemTotal;
  if (small) {
    if (raise_dyn_smem(c->device, 4, full_search_kernel<V, false, false, 128>, smem)) return JSVM_ME_ERR;
  } else if (raise_dyn_smem(c->device, GRAN4 ? 1 : 0, full_search_kernel<V, GRAN4>, smem)) return JSVM_ME_ERR;
  begin_timed(c, 1);
  if (small)
    full_search_kernel<V, false, false, 128><<<n_groups, 128, smem, c->stream>>>(
        c->d_planes, c->plane_bytes, c->plane_stride, d_groups + first_group, d_jobs, d_org,
        c->W >> 4, c->H >> 4, c->search_range, d_keys, d_aux);
  else
  full_search_kernel<V, GRAN4><<<n_groups, kK2Threads, smem, c->stream>>>(
      c->d_planes, c->plane_bytes, c->plane_stride, d_groups + first_group, d_jobs, d_org,
      c->W >> 4, c->H >> 4, c->search_range, d_keys, d_aux);
  end_timed(c);
  c->launches++;
  CU_OK(cudaGetLastError());
  return JSVM_ME_OK;
}

// The persistent producer / consumer kernel (one CTA per SM, two shared-memory stages): used when the launch is large
// enough to keep every SM busy for many groups and two stages of the tallest union window fit.
constexpr size_t kMaxDynSmem = 232448;     // 227 KB per CTA on sm_100
template <int V, bool GRAN4>
int launch_search_persistent(jsvm_me_ctx* c, const Group* d_groups, int n_groups, int max_uh,
                             const jsvm_me_job* d_jobs, const int16_t* d_org, uint32_t* d_keys, JobAux* d_aux) {
  const size_t smem = k2p_smem_bytes<V, GRAN4>(max_uh);
  if (raise_dyn_smem(c->device, GRAN4 ? 3 : 2, full_search_persistent_kernel<V, GRAN4>, smem)) return JSVM_ME_ERR;
  begin_timed(c, 1);
  full_search_persistent_kernel<V, GRAN4><<<std::min(n_groups, c->sm_count), kK2PThreads, smem, c->stream>>>(
      c->d_planes, c->plane_bytes, c->plane_stride, d_groups, n_groups, d_jobs, d_org,
      c->W >> 4, c->H >> 4, c->search_range, d_keys, d_aux, (int)k2p_stage_bytes<V, GRAN4>(max_uh));
  end_timed(c);
  c->launches++;
  CU_OK(cudaGetLastError());
  return JSVM_ME_OK;
}

// which of the two kernels serves n_groups groups whose tallest union window has max_uh rows
template <int V, bool GRAN4>
int choose_persistent(jsvm_me_ctx* c, int n_groups, int max_uh, bool* persistent) {
  // JSVM_ME_K2 = group | persistent forces one of the two kernels (A/B measurements)
  const char* force = getenv("JSVM_ME_K2");
  const bool fits = k2p_smem_bytes<V, GRAN4>(max_uh) <= kMaxDynSmem;
  // 4x4-granular groups (41 slots): staging 41 rate tables keeps the staging warps busier than the search warps; measured
  // 19 % slower than one CTA per group on C3, so the persistent kernel is only chosen for them when forced
  // Small windows (search range <= ~40: union windows of at most 96 rows) leave a group too little work to hide the hand-over
  // between stages: one CTA per group measured 8 % (CIF, SR 32) to 17 % (QCIF, SR 16) faster there.
  *persistent = n_groups >= 4 * c->sm_count && fits && !GRAN4 && max_uh > 96;
  if (force && !strcmp(force, "group")) *persistent = false;
  if (force && !strcmp(force, "persistent")) {
    if (!fits) return fail("JSVM_ME_K2=persistent: union windows of %d rows do not fit two stages", max_uh);
    *persistent = n_groups > 0;
  }
  return JSVM_ME_OK;
}

// K2 for the groups of a plan, then K3 for the jobs [job_begin, job_end) they cover (K2 reaches jobs through the groups'
// absolute first_job; all per-job arrays are indexed by the absolute job number).
// the two stages use different cost-factor classes (ENC/MotionEstimation.cpp:255, 325): SSE factor iff the stage's function is SSD
inline bool mixed_factor_classes(const jsvm_me_ctx* c) { return (c->full_dfunc == JSVM_DF_SSD) != (c->sub_dfunc == JSVM_DF_SSD); }

int sync_factor_map(jsvm_me_ctx* c) {
  if (!c->factor_map_dirty) return JSVM_ME_OK;
  if (c->d_factor_map.reserve(std::max<size_t>(1, c->factor_map.size()))) return JSVM_ME_ERR;
  if (!c->factor_map.empty())
    CU_OK(cudaMemcpyAsync(c->d_factor_map.p, c->factor_map.data(), sizeof(uint2) * c->factor_map.size(), cudaMemcpyHostToDevice, c->stream));
  CU_OK(cudaStreamSynchronize(c->stream));                  // the host vector may change again right after this call
  c->factor_map_dirty = false;
  return JSVM_ME_OK;
}

int run_search(jsvm_me_ctx* c, const jsvm_me_job* d_jobs, int job_begin, int job_end, const Group* d_groups, const jsvm_me_plan_info* info,
               const int16_t* d_org, uint32_t* d_keys, JobAux* d_aux, jsvm_me_result* d_results, uint8_t* d_preds,
               const int16_t* d_org_chroma = nullptr, const jsvm_mv* d_cands = nullptr) {
  if (!c->params_set) return fail("jsvm_me_set_params has not been called");
  // SPIRAL / LOG / FAST / TZ, and the full search with SearchRange > 64 (windows too large to stage): one warp per job (K5), no plan
  const bool fast = c->search_mode != JSVM_SEARCH_BLOCK || c->search_range > kMaxSearchRange;
  const bool refine = !fast && c->fast_bi_search;            // BLOCK_SEARCH + FastBiSearch: the jobs with a per-call range are redone by K5
  if (!fast && (info->max_uh_8x8 > kMaxUH || info->max_uh_4x4 > kMaxUH)) return fail("plan: union window taller than %d", kMaxUH);
  const int n = job_end - job_begin;
  if (n <= 0) return JSVM_ME_OK;
  if (flush_margins(c) || sync_factor_map(c)) return JSVM_ME_ERR;
  if ((fast || refine) && c->d_full_costs.cap < (size_t)job_end) return fail("internal: full-pel cost buffer not reserved");
  if (c->full_dfunc == JSVM_DF_YUV_SAD && !c->d_chroma) return fail("SearchFuncFullPel 3 (YUV_SAD) needs the chroma planes (jsvm_me_upload_chroma_*)");
  auto launch_k5 = [&](int only_refinement) -> int {
    auto k5 = c->full_dfunc == JSVM_DF_SAD ? fast_search_kernel<JSVM_DF_SAD> : c->full_dfunc == JSVM_DF_SSD ? fast_search_kernel<JSVM_DF_SSD>
            : c->full_dfunc == JSVM_DF_HADAMARD ? fast_search_kernel<JSVM_DF_HADAMARD> : fast_search_kernel<JSVM_DF_YUV_SAD>;
    begin_timed(c, 1);
    k5<<<(n + kK5WarpsPerCta - 1) / kK5WarpsPerCta, 32 * kK5WarpsPerCta, 0, c->stream>>>(
        c->d_planes, c->plane_bytes, c->plane_stride, c->d_chroma, c->chroma_bytes, c->chroma_stride,
        d_jobs + job_begin, n, d_org, d_org_chroma, d_cands ? d_cands + 3 * (size_t)job_begin : nullptr,
        c->W >> 4, c->H >> 4, c->search_mode, c->search_range, c->el_search_range ? c->el_search_range : c->search_range, c->fast_bi_search,
        only_refinement, d_keys + job_begin, c->d_full_costs.p + job_begin, d_aux + job_begin);
    end_timed(c);
    c->launches++;
    CU_OK(cudaGetLastError());
    return JSVM_ME_OK;
  };
  if (fast) {
    if (launch_k5(0)) return JSVM_ME_ERR;
    static const char* const names[5] = {"fast_search_kernel(BLOCK, SearchRange > 64)", "fast_search_kernel(SPIRAL)", "fast_search_kernel(LOG)", "fast_search_kernel(FAST)", "fast_search_kernel(TZ)"};
    c->last_k2 = names[c->search_mode];
  }
  const bool generic = !fast && c->full_dfunc != JSVM_DF_SAD;         // SSD / Hadamard / YUV-SAD full search: one CTA per job, no grouping
  if (generic) {
    if (c->d_full_costs.cap < (size_t)job_end) return fail("internal: full-pel cost buffer not reserved");
    auto k2g = c->full_dfunc == JSVM_DF_SSD ? full_search_generic_kernel<JSVM_DF_SSD>
             : c->full_dfunc == JSVM_DF_HADAMARD ? full_search_generic_kernel<JSVM_DF_HADAMARD> : full_search_generic_kernel<JSVM_DF_YUV_SAD>;
    begin_timed(c, 1);
    k2g<<<n, kK2GThreads, 0, c->stream>>>(c->d_planes, c->plane_bytes, c->plane_stride, c->d_chroma, c->chroma_bytes, c->chroma_stride,
                                          d_jobs + job_begin, n, d_org, d_org_chroma,
                                          c->W >> 4, c->H >> 4, c->search_range, d_keys + job_begin,
                                          c->d_full_costs.p + job_begin, d_aux + job_begin);
    end_timed(c);
    c->launches++;
    c->last_k2 = c->full_dfunc == JSVM_DF_SSD ? "full_search_generic_kernel<SSD>"
               : c->full_dfunc == JSVM_DF_HADAMARD ? "full_search_generic_kernel<HADAMARD>" : "full_search_generic_kernel<YUV_SAD>";
    CU_OK(cudaGetLastError());
  }
  const int n8 = generic || fast ? 0 : info->n_groups_8x8, n4 = generic || fast ? 0 : info->n_groups - info->n_groups_8x8;
  bool p8 = false, p4 = false;
  if (choose_persistent<kV8, false>(c, n8, info->max_uh_8x8, &p8) || choose_persistent<kV4, true>(c, n4, info->max_uh_4x4, &p4)) return JSVM_ME_ERR;
  // a host-buffer call is served by several launches (the first and the last of them can be small): report the largest one
  if (job_begin == 0) c->last_k2_units = 0;
  if (std::max(n8, n4) >= c->last_k2_units) {
    c->last_k2_units = std::max(n8, n4);
    if (n8 > 0) c->last_k2 = p8 ? "full_search_persistent_kernel<8,false>" : "full_search_kernel<8,false>";
    else if (n4 > 0) c->last_k2 = p4 ? "full_search_persistent_kernel<4,true>" : "full_search_kernel<4,true>";
  }
  if (n8 > 0 && (p8 ? launch_search_persistent<kV8, false>(c, d_groups, n8, info->max_uh_8x8, d_jobs, d_org, d_keys, d_aux)
                    : launch_search<kV8, false>(c, d_groups, 0, n8, info->max_uh_8x8, d_jobs, d_org, d_keys, d_aux))) return JSVM_ME_ERR;
  if (n4 > 0 && (p4 ? launch_search_persistent<kV4, true>(c, d_groups + n8, n4, info->max_uh_4x4, d_jobs, d_org, d_keys, d_aux)
                    : launch_search<kV4, true>(c, d_groups, n8, n4, info->max_uh_4x4, d_jobs, d_org, d_keys, d_aux))) return JSVM_ME_ERR;
  if (refine && launch_k5(1)) return JSVM_ME_ERR;
  begin_timed(c, 2);
  const int ctas = (n + kK3WarpsPerCta - 1) / kK3WarpsPerCta;
  auto k3 = c->sub_dfunc == JSVM_DF_HADAMARD ? subpel_kernel<JSVM_DF_HADAMARD>
          : c->sub_dfunc == JSVM_DF_SSD ? subpel_kernel<JSVM_DF_SSD> : subpel_kernel<JSVM_DF_SAD>;
  k3<<<ctas, 32 * kK3WarpsPerCta, 0, c->stream>>>(
      c->d_planes, c->plane_bytes, c->plane_stride, c->d_half, c->half_bytes, c->half_stride,
      d_jobs + job_begin, d_aux + job_begin, n, d_org, d_keys + job_begin,
      generic || fast || refine ? c->d_full_costs.p + job_begin : nullptr, d_results + job_begin,
      d_preds ? d_preds + (size_t)job_begin * 256 : nullptr,
      c->d_factor_map.p, mixed_factor_classes(c) ? (int)c->factor_map.size() : 0);
  end_timed(c);
  c->launches++;
  CU_OK(cudaGetLastError());
  return JSVM_ME_OK;
}

// What a host-buffer search needs beyond the plan when the distortion functions are not plain SAD / SAD-class:
//  * YUV_SAD: chroma planes of every slot the jobs touch, and the chroma originals (if registered) on the device;
//  * mixed factor classes: every job's full-pel factor must have a registered sub-pel partner.
int prepare_search_extras(jsvm_me_ctx* c, const jsvm_me_job* jobs, int n_jobs, int n_org_mbs, const int16_t** d_org_chroma) {
  *d_org_chroma = nullptr;
  if (c->full_dfunc == JSVM_DF_YUV_SAD) {
    if (!c->d_chroma) return fail("SearchFuncFullPel 3 (YUV_SAD) needs the chroma planes (jsvm_me_upload_chroma_*)");
    const bool have_org = c->org_chroma != nullptr;
    for (int i = 0; i < n_jobs; ++i) {
      const jsvm_me_job& j = jobs[i];
      if (j.ref_slot < 0 || j.ref_slot >= c->n_slots || j.cur_slot < 0 || j.cur_slot >= c->n_slots) return fail("job %d: plane slot out of range", i);
      if (j.org_index >= 0 && !have_org) continue;                         // luma-only job (bi-pred / unequal weights)
      if (!c->slot_chroma[j.ref_slot]) return fail("job %d: ref_slot %d has no chroma planes", i, (int)j.ref_slot);
      if (j.org_index < 0 && !c->slot_chroma[j.cur_slot]) return fail("job %d: cur_slot %d has no chroma planes", i, (int)j.cur_slot);
    }
    if (have_org && n_org_mbs > 0) {
      if (c->n_org_chroma < n_org_mbs) return fail("jsvm_me_set_org_chroma registered %d macroblocks, the call passes %d", c->n_org_chroma, n_org_mbs);
      if (c->d_org_chroma.reserve((size_t)n_org_mbs * 128)) return JSVM_ME_ERR;
      CU_OK(cudaMemcpyAsync(c->d_org_chroma.p, c->org_chroma, sizeof(int16_t) * 128 * n_org_mbs, cudaMemcpyHostToDevice, c->stream));
      *d_org_chroma = c->d_org_chroma.p;
    }
  }
  if (mixed_factor_classes(c)) {
    uint32_t last = 0; bool have_last = false;
    for (int i = 0; i < n_jobs; ++i) {
      const uint32_t f = jobs[i].cost_factor;
      if (have_last && f == last) continue;
      bool found = false;
      for (const uint2& e : c->factor_map) if (e.x == f) { found = true; break; }
      if (!found) return fail("job %d: cost factor %u has no sub-pel partner (SearchFuncFullPel %d / SearchFuncSubPel %d use different "
                              "factor classes: register the pair with jsvm_me_map_cost_factor)", i, f, c->full_dfunc, c->sub_dfunc);
      last = f; have_last = true;
    }
  }
  return JSVM_ME_OK;
}

// jsvm_me_search for SPIRAL / LOG / FAST / TZ: no plan, one batch (these searches touch a few hundred points per job).
int check_fast_jobs(jsvm_me_ctx* c, const jsvm_me_job* jobs, int n_jobs, int n_org_mbs) {
  const int mbw = c->W >> 4, mbh = c->H >> 4;
  for (int i = 0; i < n_jobs; ++i) {
    const jsvm_me_job& j = jobs[i];
    int w, h;
    if (!mode_size(j.mode, w, h)) return fail("job %d: unsupported mode %d", i, (int)(j.mode & JSVM_JOB_MODE_MASK));
    if (j.blk > 15 || partition_slot(j.mode, j.blk) < 0) return fail("job %d: blk %d is not a legal origin for mode %d", i, (int)j.blk, (int)(j.mode & JSVM_JOB_MODE_MASK));
    if (j.cur_slot < 0 || j.cur_slot >= c->n_slots || j.ref_slot < 0 || j.ref_slot >= c->n_slots) return fail("job %d: plane slot out of range", i);
    if (j.mb_x < 0 || j.mb_x >= mbw || j.mb_y < 0 || j.mb_y >= mbh) return fail("job %d: macroblock (%d,%d) outside the picture", i, j.mb_x, j.mb_y);
    if (j.search_range < 0 || j.search_range > kMaxWideRange) return fail("job %d: per-job search range %d outside [0,%d]", i, (int)j.search_range, kMaxWideRange);
    if (j.org_index >= n_org_mbs) return fail("job %d: org_index %d >= n_org_mbs %d", i, j.org_index, n_org_mbs);
    if (j.org_index < 0 && !c->slot_full[j.cur_slot]) return fail("job %d: cur_slot %d is empty", i, (int)j.cur_slot);
    if (!c->slot_half[j.ref_slot]) return fail("job %d: ref_slot %d has no half-pel plane", i, (int)j.ref_slot);
    if (c->search_mode == JSVM_SEARCH_SPIRAL && j.search_range == 0) {
      // xPelSpiralSearch limits the start to [ m_cMin + 4 SR, m_cMax - 4 SR ] and then reads (2 SR + 1)^2 positions unchecked
      const int lo_h = ((-16 * j.mb_x - 24) + c->search_range) * 4, hi_h = ((16 * (mbw - j.mb_x) + 8) - c->search_range) * 4;
      const int lo_v = ((-16 * j.mb_y - 24) + c->search_range) * 4, hi_v = ((16 * (mbh - j.mb_y) + 8) - c->search_range) * 4;
      if (lo_h > hi_h || lo_v > hi_v)
        return fail("job %d: spiral search with range %d does not fit the padded %dx%d picture (the reference reads outside its planes here)",
                    i, c->search_range, c->W, c->H);
    }
  }
  const bool need_cands = c->search_mode == JSVM_SEARCH_FAST || c->search_mode == JSVM_SEARCH_TZ;
  if (need_cands && c->cands && c->n_cands < n_jobs) return fail("jsvm_me_set_mv_candidates registered %d jobs, the call passes %d", c->n_cands, n_jobs);
  return JSVM_ME_OK;
}

int search_fast_host(jsvm_me_ctx* c, const jsvm_me_job* jobs, int n_jobs, const int16_t* org_mbs, int n_org_mbs,
                     jsvm_me_result* results, uint8_t* preds) {
  CU_OK(cudaSetDevice(c->device));
  if (check_fast_jobs(c, jobs, n_jobs, n_org_mbs)) return JSVM_ME_ERR;
  const bool need_cands = c->search_mode == JSVM_SEARCH_FAST || c->search_mode == JSVM_SEARCH_TZ;
  const bool want_preds = preds != nullptr || n_jobs <= 4096;
  if (c->d_jobs.reserve(n_jobs) || c->d_keys.reserve(n_jobs) || c->d_full_costs.reserve(n_jobs) || c->d_aux.reserve(n_jobs) ||
      c->d_results.reserve(n_jobs) || (want_preds && c->d_preds.reserve((size_t)n_jobs * 256)) ||
      (n_org_mbs > 0 && c->d_org.reserve((size_t)n_org_mbs * 256)) || (need_cands && c->cands && c->d_cands.reserve((size_t)n_jobs * 3))) return JSVM_ME_ERR;
  const int16_t* d_org_chroma = nullptr;
  if (prepare_search_extras(c, jobs, n_jobs, n_org_mbs, &d_org_chroma)) return JSVM_ME_ERR;
  if (n_org_mbs > 0) CU_OK(cudaMemcpyAsync(c->d_org.p, org_mbs, sizeof(int16_t) * 256 * n_org_mbs, cudaMemcpyHostToDevice, c->stream));
  CU_OK(cudaMemcpyAsync(c->d_jobs.p, jobs, sizeof(jsvm_me_job) * n_jobs, cudaMemcpyHostToDevice, c->stream));
  const jsvm_mv* d_cands = nullptr;
  if (need_cands && c->cands) {
    CU_OK(cudaMemcpyAsync(c->d_cands.p, c->cands, sizeof(jsvm_mv) * 3 * n_jobs, cudaMemcpyHostToDevice, c->stream));
    d_cands = c->d_cands.p;
  }
  jsvm_me_plan_info none{};
  if (run_search(c, c->d_jobs.p, 0, n_jobs, nullptr, &none, c->d_org.p, c->d_keys.p, c->d_aux.p, c->d_results.p, want_preds ? c->d_preds.p : nullptr,
                 d_org_chroma, d_cands)) { cudaStreamSynchronize(c->stream); return JSVM_ME_ERR; }
  CU_OK(cudaMemcpyAsync(results, c->d_results.p, sizeof(jsvm_me_result) * n_jobs, cudaMemcpyDeviceToHost, c->stream));
  if (preds) CU_OK(cudaMemcpyAsync(preds, c->d_preds.p, (size_t)n_jobs * 256, cudaMemcpyDeviceToHost, c->stream));
  CU_OK(cudaStreamSynchronize(c->stream));
  c->last_small = false;
  if (want_preds && n_jobs <= 4096) c->last_jobs.assign(jobs, jobs + n_jobs); else c->last_jobs.clear();
  c->last_has_preds = want_preds && n_jobs <= 4096;
  return JSVM_ME_OK;
}

// jsvm_me_search for up to 64 jobs: what MbEncoder's one-call-at-a-time use of estimateBlockWithStart turns into.
int search_small(jsvm_me_ctx* c, const jsvm_me_job* jobs, int n_jobs, const int16_t* org_mbs, int n_org_mbs,
                 jsvm_me_result* results, uint8_t* preds) {
  CU_OK(cudaSetDevice(c->device));
  if (!c->small) CU_OK(cudaHostAlloc((void**)&c->small, sizeof(jsvm_me_ctx::SmallArena), cudaHostAllocMapped));
  jsvm_me_ctx::SmallArena* a = c->small;                     // unified addressing: the same pointer is valid on the device
  const bool fast = c->search_mode != JSVM_SEARCH_BLOCK || c->search_range > kMaxSearchRange;
  PlanChunk pc;
  const jsvm_mv* d_cands = nullptr;
  if (fast) {                                                // SPIRAL / LOG / FAST / TZ: no plan; the neighbours' MVs travel in the arena
    if (check_fast_jobs(c, jobs, n_jobs, n_org_mbs)) return JSVM_ME_ERR;
    if ((c->search_mode == JSVM_SEARCH_FAST || c->search_mode == JSVM_SEARCH_TZ) && c->cands) {
      memcpy(a->cands, c->cands, sizeof(jsvm_mv) * 3 * n_jobs);
      d_cands = a->cands;
    }
  } else {
    plan_range(c->W, c->H, c->n_slots, c->search_range, jobs, 0, n_jobs, pc);
    if (pc.err_job >= 0) return fail("job %d: %s", pc.err_job, pc.err);
  }
  jsvm_me_plan_info info{};
  info.n_groups_8x8 = (int32_t)pc.g8.size(); info.n_groups = (int32_t)(pc.g8.size() + pc.g4.size()); info.n_positions = pc.npos;
  int ng = 0;
  for (const std::vector<Group>* v : {&pc.g8, &pc.g4})
    for (const Group& g : *v) {
      if (g.org_index >= n_org_mbs) return fail("job %d: org_index %d >= n_org_mbs %d", g.first_job, g.org_index, n_org_mbs);
      if (g.org_index < 0 && !c->slot_full[g.cur_slot]) return fail("job %d: cur_slot %d is empty", g.first_job, g.cur_slot);
      if (!c->slot_half[g.ref_slot]) return fail("job %d: ref_slot %d has no half-pel plane", g.first_job, g.ref_slot);
      int32_t& mx = g.gran4 ? info.max_uh_4x4 : info.max_uh_8x8;
      mx = std::max<int32_t>(mx, g.uh);
      a->group[ng++] = g;
    }
  memcpy(a->job, jobs, sizeof(jsvm_me_job) * n_jobs);
  if (n_org_mbs > 0) memcpy(a->org, org_mbs, sizeof(int16_t) * 256 * n_org_mbs);
  if (c->d_keys.reserve(64) || c->d_full_costs.reserve(64) || c->d_aux.reserve(64)) return JSVM_ME_ERR;
  const int16_t* d_org_chroma = nullptr;
  if (prepare_search_extras(c, jobs, n_jobs, n_org_mbs, &d_org_chroma)) return JSVM_ME_ERR;
  if (run_search(c, a->job, 0, n_jobs, a->group, &info, a->org, c->d_keys.p, c->d_aux.p, a->result, a->pred, d_org_chroma, d_cands)) return JSVM_ME_ERR;
  CU_OK(cudaStreamSynchronize(c->stream));
  memcpy(results, a->result, sizeof(jsvm_me_result) * n_jobs);
  if (preds) memcpy(preds, a->pred, (size_t)n_jobs * 256);
  c->last_jobs.assign(jobs, jobs + n_jobs);
  c->last_has_preds = true; c->last_small = true;
  return JSVM_ME_OK;
}

}  // namespace

extern "C" {

const char* jsvm_me_last_error(void) { return g_err; }

int jsvm_me_create(jsvm_me_ctx** out, int device) {
  if (!out) return fail("null output pointer");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail("no CUDA device available (%s); this library has no CPU path", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail("device %d out of range [0,%d)", device, n);
  cudaDeviceProp p;
  CU_OK(cudaGetDeviceProperties(&p, device));
  if (p.major != 10) return fail("device %d is sm_%d%d; the kernels are built for sm_100a only", device, p.major, p.minor);
  CU_OK(cudaSetDevice(device));
  jsvm_me_ctx* c = new jsvm_me_ctx();
  c->device = device;
  c->sm_count = p.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || !fn) { delete c; return fail("cuTensorMapEncodeTiled not available from the driver"); }
  c->encode = (EncodeTiledFn)fn;
  if (cudaMalloc((void**)&c->d_bad, sizeof(int)) != cudaSuccess) { delete c; return fail("cudaMalloc failed"); }
  *out = c;
  return JSVM_ME_OK;
}

int jsvm_me_destroy(jsvm_me_ctx* c) {
  if (!c) return JSVM_ME_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  jsvm_me_comm_destroy(c);
  if (c->pic_graph) cudaGraphExecDestroy(c->pic_graph);
  if (c->pic_stream) cudaStreamDestroy(c->pic_stream);
  for (auto& e : c->pic_ev) if (e) cudaEventDestroy(e);
  if (c->d_planes) cudaFree(c->d_planes);
  if (c->d_half) cudaFree(c->d_half);
  if (c->d_chroma) cudaFree(c->d_chroma);
  if (c->d_slot_ids) cudaFree(c->d_slot_ids);
  if (c->d_dirty_ids) cudaFree(c->d_dirty_ids);
  if (c->d_bad) cudaFree(c->d_bad);
  c->d_mu_base.release(); c->d_mu_out.release(); c->d_rs_src.release(); c->d_rs_dst.release(); c->d_rs_t8x8.release(); c->d_wave_mbs.release(); c->d_pic_tasks.release(); c->d_pic_field.release(); c->d_pic_temp.release(); c->d_pic_results[0].release(); c->d_pic_results[1].release();
  c->d_cands.release(); c->d_factor_map.release(); c->d_org_chroma.release(); c->d_mc_jobs.release(); c->d_mc_pred.release();
  c->d_jobs.release(); c->d_groups.release(); c->d_keys.release(); c->d_full_costs.release(); c->d_aux.release();
  c->d_results.release(); c->d_preds.release(); c->d_org.release(); c->d_stage.release();
  for (auto& t : c->timed) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); }
  for (auto& e : c->chunk_done) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : {c->ev_entry, c->ev_kernels}) if (e) cudaEventDestroy(e);
  if (c->up_stream) cudaStreamDestroy(c->up_stream);
  if (c->down_stream) cudaStreamDestroy(c->down_stream);
  for (auto& pg : c->pinned_groups) if (pg) cudaFreeHost(pg);
  if (c->small) cudaFreeHost(c->small);
  delete c;
  return JSVM_ME_OK;
}

int jsvm_me_configure(jsvm_me_ctx* c, int width, int height, int n_slots) {
  if (!c) return fail("null context");
  if (width <= 0 || height <= 0 || (width & 15) || (height & 15)) return fail("picture size %dx%d must be a positive multiple of 16", width, height);
  if (n_slots <= 0 || n_slots > 32767) return fail("n_slots %d out of range", n_slots);
  CU_OK(cudaSetDevice(c->device));
  CU_OK(cudaStreamSynchronize(c->stream));
  if (c->d_planes) { cudaFree(c->d_planes); c->d_planes = nullptr; }
  if (c->d_half) { cudaFree(c->d_half); c->d_half = nullptr; }
  if (c->d_chroma) { cudaFree(c->d_chroma); c->d_chroma = nullptr; }
  if (c->d_slot_ids) { cudaFree(c->d_slot_ids); c->d_slot_ids = nullptr; }
  if (c->d_dirty_ids) { cudaFree(c->d_dirty_ids); c->d_dirty_ids = nullptr; }
  c->margin_dirty.clear();
  c->W = width; c->H = height; c->n_slots = n_slots;
  c->wave_mbs.clear(); c->wave_off.clear(); c->pic_n_last = 0;
  if (c->pic_graph) { cudaGraphExecDestroy(c->pic_graph); c->pic_graph = nullptr; }
  c->cands = nullptr; c->n_cands = 0; c->d_cands_user = nullptr; c->org_chroma = nullptr; c->n_org_chroma = 0;   // per-call inputs of the old geometry
  c->plane_stride = width + 2 * kMargin; c->plane_rows = height + 2 * kMargin;
  c->plane_bytes = (size_t)c->plane_stride * c->plane_rows;
  c->half_stride = 2 * c->plane_stride; c->half_rows = 2 * c->plane_rows;
  c->half_bytes = (size_t)c->half_stride * c->half_rows;
  CU_OK(cudaMalloc((void**)&c->d_planes, c->plane_bytes * n_slots + 256));
  CU_OK(cudaMalloc((void**)&c->d_half, c->half_bytes * n_slots));
  CU_OK(cudaMalloc((void**)&c->d_slot_ids, sizeof(int) * n_slots));
  CU_OK(cudaMalloc((void**)&c->d_dirty_ids, sizeof(int) * n_slots));
  CU_OK(cudaMemsetAsync(c->d_planes, 0, c->plane_bytes * n_slots + 256, c->stream));
  CU_OK(cudaMemsetAsync(c->d_half, 0, c->half_bytes * n_slots, c->stream));     // samples outside the written region stay 0 (reference: memset at allocation)
  c->slot_full.assign(n_slots, 0); c->slot_half.assign(n_slots, 0); c->slot_chroma.assign(n_slots, 0);
  c->chroma_stride = width / 2 + 2 * kCMargin; c->chroma_rows = height / 2 + 2 * kCMargin;
  c->chroma_bytes = (size_t)c->chroma_stride * c->chroma_rows;
  if (make_map(c, &c->map_k1, c->d_planes, kK1BoxW, kK1BoxH)) return JSVM_ME_ERR;
  return JSVM_ME_OK;
}

int jsvm_me_set_params(jsvm_me_ctx* c, int search_mode, int full_pel_dfunc, int sub_pel_dfunc, int search_range) {
  if (!c) return fail("null context");
  // MotionVectorSearchParams::check (ENC/CodingParameter.cpp:17-36)
  if (search_mode < JSVM_SEARCH_BLOCK || search_mode > JSVM_SEARCH_TZ) return fail("No Such Search Mode %d: 0==Block, 1==Spiral, 2==Log, 3==Fast, 4==TZ", search_mode);
  if ((search_mode == JSVM_SEARCH_LOG || search_mode == JSVM_SEARCH_FAST) && full_pel_dfunc == JSVM_DF_YUV_SAD)
    return fail("Log and Fast search not possible in comb. with distortion measure 3 (ENC/CodingParameter.cpp:21; xPelLogSearch never sets the "
                "chroma search pointers, ENC/MotionEstimation.cpp:472-709)");
  if (full_pel_dfunc < 0 || full_pel_dfunc > 3) return fail("No Such Search Func (Full Pel) %d: 0==SAD, 1==SSE, 2==Hadamard, 3==YUV-SAD", full_pel_dfunc);
  if (sub_pel_dfunc < 0 || sub_pel_dfunc > 2) return fail("No Such Search Func (Sub Pel) %d: 0==SAD, 1==SSE, 2==Hadamard", sub_pel_dfunc);
  if (search_range < 1 || search_range > kMaxWideRange) return fail("SearchRange %d outside [1,%d]", search_range, kMaxWideRange);
  c->search_range = search_range; c->sub_dfunc = sub_pel_dfunc; c->full_dfunc = full_pel_dfunc; c->search_mode = search_mode; c->params_set = true;
  return JSVM_ME_OK;
}

int jsvm_me_set_fast_search_params(jsvm_me_ctx* c, int el_search_range, int fast_bi_search) {
  if (!c) return fail("null context");
  if (el_search_range < 0 || el_search_range > kMaxWideRange) return fail("ELSearchRange %d outside [0,%d]", el_search_range, kMaxWideRange);
  c->el_search_range = el_search_range; c->fast_bi_search = fast_bi_search != 0;
  return JSVM_ME_OK;
}

int jsvm_me_set_mv_candidates(jsvm_me_ctx* c, const jsvm_mv* abc, int n_jobs) {
  if (!c) return fail("null context");
  if (abc && n_jobs <= 0) return fail("n_jobs %d", n_jobs);
  c->cands = abc; c->n_cands = abc ? n_jobs : 0;
  return JSVM_ME_OK;
}

int jsvm_me_set_mv_candidates_device(jsvm_me_ctx* c, const jsvm_mv* d_abc) {
  if (!c) return fail("null context");
  c->d_cands_user = d_abc;
  return JSVM_ME_OK;
}

int jsvm_me_set_stream(jsvm_me_ctx* c, void* s) {
  if (!c) return fail("null context");
  if (c->stream != (cudaStream_t)s && !c->margin_dirty.empty()) {      // pending fillMargin launches stay behind their uploads, on the old stream
    CU_OK(cudaSetDevice(c->device));
    if (flush_margins(c)) return JSVM_ME_ERR;
  }
  c->stream = (cudaStream_t)s;
  return JSVM_ME_OK;
}

int jsvm_me_upload_plane_xpel(jsvm_me_ctx* c, int slot, const int16_t* origin, int width, int height, int stride) {
  if (check_slot(c, slot)) return JSVM_ME_ERR;
  if (!origin) return fail("null plane");
  if (width != c->W || height != c->H) return fail("plane is %dx%d but the context is configured for %dx%d", width, height, c->W, c->H);
  if (stride < c->plane_stride) return fail("stride %d smaller than width + 64", stride);
  CU_OK(cudaSetDevice(c->device));
  const size_t n16 = (size_t)c->plane_stride * c->plane_rows;
  if (c->d_stage.reserve(n16 * 2)) return JSVM_ME_ERR;
  const int16_t* top = origin - (ptrdiff_t)kMargin * stride - kMargin;
  CU_OK(cudaMemcpy2DAsync(c->d_stage.p, (size_t)c->plane_stride * 2, top, (size_t)stride * 2, (size_t)c->plane_stride * 2, c->plane_rows, cudaMemcpyHostToDevice, c->stream));
  CU_OK(cudaMemsetAsync(c->d_bad, 0, sizeof(int), c->stream));
  dim3 grid((c->plane_stride + 255) / 256, c->plane_rows), block(256);
  xpel_to_u8_kernel<<<grid, block, 0, c->stream>>>((const int16_t*)c->d_stage.p, c->plane_stride, c->d_planes + (size_t)slot * c->plane_bytes, c->plane_stride, c->plane_stride, c->plane_rows, c->d_bad);
  c->launches++;
  CU_OK(cudaGetLastError());
  int bad = 0;
  CU_OK(cudaMemcpyAsync(&bad, c->d_bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU_OK(cudaStreamSynchronize(c->stream));
  if (bad) return fail("plane holds %d samples outside [0,255]; the 8-bit search path cannot represent them", bad);
  c->margin_dirty.erase(std::remove(c->margin_dirty.begin(), c->margin_dirty.end(), slot), c->margin_dirty.end());   // the margins came with the plane
  c->slot_full[slot] = 1; c->slot_half[slot] = 0;
  return JSVM_ME_OK;
}

static int set_plane_u8(jsvm_me_ctx* c, int slot, const uint8_t* src, int width, int height, int stride, cudaMemcpyKind kind) {
  if (check_slot(c, slot)) return JSVM_ME_ERR;
  if (!src) return fail("null plane");
  if (width != c->W || height != c->H) return fail("plane is %dx%d but the context is configured for %dx%d", width, height, c->W, c->H);
  if (stride < width) return fail("stride %d smaller than width", stride);
  CU_OK(cudaSetDevice(c->device));
  uint8_t* dst = c->d_planes + (size_t)slot * c->plane_bytes + (size_t)kMargin * c->plane_stride + kMargin;
  CU_OK(cudaMemcpy2DAsync(dst, c->plane_stride, src, stride, width, height, kind, c->stream));
  return JSVM_ME_OK;
}

int jsvm_me_upload_plane_u8(jsvm_me_ctx* c, int slot, const uint8_t* host, int width, int height, int stride) {
  if (set_plane_u8(c, slot, host, width, height, stride, cudaMemcpyHostToDevice)) return JSVM_ME_ERR;
  mark_margin_dirty(c, slot);                                  // fillMargin runs in front of the first reader (flush_margins)
  c->slot_full[slot] = 1; c->slot_half[slot] = 0;
  return JSVM_ME_OK;
}

int jsvm_me_set_plane_device_u8(jsvm_me_ctx* c, int slot, const uint8_t* dev, int width, int height, int stride) {
  if (set_plane_u8(c, slot, dev, width, height, stride, cudaMemcpyDeviceToDevice)) return JSVM_ME_ERR;
  mark_margin_dirty(c, slot);
  c->slot_full[slot] = 1; c->slot_half[slot] = 0;
  return JSVM_ME_OK;
}

// ---- chroma planes (YUV_SAD search, motion-compensated prediction) ----
static int ensure_chroma(jsvm_me_ctx* c) {
  if (c->d_chroma) return JSVM_ME_OK;
  CU_OK(cudaMalloc((void**)&c->d_chroma, c->chroma_bytes * 2 * c->n_slots));
  CU_OK(cudaMemsetAsync(c->d_chroma, 0, c->chroma_bytes * 2 * c->n_slots, c->stream));
  return JSVM_ME_OK;
}

int jsvm_me_upload_chroma_u8(jsvm_me_ctx* c, int slot, const uint8_t* cb, const uint8_t* cr, int stride) {
  if (check_slot(c, slot)) return JSVM_ME_ERR;
  if (!cb || !cr) return fail("null chroma plane");
  const int cw = c->W / 2, ch = c->H / 2;
  if (stride < cw) return fail("stride %d smaller than the chroma width %d", stride, cw);
  CU_OK(cudaSetDevice(c->device));
  if (ensure_chroma(c)) return JSVM_ME_ERR;
  for (int comp = 0; comp < 2; ++comp) {
    uint8_t* plane = c->d_chroma + ((size_t)slot * 2 + comp) * c->chroma_bytes;
    CU_OK(cudaMemcpy2DAsync(plane + (size_t)kCMargin * c->chroma_stride + kCMargin, c->chroma_stride, comp ? cr : cb, stride, cw, ch,
                            cudaMemcpyHostToDevice, c->stream));
    dim3 grid(1, c->chroma_rows), block(256);
    plane_margin_fill_kernel<<<grid, block, 0, c->stream>>>(plane, c->chroma_stride, cw, ch, kCMargin);
    c->launches++;
    CU_OK(cudaGetLastError());
  }
  c->slot_chroma[slot] = 1;
  return JSVM_ME_OK;
}

int jsvm_me_upload_chroma_xpel(jsvm_me_ctx* c, int slot, const int16_t* cb_origin, const int16_t* cr_origin, int stride) {
  if (check_slot(c, slot)) return JSVM_ME_ERR;
  if (!cb_origin || !cr_origin) return fail("null chroma plane");
  if (stride < c->chroma_stride) return fail("stride %d smaller than width / 2 + 32", stride);
  CU_OK(cudaSetDevice(c->device));
  if (ensure_chroma(c)) return JSVM_ME_ERR;
  const size_t n16 = (size_t)c->chroma_stride * c->chroma_rows;
  if (c->d_stage.reserve(n16 * 2)) return JSVM_ME_ERR;
  CU_OK(cudaMemsetAsync(c->d_bad, 0, sizeof(int), c->stream));
  for (int comp = 0; comp < 2; ++comp) {
    const int16_t* top = (comp ? cr_origin : cb_origin) - (ptrdiff_t)kCMargin * stride - kCMargin;
    CU_OK(cudaMemcpy2DAsync(c->d_stage.p, (size_t)c->chroma_stride * 2, top, (size_t)stride * 2, (size_t)c->chroma_stride * 2, c->chroma_rows,
                            cudaMemcpyHostToDevice, c->stream));
    dim3 grid((c->chroma_stride + 255) / 256, c->chroma_rows), block(256);
    xpel_to_u8_kernel<<<grid, block, 0, c->stream>>>((const int16_t*)c->d_stage.p, c->chroma_stride,
                                                     c->d_chroma + ((size_t)slot * 2 + comp) * c->chroma_bytes, c->chroma_stride,
                                                     c->chroma_stride, c->chroma_rows, c->d_bad);
    c->launches++;
    CU_OK(cudaGetLastError());
  }
  int bad = 0;
  CU_OK(cudaMemcpyAsync(&bad, c->d_bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU_OK(cudaStreamSynchronize(c->stream));
  if (bad) return fail("chroma planes hold %d samples outside [0,255]", bad);
  c->slot_chroma[slot] = 1;
  return JSVM_ME_OK;
}

int jsvm_me_set_org_chroma(jsvm_me_ctx* c, const int16_t* org_chroma_mbs, int n_org_mbs) {
  if (!c) return fail("null context");
  if (org_chroma_mbs && n_org_mbs <= 0) return fail("n_org_mbs %d", n_org_mbs);
  c->org_chroma = org_chroma_mbs; c->n_org_chroma = org_chroma_mbs ? n_org_mbs : 0;
  return JSVM_ME_OK;
}

int jsvm_me_map_cost_factor(jsvm_me_ctx* c, uint32_t factor_full_pel, uint32_t factor_sub_pel) {
  if (!c) return fail("null context");
  for (uint2& e : c->factor_map)
    if (e.x == factor_full_pel) { if (e.y != factor_sub_pel) { e.y = factor_sub_pel; c->factor_map_dirty = true; } return JSVM_ME_OK; }
  if (c->factor_map.size() >= 256) return fail("more than 256 cost-factor pairs registered");
  c->factor_map.push_back(make_uint2(factor_full_pel, factor_sub_pel));
  c->factor_map_dirty = true;
  return JSVM_ME_OK;
}

// ---- K4: motion-compensated prediction ----
int jsvm_me_compensate_device(jsvm_me_ctx* c, const jsvm_mc_job* d_jobs, int n_jobs, uint8_t* d_pred_y, uint8_t* d_pred_cb, uint8_t* d_pred_cr) {
  if (!c || !d_jobs || !d_pred_y) return fail("null argument");
  if (!c->d_planes) return fail("jsvm_me_configure has not been called");
  if ((d_pred_cb == nullptr) != (d_pred_cr == nullptr)) return fail("pred_cb and pred_cr must both be given or both be NULL");
  if (d_pred_cb && !c->d_chroma) return fail("chroma prediction needs the chroma planes (jsvm_me_upload_chroma_*)");
  if (n_jobs <= 0) return JSVM_ME_OK;
  CU_OK(cudaSetDevice(c->device));
  begin_timed(c, 3);
  compensate_kernel<<<n_jobs, kK4Threads, 0, c->stream>>>(c->d_half, c->half_bytes, c->half_stride, c->d_chroma, c->chroma_bytes, c->chroma_stride,
                                                          d_jobs, n_jobs, c->W >> 4, c->H >> 4, d_pred_y, d_pred_cb, d_pred_cr);
  end_timed(c);
  c->launches++;
  CU_OK(cudaGetLastError());
  return JSVM_ME_OK;
}

int jsvm_me_compensate(jsvm_me_ctx* c, const jsvm_mc_job* jobs, int n_jobs, uint8_t* pred_y, uint8_t* pred_cb, uint8_t* pred_cr) {
  if (!c || !jobs || !pred_y) return fail("null argument");
  if (!c->d_planes) return fail("jsvm_me_configure has not been called");
  if ((pred_cb == nullptr) != (pred_cr == nullptr)) return fail("pred_cb and pred_cr must both be given or both be NULL");
  if (n_jobs <= 0) return JSVM_ME_OK;
  const int mbw = c->W >> 4, mbh = c->H >> 4;
  for (int i = 0; i < n_jobs; ++i) {
    const jsvm_mc_job& j = jobs[i];
    if (j.mb_x < 0 || j.mb_x >= mbw || j.mb_y < 0 || j.mb_y >= mbh) return fail("job %d: macroblock (%d,%d) outside the picture", i, j.mb_x, j.mb_y);
    if (j.blk > 15) return fail("job %d: blk %d out of range", i, (int)j.blk);
    const bool sx = j.size_x == 16 || j.size_x == 8 || j.size_x == 4, sy = j.size_y == 16 || j.size_y == 8 || j.size_y == 4;
    if (!sx || !sy || 4 * (j.blk & 3) + j.size_x > 16 || 4 * (j.blk >> 2) + j.size_y > 16)
      return fail("job %d: block %dx%d at blk %d does not fit the macroblock", i, (int)j.size_x, (int)j.size_y, (int)j.blk);
    if (j.ref_slot[0] < 0 && j.ref_slot[1] < 0) return fail("job %d: no reference list used", i);
    if (j.log_denom[0] > 7 || j.log_denom[1] > 7) return fail("job %d: log2 weight denominator above 7", i);
    for (int l = 0; l < 2; ++l) {
      if (j.ref_slot[l] < 0) continue;
      if (j.ref_slot[l] >= c->n_slots) return fail("job %d: ref_slot %d out of range", i, (int)j.ref_slot[l]);
      if (!c->slot_half[j.ref_slot[l]]) return fail("job %d: ref_slot %d has no half-pel plane", i, (int)j.ref_slot[l]);
      if (pred_cb && !c->slot_chroma[j.ref_slot[l]]) return fail("job %d: ref_slot %d has no chroma planes", i, (int)j.ref_slot[l]);
    }
  }
  CU_OK(cudaSetDevice(c->device));
  const size_t per_job = pred_cb ? 384 : 256;
  if (c->d_mc_jobs.reserve(n_jobs) || c->d_mc_pred.reserve((size_t)n_jobs * per_job)) return JSVM_ME_ERR;
  CU_OK(cudaMemcpyAsync(c->d_mc_jobs.p, jobs, sizeof(jsvm_mc_job) * n_jobs, cudaMemcpyHostToDevice, c->stream));
  uint8_t* dy = c->d_mc_pred.p;
  uint8_t* dcb = pred_cb ? dy + (size_t)n_jobs * 256 : nullptr;
  uint8_t* dcr = pred_cb ? dcb + (size_t)n_jobs * 64 : nullptr;
  if (jsvm_me_compensate_device(c, c->d_mc_jobs.p, n_jobs, dy, dcb, dcr)) return JSVM_ME_ERR;
  CU_OK(cudaMemcpyAsync(pred_y, dy, (size_t)n_jobs * 256, cudaMemcpyDeviceToHost, c->stream));
  if (pred_cb) {
    CU_OK(cudaMemcpyAsync(pred_cb, dcb, (size_t)n_jobs * 64, cudaMemcpyDeviceToHost, c->stream));
    CU_OK(cudaMemcpyAsync(pred_cr, dcr, (size_t)n_jobs * 64, cudaMemcpyDeviceToHost, c->stream));
  }
  CU_OK(cudaStreamSynchronize(c->stream));
  return JSVM_ME_OK;
}

int jsvm_me_interp_halfpel_batch(jsvm_me_ctx* c, const int* slots, int n) {
  if (!c) return fail("null context");
  if (n <= 0) return JSVM_ME_OK;
  if (n > c->n_slots) return fail("batch of %d slots exceeds n_slots %d", n, c->n_slots);
  for (int i = 0; i < n; ++i) {
    if (check_slot(c, slots[i])) return JSVM_ME_ERR;
    if (!c->slot_full[slots[i]]) return fail("slot %d holds no full-pel plane", slots[i]);
  }
  CU_OK(cudaSetDevice(c->device));
  if (flush_margins(c)) return JSVM_ME_ERR;
  CU_OK(cudaMemcpyAsync(c->d_slot_ids, slots, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
  dim3 grid((c->W + 2 * kMargin + kK1TileX - 1) / kK1TileX, (c->H + 2 * (kMargin - 4) + kK1TileY - 1) / kK1TileY, n);
  begin_timed(c, 0);
  halfpel_kernel<<<grid, kK1Threads, 0, c->stream>>>(c->map_k1, c->d_half, c->half_bytes, c->half_stride, c->d_slot_ids, c->W, c->H);
  end_timed(c);
  c->launches++;
  CU_OK(cudaGetLastError());
  for (int i = 0; i < n; ++i) c->slot_half[slots[i]] = 1;
  return JSVM_ME_OK;
}

int jsvm_me_interp_halfpel(jsvm_me_ctx* c, int slot) { return jsvm_me_interp_halfpel_batch(c, &slot, 1); }

int jsvm_me_download_halfpel_u8(jsvm_me_ctx* c, int slot, uint8_t* dst, int dst_stride) {
  if (check_slot(c, slot)) return JSVM_ME_ERR;
  if (!c->slot_half[slot]) return fail("slot %d holds no half-pel plane", slot);
  const int m = 2 * (kMargin - 4);
  const int rows = 2 * c->H + 2 * m, cols = 2 * c->W + 2 * m;
  if (dst_stride < cols) return fail("dst_stride %d smaller than %d", dst_stride, cols);
  CU_OK(cudaSetDevice(c->device));
  const uint8_t* src = c->d_half + (size_t)slot * c->half_bytes + (size_t)(kHpOrigin - m) * c->half_stride + (kHpOrigin - m);
  CU_OK(cudaMemcpy2DAsync(dst, dst_stride, src, c->half_stride, cols, rows, cudaMemcpyDeviceToHost, c->stream));
  CU_OK(cudaStreamSynchronize(c->stream));
  return JSVM_ME_OK;
}

int jsvm_me_download_halfpel_xpel(jsvm_me_ctx* c, int slot, int16_t* origin, int stride) {
  if (check_slot(c, slot)) return JSVM_ME_ERR;
  if (!c->slot_half[slot]) return fail("slot %d holds no half-pel plane", slot);
  const int m = 2 * (kMargin - 4);
  const int rows = 2 * c->H + 2 * m, cols = 2 * c->W + 2 * m;
  if (stride < cols) return fail("stride %d smaller than %d", stride, cols);
  CU_OK(cudaSetDevice(c->device));
  if (c->d_stage.reserve((size_t)rows * cols * 2)) return JSVM_ME_ERR;
  const uint8_t* src = c->d_half + (size_t)slot * c->half_bytes + (size_t)(kHpOrigin - m) * c->half_stride + (kHpOrigin - m);
  dim3 grid((cols + 255) / 256, rows), block(256);
  u8_to_xpel_kernel<<<grid, block, 0, c->stream>>>(src, c->half_stride, (int16_t*)c->d_stage.p, cols, cols, rows);
  c->launches++;
  CU_OK(cudaGetLastError());
  CU_OK(cudaMemcpy2DAsync(origin - (ptrdiff_t)m * stride - m, (size_t)stride * 2, c->d_stage.p, (size_t)cols * 2, (size_t)cols * 2, rows, cudaMemcpyDeviceToHost, c->stream));
  CU_OK(cudaStreamSynchronize(c->stream));
  return JSVM_ME_OK;
}

size_t jsvm_me_group_size(void) { return sizeof(Group); }

int jsvm_me_plan(jsvm_me_ctx* c, const jsvm_me_job* jobs, int n_jobs, void* groups_out, jsvm_me_plan_info* info) {
  if (!c || !jobs || !groups_out || !info) return fail("null argument");
  if (!c->d_planes) return fail("jsvm_me_configure has not been called");
  if (!c->params_set) return fail("jsvm_me_set_params has not been called");
  return jsvm_me_plan_host(c->W, c->H, c->n_slots, c->search_range, jobs, n_jobs, groups_out, info);
}

int jsvm_me_plan_host(int width, int height, int n_slots, int search_range, const jsvm_me_job* jobs, int n_jobs,
                      void* groups_out, jsvm_me_plan_info* info) {
  if (!jobs || !groups_out || !info) return fail("null argument");
  if (width <= 0 || height <= 0 || (width & 15) || (height & 15)) return fail("picture size %dx%d must be a positive multiple of 16", width, height);
  if (search_range < 1 || search_range > kMaxSearchRange) return fail("SearchRange %d outside [1,%d]", search_range, kMaxSearchRange);
  std::vector<Group> g;
  if (plan_jobs(width, height, n_slots, search_range, jobs, n_jobs, g, info)) return JSVM_ME_ERR;
  memcpy(groups_out, g.data(), g.size() * sizeof(Group));
  return JSVM_ME_OK;
}

int jsvm_me_search_device(jsvm_me_ctx* c, const jsvm_me_job* d_jobs, int n_jobs, const void* d_groups, const jsvm_me_plan_info* info,
                          const int16_t* d_org_mbs, int n_org_mbs, jsvm_me_result* d_results, uint8_t* d_preds) {
  (void)n_org_mbs;
  if (!c || !d_jobs || !d_results) return fail("null argument");
  const bool fast = c->search_mode != JSVM_SEARCH_BLOCK || c->search_range > kMaxSearchRange;     // these searches need no plan
  if (!fast && (!d_groups || !info)) return fail("null argument");
  if (n_jobs <= 0) return JSVM_ME_OK;
  CU_OK(cudaSetDevice(c->device));
  if (c->d_keys.reserve(n_jobs) || c->d_full_costs.reserve(n_jobs) || c->d_aux.reserve(n_jobs)) return JSVM_ME_ERR;
  jsvm_me_plan_info none{};
  return run_search(c, d_jobs, 0, n_jobs, (const Group*)d_groups, info ? info : &none, d_org_mbs, c->d_keys.p, c->d_aux.p, d_results, d_preds,
                    nullptr, c->d_cands_user);
}

int jsvm_me_search(jsvm_me_ctx* c, const jsvm_me_job* jobs, int n_jobs, const int16_t* org_mbs, int n_org_mbs,
                   jsvm_me_result* results, uint8_t* preds) {
  if (!c || !jobs || !results) return fail("null argument");
  if (!c->d_planes) return fail("jsvm_me_configure has not been called");
  if (!c->params_set) return fail("jsvm_me_set_params has not been called");
  if (n_jobs <= 0) return JSVM_ME_OK;
  c->last_jobs.clear(); c->last_has_preds = false;     // a failed call leaves no prediction to fetch (compensateBlock contract)
  if (n_org_mbs > 0) {
    if (!org_mbs) return fail("org_mbs is null");
    for (size_t i = 0; i < (size_t)n_org_mbs * 256; ++i)
      if (org_mbs[i] < -1024 || org_mbs[i] > 1023)      // residual-prediction originals span [-255, 510] (N3); the key widths assume |v| < 1024
        return fail("org_mbs sample %zu = %d outside [-1024,1023]", i, (int)org_mbs[i]);
  }
  if (n_jobs <= 64 && n_org_mbs <= 64) return search_small(c, jobs, n_jobs, org_mbs, n_org_mbs, results, preds);
  if (c->search_mode != JSVM_SEARCH_BLOCK || c->search_range > kMaxSearchRange) return search_fast_host(c, jobs, n_jobs, org_mbs, n_org_mbs, results, preds);
  c->last_small = false;
  CU_OK(cudaSetDevice(c->device));
  // every error exit below: nothing may still be copying into the caller's buffers, and no prediction of an earlier call may be served
  struct ExitGuard {
    jsvm_me_ctx* c; bool ok = false;
    ~ExitGuard() {
      if (ok) return;
      cudaStreamSynchronize(c->stream);
      if (c->up_stream) cudaStreamSynchronize(c->up_stream);
      if (c->down_stream) cudaStreamSynchronize(c->down_stream);
      c->last_jobs.clear(); c->last_has_preds = false;
    }
  } guard{c};
  const auto t_entry = std::chrono::steady_clock::now();
  const bool want_preds = preds != nullptr || n_jobs <= 4096;
  if (c->d_jobs.reserve(n_jobs) || c->d_keys.reserve(n_jobs) || c->d_full_costs.reserve(n_jobs) || c->d_aux.reserve(n_jobs) ||
      c->d_results.reserve(n_jobs) || (want_preds && c->d_preds.reserve((size_t)n_jobs * 256)) ||
      (n_org_mbs > 0 && c->d_org.reserve((size_t)n_org_mbs * 256))) return JSVM_ME_ERR;
  if (n_org_mbs > 0) CU_OK(cudaMemcpyAsync(c->d_org.p, org_mbs, sizeof(int16_t) * 256 * n_org_mbs, cudaMemcpyHostToDevice, c->stream));
  const int16_t* d_org_chroma = nullptr;
  if (prepare_search_extras(c, jobs, n_jobs, n_org_mbs, &d_org_chroma)) return JSVM_ME_ERR;
  if (!c->up_stream) {
    CU_OK(cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking));
    CU_OK(cudaStreamCreateWithFlags(&c->down_stream, cudaStreamNonBlocking));
    for (cudaEvent_t* e : {&c->ev_entry, &c->ev_kernels}) CU_OK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  }
  // the side streams start after whatever the caller queued on the ctx stream before this call (plane uploads, K1)
  CU_OK(cudaEventRecord(c->ev_entry, c->stream));
  CU_OK(cudaStreamWaitEvent(c->up_stream, c->ev_entry, 0));
  CU_OK(cudaStreamWaitEvent(c->down_stream, c->ev_entry, 0));
  // Pipeline.  PLANNING happens in units of ~16 K jobs cut at macroblock boundaries, by worker threads, in order.
  // LAUNCHES take every unit that is ready, up to kMaxUnits of them, as one batch: the calling thread never waits for more than
  // the one unit it needs next, and once the planners are ahead of the GPU the launches grow to 128 K jobs (a persistent K2
  // launch loses its pipeline fill and its last-wave imbalance once per launch: ~1 group of ~12 per SM at 16 K jobs, of ~100
  // at 128 K).  Job / group uploads and result copy-backs run on their own streams next to the kernels.
  int kUnit = 16384;
  if (const char* e = getenv("JSVM_ME_CHUNK")) kUnit = std::max(1024, std::min(1 << 17, atoi(e)));     // tuning knob (jobs per planning unit)
  constexpr int kMaxUnits = 8;
  const bool merge_launches = getenv("JSVM_ME_NO_MERGE") == nullptr;       // A/B knob
  std::vector<int> cut(1, 0);
  while (cut.back() < n_jobs) {
    // the first units are small (2 K, 4 K, 8 K jobs): the GPU has nothing to do until unit 0 is planned and uploaded
    const int size = std::min(kUnit, 2048 << std::min<size_t>(cut.size() - 1, 6));
    int end = std::min(n_jobs, cut.back() + size);
    while (end < n_jobs && same_mb(jobs[end], jobs[end - 1])) ++end;
    cut.push_back(end);
  }
  const int n_units = (int)cut.size() - 1;
  if (c->d_groups.reserve((size_t)n_jobs)) return JSVM_ME_ERR;                      // upper bound: one group per job
  const size_t ring_cap = (size_t)kMaxUnits * ((size_t)kUnit + 256);      // a unit never holds more groups than jobs (+ the MB it finishes)
  if (c->pinned_cap < ring_cap) {
    for (auto& pg : c->pinned_groups) { if (pg) cudaFreeHost(pg); pg = nullptr; }
    for (auto& pg : c->pinned_groups) CU_OK(cudaMallocHost((void**)&pg, ring_cap * sizeof(Group)));
    c->pinned_cap = ring_cap;
  }
  std::vector<PlanChunk> plans(n_units);
  std::vector<int> status(n_units, 0);                      // 0 = pending, 1 = planned, -1 = failed
  std::mutex mu; std::condition_variable cv;
  // Planner threads: one thread plans 2.2x faster than the GPU searches (150 vs 67 jobs/us on C4), so a few are enough; a thread per
  // core (16 on the measured box) starves this thread and the driver's copy threads: 19.4 instead of 19.0 ms per 16-pair call.
  unsigned max_workers = 4;
  if (const char* e = getenv("JSVM_ME_PLAN_THREADS")) max_workers = (unsigned)std::max(1, std::min(64, atoi(e)));      // tuning knob
  const int n_workers = n_units <= 1 ? 0 : (int)std::min<unsigned>((unsigned)n_units, std::min(max_workers, std::max(1u, std::thread::hardware_concurrency())));
  auto plan_unit = [&](int k) {
    plan_range(c->W, c->H, c->n_slots, c->search_range, jobs, cut[k], cut[k + 1], plans[k]);
    if (getenv("JSVM_ME_TRACE") && (k < 12 || k % 8 == 0))
      fprintf(stderr, "[jsvm_me trace] host %8.3f ms: unit %d planned\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count(), k);
    { std::lock_guard<std::mutex> lk(mu); status[k] = plans[k].err_job >= 0 ? -1 : 1; }
    cv.notify_all();
  };
  std::vector<std::thread> workers;
  for (int t = 0; t < n_workers; ++t)
    workers.emplace_back([&, t]() { for (int k = t; k < n_units; k += n_workers) plan_unit(k); });
  struct Joiner { std::vector<std::thread>& w; ~Joiner() { for (auto& x : w) if (x.joinable()) x.join(); } } joiner{workers};

  size_t groups_done = 0;
  int launch = 0;
  for (int k = 0; k < n_units; ++launch) {
    if (n_workers == 0) plan_unit(k);
    int m = k;                                              // units [k, m) go into this launch
    {
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] { return status[k] != 0; });
      // (the launch list below holds the 8x8-granular groups of all its units first, then the 4x4-granular ones)
      const bool merge_g4 = getenv("JSVM_ME_NO_MERGE_G4") == nullptr;       // A/B knob
      auto take_ready = [&] {
        while (m < n_units && m - k < kMaxUnits && status[m] != 0 && (m == k || merge_g4 || (plans[m].g4.empty() && plans[k].g4.empty()))) {
          if (status[m] < 0) break;
          ++m;
        }
      };
      take_ready();
      // While the GPU still has work queued on the ctx stream (the plane uploads and K1 in front of the first launch, the previous
      // launch later) there is no hurry: keep collecting planned units, so that the launch is larger (every persistent K2 launch pays
      // its pipeline fill and its last-wave imbalance once).
      if (merge_launches && n_workers > 0)
        while (m > k && m < n_units && m - k < kMaxUnits && status[m] == 0 && (merge_g4 || plans[k].g4.empty())) {
          lk.unlock();
          const bool busy = cudaStreamQuery(c->stream) == cudaErrorNotReady;
          lk.lock();
          if (!busy) break;
          cv.wait_for(lk, std::chrono::microseconds(30), [&] { return status[m] != 0; });
          take_ready();
        }
      if (m == k) m = k + 1;                                // unit k failed: reported below
    }
    for (int u = k; u < m; ++u)
      if (status[u] < 0) return fail("job %d: %s", plans[u].err_job, plans[u].err);
    const int begin = cut[k], end = cut[m];
    if (getenv("JSVM_ME_TRACE"))
      fprintf(stderr, "[jsvm_me trace] host %8.3f ms: launch %d takes units %d..%d (jobs %d..%d)\n",
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count(), launch, k, m - 1, begin, end);
    Group* pg = c->pinned_groups[launch & 3];
    if (launch >= 4) CU_OK(cudaEventSynchronize(c->chunk_done[launch & 3]));        // the slot's previous upload has been consumed
    jsvm_me_plan_info info{};
    size_t ng = 0;
    for (int pass = 0; pass < 2; ++pass)                    // all 8x8-granular groups first, then the 4x4-granular ones
      for (int u = k; u < m; ++u) {
        const std::vector<Group>& v = pass == 0 ? plans[u].g8 : plans[u].g4;
        if (ng + v.size() > ring_cap) return fail("internal: launch plan larger than its ring slot");
        for (const Group& g : v) {                          // one check per group: planes present, originals in range
          if (g.org_index >= n_org_mbs) return fail("job %d: org_index %d >= n_org_mbs %d", g.first_job, g.org_index, n_org_mbs);
          if (g.org_index < 0 && !c->slot_full[g.cur_slot]) return fail("job %d: cur_slot %d is empty", g.first_job, g.cur_slot);
          if (!c->slot_half[g.ref_slot]) return fail("job %d: ref_slot %d has no half-pel plane", g.first_job, g.ref_slot);
          int32_t& mx = g.gran4 ? info.max_uh_4x4 : info.max_uh_8x8;
          mx = std::max<int32_t>(mx, g.uh);
        }
        if (!v.empty()) memcpy(pg + ng, v.data(), sizeof(Group) * v.size());
        ng += v.size();
        if (pass == 0) info.n_groups_8x8 += (int32_t)v.size();
        info.n_positions += pass == 0 ? plans[u].npos : 0;
      }
    info.n_groups = (int32_t)ng;
    for (int u = k; u < m; ++u) { std::vector<Group>().swap(plans[u].g8); std::vector<Group>().swap(plans[u].g4); }
    CU_OK(cudaMemcpyAsync(c->d_jobs.p + begin, jobs + begin, sizeof(jsvm_me_job) * (end - begin), cudaMemcpyHostToDevice, c->up_stream));
    CU_OK(cudaMemcpyAsync(c->d_groups.p + groups_done, pg, sizeof(Group) * ng, cudaMemcpyHostToDevice, c->up_stream));
    if (!c->chunk_done[launch & 3]) CU_OK(cudaEventCreateWithFlags(&c->chunk_done[launch & 3], cudaEventDisableTiming));
    CU_OK(cudaEventRecord(c->chunk_done[launch & 3], c->up_stream));                // ring slot consumed; the batch is on the device
    CU_OK(cudaStreamWaitEvent(c->stream, c->chunk_done[launch & 3], 0));
    if (run_search(c, c->d_jobs.p, begin, end, c->d_groups.p + groups_done, &info, c->d_org.p, c->d_keys.p, c->d_aux.p, c->d_results.p,
                   want_preds ? c->d_preds.p : nullptr, d_org_chroma)) return JSVM_ME_ERR;
    CU_OK(cudaEventRecord(c->ev_kernels, c->stream));
    CU_OK(cudaStreamWaitEvent(c->down_stream, c->ev_kernels, 0));
    CU_OK(cudaMemcpyAsync(results + begin, c->d_results.p + begin, sizeof(jsvm_me_result) * (end - begin), cudaMemcpyDeviceToHost, c->down_stream));
    groups_done += ng;
    k = m;
  }
  if (preds) CU_OK(cudaMemcpyAsync(preds, c->d_preds.p, (size_t)n_jobs * 256, cudaMemcpyDeviceToHost, c->stream));
  if (getenv("JSVM_ME_TRACE"))
    fprintf(stderr, "[jsvm_me trace] host %8.3f ms: all launches queued\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count());
  CU_OK(cudaStreamSynchronize(c->stream));
  CU_OK(cudaStreamSynchronize(c->down_stream));
  if (getenv("JSVM_ME_TRACE"))
    fprintf(stderr, "[jsvm_me trace] host %8.3f ms: results on the host\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count());
  if (want_preds && n_jobs <= 4096) c->last_jobs.assign(jobs, jobs + n_jobs); else c->last_jobs.clear();
  c->last_has_preds = want_preds && n_jobs <= 4096;
  guard.ok = true;
  return JSVM_ME_OK;
}

int jsvm_me_fetch_pred(jsvm_me_ctx* c, int job_idx, int16_t* dst, int dst_stride) {
  if (!c || !dst) return fail("null argument");
  if (job_idx < 0 || job_idx >= (int)c->last_jobs.size() || !c->last_has_preds)
    return fail("job index %d: no prediction block kept from the last jsvm_me_search call (kept for calls of up to 4096 jobs)", job_idx);
  int w, h;
  mode_size(c->last_jobs[job_idx].mode, w, h);
  uint8_t blk[256];
  if (c->last_small) memcpy(blk, c->small->pred + (size_t)job_idx * 256, 256);
  else {
    CU_OK(cudaSetDevice(c->device));
    CU_OK(cudaMemcpyAsync(blk, c->d_preds.p + (size_t)job_idx * 256, 256, cudaMemcpyDeviceToHost, c->stream));
    CU_OK(cudaStreamSynchronize(c->stream));
  }
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) dst[y * dst_stride + x] = blk[y * 16 + x];
  return JSVM_ME_OK;
}

// ---- K6: whole pictures, MV prediction and macroblock scheduling on the device ----
int jsvm_me_estimate_pictures(jsvm_me_ctx* c, const jsvm_pic_task* tasks, int n_pics, jsvm_pic_mb* out) {
  if (!c || !tasks) return fail("null argument");
  if (!c->d_planes) return fail("jsvm_me_configure has not been called");
  if (!c->params_set) return fail("jsvm_me_set_params has not been called");
  if (c->search_mode != JSVM_SEARCH_BLOCK || c->full_dfunc != JSVM_DF_SAD || c->search_range > kMaxSearchRange || c->fast_bi_search)
    return fail("jsvm_me_estimate_pictures needs SearchMode 0, SearchFuncFullPel 0, SearchRange <= %d, FastBiSearch 0", kMaxSearchRange);
  if (mixed_factor_classes(c)) return fail("jsvm_me_estimate_pictures: one cost factor per picture (SearchFuncSubPel 1 needs the SSE factor)");
  if (n_pics <= 0) return JSVM_ME_OK;
  const int R = tasks[0].n_refs;
  if (R < 1 || R > JSVM_PIC_MAX_REFS) return fail("n_refs %d outside [1,%d]", R, JSVM_PIC_MAX_REFS);
  for (int p = 0; p < n_pics; ++p) {
    const jsvm_pic_task& t = tasks[p];
    if (t.n_refs != R) return fail("picture %d: n_refs %d differs from picture 0 (%d)", p, (int)t.n_refs, R);
    if (t.cur_slot < 0 || t.cur_slot >= c->n_slots || !c->slot_full[t.cur_slot]) return fail("picture %d: cur_slot %d is empty or out of range", p, (int)t.cur_slot);
    for (int r = 0; r < R; ++r)
      if (t.ref_slot[r] < 0 || t.ref_slot[r] >= c->n_slots || !c->slot_half[t.ref_slot[r]])
        return fail("picture %d: ref_slot %d has no half-pel plane or is out of range", p, (int)t.ref_slot[r]);
  }
  CU_OK(cudaSetDevice(c->device));
  if (flush_margins(c)) return JSVM_ME_ERR;
  const int mbw = c->W >> 4, mbh = c->H >> 4, n_mb = mbw * mbh;
  if (c->wave_mbs.empty()) {
    const int n_waves = mbw + 2 * (mbh - 1);
    c->wave_off.assign(n_waves + 1, 0);
    for (int y = 0; y < mbh; ++y) for (int x = 0; x < mbw; ++x) c->wave_off[x + 2 * y + 1]++;
    for (int w = 0; w < n_waves; ++w) c->wave_off[w + 1] += c->wave_off[w];
    c->wave_mbs.resize(n_mb);
    std::vector<int> fill(c->wave_off.begin(), c->wave_off.end() - 1);
    for (int y = 0; y < mbh; ++y) for (int x = 0; x < mbw; ++x) c->wave_mbs[fill[x + 2 * y]++] = make_int2(x, y);
    if (c->d_wave_mbs.reserve(n_mb)) return JSVM_ME_ERR;
    CU_OK(cudaMemcpyAsync(c->d_wave_mbs.p, c->wave_mbs.data(), sizeof(int2) * n_mb, cudaMemcpyHostToDevice, c->stream));
    CU_OK(cudaStreamSynchronize(c->stream));
  }
  const int n_waves = (int)c->wave_off.size() - 1;
  int max_wave = 0;
  for (int w = 0; w < n_waves; ++w) max_wave = std::max(max_wave, c->wave_off[w + 1] - c->wave_off[w]);
  const size_t max_jobs = (size_t)n_pics * max_wave * R * 4;
  if (c->d_pic_tasks.reserve(n_pics) || c->d_pic_field.reserve((size_t)n_pics * n_mb) || c->d_pic_temp.reserve((size_t)n_pics * n_mb) ||
      c->d_pic_results[0].reserve(max_jobs) || c->d_pic_results[1].reserve(max_jobs) || c->d_jobs.reserve(max_jobs) || c->d_groups.reserve(max_jobs) ||
      c->d_keys.reserve(max_jobs) || c->d_aux.reserve(max_jobs)) return JSVM_ME_ERR;
  CU_OK(cudaMemcpyAsync(c->d_pic_tasks.p, tasks, sizeof(jsvm_pic_task) * n_pics, cudaMemcpyHostToDevice, c->stream));
  if (sync_factor_map(c)) return JSVM_ME_ERR;
  // union windows of a step: at most the window of one job plus 32 rows of spread between the partitions' predictors
  const int cap_uh = std::min(kMaxUH, 2 * c->search_range + 1 + 32);
  const size_t smem = k2_smem_bytes<kV8, false>(cap_uh);
  if (raise_dyn_smem(c->device, 0, full_search_kernel<kV8, false>, smem)) return JSVM_ME_ERR;
  {
    // the single-8x8 variant for steps 3 and 4 (its own function, hence its own shared-memory attribute)
    static std::mutex mu; static size_t configured[kMaxDevices] = {};
    std::lock_guard<std::mutex> lk(mu);
    if (smem > configured[c->device % kMaxDevices]) {
      CU_OK(cudaFuncSetAttribute(full_search_kernel<kV8, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured[c->device % kMaxDevices] = smem;
    }
  }
  auto k3 = c->sub_dfunc == JSVM_DF_HADAMARD ? subpel_kernel<JSVM_DF_HADAMARD>
          : c->sub_dfunc == JSVM_DF_SSD ? subpel_kernel<JSVM_DF_SSD> : subpel_kernel<JSVM_DF_SAD>;
  // The wave loop: 13 small dependent launches per wave.  Issued directly when kernel timing is on (events cannot be timed inside a
  // graph); otherwise captured ONCE into a CUDA graph on a private stream and replayed -- the launch sequence is a pure function of
  // the picture size, the batch size and the buffer addresses (PicKey).
  auto issue = [&](cudaStream_t stream) -> int {
    int flip = 0;
    for (int w = 0; w < n_waves; ++w) {
      const int n_wave = c->wave_off[w + 1] - c->wave_off[w];
      const int n_thr = n_pics * n_wave;
      for (int step = 0; step <= 4; ++step) {
        picture_step_kernel<<<(n_thr + 127) / 128, 128, 0, stream>>>(
            c->d_pic_tasks.p, n_pics, mbw, mbh, c->d_wave_mbs.p + c->wave_off[w], n_wave, step, R, c->search_range, cap_uh,
            c->d_pic_field.p, c->d_pic_temp.p, c->d_pic_results[flip ^ 1].p, c->d_jobs.p, c->d_groups.p);
        c->launches++;
        if (step == 4) break;
        const int n_jobs = n_thr * R * pic_step_jobs(step);
        begin_timed(c, 1);
        if (step < 2)
          full_search_kernel<kV8, false><<<n_jobs, kK2Threads, smem, stream>>>(
              c->d_planes, c->plane_bytes, c->plane_stride, c->d_groups.p, c->d_jobs.p, nullptr, mbw, mbh, c->search_range, c->d_keys.p, c->d_aux.p);
        else          // steps 3 and 4: one 8x8 job per group
          full_search_kernel<kV8, false, true><<<n_jobs, kK2Threads, smem, stream>>>(
              c->d_planes, c->plane_bytes, c->plane_stride, c->d_groups.p, c->d_jobs.p, nullptr, mbw, mbh, c->search_range, c->d_keys.p, c->d_aux.p);
        end_timed(c);
        begin_timed(c, 2);
        k3<<<(n_jobs + kK3WarpsPerCta - 1) / kK3WarpsPerCta, 32 * kK3WarpsPerCta, 0, stream>>>(
            c->d_planes, c->plane_bytes, c->plane_stride, c->d_half, c->half_bytes, c->half_stride,
            c->d_jobs.p, c->d_aux.p, n_jobs, nullptr, c->d_keys.p, nullptr, c->d_pic_results[flip].p, nullptr, c->d_factor_map.p, 0);
        end_timed(c);
        c->launches += 2;
        flip ^= 1;
      }
    }
    CU_OK(cudaGetLastError());
    return JSVM_ME_OK;
  };
  if (c->timing || getenv("JSVM_ME_NO_GRAPH")) {
    if (issue(c->stream)) return JSVM_ME_ERR;
  } else {
    jsvm_me_ctx::PicKey key{n_pics, R, c->sub_dfunc, c->search_range,
                            {c->d_pic_tasks.p, c->d_pic_field.p, c->d_pic_temp.p, c->d_pic_results[0].p, c->d_pic_results[1].p, c->d_jobs.p, c->d_groups.p, c->d_keys.p}};
    if (!c->pic_stream) {
      CU_OK(cudaStreamCreateWithFlags(&c->pic_stream, cudaStreamNonBlocking));
      for (auto& e : c->pic_ev) CU_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    if (!c->pic_graph || memcmp(&key, &c->pic_key, sizeof(key)) != 0 || c->d_aux.p != c->pic_aux) {
      if (c->pic_graph) { cudaGraphExecDestroy(c->pic_graph); c->pic_graph = nullptr; }
      const uint64_t launches_before = c->launches;
      cudaGraph_t graph = nullptr;
      CU_OK(cudaStreamBeginCapture(c->pic_stream, cudaStreamCaptureModeThreadLocal));
      const int rc = issue(c->pic_stream);
      const cudaError_t e = cudaStreamEndCapture(c->pic_stream, &graph);
      if (rc || e != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return rc ? JSVM_ME_ERR : fail("cudaStreamEndCapture failed: %s", cudaGetErrorString(e)); }
      const cudaError_t ei = cudaGraphInstantiate(&c->pic_graph, graph, 0);
      cudaGraphDestroy(graph);
      if (ei != cudaSuccess) { c->pic_graph = nullptr; return fail("cudaGraphInstantiate failed: %s", cudaGetErrorString(ei)); }
      c->pic_key = key; c->pic_aux = c->d_aux.p;
      c->pic_launches = c->launches - launches_before;
      c->launches = launches_before;                         // capturing launched nothing
    }
    CU_OK(cudaEventRecord(c->pic_ev[0], c->stream));
    CU_OK(cudaStreamWaitEvent(c->pic_stream, c->pic_ev[0], 0));
    CU_OK(cudaGraphLaunch(c->pic_graph, c->pic_stream));
    CU_OK(cudaEventRecord(c->pic_ev[1], c->pic_stream));
    CU_OK(cudaStreamWaitEvent(c->stream, c->pic_ev[1], 0));
    c->launches += c->pic_launches;                          // kernels the graph runs
  }
  c->last_k2 = "full_search_kernel<8,false>";
  c->pic_n_last = n_pics;
  if (out) return jsvm_me_download_pictures(c, out, n_pics);
  return JSVM_ME_OK;
}

int jsvm_me_download_pictures(jsvm_me_ctx* c, jsvm_pic_mb* out, int n_pics) {
  if (!c || !out) return fail("null argument");
  if (n_pics <= 0 || n_pics > c->pic_n_last) return fail("n_pics %d: the last jsvm_me_estimate_pictures call covered %d pictures", n_pics, c->pic_n_last);
  CU_OK(cudaSetDevice(c->device));
  CU_OK(cudaMemcpyAsync(out, c->d_pic_field.p, sizeof(jsvm_pic_mb) * (size_t)n_pics * (c->W >> 4) * (c->H >> 4), cudaMemcpyDeviceToHost, c->stream));
  CU_OK(cudaStreamSynchronize(c->stream));
  return JSVM_ME_OK;
}

// ---- K7: inter-layer resampling ----
namespace {
int ceil_log2(int i) { int s = 0; --i; while (i > 0) { ++s; i >>= 1; } return s; }          // include/ResizeParameters.h:30-39

// xCompResidualUpsampling (COM/DownConvert.cpp:2733-2817) for progressive layers: the arguments of xBasicResidualUpsampling
ResamplePlane residual_plane(const jsvm_resize_params& rp, bool chroma) {
  const int f = chroma ? 2 : 1;
  const int ref_phase_x = chroma ? rp.ref_chroma_phase_x : 0, ref_phase_y = chroma ? rp.ref_chroma_phase_y : 0;
  const int phase_x = chroma ? rp.chroma_phase_x : 0, phase_y = chroma ? rp.chroma_phase_y : 0;
  const int ref_w = rp.ref_frame_width / f, ref_h = rp.ref_frame_height / f;
  const int out_w = rp.scaled_ref_width / f, out_h = rp.scaled_ref_height / f;
  ResamplePlane P;
  P.base_w = ref_w; P.base_h = ref_h; P.cur_w = rp.frame_width / f; P.cur_h = rp.frame_height / f;
  P.l_off = rp.left_offset / f; P.t_off = rp.top_offset / f;
  P.r_off = P.cur_w - P.l_off - out_w; P.b_off = P.cur_h - P.t_off - out_h;
  const int shift_x = rp.level_idc <= 30 ? 16 : 31 - ceil_log2(ref_w), shift_y = rp.level_idc <= 30 ? 16 : 31 - ceil_log2(ref_h);
  P.shift_x_m4 = shift_x - 4; P.shift_y_m4 = shift_y - 4;
  P.scale_x = (int)((((uint32_t)ref_w << shift_x) + (uint32_t)(out_w >> 1)) / (uint32_t)out_w);
  P.scale_y = (int)((((uint32_t)ref_h << shift_y) + (uint32_t)(out_h >> 1)) / (uint32_t)out_h);
  P.offset_x = P.l_off; P.offset_y = P.t_off;
  P.add_x = (((ref_w * (2 + phase_x)) << (shift_x - 2)) + (out_w >> 1)) / out_w + (1 << (shift_x - 5));
  P.add_y = (((ref_h * (2 + phase_y)) << (shift_y - 2)) + (out_h >> 1)) / out_h + (1 << (shift_y - 5));
  P.delta_x = 4 * (2 + ref_phase_x); P.delta_y = 4 * (2 + ref_phase_y);
  P.mb_size = chroma ? 8 : 16; P.mbs_per_row = rp.ref_frame_width / 16; P.chroma = chroma ? 1 : 0;
  return P;
}
}  // namespace

int jsvm_me_upsample_residual(jsvm_me_ctx* c, const jsvm_resize_params* rp, const int16_t* base_y, const int16_t* base_cb, const int16_t* base_cr,
                              const uint8_t* base_transform8x8, int16_t* out_y, int16_t* out_cb, int16_t* out_cr) {
  if (!c || !rp || !base_y || !base_transform8x8 || !out_y) return fail("null argument");
  const bool chroma = base_cb || base_cr || out_cb || out_cr;
  if (chroma && !(base_cb && base_cr && out_cb && out_cr)) return fail("chroma planes: all four pointers or none");
  if (rp->frame_width <= 0 || rp->frame_height <= 0 || (rp->frame_width & 15) || (rp->frame_height & 15) ||
      rp->ref_frame_width <= 0 || rp->ref_frame_height <= 0 || (rp->ref_frame_width & 15) || (rp->ref_frame_height & 15))
    return fail("layer sizes must be positive multiples of 16");
  if (rp->scaled_ref_width <= 0 || rp->scaled_ref_height <= 0 || rp->left_offset < 0 || rp->top_offset < 0 ||
      rp->left_offset + rp->scaled_ref_width > rp->frame_width || rp->top_offset + rp->scaled_ref_height > rp->frame_height ||
      (rp->scaled_ref_width & 1) || (rp->scaled_ref_height & 1) || (rp->left_offset & 1) || (rp->top_offset & 1))
    return fail("scaled reference window %dx%d at (%d,%d) does not fit the %dx%d frame (even sizes and offsets)", rp->scaled_ref_width,
                rp->scaled_ref_height, rp->left_offset, rp->top_offset, rp->frame_width, rp->frame_height);
  CU_OK(cudaSetDevice(c->device));
  const size_t n_src = (size_t)rp->ref_frame_width * rp->ref_frame_height, n_dst = (size_t)rp->frame_width * rp->frame_height;
  const size_t n_mb = n_src / 256;
  if (c->d_rs_src.reserve(n_src * 3 / 2) || c->d_rs_dst.reserve(n_dst * 3 / 2) || c->d_rs_t8x8.reserve(n_mb)) return JSVM_ME_ERR;
  CU_OK(cudaMemcpyAsync(c->d_rs_t8x8.p, base_transform8x8, n_mb, cudaMemcpyHostToDevice, c->stream));
  const int16_t* srcs[3] = { base_y, base_cb, base_cr };
  int16_t* dsts[3] = { out_y, out_cb, out_cr };
  size_t so = 0, dof = 0;
  for (int k = 0; k < (chroma ? 3 : 1); ++k) {
    const ResamplePlane P = residual_plane(*rp, k > 0);
    const size_t ns = (size_t)P.base_w * P.base_h, nd = (size_t)P.cur_w * P.cur_h;
    CU_OK(cudaMemcpyAsync(c->d_rs_src.p + so, srcs[k], ns * sizeof(int16_t), cudaMemcpyHostToDevice, c->stream));
    dim3 grid((P.cur_w + 255) / 256, P.cur_h), block(256);
    begin_timed(c, 3);
    residual_upsample_kernel<<<grid, block, 0, c->stream>>>(P, c->d_rs_src.p + so, c->d_rs_t8x8.p, c->d_rs_dst.p + dof);
    end_timed(c);
    c->launches++;
    CU_OK(cudaGetLastError());
    CU_OK(cudaMemcpyAsync(dsts[k], c->d_rs_dst.p + dof, nd * sizeof(int16_t), cudaMemcpyDeviceToHost, c->stream));
    so += ns; dof += nd;
  }
  CU_OK(cudaStreamSynchronize(c->stream));
  return JSVM_ME_OK;
}

int jsvm_me_upsample_motion(jsvm_me_ctx* c, const jsvm_motion_upsample_params* p, const jsvm_mb_motion* base, jsvm_mb_motion* out) {
  if (!c || !p || !base || !out) return fail("null argument");
  const jsvm_resize_params& r = p->resize;
  if (r.frame_width <= 0 || r.frame_height <= 0 || (r.frame_width & 15) || (r.frame_height & 15) ||
      r.ref_frame_width <= 0 || r.ref_frame_height <= 0 || (r.ref_frame_width & 15) || (r.ref_frame_height & 15))
    return fail("layer sizes must be positive multiples of 16");
  if (r.scaled_ref_width <= 0 || r.scaled_ref_height <= 0 || r.left_offset < 0 || r.top_offset < 0 ||
      r.left_offset + r.scaled_ref_width > r.frame_width || r.top_offset + r.scaled_ref_height > r.frame_height)
    return fail("scaled reference window %dx%d at (%d,%d) does not fit the %dx%d frame", r.scaled_ref_width, r.scaled_ref_height, r.left_offset,
                r.top_offset, r.frame_width, r.frame_height);
  if (p->slice_type < 0 || p->slice_type > 2) return fail("slice_type %d: 0 P, 1 B, 2 I", p->slice_type);
  const int nb = (r.ref_frame_width >> 4) * (r.ref_frame_height >> 4), nc = (r.frame_width >> 4) * (r.frame_height >> 4);
  for (int i = 0; i < nb; ++i) {
    const jsvm_mb_motion& b = base[i];
    if (b.mb_mode >= 6) continue;
    bool ok = b.slice_type <= 2;
    for (int k = 0; k < 4; ++k) ok = ok && (b.blk_mode[k] == 0 || (b.blk_mode[k] >= 8 && b.blk_mode[k] <= 11));
    if (!ok) return fail("base macroblock %d: illegal slice type or sub-macroblock mode", i);
  }
  MotionUpsample M;
  // getSpatialResolutionChangeFlag / getRestrictedSpatialResolutionChangeFlag (include/ResizeParameters.h:130-153), progressive layers
  const bool change = (r.left_offset % 16) || (r.top_offset % 16) || r.ref_frame_width != r.scaled_ref_width || r.ref_frame_height != r.scaled_ref_height ||
                      r.ref_chroma_phase_x != r.chroma_phase_x || r.ref_chroma_phase_y != r.chroma_phase_y;
  M.rs_change = !change || (!(r.left_offset % 16) && !(r.top_offset % 16) &&
                            (r.ref_frame_height == r.scaled_ref_height || 2 * r.ref_frame_height == r.scaled_ref_height) &&
                            (r.ref_frame_width == r.scaled_ref_width || 2 * r.ref_frame_width == r.scaled_ref_width));
  M.frame_mbw = r.frame_width >> 4; M.frame_mbh = r.frame_height >> 4; M.ref_w = r.ref_frame_width; M.ref_h = r.ref_frame_height;
  M.crop_x0 = (r.left_offset + 15) / 16; M.crop_y0 = (r.top_offset + 15) / 16;
  M.crop_x1 = (r.left_offset + r.scaled_ref_width) / 16; M.crop_y1 = (r.top_offset + r.scaled_ref_height) / 16;
  M.shift_x = r.level_idc <= 30 ? 16 : 31 - ceil_log2(M.ref_w); M.shift_y = r.level_idc <= 30 ? 16 : 31 - ceil_log2(M.ref_h);
  M.scale_x = (int)((((uint32_t)M.ref_w << M.shift_x) + (uint32_t)(r.scaled_ref_width >> 1)) / (uint32_t)r.scaled_ref_width);
  M.scale_y = (int)((((uint32_t)M.ref_h << M.shift_y) + (uint32_t)(r.scaled_ref_height >> 1)) / (uint32_t)r.scaled_ref_height);
  M.add_x = (1 << (M.shift_x - 1)) - r.left_offset * M.scale_x; M.add_y = (1 << (M.shift_y - 1)) - r.top_offset * M.scale_y;
  M.mv_scale_x = (int)((((uint64_t)r.scaled_ref_width << 16) + (uint64_t)(M.ref_w >> 1)) / (uint64_t)M.ref_w);
  M.mv_scale_y = (int)((((uint64_t)r.scaled_ref_height << 16) + (uint64_t)(M.ref_h >> 1)) / (uint64_t)M.ref_h);
  M.max_list = p->slice_type == 1 ? 2 : (p->slice_type == 2 ? 0 : 1);
  M.slice_type = p->slice_type; M.ref_layer_dq_id = p->ref_layer_dq_id; M.direct_8x8_inference = p->direct_8x8_inference;
  M.check_residual_pred = p->check_residual_pred; M.mv_threshold = p->mv_threshold;
  CU_OK(cudaSetDevice(c->device));
  if (c->d_mu_base.reserve(nb) || c->d_mu_out.reserve(nc)) return JSVM_ME_ERR;
  CU_OK(cudaMemcpyAsync(c->d_mu_base.p, base, sizeof(jsvm_mb_motion) * nb, cudaMemcpyHostToDevice, c->stream));
  begin_timed(c, 3);
  motion_upsample_kernel<<<(nc + 63) / 64, 64, 0, c->stream>>>(M, c->d_mu_base.p, c->d_mu_out.p);
  end_timed(c);
  c->launches++;
  CU_OK(cudaGetLastError());
  CU_OK(cudaMemcpyAsync(out, c->d_mu_out.p, sizeof(jsvm_mb_motion) * nc, cudaMemcpyDeviceToHost, c->stream));
  CU_OK(cudaStreamSynchronize(c->stream));
  return JSVM_ME_OK;
}

// ---- multi-GPU gather over NCCL (resolved at run time: no link-time dependency, single-GPU use never loads it) ----
namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;
int load_nccl() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.lib) return JSVM_ME_OK;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);      // the copy the process already mapped (torch's), else the system one
  if (!lib) return fail("libnccl.so.2 not found (%s): the multi-GPU gather needs NCCL", dlerror());
  #define SYM(field, name) do { *(void**)(&g_nccl.field) = dlsym(lib, name); if (!g_nccl.field) return fail("NCCL symbol %s missing", name); } while (0)
  SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
  SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv");
  SYM(GetErrorString, "ncclGetErrorString");
  #undef SYM
  g_nccl.lib = lib;
  return JSVM_ME_OK;
}
#define NCCL_OK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return fail("%s failed: %s", #call, g_nccl.GetErrorString(r_)); } while (0)
}  // namespace

int jsvm_me_comm_unique_id(uint8_t id[128]) {
  if (!id) return fail("null argument");
  if (load_nccl()) return JSVM_ME_ERR;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId u;
  NCCL_OK(g_nccl.GetUniqueId(&u));
  memcpy(id, &u, 128);
  return JSVM_ME_OK;
}

int jsvm_me_comm_init(jsvm_me_ctx* c, int rank, int world, const uint8_t id[128]) {
  if (!c || !id) return fail("null argument");
  if (world < 1 || rank < 0 || rank >= world) return fail("rank %d / world %d", rank, world);
  if (c->comm) return fail("communicator already initialised");
  if (load_nccl()) return JSVM_ME_ERR;
  CU_OK(cudaSetDevice(c->device));
  ncclUniqueId u;
  memcpy(&u, id, 128);
  NCCL_OK(g_nccl.CommInitRank(&c->comm, world, u, rank));
  c->comm_rank = rank; c->comm_world = world;
  return JSVM_ME_OK;
}

int jsvm_me_gather_results_device(jsvm_me_ctx* c, const jsvm_me_result* d_local, int n_local, const int* n_per_rank,
                                  jsvm_me_result* d_all, int n_max, int root) {
  if (!c || !c->comm) return fail("jsvm_me_comm_init has not been called");
  if (!n_per_rank || root < 0 || root >= c->comm_world) return fail("bad argument");
  if (n_per_rank[c->comm_rank] != n_local || n_local > n_max) return fail("rank %d: n_local %d does not match n_per_rank / n_max", c->comm_rank, n_local);
  if (c->comm_rank == root && !d_all) return fail("null receive buffer on the root");
  CU_OK(cudaSetDevice(c->device));
  NCCL_OK(g_nccl.GroupStart());
  if (n_local > 0 && c->comm_rank != root) NCCL_OK(g_nccl.Send(d_local, (size_t)n_local * sizeof(jsvm_me_result), ncclUint8, root, c->comm, c->stream));
  if (c->comm_rank == root)
    for (int r = 0; r < c->comm_world; ++r) {
      if (r == root || n_per_rank[r] <= 0) continue;
      NCCL_OK(g_nccl.Recv(d_all + (size_t)r * n_max, (size_t)n_per_rank[r] * sizeof(jsvm_me_result), ncclUint8, r, c->comm, c->stream));
    }
  NCCL_OK(g_nccl.GroupEnd());
  if (c->comm_rank == root && n_local > 0 && d_local != d_all + (size_t)root * n_max)
    CU_OK(cudaMemcpyAsync(d_all + (size_t)root * n_max, d_local, (size_t)n_local * sizeof(jsvm_me_result), cudaMemcpyDeviceToDevice, c->stream));
  return JSVM_ME_OK;
}

int jsvm_me_comm_destroy(jsvm_me_ctx* c) {
  if (!c) return fail("null context");
  if (c->comm) { cudaStreamSynchronize(c->stream); g_nccl.CommDestroy(c->comm); c->comm = nullptr; }
  return JSVM_ME_OK;
}

int jsvm_me_sync(jsvm_me_ctx* c) {
  if (!c) return fail("null context");
  CU_OK(cudaSetDevice(c->device));
  CU_OK(cudaStreamSynchronize(c->stream));
  return JSVM_ME_OK;
}

uint64_t jsvm_me_launch_count(jsvm_me_ctx* c) { return c ? c->launches : 0; }
const char* jsvm_me_last_search_kernel(jsvm_me_ctx* c) { return c ? c->last_k2 : ""; }

int jsvm_me_timing_begin(jsvm_me_ctx* c) {
  if (!c) return fail("null context");
  for (auto& t : c->timed) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); }
  c->timed.clear();
  c->timing = true;
  return JSVM_ME_OK;
}

int jsvm_me_timing_end(jsvm_me_ctx* c, double* ms_interp, double* ms_search, double* ms_subpel,
                       uint64_t* n_interp, uint64_t* n_search, uint64_t* n_subpel) {
  if (!c) return fail("null context");
  CU_OK(cudaSetDevice(c->device));
  CU_OK(cudaStreamSynchronize(c->stream));
  double ms[4] = {0, 0, 0, 0}; uint64_t n[4] = {0, 0, 0, 0};
  const bool trace = getenv("JSVM_ME_TRACE") != nullptr;      // debugging aid: the timeline of the timed launches on stderr
  for (auto& t : c->timed) {
    float f = 0;
    CU_OK(cudaEventElapsedTime(&f, t.e0, t.e1));
    if (trace) {
      float a = 0;
      cudaEventElapsedTime(&a, c->timed.front().e0, t.e0);
      fprintf(stderr, "[jsvm_me trace] kind %d  start %8.3f ms  end %8.3f ms  (%.3f ms)\n", t.kind, a, a + f, f);
    }
    ms[t.kind] += f; n[t.kind]++;
  }
  for (auto& t : c->timed) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); }
  c->timed.clear();
  c->timing = false;
  if (ms_interp) *ms_interp = ms[0]; if (ms_search) *ms_search = ms[1]; if (ms_subpel) *ms_subpel = ms[2];
  if (n_interp) *n_interp = n[0]; if (n_search) *n_search = n[1]; if (n_subpel) *n_subpel = n[2];
  c->last_mc_ms = ms[3]; c->last_mc_n = n[3];
  return JSVM_ME_OK;
}

int jsvm_me_timing_compensate(jsvm_me_ctx* c, double* ms_compensate, uint64_t* n_compensate) {
  if (!c) return fail("null context");
  if (ms_compensate) *ms_compensate = c->last_mc_ms;
  if (n_compensate) *n_compensate = c->last_mc_n;
  return JSVM_ME_OK;
}

#ifdef K2_PRUNE_STATS
int jsvm_me_debug_prune_stats(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, jsvm::g_prune_stats, 32);
  if (reset) { unsigned long long z[4] = {0, 0, 0, 0}; cudaMemcpyToSymbol(jsvm::g_prune_stats, z, 32); }
  return 0;
}
#endif

}  // extern "C"
