// tools/int32_peak.cu -- measures the INT32-pipe denominators the K2/K3 roofline is quoted
// against (MEASURED_PEAKS.json has no INT32 entry, so it must be measured
// on the same box).  Every kernel is ILP-8 dependent chains of ONE instruction kind (or a
// fixed mix), all 148 SMs, 8 CTAs x 256 threads per SM.  Prints one JSON object.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/int32_peak tools/int32_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ILP = 8;
constexpr int UNROLL = 16;

enum Kind { K_IADD3, K_VABSDIFF4, K_IMAD, K_LOP3, K_VIMNMX, K_LEA, K_SHF, K_PRMT, K_DP4A,
            K_MIX_VABS_IMAD, K_MIX_VABS_IADD, K_MIX_VABS2_IMAD1, K_MIX_VABS_LEA_MIN, K_SAD32, K_NKINDS };
static const char* kKindName[] = { "iadd3", "vabsdiff4_acc", "imad", "lop3", "vimnmx", "lea", "shf", "prmt", "dp4a",
            "mix_vabsdiff4_imad_1to1", "mix_vabsdiff4_iadd3_1to1", "mix_vabsdiff4_imad_2to1", "mix_vabsdiff4_lea_min_2to1to1", "sad32" };
// lane-ops counted per inner step per chain
static const int kOpsPerStep[] = { 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 3, 4, 1 };

template <int KIND>
__global__ void __launch_bounds__(256) bench(uint32_t* out, uint32_t seed, int iters) {
  uint32_t a[ILP], b[ILP];
  uint32_t c = seed ^ threadIdx.x, d = seed * 2654435761u + blockIdx.x;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { a[i] = c + i * 0x01020304u; b[i] = d ^ (i * 0x9e3779b9u); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (KIND == K_IADD3)      asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a[i]) : "r"(b[i]), "r"(c));
        if (KIND == K_VABSDIFF4)  asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
        if (KIND == K_IMAD)       asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
        if (KIND == K_LOP3)       asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
        if (KIND == K_VIMNMX)     asm volatile("min.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
        if (KIND == K_LEA)        asm volatile("{ .reg .u32 t; shl.b32 t, %0, 15; add.u32 %0, t, %1; }" : "+r"(a[i]) : "r"(b[i]));
        if (KIND == K_SHF)        asm volatile("shf.r.wrap.b32 %0, %0, %1, 8;" : "+r"(a[i]) : "r"(b[i]));
        if (KIND == K_PRMT)       asm volatile("prmt.b32 %0, %0, %1, 0x4321;" : "+r"(a[i]) : "r"(b[i]));
        if (KIND == K_DP4A)       asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
        if (KIND == K_SAD32)      asm volatile("sad.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
        if (KIND == K_MIX_VABS_IMAD) {
          asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(d), "r"(c));
        }
        if (KIND == K_MIX_VABS_IADD) {
          asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
          asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(b[i]) : "r"(d), "r"(c));
        }
        if (KIND == K_MIX_VABS2_IMAD1) {
          asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
          asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(d), "r"(b[i]));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(d), "r"(c));
        }
        if (KIND == K_MIX_VABS_LEA_MIN) {
          uint32_t t;
          asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
          asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(d), "r"(b[i]));
          asm volatile("{ .reg .u32 s; shl.b32 s, %1, 15; add.u32 %0, s, %2; }" : "=r"(t) : "r"(a[i]), "r"(d));
          asm volatile("min.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(t));
        }
      }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r ^= a[i] ^ b[i];
  if (r == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = r;  // never true in practice; keeps the chains live
}

// shared-memory load throughput next to the ALU pipe: each thread does W-byte loads
template <int WORDS>
__global__ void __launch_bounds__(256) lds_bench(uint32_t* out, int iters) {
  __shared__ __align__(16) uint32_t sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 2654435761u;
  __syncthreads();
  uint32_t acc[4] = {0, 0, 0, 0};
  int idx = (threadIdx.x * WORDS) & 4095;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      int j = (idx + u * 256) & 4095;
      if (WORDS == 1) { acc[0] += sm[j]; }
      if (WORDS == 2) { uint2 v = *reinterpret_cast<uint2*>(&sm[j]); acc[0] += v.x; acc[1] += v.y; }
      if (WORDS == 4) { uint4 v = *reinterpret_cast<uint4*>(&sm[j]); acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w; }
    }
    idx = (idx + acc[0]) & 4095 & ~(WORDS - 1);
  }
  if ((acc[0] ^ acc[1] ^ acc[2] ^ acc[3]) == 0x12345678u) out[threadIdx.x] = acc[0];
}

template <int KIND>
static double run_kind(uint32_t* d_out, int n_sm, int iters, float min_seconds, double* sustained) {
  dim3 grid(n_sm * 8), block(256);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; ++w) bench<KIND><<<grid, block>>>(d_out, 12345u + w, iters);
  CK(cudaDeviceSynchronize());
  double best = 0.0;
  double ops_per_launch = double(grid.x) * block.x * double(iters) * UNROLL * ILP * kOpsPerStep[KIND];
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0));
    bench<KIND><<<grid, block>>>(d_out, 777u + rep, iters);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double r = ops_per_launch / (ms * 1e-3);
    if (r > best) best = r;
  }
  *sustained = 0.0;
  if (min_seconds > 0) {   // back-to-back launches for >= min_seconds (SURVEY §8d asks for a >=4 s run)
    int launches = 0; float total_ms = 0;
    CK(cudaEventRecord(e0));
    do {
      for (int k = 0; k < 8; ++k) bench<KIND><<<grid, block>>>(d_out, 99u + launches + k, iters);
      launches += 8;
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&total_ms, e0, e1));
    } while (total_ms < min_seconds * 1e3f);
    *sustained = ops_per_launch * launches / (total_ms * 1e-3);
  }
  CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
  return best;
}

template <int WORDS>
static double run_lds(uint32_t* d_out, int n_sm, int iters) {
  dim3 grid(n_sm * 8), block(256);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 2; ++w) lds_bench<WORDS><<<grid, block>>>(d_out, iters);
  CK(cudaDeviceSynchronize());
  double best = 0;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    lds_bench<WORDS><<<grid, block>>>(d_out, iters);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double bytes = double(grid.x) * block.x * double(iters) * 16 * WORDS * 4;
    double r = bytes / (ms * 1e-3);
    if (r > best) best = r;
  }
  return best;
}

int main(int argc, char** argv) {
  float sustain_s = argc > 1 ? atof(argv[1]) : 4.0f;
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int n_sm = p.multiProcessorCount;
  uint32_t* d_out; CK(cudaMalloc(&d_out, size_t(n_sm) * 8 * 256 * 4));
  int iters = 2048;
  double burst[K_NKINDS], sust[K_NKINDS];
#define RUN(K, S) burst[K] = run_kind<K>(d_out, n_sm, iters, S, &sust[K]);
  RUN(K_IADD3, sustain_s) RUN(K_VABSDIFF4, sustain_s) RUN(K_IMAD, 0) RUN(K_LOP3, 0) RUN(K_VIMNMX, 0) RUN(K_LEA, 0)
  RUN(K_SHF, 0) RUN(K_PRMT, 0) RUN(K_DP4A, 0) RUN(K_MIX_VABS_IMAD, 0) RUN(K_MIX_VABS_IADD, 0) RUN(K_MIX_VABS2_IMAD1, 0)
  RUN(K_MIX_VABS_LEA_MIN, 0) RUN(K_SAD32, 0)
  double lds1 = run_lds<1>(d_out, n_sm, 512), lds2 = run_lds<2>(d_out, n_sm, 512), lds4 = run_lds<4>(d_out, n_sm, 512);
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_max_mhz\": %d, \"unit\": \"lane-ops/s\", \"burst\": {", p.name, n_sm, clk_khz / 1000);
  for (int k = 0; k < K_NKINDS; ++k) printf("%s\"%s\": %.4e", k ? ", " : "", kKindName[k], burst[k]);
  printf("}, \"sustained_%gs\": {\"iadd3\": %.4e, \"vabsdiff4_acc\": %.4e}, ", sustain_s, sust[K_IADD3], sust[K_VABSDIFF4]);
  printf("\"lanes_per_clk_per_sm_at_max_clock\": {");
  for (int k = 0; k < K_NKINDS; ++k) printf("%s\"%s\": %.2f", k ? ", " : "", kKindName[k], burst[k] / (double(n_sm) * clk_khz * 1e3));
  printf("}, \"lds_bytes_per_s\": {\"lds32\": %.4e, \"lds64\": %.4e, \"lds128\": %.4e}}\n", lds1, lds2, lds4);
  return 0;
}
