// tools/tma_probe.cu -- minimal TMA tiled-load probe used while bringing up K1/K2 (debug aid).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include "../jsvm-code-notes_b200/csrc/k2_fullsearch.cuh"
using namespace jsvm;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe3d(const __grid_constant__ CUtensorMap map, uint8_t* out, int bytes, int x, int y, int z) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_expect_tx(&bar, bytes); tma_load_3d(sm, &map, &bar, x, y, z); }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char** argv) {
  int bw = argc > 1 ? atoi(argv[1]) : 80, bh = argc > 2 ? atoi(argv[2]) : 37;
  int x = argc > 3 ? atoi(argv[3]) : 10, y = argc > 4 ? atoi(argv[4]) : 2, z = argc > 5 ? atoi(argv[5]) : 1;
  int W = 240, H = 208, N = 2;
  std::vector<uint8_t> h((size_t)W * H * N);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + i / W);
  uint8_t *d, *o; cudaMalloc(&d, h.size()); cudaMalloc(&o, 65536);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap map;
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N}; cuuint64_t strides[2] = {(cuuint64_t)W, (cuuint64_t)W * H};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
  CUresult r = ((EncodeTiledFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode=%d box=%dx%d at (%d,%d,%d)\n", (int)r, bw, bh, x, y, z);
  int bytes = bw * bh;
  probe3d<<<1, 128, bytes>>>(map, o, bytes, x, y, z);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<uint8_t> g(bytes); cudaMemcpy(g.data(), o, bytes, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r2 = 0; r2 < bh; ++r2) for (int c = 0; c < bw; ++c) {
    int gx = x + c, gy = y + r2; uint8_t want = (gx < 0 || gx >= W || gy < 0 || gy >= H) ? 0 : h[(size_t)z * W * H + (size_t)gy * W + gx];
    if (g[r2 * bw + c] != want) ++bad;
  }
  printf("mismatches=%d\n", bad);
  return 0;
}
