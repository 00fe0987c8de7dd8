// Micro-test: which fp32 / SWIZZLE_NONE TMA box shapes and coordinates load correctly?   usage: tma_f32_test <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "hk_ptx.cuh"
using namespace hk;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap m, int c0, int c1, int c2, int c3, int bytes, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(&bar, bytes);
    ptx::tma_load_4d(smem, &m, &bar, c0, c1, c2, c3);
  }
  ptx::mbar_wait(&bar, 0, 1);
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}
int main(int argc, char** argv) {
  const int v = argc > 1 ? atoi(argv[1]) : 0;
  const int W = 128, H = 96, C = 3, B = 1;
  int bw = 132, bh = 2, bc = 3, x0 = -5, y0 = -4;
  if (v == 1) { x0 = 0; y0 = 0; }
  if (v == 2) { bw = 128; }
  if (v == 3) { bc = 1; }
  if (v == 4) { bw = 64; x0 = 0; y0 = 0; }
  if (v == 5) { y0 = 0; }
  if (v == 6) { x0 = -4; }
  if (v == 7) { x0 = 251; y0 = 10; }
  std::vector<float> h((size_t)B * C * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 1000) + 1.f;
  float *d, *out;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 65536);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn encode = (EncodeTiledFn)p;
  CUtensorMap m;
  const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
  const cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bc, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d box {%d,%d,%d,1} coord (%d,%d,0,0): encode rc %d; ", v, bw, bh, bc, x0, y0, (int)r);
  const int bytes = bw * bh * bc * 4;
  k<<<1, 128, 32768>>>(m, x0, y0, 0, 0, bytes, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("run: %s; ", cudaGetErrorString(e));
  if (e != cudaSuccess) { printf("\n"); return 1; }
  std::vector<float> o(bytes / 4);
  cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int c = 0; c < bc; ++c) for (int y = 0; y < bh; ++y) for (int x = 0; x < bw; ++x) {
    const int gx = x0 + x, gy = y0 + y;
    const float ref = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[((size_t)c * H + gy) * W + gx] : 0.f;
    if (o[(c * bh + y) * bw + x] != ref) ++bad;
  }
  printf("mismatches %d of %d\n", bad, bytes / 4);
  return 0;
}
