// Micro-test: tcgen05.mma (M=128, N=64, K=16, bf16) with a K-major NO-SWIZZLE A operand whose core matrices OVERLAP
// (row i, 16-byte K chunk j at byte 16*(i+j): LBO = 16 B, SBO = 128 B) -- the layout the row-streaming stem relies on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I hulk_keypoints_b200/csrc -o tools/bin/umma_nosw_test tools/umma_nosw_test.cu
//   usage: umma_nosw_test <mode>   mode 0: canonical non-overlapping A (LBO=128,SBO=256); 1: overlapping (LBO=16,SBO=128)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "hk_ptx.cuh"
using namespace hk;

__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__global__ void k(const __nv_bfloat16* a_img, int a_bytes, const __nv_bfloat16* b_img, int b_bytes, uint32_t lbo_a, uint32_t sbo_a, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + 8192;
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < a_bytes / 2; i += blockDim.x) reinterpret_cast<__nv_bfloat16*>(sA)[i] = a_img[i];
  for (int i = tid; i < b_bytes / 2; i += blockDim.x) reinterpret_cast<__nv_bfloat16*>(sB)[i] = b_img[i];
  if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(&tptr, 64); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tb = tptr;
  if (tid == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16_f32(128, 64);
    ptx::umma_bf16(tb, desc_nosw(ptx::smem_u32(sA), lbo_a, sbo_a), desc_nosw(ptx::smem_u32(sB), 128, 256), idesc, 0u);
    ptx::umma_commit(&bar);
  }
  ptx::mbar_wait(&bar, 0, 99);
  ptx::tc_fence_after();
  uint32_t r[32];
  for (int half = 0; half < 2; ++half) {
    ptx::tmem_ld_32x32(tb + ((uint32_t)(warp * 32) << 16) + half * 32, r);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(size_t)tid * 64 + half * 32 + j] = __uint_as_float(r[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tb, 64); }
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 1;
  // logical A[128][16], B[64][16]
  std::vector<float> A(128 * 16), B(64 * 16);
  std::vector<__nv_bfloat16> a_img(4096, __float2bfloat16(0.f)), b_img(64 * 16);
  srand(1);
  uint32_t lbo, sbo;
  if (mode == 0) {
    lbo = 128; sbo = 256;
    for (int i = 0; i < 128; ++i) for (int kk = 0; kk < 16; ++kk) {
      float v = (float)((rand() % 17) - 8) / 8.f;
      A[i * 16 + kk] = v;
      const int j = kk / 8, e = kk % 8;
      a_img[((i / 8) * sbo + j * lbo + (i % 8) * 16) / 2 + e] = __float2bfloat16(v);
    }
  } else {
    lbo = 16; sbo = 128;
    std::vector<float> base(8 * 128 + 16);
    for (auto& v : base) v = (float)((rand() % 17) - 8) / 8.f;
    for (size_t i = 0; i < base.size(); ++i) a_img[i] = __float2bfloat16(base[i]);
    for (int i = 0; i < 128; ++i) for (int kk = 0; kk < 16; ++kk) A[i * 16 + kk] = base[8 * i + kk];   // byte 16*(i+j) + 2e, k = 8j+e
  }
  for (int n = 0; n < 64; ++n) for (int kk = 0; kk < 16; ++kk) {
    float v = (float)((rand() % 13) - 6) / 4.f;
    B[n * 16 + kk] = v;
    const int j = kk / 8, e = kk % 8;
    b_img[((n / 8) * 256 + j * 128 + (n % 8) * 16) / 2 + e] = __float2bfloat16(v);
  }
  __nv_bfloat16 *da, *db; float* dout;
  cudaMalloc(&da, a_img.size() * 2); cudaMalloc(&db, b_img.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(da, a_img.data(), a_img.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b_img.data(), b_img.size() * 2, cudaMemcpyHostToDevice);
  k<<<1, 128, 16384>>>(da, (int)a_img.size() * 2, db, (int)b_img.size() * 2, lbo, sbo, dout);
  cudaError_t e = cudaDeviceSynchronize();
  printf("mode %d (LBO %u SBO %u): %s\n", mode, lbo, sbo, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> out(128 * 64);
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int i = 0; i < 128; ++i) for (int n = 0; n < 64; ++n) {
    double ref = 0;
    for (int kk = 0; kk < 16; ++kk) ref += (double)A[i * 16 + kk] * B[n * 16 + kk];
    worst = fmax(worst, fabs(ref - out[i * 64 + n]));
  }
  printf("max |D - ref| = %g  -> %s\n", worst, worst < 1e-3 ? "OK" : "MISMATCH");
  return worst < 1e-3 ? 0 : 2;
}
