// Micro-test: tcgen05.mma (M=128, N=64, K=64 as 4 x K16, bf16) with a K-major SWIZZLE_128B A operand whose START ADDRESS is shifted by
// whole 128-byte rows (pixels) inside the 1024-byte swizzle atom, the way a horizontal filter tap would address ONE haloed activation
// box [image row][16 px][64 ch] (8 output pixels + halo per image row, SBO = one image row = 2048 B).  Which value of the
// descriptor's base-offset field (bits 49..51) makes the tensor core un-swizzle correctly?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I hulk_keypoints_b200/csrc -o tools/bin/umma_sw128_offset_test tools/umma_sw128_offset_test.cu
//   usage: umma_sw128_offset_test        (runs shift 0..7 x base-offset policy 0..2)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include "hk_ptx.cuh"
using namespace hk;

constexpr int ROWS = 18, MAX_PITCH = 16, PX = ROWS * MAX_PITCH + 8;   // pixels x 128 B
constexpr int A_BYTES = PX * 128, B_BYTES = 64 * 128;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void k(const uint8_t* a_img, const uint8_t* b_img, int shift_px, int row0, uint32_t base_off, int PITCH_PX, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + ((A_BYTES + 1023) / 1024) * 1024;
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < A_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sA)[i] = reinterpret_cast<const uint32_t*>(a_img)[i];
  for (int i = tid; i < B_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sB)[i] = reinterpret_cast<const uint32_t*>(b_img)[i];
  if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(&tptr, 64); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tb = tptr;
  if (tid == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16_f32(128, 64);
    const uint32_t a0 = ptx::smem_u32(sA) + (uint32_t)(row0 * PITCH_PX + shift_px) * 128u;
    const uint64_t ad = desc_sw128(a0, PITCH_PX * 128, base_off);
    const uint64_t bd = desc_sw128(ptx::smem_u32(sB), 1024, 0);
    for (int kk = 0; kk < 4; ++kk) ptx::umma_bf16(tb, ad + 2 * kk, bd + 2 * kk, idesc, kk ? 1u : 0u);
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

int main() {
  // X[pixel][64 ch] stored as TMA SWIZZLE_128B stores it at a 1024-aligned destination: 16-byte chunk c8 of pixel p at chunk (c8 ^ (p & 7))
  std::vector<float> X(PX * 64), W(64 * 64);
  std::vector<__nv_bfloat16> a_img(A_BYTES / 2), b_img(B_BYTES / 2);
  srand(3);
  for (int p = 0; p < PX; ++p) for (int c = 0; c < 64; ++c) {
    const float v = (float)((rand() % 17) - 8) / 8.f;
    X[p * 64 + c] = v;
    a_img[(p * 128 + (((c >> 3) ^ (p & 7)) << 4)) / 2 + (c & 7)] = __float2bfloat16(v);
  }
  for (int n = 0; n < 64; ++n) for (int c = 0; c < 64; ++c) {
    const float v = (float)((rand() % 13) - 6) / 4.f;
    W[n * 64 + c] = v;
    b_img[(n * 128 + (((c >> 3) ^ (n & 7)) << 4)) / 2 + (c & 7)] = __float2bfloat16(v);
  }
  uint8_t *da, *db; float* dout;
  cudaMalloc(&da, A_BYTES); cudaMalloc(&db, B_BYTES); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(da, a_img.data(), A_BYTES, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b_img.data(), B_BYTES, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> out(128 * 64);
  for (int PITCH_PX : {16, 12, 10, 9}) {
  int ok_policy[3] = {0, 0, 0};
  for (int row0 = 0; row0 <= 2; ++row0)
    for (int shift = 0; shift < 8; ++shift)
      for (int policy = 0; policy < 3; ++policy) {
        const uint32_t bo = policy == 0 ? 0u : policy == 1 ? (uint32_t)shift : (uint32_t)((8 - shift) & 7);
        k<<<1, 128, 64 * 1024>>>(da, db, shift, row0, bo, PITCH_PX, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("row0 %d shift %d policy %d: %s\n", row0, shift, policy, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int m = 0; m < 128; ++m) {
          const int p = (row0 + m / 8) * PITCH_PX + (m % 8) + shift;   // 8 output pixels per image row (one 8-row group), SBO = one image row of PITCH_PX pixels
          for (int n = 0; n < 64; ++n) {
            double ref = 0;
            for (int c = 0; c < 64; ++c) ref += (double)X[p * 64 + c] * W[n * 64 + c];
            worst = fmax(worst, fabs(ref - out[m * 64 + n]));
          }
        }
        const bool ok = worst < 1e-2;
        ok_policy[policy] += ok;
        if (policy == 0 && !ok) printf("pitch %d row0 %d shift %d base_offset 0: max err %.4g MISMATCH\n", PITCH_PX, row0, shift, worst);
      }
  printf("pitch %2d px (SBO %4d B): OK counts of 24: base_offset=0: %d, base_offset=shift: %d, base_offset=8-shift: %d\n", PITCH_PX, PITCH_PX * 128,
         ok_policy[0], ok_policy[1], ok_policy[2]);
  }
  return 0;
}
