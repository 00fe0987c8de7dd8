// Micro-benchmark: issue rate of tcgen05.mma kind::f16 (bf16 -> fp32) from shared memory, per shape, no TMA traffic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I hulk_keypoints_b200/csrc tools/mma_rate.cu -o tools/bin/mma_rate
// Prints SM clocks per MMA instruction for (cta_group, M, N, A-major, B-major).  All 148 SMs run the same loop.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "hk_ptx.cuh"
using namespace hk;

__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void csync() { asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory"); }

template <int CG>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if constexpr (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit_mc3(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(ptx::smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint64_t* bar) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ptx::smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(ptx::smem_u32(bar)), "h"((uint16_t)1) : "memory");
}

// STAGES x (A 16 KB + B up to 32 KB)
// variant bits: 1 = commit to a stage barrier after every 4 MMAs (as a pipelined mainloop does), 2 = the other 96 threads spin on an
// mbarrier while the MMAs run, 4 = tcgen05.fence::after_thread_sync before every k-block
template <int CG, int M, int N, int AMN, int BMN>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters, int variant) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (ptx::smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint64_t stage_bar[4];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tptr;
  constexpr int STAGES = 4, A_BYTES = 16384, B_BYTES = 32768, STAGE = A_BYTES + B_BYTES;
  for (int i = threadIdx.x; i < STAGES * STAGE / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u ^ ((i * 2654435761u) & 0x03ff03ffu);  // bf16 values around 0.0078..0.0156
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::mbar_init(&done_bar, 1); for (int i = 0; i < 4; ++i) ptx::mbar_init(&stage_bar[i], 1); ptx::fence_mbar_init(); }
  if (warp == 1) {
    if constexpr (CG == 1) { ptx::tmem_alloc(&tptr, 512); ptx::tmem_relinquish(); }
    else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(&tptr)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) csync();
  ptx::tc_fence_after();
  const uint32_t tmem = tptr;
  const bool issuer = threadIdx.x == 0 && (CG == 1 || ctarank() == 0);
  if (issuer) {
    uint32_t idesc = ptx::make_idesc_bf16_f32(M, N) | (uint32_t(AMN) << 15) | (uint32_t(BMN) << 16);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t sa = ptx::smem_u32(smem + (it % STAGES) * STAGE);
      uint64_t ad = ptx::make_smem_desc_sw128(sa), bd = ptx::make_smem_desc_sw128(sa + A_BYTES);
      if (AMN) ad |= (uint64_t)(8192 >> 4) << 16;  // LBO: next 64-element group along M
      if (BMN) bd |= (uint64_t)(8192 >> 4) << 16;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        mma<CG>(tmem + (it & 1) * N, ad + (AMN ? 128 * k : 2 * k), bd + (BMN ? 128 * k : 2 * k), idesc, (it > 1 || k) ? 1u : 0u);
      if (variant & 1) { if (CG == 2 && (variant & 8)) commit_mc3(&stage_bar[it % STAGES]); else commit<CG>(&stage_bar[it % STAGES]); }
      if (variant & 4) ptx::tc_fence_after();
    }
    commit<CG>(&bar);
    ptx::mbar_wait(&bar, 0, 99);
    const long long t1 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; }
    ptx::mbar_arrive(&done_bar);
    if constexpr (CG == 2)
      asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(ptx::smem_u32(&done_bar)), "r"(1) : "memory");
  } else if ((variant & 2) && warp != 1) {
    while (!ptx::mbar_try_wait(&done_bar, 0)) {}
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) csync();
  if (warp == 1) {
    ptx::tc_fence_after();
    if constexpr (CG == 1) ptx::tmem_dealloc(tmem, 512);
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

static int g_iters = 2000;
template <int CG, int M, int N, int AMN, int BMN>
void run(long long* dout, int grid, int variant = 0) {
  auto k = rate_kernel<CG, M, N, AMN, BMN>;
  const int smem = 4 * 49152 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = g_iters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  long long best = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, dout, iters, variant);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) { printf("cg%d M%d N%d a%d b%d: ERROR %s / %s\n", CG, M, N, AMN, BMN, cudaGetErrorString(e), cudaGetErrorString(e2)); exit(1); }
    long long clk; cudaMemcpy(&clk, dout, 8, cudaMemcpyDeviceToHost);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep == 2) {
      const double per = (double)clk / (iters * 4);
      const double flops = 2.0 * M * N * 16 * iters * 4 * (grid / CG);
      printf("v%d cta_group=%d M=%3d N=%3d Amajor=%s Bmajor=%s grid=%3d: %7.1f clk/MMA (ideal %5.1f)  %6.1f%% of math peak, %7.1f TF/s wall\n", variant, CG, M, N, AMN ? "MN" : "K ",
             BMN ? "MN" : "K ", grid, per, (double)(M / CG) * N * 16 / 4096.0, 100.0 * ((double)(M / CG) * N * 16 / 4096.0) / per, flops / (ms * 1e-3) / 1e12);
    }
    best = clk;
  }
  (void)best;
}

int main(int argc, char** argv) {
  setvbuf(stdout, nullptr, _IOLBF, 0);
  long long* dout; cudaMalloc(&dout, 64);
  if (argc > 2) {  // sustained run: power-capped clocks
    g_iters = atoi(argv[2]);
    for (int rep = 0; rep < 3; ++rep) {
      run<2, 256, 256, 0, 0>(dout, 148, 1);
      run<2, 256, 128, 0, 0>(dout, 148, 1);
    }
    return 0;
  }
  if (argc > 1) {  // variants study
    for (int v : {1, 9, 15}) {
      run<2, 256, 128, 0, 0>(dout, 148, v);
      run<2, 256, 256, 0, 0>(dout, 148, v);
      run<1, 128, 128, 0, 0>(dout, 148, v);
    }
    return 0;
  }
  for (int grid : {2, 148}) {
    run<1, 128, 64, 0, 0>(dout, grid);
    run<1, 128, 128, 0, 0>(dout, grid);
    run<1, 128, 256, 0, 0>(dout, grid);
    run<1, 64, 256, 0, 0>(dout, grid);
    run<2, 256, 64, 0, 0>(dout, grid);
    run<2, 256, 128, 0, 0>(dout, grid);
    run<2, 256, 256, 0, 0>(dout, grid);
    run<2, 128, 256, 0, 0>(dout, grid);
    run<1, 128, 256, 1, 1>(dout, grid);
    run<1, 128, 128, 1, 1>(dout, grid);
    run<2, 256, 256, 1, 1>(dout, grid);
    run<1, 128, 256, 1, 0>(dout, grid);
    run<1, 128, 256, 0, 1>(dout, grid);
  }
  return 0;
}
