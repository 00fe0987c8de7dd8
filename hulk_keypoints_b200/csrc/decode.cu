// Per-keypoint heatmap argmax.  HBM-bound: reads 4*K*H*W bytes per image once.
//
// Replaces `np.unravel_index(h.argmax(), h.shape)` (reference src/prediction.py:46): the winner is
// the FIRST flat index holding the maximum; NaN beats every number (numpy semantics).
// Two launches, no atomics, deterministic:
//   1. argmax_partial_kernel: grid (chunks, B*K); each CTA streams one chunk of one map with 128-bit
//      loads, keeps a per-thread (value, index) pair, then warp-shuffle + smem reduction.
//   2. argmax_final_kernel: one warp per map folds the per-chunk partials and writes (y, x).
#include "hk_common.cuh"

namespace hk {

struct ArgPair {
  float v;
  int i;
};

// true when candidate b must replace a
__device__ __forceinline__ bool arg_better(float bv, int bi, float av, int ai) {
  const bool bnan = bv != bv, anan = av != av;
  if (bnan || anan) return bnan && (!anan || bi < ai);
  return bv > av || (bv == av && bi < ai);
}

__device__ __forceinline__ void warp_arg_reduce(float& v, int& i) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, off);
    const int oi = __shfl_xor_sync(0xffffffffu, i, off);
    if (arg_better(ov, oi, v, i)) {
      v = ov;
      i = oi;
    }
  }
}

constexpr int kArgThreads = 256;

__global__ void __launch_bounds__(kArgThreads)
argmax_partial_kernel(const float* __restrict__ heat, int hw, int chunk_len, int chunks, ArgPair* __restrict__ partial) {
  const int map = blockIdx.y;
  const int chunk = blockIdx.x;
  const float* src = heat + (size_t)map * hw;
  const int begin = chunk * chunk_len;
  const int end = min(hw, begin + chunk_len);

  float bv = -INFINITY;
  int bi = 0x7fffffff;
  bool have = false;
  auto consider = [&](float v, int idx) {
    // indices arrive in increasing order per thread: strict '>' keeps the first occurrence
    if (!have || (v > bv) || (v != v && bv == bv)) {
      bv = v;
      bi = idx;
      have = true;
    }
  };

  const bool vec_ok = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((begin & 3) == 0);
  if (vec_ok) {
    const int nvec = (end - begin) >> 2;
    const float4* src4 = reinterpret_cast<const float4*>(src + begin);
    for (int j = threadIdx.x; j < nvec; j += kArgThreads) {
      const float4 q = __ldcs(src4 + j);  // streaming: each element is read exactly once
      const int base = begin + (j << 2);
      consider(q.x, base);
      consider(q.y, base + 1);
      consider(q.z, base + 2);
      consider(q.w, base + 3);
    }
    for (int idx = begin + (nvec << 2) + threadIdx.x; idx < end; idx += kArgThreads) consider(src[idx], idx);
  } else {
    for (int idx = begin + threadIdx.x; idx < end; idx += kArgThreads) consider(src[idx], idx);
  }
  if (!have) {
    bv = -INFINITY;
    bi = 0x7fffffff;
  }
  warp_arg_reduce(bv, bi);

  __shared__ float sv[kArgThreads / 32];
  __shared__ int si[kArgThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    sv[warp] = bv;
    si[warp] = bi;
  }
  __syncthreads();
  if (warp == 0) {
    bv = lane < kArgThreads / 32 ? sv[lane] : -INFINITY;
    bi = lane < kArgThreads / 32 ? si[lane] : 0x7fffffff;
    warp_arg_reduce(bv, bi);
    if (lane == 0) {
      ArgPair r;
      r.v = bv;
      r.i = bi;
      partial[(size_t)map * chunks + chunk] = r;
    }
  }
}

__global__ void __launch_bounds__(128)
argmax_final_kernel(const ArgPair* __restrict__ partial, int maps, int chunks, int w, int32_t* __restrict__ yx,
                    float* __restrict__ maxval) {
  const int map = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (map >= maps) return;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = lane; c < chunks; c += 32) {
    const ArgPair r = partial[(size_t)map * chunks + c];
    if (r.i != 0x7fffffff && (bi == 0x7fffffff || arg_better(r.v, r.i, bv, bi))) {
      bv = r.v;
      bi = r.i;
    }
  }
  warp_arg_reduce(bv, bi);
  if (lane == 0) {
    if (bi == 0x7fffffff) bi = 0;
    yx[2 * map + 0] = bi / w;
    yx[2 * map + 1] = bi % w;
    if (maxval) maxval[map] = bv;
  }
}

static void argmax_plan(int hw, int* chunk_len, int* chunks) {
  int c = ceil_div(hw, 8192);
  if (c < 1) c = 1;
  if (c > 64) c = 64;
  int len = ceil_div(hw, c);
  len = (len + 3) & ~3;  // keep chunk starts 16-byte aligned for the float4 path
  *chunk_len = len;
  *chunks = ceil_div(hw, len);
}

}  // namespace hk

extern "C" {

size_t hk_argmax_workspace_bytes(int B, int K, int H, int W) {
  if (B <= 0 || K <= 0 || H <= 0 || W <= 0) return 0;
  int len, chunks;
  hk::argmax_plan(H * W, &len, &chunks);
  return (size_t)B * K * chunks * sizeof(hk::ArgPair);
}

int hk_argmax_decode(const float* heat, int B, int K, int H, int W, int32_t* yx, float* maxval_or_null, void* ws,
                     size_t ws_bytes, void* stream) {
  using namespace hk;
  HK_REQUIRE(heat && yx && ws, "hk_argmax_decode: null pointer");
  HK_REQUIRE(B > 0 && K > 0 && H > 0 && W > 0, "hk_argmax_decode: bad shape B=%d K=%d H=%d W=%d", B, K, H, W);
  HK_REQUIRE((long long)H * W < 0x7fffffffLL, "hk_argmax_decode: map too large");
  HK_REQUIRE((long long)B * K <= 65535, "hk_argmax_decode: B*K=%lld exceeds grid.y", (long long)B * K);
  HK_REQUIRE(ws_bytes >= hk_argmax_workspace_bytes(B, K, H, W), "hk_argmax_decode: workspace too small");
  int len, chunks;
  argmax_plan(H * W, &len, &chunks);
  const int maps = B * K;
  cudaStream_t s = as_stream(stream);
  argmax_partial_kernel<<<dim3(chunks, maps), kArgThreads, 0, s>>>(heat, H * W, len, chunks, static_cast<ArgPair*>(ws));
  int rc = check_launch("argmax_partial_kernel");
  if (rc) return rc;
  argmax_final_kernel<<<ceil_div(maps, 4), 128, 0, s>>>(static_cast<const ArgPair*>(ws), maps, chunks, W, yx, maxval_or_null);
  return check_launch("argmax_final_kernel");
}

}  // extern "C"
