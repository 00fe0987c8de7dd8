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
    // four independent 16-byte loads in flight per thread (the kernel is a pure stream: its throughput is the bytes it keeps in flight);
    // a thread still sees its indices in increasing order
    int j = threadIdx.x;
    for (; j + 3 * kArgThreads < nvec; j += 4 * kArgThreads) {
      float4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = __ldcs(src4 + j + u * kArgThreads);  // streaming: each element is read exactly once
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int base = begin + ((j + u * kArgThreads) << 2);
        consider(q[u].x, base);
        consider(q[u].y, base + 1);
        consider(q[u].z, base + 2);
        consider(q[u].w, base + 3);
      }
    }
    for (; j < nvec; j += kArgThreads) {
      const float4 q = __ldcs(src4 + j);
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

// elements per CTA: 32 float4 loads per thread (eight batches of four) before the block reduction.  Measured stand-alone at 64x4x480x640
// (tools/diag_decode.py): 8 K elements per CTA and one load in flight (round 1) 71.5 us = 4.4 TB/s; four loads in flight: 8 K 66.4 us,
// 16 K 60.5 us, 32 K 58.4 us = 5.4 TB/s; at 16x32x960x1280: 376 -> 353 us = 7.1 TB/s (a read-only stream beats the copy figure).
#ifndef HK_ARGMAX_CHUNK
#define HK_ARGMAX_CHUNK 32768
#endif
static void argmax_plan(int hw, int* chunk_len, int* chunks) {
  int c = ceil_div(hw, HK_ARGMAX_CHUNK);
  if (c < 1) c = 1;
  if (c > 64) c = 64;
  int len = ceil_div(hw, c);
  len = (len + 3) & ~3;  // keep chunk starts 16-byte aligned for the float4 path
  *chunk_len = len;
  *chunks = ceil_div(hw, len);
}


// ------------------------------------------------------------------------------------------------ soft-argmax
// `Prediction.expectation` (reference src/prediction.py:31-38, called at :45): softmax over one (H, W) map, then the expected
// (x, y) index -- with the reference's flattening quirk kept: the map is flattened as `d.T.ravel()` (column-major: flat index
// i = c*H + r for element d[r, c]) while the index arrays assume row-major order (x' = i % W, y' = i // W).  One pass over the
// heatmap (HBM-bound: 4*H*W bytes per map): each thread keeps an online-softmax state (running max m, and S = sum e,
// Sx = sum e*x', Sy = sum e*y' relative to m, e = expf(v - m) in fp32 like numpy's float32 exp, sums in fp64 like the
// reference's float64 dot); fixed-order warp / CTA / chunk reductions (deterministic, no atomics).
struct SoftPartial {
  float m;
  float pad;
  double s, sx, sy;
};

__device__ __forceinline__ void soft_merge(float& m, double& s, double& sx, double& sy, float om, double os, double osx, double osy) {
  // combine two online-softmax states; an empty state has m = -inf and zero sums
  const float nm = fmaxf(m, om);
  const double fa = (m == nm) ? 1.0 : (double)expf(m - nm);   // m = -inf -> 0
  const double fb = (om == nm) ? 1.0 : (double)expf(om - nm);
  s = s * fa + os * fb;
  sx = sx * fa + osx * fb;
  sy = sy * fa + osy * fb;
  m = nm;
}

constexpr int kSoftThreads = 256;

__global__ void __launch_bounds__(kSoftThreads)
soft_argmax_partial_kernel(const float* __restrict__ heat, int H, int W, int rows_per_chunk, int chunks, SoftPartial* __restrict__ partial) {
  const int map = blockIdx.y, chunk = blockIdx.x;
  const float* src = heat + (size_t)map * H * W;
  const int r0 = chunk * rows_per_chunk, r1 = min(H, r0 + rows_per_chunk);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float m = -INFINITY;
  double s = 0.0, sx = 0.0, sy = 0.0;
  const int hx = H % W, hy = H / W;   // flat index step between neighbouring columns of one row: i += H
  const bool vec_ok = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  for (int r = r0 + warp; r < r1; r += kSoftThreads / 32) {
    const float* row = src + (size_t)r * W;
    for (int c = lane * 4; c < W; c += 128) {
      float v[4];
      int n = 4;
      if (vec_ok) {
        const float4 q = __ldcs(reinterpret_cast<const float4*>(row + c));
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
        n = min(4, W - c);
        for (int j = 0; j < 4; ++j) v[j] = j < n ? row[c + j] : -INFINITY;
      }
      float vm = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
      if (vm > m) {   // rescale the running sums once per 4 elements at most
        const double f = (double)expf(m - vm);   // m = -inf -> 0
        s *= f; sx *= f; sy *= f;
        m = vm;
      }
      const unsigned i0 = (unsigned)c * (unsigned)H + (unsigned)r;   // flat index of d.T.ravel() (< 2^31, checked on the host)
      int yq = (int)(i0 / (unsigned)W), xq = (int)(i0 - (unsigned)yq * (unsigned)W);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < n) {
          const double e = (double)expf(v[j] - m);
          s += e;
          sx += e * (double)xq;
          sy += e * (double)yq;
        }
        xq += hx; yq += hy;
        if (xq >= W) { xq -= W; ++yq; }
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, off);
    const double os = __shfl_xor_sync(0xffffffffu, s, off), osx = __shfl_xor_sync(0xffffffffu, sx, off),
                 osy = __shfl_xor_sync(0xffffffffu, sy, off);
    soft_merge(m, s, sx, sy, om, os, osx, osy);
  }
  __shared__ SoftPartial sp[kSoftThreads / 32];
  if (lane == 0) { sp[warp].m = m; sp[warp].s = s; sp[warp].sx = sx; sp[warp].sy = sy; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kSoftThreads / 32; ++w) soft_merge(m, s, sx, sy, sp[w].m, sp[w].s, sp[w].sx, sp[w].sy);
    SoftPartial out;
    out.m = m; out.pad = 0.f; out.s = s; out.sx = sx; out.sy = sy;
    partial[(size_t)map * chunks + chunk] = out;
  }
}

__global__ void soft_argmax_final_kernel(const SoftPartial* __restrict__ partial, int maps, int chunks, double* __restrict__ exp_xy,
                                         int32_t* __restrict__ exp_int) {
  const int map = blockIdx.x * blockDim.x + threadIdx.x;
  if (map >= maps) return;
  float m = -INFINITY;
  double s = 0.0, sx = 0.0, sy = 0.0;
  for (int c = 0; c < chunks; ++c) {
    const SoftPartial p = partial[(size_t)map * chunks + c];
    soft_merge(m, s, sx, sy, p.m, p.s, p.sx, p.sy);
  }
  const double ex = sx / s, ey = sy / s;
  exp_xy[2 * map + 0] = ex;
  exp_xy[2 * map + 1] = ey;
  if (exp_int) {   // int(np.dot(...)) of the reference: truncation toward zero (values are >= 0)
    exp_int[2 * map + 0] = (int32_t)ex;
    exp_int[2 * map + 1] = (int32_t)ey;
  }
}

static void soft_plan(int H, int* rows_per_chunk, int* chunks) {
  int c = ceil_div(H, 16);   // >= 16 rows per CTA (two per warp); at most 64 chunks per map
  if (c > 64) c = 64;
  if (c < 1) c = 1;
  *rows_per_chunk = ceil_div(H, c);
  *chunks = ceil_div(H, *rows_per_chunk);
}

}  // namespace hk

extern "C" {

size_t hk_soft_argmax_workspace_bytes(int maps, int H, int W) {
  if (maps <= 0 || H <= 0 || W <= 0) return 0;
  int rows, chunks;
  hk::soft_plan(H, &rows, &chunks);
  return (size_t)maps * chunks * sizeof(hk::SoftPartial);
}

int hk_soft_argmax(const float* heat, int maps, int H, int W, double* exp_xy, int32_t* exp_int_or_null, void* ws, size_t ws_bytes,
                   void* stream) {
  using namespace hk;
  HK_REQUIRE(heat && exp_xy && ws, "hk_soft_argmax: null pointer");
  HK_REQUIRE(maps > 0 && H > 0 && W > 0 && maps <= 65535, "hk_soft_argmax: bad shape maps=%d H=%d W=%d", maps, H, W);
  HK_REQUIRE((long long)H * W < 0x7fffffffLL, "hk_soft_argmax: map too large");
  HK_REQUIRE(ws_bytes >= hk_soft_argmax_workspace_bytes(maps, H, W), "hk_soft_argmax: workspace too small");
  HK_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 7) == 0, "hk_soft_argmax: workspace must be 8-byte aligned");
  int rows, chunks;
  soft_plan(H, &rows, &chunks);
  cudaStream_t s = as_stream(stream);
  soft_argmax_partial_kernel<<<dim3(chunks, maps), kSoftThreads, 0, s>>>(heat, H, W, rows, chunks, static_cast<SoftPartial*>(ws));
  int rc = check_launch("soft_argmax_partial_kernel");
  if (rc) return rc;
  soft_argmax_final_kernel<<<ceil_div(maps, 64), 64, 0, s>>>(static_cast<const SoftPartial*>(ws), maps, chunks, exp_xy, exp_int_or_null);
  return check_launch("soft_argmax_final_kernel");
}

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
