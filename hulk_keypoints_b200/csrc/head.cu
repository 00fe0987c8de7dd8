// Heatmap head (SURVEY.md k8 + k9): scoring 1x1 conv on the K live rows, bilinear x8 upsample with
// align_corners=True, sigmoid.  HBM-bound: reads the (B,h,w,C) feature map once, writes (B,K,H,W) fp32 once.
//
// Reference call sites: src/resnet.py:215 (fc applied as a 1x1 conv), src/resnet_dilated.py:27
// (upsample_bilinear to the input size), src/model.py:21 (slice [:, :K], sigmoid).  The reference runs
// the conv and the upsample on all 1000 channels and slices afterwards; both ops act per output
// channel, so only the K kept rows are computed here.
//
//   1. head_logits_kernel: one warp per low-res pixel; 128-bit loads of the C-vector, K dot products
//      against fc rows staged in shared memory, warp-shuffle reduction -> (B,K,h,w) fp32 logits (L2-resident).
//   2. head_upsample_sigmoid_kernel: one thread per 4 consecutive output x; ATen's align_corners
//      arithmetic (scale=(in-1)/(out-1), src=scale*dst, lambda=src-floor) in fp32, sigmoid, float4 store.
#include "hk_common.cuh"

namespace hk {

constexpr int kHeadThreads = 256;
constexpr int kHeadMaxPerLane = 16;  // C <= 512

template <typename FeatT>
__device__ __forceinline__ void load8(const FeatT* p, float* f);
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float* f) {
  const uint4 q = __ldcs(reinterpret_cast<const uint4*>(p));
  unpack_bf16x2(q.x, f[0], f[1]);
  unpack_bf16x2(q.y, f[2], f[3]);
  unpack_bf16x2(q.z, f[4], f[5]);
  unpack_bf16x2(q.w, f[6], f[7]);
}
template <>
__device__ __forceinline__ void load8<float>(const float* p, float* f) {
  const float4 a = __ldcs(reinterpret_cast<const float4*>(p));
  const float4 b = __ldcs(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

template <typename FeatT>
__global__ void __launch_bounds__(kHeadThreads)
head_logits_kernel(const FeatT* __restrict__ feat, const float* __restrict__ w_fc, const float* __restrict__ b_fc,
                   float* __restrict__ logits, long long pixels, int hw, int K, int C) {
  extern __shared__ float sw[];  // K*C fc rows
  for (int i = threadIdx.x; i < K * C; i += kHeadThreads) sw[i] = w_fc[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = kHeadThreads >> 5;
  const int slabs = C >> 8;  // 256 channels per pass (32 lanes x 8)
  for (long long pix = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); pix < pixels;
       pix += (long long)gridDim.x * warps_per_block) {
    float f[kHeadMaxPerLane];
    const FeatT* src = feat + pix * C;
#pragma unroll
    for (int s = 0; s < kHeadMaxPerLane / 8; ++s)
      if (s < slabs) load8<FeatT>(src + s * 256 + lane * 8, f + s * 8);
    const long long b = pix / hw;
    const int rem = (int)(pix - b * hw);
    for (int k = 0; k < K; ++k) {
      float acc = 0.f;
      const float* wk = sw + k * C;
#pragma unroll
      for (int s = 0; s < kHeadMaxPerLane / 8; ++s) {
        if (s < slabs) {
          const float4 w0 = *reinterpret_cast<const float4*>(wk + s * 256 + lane * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(wk + s * 256 + lane * 8 + 4);
          acc = fmaf(f[s * 8 + 0], w0.x, acc);
          acc = fmaf(f[s * 8 + 1], w0.y, acc);
          acc = fmaf(f[s * 8 + 2], w0.z, acc);
          acc = fmaf(f[s * 8 + 3], w0.w, acc);
          acc = fmaf(f[s * 8 + 4], w1.x, acc);
          acc = fmaf(f[s * 8 + 5], w1.y, acc);
          acc = fmaf(f[s * 8 + 6], w1.z, acc);
          acc = fmaf(f[s * 8 + 7], w1.w, acc);
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (lane == 0) logits[(b * K + k) * hw + rem] = acc + __ldg(b_fc + k);
    }
  }
}

__device__ __forceinline__ float sigmoidf_ref(float v) { return 1.0f / (1.0f + expf(-v)); }

__global__ void __launch_bounds__(256)
head_upsample_sigmoid_kernel(const float* __restrict__ logits, float* __restrict__ heat, int maps, int h, int w, int H,
                             int W, float ry, float rx) {
  const int wq = (W + 3) >> 2;
  const long long total = (long long)maps * H * wq;
  const bool vec = (W & 3) == 0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int xq = (int)(t % wq);
    const long long r = t / wq;
    const int Y = (int)(r % H);
    const int map = (int)(r / H);
    const float sy = ry * (float)Y;
    const int y0 = min((int)sy, h - 1);
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0);
    const float ly = sy - (float)y0, hy = 1.0f - ly;
    const float* row0 = logits + ((size_t)map * h + y0) * w;
    const float* row1 = logits + ((size_t)map * h + y1) * w;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int X = min(xq * 4 + j, W - 1);
      const float sx = rx * (float)X;
      const int x0 = min((int)sx, w - 1);
      const int x1 = x0 + (x0 < w - 1 ? 1 : 0);
      const float lx = sx - (float)x0, hx = 1.0f - lx;
      const float v = hy * (hx * __ldg(row0 + x0) + lx * __ldg(row0 + x1)) + ly * (hx * __ldg(row1 + x0) + lx * __ldg(row1 + x1));
      o[j] = sigmoidf_ref(v);
    }
    float* dst = heat + ((size_t)map * H + Y) * W + xq * 4;
    if (vec) {
      __stcs(reinterpret_cast<float4*>(dst), make_float4(o[0], o[1], o[2], o[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (xq * 4 + j < W) dst[j] = o[j];
    }
  }
}

}  // namespace hk

extern "C" int hk_head_fwd(const void* feat, int feat_dtype, const float* w_fc, const float* b_fc, float* logits_ws,
                           float* heat, int B, int K, int C, int h, int w, int H, int W, void* stream) {
  using namespace hk;
  HK_REQUIRE(feat && w_fc && b_fc && logits_ws && heat, "hk_head_fwd: null pointer");
  HK_REQUIRE(B > 0 && K > 0 && h > 0 && w > 0 && H > 0 && W > 0, "hk_head_fwd: bad shape");
  HK_REQUIRE(C % 256 == 0 && C <= 32 * kHeadMaxPerLane, "hk_head_fwd: C=%d must be a multiple of 256 and <= 512", C);
  HK_REQUIRE((size_t)K * C * sizeof(float) <= 200 * 1024, "hk_head_fwd: K=%d too large for shared memory", K);
  HK_REQUIRE(feat_dtype == HK_BF16 || feat_dtype == HK_F32, "hk_head_fwd: feat dtype must be bf16 or f32");
  HK_REQUIRE((reinterpret_cast<uintptr_t>(feat) & 15) == 0 && (reinterpret_cast<uintptr_t>(heat) & 15) == 0,
             "hk_head_fwd: feat/heat must be 16-byte aligned");
  cudaStream_t s = as_stream(stream);
  const long long pixels = (long long)B * h * w;
  const size_t smem = (size_t)K * C * sizeof(float);
  int blocks = (int)ceil_div_ll(pixels, (kHeadThreads / 32) * 4);
  const int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cudaError_t e;
  if (feat_dtype == HK_BF16) {
    if (smem > 48 * 1024) {
      e = cudaFuncSetAttribute(head_logits_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return fail(HK_ERR_CUDA, "hk_head_fwd: smem attribute: %s", cudaGetErrorString(e));
    }
    head_logits_kernel<__nv_bfloat16><<<blocks, kHeadThreads, smem, s>>>(static_cast<const __nv_bfloat16*>(feat), w_fc, b_fc,
                                                                        logits_ws, pixels, h * w, K, C);
  } else {
    if (smem > 48 * 1024) {
      e = cudaFuncSetAttribute(head_logits_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return fail(HK_ERR_CUDA, "hk_head_fwd: smem attribute: %s", cudaGetErrorString(e));
    }
    head_logits_kernel<float><<<blocks, kHeadThreads, smem, s>>>(static_cast<const float*>(feat), w_fc, b_fc, logits_ws, pixels,
                                                                h * w, K, C);
  }
  int rc = check_launch("head_logits_kernel");
  if (rc) return rc;
  // ATen area_pixel_compute_scale(align_corners=True): (in - 1) / (out - 1) in float, 0 when out == 1
  const float ry = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
  const float rx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const long long total = (long long)B * K * H * ((W + 3) / 4);
  int ublocks = (int)ceil_div_ll(total, 256);
  const int ucap = sm_count() * 16;
  if (ublocks > ucap) ublocks = ucap;
  head_upsample_sigmoid_kernel<<<ublocks, 256, 0, s>>>(logits_ws, heat, B * K, h, w, H, W, ry, rx);
  return check_launch("head_upsample_sigmoid_kernel");
}
