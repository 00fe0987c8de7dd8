// Gaussian heatmap targets and BCE forward/backward (k12/k13).  All HBM-bound.
//
// Reference arithmetic being reproduced:
//   gauss_2d_batch (src/dataset.py:36-44):  G = exp(-((X-U)^2 + (Y-V)^2) / (2.0*sigma**2)) evaluated in
//     float32 op by op (no fused multiply-add), then widened with .double().
//   train.py:21,25:  BCELoss(mean) on pred.double(): -(t*max(log p,-100) + (1-t)*max(log(1-p),-100)) in f64.
//   autograd of train.py:35 down to the logits: ATen binary_cross_entropy_backward
//     g_p = (1/N) * (p - t) / max((1-p)*p, 1e-12)  in f64, cast to f32, then sigmoid_backward
//     g_z = (g_p * (1 - p)) * p in f32.
#include "hk_common.cuh"

namespace hk {

// One fp32 Gaussian value with the reference's rounding sequence (explicit _rn ops forbid FMA contraction).
__device__ __forceinline__ float gauss_value(float x, float y, float u, float v, float denom) {
  const float dx = __fsub_rn(x, u);
  const float dy = __fsub_rn(y, v);
  const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
  return expf(__fdiv_rn(-d2, denom));
}

// grid: (ceil(W/4 * H / threads), B*K).  Each thread produces 4 values.  fp32 output: 4 consecutive x (one 16-byte store; a warp writes
// 512 contiguous bytes).  fp64 output: two PAIRS 64 elements apart inside the warp's 128-element run, so that each of the two 16-byte
// store instructions of a warp covers 512 contiguous bytes -- with 4 consecutive doubles per thread every store instruction filled only
// half of each 32-byte sector it touched: 169 us = 3.7 TB/s at 64x4x480x640, now 127 us = 5.0 TB/s (tools/diag_decode.py).
template <typename OutT>
__global__ void __launch_bounds__(256)
gauss_targets_kernel(const float* __restrict__ uv, int H, int W, float denom, OutT* __restrict__ out) {
  const int map = blockIdx.y;
  const float u = __ldg(uv + 2 * map), v = __ldg(uv + 2 * map + 1);
  const int wq = (W + 3) >> 2;
  const unsigned total = (unsigned)wq * (unsigned)H;
  OutT* dst = out + (size_t)map * H * W;
  const bool vec = (W & 3) == 0;
  if constexpr (sizeof(OutT) == 8) if ((W & 1) == 0) {
    // element pairs; a warp owns 128 consecutive elements of the flattened map per iteration (pairs may fall into different rows)
    const unsigned hw = (unsigned)H * (unsigned)W;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned warps = (gridDim.x * blockDim.x) >> 5;
    for (unsigned base = (((blockIdx.x * blockDim.x) + threadIdx.x) >> 5) * 128u; base < hw; base += warps * 128u) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const unsigned e = base + (unsigned)h * 64u + 2u * lane;
        if (e < hw) {
          const int y = (int)(e / (unsigned)W);
          const int x = (int)(e - (unsigned)y * (unsigned)W);   // even, and x + 1 < W because W is even
          const float g0 = gauss_value((float)x, (float)y, u, v, denom);
          const float g1 = gauss_value((float)(x + 1), (float)y, u, v, denom);
          __stcs(reinterpret_cast<double2*>(dst + e), make_double2((double)g0, (double)g1));
        }
      }
    }
    return;
  }
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int y = (int)(t / (unsigned)wq);
    const int x0 = (int)(t - (unsigned)y * (unsigned)wq) << 2;
    float g[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) g[j] = gauss_value((float)(x0 + j), (float)y, u, v, denom);
    OutT* p = dst + (size_t)y * W + x0;
    if (vec) {
      if constexpr (sizeof(OutT) == 8) {
        double2* p2 = reinterpret_cast<double2*>(p);
        __stcs(p2, make_double2((double)g[0], (double)g[1]));
        __stcs(p2 + 1, make_double2((double)g[2], (double)g[3]));
      } else {
        __stcs(reinterpret_cast<float4*>(p), make_float4(g[0], g[1], g[2], g[3]));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (x0 + j < W) p[j] = (OutT)g[j];
    }
  }
}

// `normalize` of reference src/dataset.py:33-34 (F.normalize(x, p=1): L1 over dim 1 = the H axis of a (K,H,W) tensor, eps 1e-12),
// applied by gauss_2d_batch(normalize_dist=True) (dataset.py:42-43; unused by the reference's own calls).  One thread per (k, w)
// column: coalesced across the warp, sequential fp32 sum over h, then the divide; fp32 in, fp64 out (the `.double()` of :44).
__global__ void l1_normalize_dim1_kernel(const float* __restrict__ x, int H, int W, double* __restrict__ out) {
  const int k = blockIdx.y, w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  const float* src = x + (size_t)k * H * W + w;
  float acc = 0.f;
  for (int h = 0; h < H; ++h) acc = __fadd_rn(acc, fabsf(__ldg(src + (size_t)h * W)));
  const float denom = fmaxf(acc, 1e-12f);
  double* dst = out + (size_t)k * H * W + w;
  for (int h = 0; h < H; ++h) dst[(size_t)h * W] = (double)__fdiv_rn(__ldg(src + (size_t)h * W), denom);
}

constexpr int kBceThreads = 256;
constexpr int kBceMaxBlocks = 148 * 8;

// TargetMode: 0 = f64 tensor, 1 = f32 tensor, 2 = generated from uv on the fly.
template <int TargetMode, bool FromLogits>
__global__ void __launch_bounds__(kBceThreads)
bce_fwd_bwd_kernel(const float* __restrict__ pred, const void* __restrict__ target, const float* __restrict__ uv,
                   long long n, int H, int W, float denom, double inv_n, double* __restrict__ block_sums,
                   float* __restrict__ grad_logits) {
  double acc = 0.0;
  const long long nq = n >> 2;  // host guarantees n % 4 == 0 and (for mode 2) W % 4 == 0
  const int hw = H * W;
  for (long long q = (long long)blockIdx.x * kBceThreads + threadIdx.x; q < nq; q += (long long)gridDim.x * kBceThreads) {
    const float4 p4 = __ldcs(reinterpret_cast<const float4*>(pred) + q);
    float pf[4] = {p4.x, p4.y, p4.z, p4.w};
    if constexpr (FromLogits) {  // model.py:21 sigmoid fused in: the heatmap is never written for the loss
#pragma unroll
      for (int j = 0; j < 4; ++j) pf[j] = 1.0f / (1.0f + expf(-pf[j]));
    }
    double t[4];
    if constexpr (TargetMode == 0) {
      const double2* tp = reinterpret_cast<const double2*>(target) + 2 * q;
      const double2 a = __ldcs(tp), b = __ldcs(tp + 1);
      t[0] = a.x; t[1] = a.y; t[2] = b.x; t[3] = b.y;
    } else if constexpr (TargetMode == 1) {
      const float4 a = __ldcs(reinterpret_cast<const float4*>(target) + q);
      t[0] = a.x; t[1] = a.y; t[2] = a.z; t[3] = a.w;
    } else {
      const unsigned e = (unsigned)q << 2;  // host guarantees n < 2^32 in label mode
      const int map = (int)(e / (unsigned)hw);
      const int rem = (int)(e - (unsigned)map * (unsigned)hw);
      const int y = rem / W, x0 = rem - y * W;
      const float u = __ldg(uv + 2 * map), v = __ldg(uv + 2 * map + 1);
#pragma unroll
      for (int j = 0; j < 4; ++j) t[j] = (double)gauss_value((float)(x0 + j), (float)y, u, v, denom);
    }
    float g[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double p = (double)pf[j];
      // Each fp64 log costs ~45 double-precision instructions and this kernel is bound by them.  Where the target is exactly 0 (the fp32
      // Gaussian underflows beyond ~115 px of the keypoint: ~86 % of a 480x640 map) the term t*log(p) is 0 * (finite, clamped) = 0 and adding
      // it changes no bit of the sum, so log(p) is not evaluated; likewise log(1-p) where t is exactly 1.  Only with the in-kernel sigmoid
      // (0 < p < 1 guaranteed); a caller-supplied p keeps both logs so that invalid probabilities still poison the loss as in torch.
      const bool skip_lp = FromLogits && t[j] == 0.0, skip_l1p = FromLogits && t[j] == 1.0;
      const double lp = skip_lp ? 0.0 : fmax(log(p), -100.0);
      const double l1p = skip_l1p ? 0.0 : fmax(log(1.0 - p), -100.0);
      acc -= t[j] * lp + (1.0 - t[j]) * l1p;
      const double gp = inv_n * (p - t[j]) / fmax((1.0 - p) * p, 1e-12);
      g[j] = __fmul_rn(__fmul_rn((float)gp, __fsub_rn(1.0f, pf[j])), pf[j]);
    }
    if (grad_logits) __stcs(reinterpret_cast<float4*>(grad_logits) + q, make_float4(g[0], g[1], g[2], g[3]));
  }
  // block reduction in a fixed order -> deterministic partial
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  __shared__ double sacc[kBceThreads / 32];
  if ((threadIdx.x & 31) == 0) sacc[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kBceThreads / 32; ++i) s += sacc[i];
    block_sums[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256)
bce_finish_kernel(const double* __restrict__ block_sums, int nblocks, double inv_n, double* __restrict__ loss) {
  __shared__ double s[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 256) acc += block_sums[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) s[threadIdx.x] += s[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = s[0] * inv_n;
}

static int bce_blocks(long long n) {
  long long b = ceil_div_ll(n >> 2, kBceThreads * 4);
  if (b < 1) b = 1;
  if (b > kBceMaxBlocks) b = kBceMaxBlocks;
  return (int)b;
}

}  // namespace hk

namespace hk {
// sigmoid of model.py:21 and its autograd (sigmoid_backward: g_z = (g_p * (1 - p)) * p, ATen's operation order), for the split
// train step where the loss is computed by the caller (an unmodified train.py: nn.BCELoss on model.forward(img).double()).
__global__ void __launch_bounds__(256) sigmoid_fwd_kernel(const float* __restrict__ z, float* __restrict__ p, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = 1.0f / (1.0f + expf(-z[i]));
}
__global__ void __launch_bounds__(256) sigmoid_bwd_kernel(const float* __restrict__ p, const float* __restrict__ gp, float* __restrict__ gz,
                                                         long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pv = p[i];
    gz[i] = __fmul_rn(__fmul_rn(gp[i], __fsub_rn(1.0f, pv)), pv);
  }
}
}  // namespace hk

extern "C" {

int hk_sigmoid_fwd(const float* logits, float* heat, long long n, void* stream) {
  using namespace hk;
  HK_REQUIRE(logits && heat && n > 0, "hk_sigmoid_fwd: bad argument");
  long long blocks = ceil_div_ll(n, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  sigmoid_fwd_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(logits, heat, n);
  return check_launch("sigmoid_fwd_kernel");
}

int hk_sigmoid_bwd(const float* heat, const float* grad_heat, float* grad_logits, long long n, void* stream) {
  using namespace hk;
  HK_REQUIRE(heat && grad_heat && grad_logits && n > 0, "hk_sigmoid_bwd: bad argument");
  long long blocks = ceil_div_ll(n, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  sigmoid_bwd_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(heat, grad_heat, grad_logits, n);
  return check_launch("sigmoid_bwd_kernel");
}

int hk_gauss_targets(const float* uv, int B, int K, int H, int W, float sigma, void* out, int out_dtype, void* stream) {
  using namespace hk;
  HK_REQUIRE(uv && out, "hk_gauss_targets: null pointer");
  HK_REQUIRE(B > 0 && K > 0 && H > 0 && W > 0 && sigma > 0.f, "hk_gauss_targets: bad shape/sigma");
  HK_REQUIRE((long long)B * K <= 65535, "hk_gauss_targets: B*K exceeds grid.y");
  HK_REQUIRE(out_dtype == HK_F64 || out_dtype == HK_F32, "hk_gauss_targets: out_dtype must be HK_F64 or HK_F32");
  HK_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "hk_gauss_targets: out must be 16-byte aligned");
  // python: 2.0*sigma**2 evaluated in double, then the tensor/scalar division runs in float32
  const float denom = (float)(2.0 * (double)sigma * (double)sigma);
  const long long per_map = (long long)((W + 3) / 4) * H;
  int bx = (int)ceil_div_ll(per_map, 256);
  const int cap = ceil_div(148 * 16, B * K) > 1 ? ceil_div(148 * 16, B * K) : 1;
  if (bx > cap) bx = cap;  // a few waves in total; the grid-stride loop covers the rest
  dim3 grid(bx, B * K);
  if (out_dtype == HK_F64)
    gauss_targets_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(uv, H, W, denom, static_cast<double*>(out));
  else
    gauss_targets_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(uv, H, W, denom, static_cast<float*>(out));
  return check_launch("gauss_targets_kernel");
}

int hk_l1_normalize_dim1(const float* x, int K, int H, int W, double* out, void* stream) {
  using namespace hk;
  HK_REQUIRE(x && out, "hk_l1_normalize_dim1: null pointer");
  HK_REQUIRE(K > 0 && H > 0 && W > 0 && K <= 65535, "hk_l1_normalize_dim1: bad shape");
  l1_normalize_dim1_kernel<<<dim3(ceil_div(W, 128), K), 128, 0, as_stream(stream)>>>(x, H, W, out);
  return check_launch("l1_normalize_dim1_kernel");
}

size_t hk_bce_workspace_bytes(long long n) {
  if (n <= 0) return 0;
  return (size_t)hk::bce_blocks(n) * sizeof(double);
}

int hk_bce_fwd_bwd(const float* pred, int pred_is_logits, const void* target_or_null, int target_dtype,
                   const float* uv_or_null, int B, int K, int H, int W, float sigma, double* loss, float* grad_logits_or_null, void* ws, size_t ws_bytes,
                   void* stream) {
  using namespace hk;
  HK_REQUIRE(pred && loss && ws, "hk_bce_fwd_bwd: null pointer");
  HK_REQUIRE(B > 0 && K > 0 && H > 0 && W > 0, "hk_bce_fwd_bwd: bad shape");
  HK_REQUIRE((target_or_null != nullptr) != (uv_or_null != nullptr), "hk_bce_fwd_bwd: pass exactly one of target / uv");
  const long long n = (long long)B * K * H * W;
  HK_REQUIRE((n & 3) == 0 && (W & 3) == 0, "hk_bce_fwd_bwd: W must be a multiple of 4");
  HK_REQUIRE((reinterpret_cast<uintptr_t>(pred) & 15) == 0, "hk_bce_fwd_bwd: pred must be 16-byte aligned");
  HK_REQUIRE(!grad_logits_or_null || (reinterpret_cast<uintptr_t>(grad_logits_or_null) & 15) == 0, "hk_bce_fwd_bwd: grad must be 16-byte aligned");
  HK_REQUIRE(ws_bytes >= hk_bce_workspace_bytes(n), "hk_bce_fwd_bwd: workspace too small");
  const int blocks = bce_blocks(n);
  const double inv_n = 1.0 / (double)n;
  const float denom = (float)(2.0 * (double)sigma * (double)sigma);
  double* sums = static_cast<double*>(ws);
  cudaStream_t s = as_stream(stream);
#define HK_BCE_LAUNCH(MODE, TGT, UV)                                                                                   \
  do {                                                                                                                  \
    if (pred_is_logits)                                                                                                 \
      bce_fwd_bwd_kernel<MODE, true><<<blocks, kBceThreads, 0, s>>>(pred, TGT, UV, n, H, W, denom, inv_n, sums, grad_logits_or_null); \
    else                                                                                                                \
      bce_fwd_bwd_kernel<MODE, false><<<blocks, kBceThreads, 0, s>>>(pred, TGT, UV, n, H, W, denom, inv_n, sums, grad_logits_or_null); \
  } while (0)
  if (target_or_null) {
    HK_REQUIRE(target_dtype == HK_F64 || target_dtype == HK_F32, "hk_bce_fwd_bwd: target dtype must be f64 or f32");
    HK_REQUIRE((reinterpret_cast<uintptr_t>(target_or_null) & 15) == 0, "hk_bce_fwd_bwd: target must be 16-byte aligned");
    if (target_dtype == HK_F64) HK_BCE_LAUNCH(0, target_or_null, nullptr);
    else HK_BCE_LAUNCH(1, target_or_null, nullptr);
  } else {
    HK_REQUIRE(sigma > 0.f, "hk_bce_fwd_bwd: sigma must be positive");
    HK_REQUIRE(n < 0xffffffffLL, "hk_bce_fwd_bwd: label mode supports up to 2^32 elements");
    HK_BCE_LAUNCH(2, nullptr, uv_or_null);
  }
#undef HK_BCE_LAUNCH
  int rc = check_launch("bce_fwd_bwd_kernel");
  if (rc) return rc;
  bce_finish_kernel<<<1, 256, 0, s>>>(sums, blocks, inv_n, loss);
  return check_launch("bce_finish_kernel");
}

}  // extern "C"
