// Train-mode BatchNorm2d (+ residual add, ReLU) forward and backward on NHWC bf16 activations.
//
// Reference: nn.BatchNorm2d in train() mode at src/resnet.py:46,49,139,187 as called from BasicBlock.forward
// (src/resnet.py:53-69) and ResNet.forward (:199-201), with autograd's backward (train.py:35).
//   forward : mu = mean_p y, var = mean_p (y-mu)^2 (biased), xhat = (y-mu)*rsqrt(var+eps),
//             out = relu(gamma*xhat + beta [+ residual]); running stats updated with momentum (unbiased var)
//   backward: d' = dout * [out > 0];  dbeta = sum d';  dgamma = sum d'*xhat;
//             dy = gamma*invstd * (d' - dbeta/N - xhat*dgamma/N)
// All four kernels are HBM-bound streaming passes (16-byte vectors of 8 bf16 channels); the reductions are two-level
// (fixed grid of per-block fp32 partials, then a double-precision finalize) and therefore deterministic -- no atomics.
#include "hk_common.cuh"

namespace hk {

constexpr int BN_THREADS = 256;
constexpr int BN_MAX_BLOCKS = 592;  // 148 SMs x 4

struct Vec8 {
  float v[8];
};
__device__ __forceinline__ Vec8 load8(const __nv_bfloat16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  Vec8 r;
  unpack_bf16x2(u.x, r.v[0], r.v[1]);
  unpack_bf16x2(u.y, r.v[2], r.v[3]);
  unpack_bf16x2(u.z, r.v[4], r.v[5]);
  unpack_bf16x2(u.w, r.v[6], r.v[7]);
  return r;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

// Block-level reduction of two per-thread 8-channel accumulators over the row lanes of the block, written as
// partial[blockIdx][0..1][C].  Thread layout: cg = tid % CG (channel group of 8), ty = tid / CG.
__device__ __forceinline__ void block_reduce_2x8(const float (&a)[8], const float (&b)[8], int C, float* __restrict__ partial) {
  __shared__ float sh[2 * BN_THREADS * 8];  // [2][rows][C] with rows*C = 256*8
  const int CG = C >> 3, cg = threadIdx.x % CG, ty = threadIdx.x / CG, rows = BN_THREADS / CG;
  float* s0 = sh + (ty * C + cg * 8);
  float* s1 = s0 + BN_THREADS * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) { s0[j] = a[j]; s1[j] = b[j]; }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += BN_THREADS) {
    const int which = c / C, ch = c - which * C;
    const float* src = sh + which * BN_THREADS * 8 + ch;
    float acc = 0.f;
    for (int r = 0; r < rows; ++r) acc += src[r * C];
    partial[(size_t)blockIdx.x * 2 * C + c] = acc;
  }
}

// ---- forward statistics ----
__global__ void __launch_bounds__(BN_THREADS) bn_stats_partial_kernel(const __nv_bfloat16* __restrict__ y, long long P, int C,
                                                                     float* __restrict__ partial) {
  const int CG = C >> 3, cg = threadIdx.x % CG, ty = threadIdx.x / CG, rows = BN_THREADS / CG;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long r = (long long)blockIdx.x * rows + ty; r < P; r += (long long)gridDim.x * rows) {
    const Vec8 v = load8(y + r * C + cg * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += v.v[j]; q[j] = fmaf(v.v[j], v.v[j], q[j]); }
  }
  block_reduce_2x8(s, q, C, partial);
}

// Finalize kernels: 1024 threads = 32 channels x 32 slices of the per-block partials (coalesced across channels; every thread has
// up to 8 independent loads in flight and walks <= 19 of the 592 partials -- one thread per channel walking them serially cost
// 170 us per launch, 8 slices with 2-deep unrolling still 15 us: the loop is pure L2 latency).  Fixed order => deterministic.
constexpr int BN_FIN_SLICES = 32;
constexpr int BN_FIN_THREADS = 32 * BN_FIN_SLICES;
__device__ __forceinline__ void bn_sum_partials(const float* __restrict__ partial, int nblocks, int C, int c, double& s, double& q) {
  __shared__ double sh[2][BN_FIN_SLICES][33];
  const int lane_c = threadIdx.x & 31, slice = threadIdx.x >> 5;
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    int blk = slice;
    for (; blk + 3 * BN_FIN_SLICES < nblocks; blk += 4 * BN_FIN_SLICES) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] += partial[(size_t)(blk + u * BN_FIN_SLICES) * 2 * C + c];
        b[u] += partial[(size_t)(blk + u * BN_FIN_SLICES) * 2 * C + C + c];
      }
    }
    for (; blk < nblocks; blk += BN_FIN_SLICES) {
      a[0] += partial[(size_t)blk * 2 * C + c];
      b[0] += partial[(size_t)blk * 2 * C + C + c];
    }
  }
  sh[0][slice][lane_c] = ((double)a[0] + (double)a[1]) + ((double)a[2] + (double)a[3]);
  sh[1][slice][lane_c] = ((double)b[0] + (double)b[1]) + ((double)b[2] + (double)b[3]);
  __syncthreads();
  double s4[4] = {0.0, 0.0, 0.0, 0.0}, q4[4] = {0.0, 0.0, 0.0, 0.0};
  if (threadIdx.x < 32) {
#pragma unroll
    for (int i = 0; i < BN_FIN_SLICES; ++i) { s4[i & 3] += sh[0][i][lane_c]; q4[i & 3] += sh[1][i][lane_c]; }
  }
  s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
  q = (q4[0] + q4[1]) + (q4[2] + q4[3]);
}

// mean / invstd, the fused affine (scale, shift) of the apply pass, running statistics
__global__ void __launch_bounds__(BN_FIN_THREADS) bn_stats_finalize_kernel(const float* __restrict__ partial, int nblocks, long long P, int C, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, float* __restrict__ running_mean,
                                         float* __restrict__ running_var, float momentum, float eps, float* __restrict__ mean_out,
                                         float* __restrict__ invstd_out, float* __restrict__ scale_out, float* __restrict__ shift_out) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  double s, q;
  bn_sum_partials(partial, nblocks, C, c, s, q);
  if (c >= C || threadIdx.x >= 32) return;
  const double n = (double)P;
  const double mean = s / n;
  double var = q / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
  mean_out[c] = (float)mean;
  invstd_out[c] = invstd;
  scale_out[c] = g * invstd;
  shift_out[c] = bt - (float)mean * g * invstd;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
  if (running_var) {
    const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// ---- forward apply: out = relu?(y*scale + shift [+ residual]) ----
__global__ void __launch_bounds__(BN_THREADS) bn_apply_fwd_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                                                                 const float* __restrict__ shift,
                                                                 const __nv_bfloat16* __restrict__ residual, int relu,
                                                                 __nv_bfloat16* __restrict__ out, uint8_t* __restrict__ relu_bits,
                                                                 long long nvec, int C) {
  // the grid stride (gridDim*256) is a multiple of C/8, so a thread keeps its 8 channels: coefficients live in registers
  const int cg = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) % (C >> 3));
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = __ldg(scale + cg * 8 + j); sh[j] = __ldg(shift + cg * 8 + j); }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const Vec8 v = load8(y + i * 8);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(v.v[j], sc[j], sh[j]);
    if (residual) {
      const Vec8 r = load8(residual + i * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += r.v[j];
    }
    if (relu) {
      if (relu_bits) {  // bit j = [out_j > 0]: the ReLU mask of the backward pass at 1/16 of the bytes of `out`
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) m |= (o[j] > 0.f ? 1u : 0u) << j;
        relu_bits[i] = (uint8_t)m;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
    }
    store8(out + i * 8, o);
  }
}

// ---- backward reduce: s1 = sum d', s2 = sum d'*xhat with d' = dout*[out>0] ----
template <bool BITS>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_partial_kernel(const __nv_bfloat16* __restrict__ dout,
                                                                   const void* __restrict__ out_mask,
                                                                   const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                                                                   const float* __restrict__ invstd, long long P, int C,
                                                                   float* __restrict__ partial) {
  const int CG = C >> 3, cg = threadIdx.x % CG, ty = threadIdx.x / CG, rows = BN_THREADS / CG;
  float mu[8], is[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { mu[j] = mean[cg * 8 + j]; is[j] = invstd[cg * 8 + j]; }
  float s1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long r = (long long)blockIdx.x * rows + ty; r < P; r += (long long)gridDim.x * rows) {
    const long long off = r * C + cg * 8;
    Vec8 d = load8(dout + off);
    if (out_mask) {
      if constexpr (BITS) {
        const uint32_t m = static_cast<const uint8_t*>(out_mask)[off >> 3];
#pragma unroll
        for (int j = 0; j < 8; ++j) d.v[j] = (m >> j) & 1u ? d.v[j] : 0.f;
      } else {
        const Vec8 m = load8(static_cast<const __nv_bfloat16*>(out_mask) + off);
#pragma unroll
        for (int j = 0; j < 8; ++j) d.v[j] = m.v[j] > 0.f ? d.v[j] : 0.f;
      }
    }
    const Vec8 v = load8(y + off);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s1[j] += d.v[j];
      s2[j] = fmaf(d.v[j], (v.v[j] - mu[j]) * is[j], s2[j]);
    }
  }
  block_reduce_2x8(s1, s2, C, partial);
}

// dgamma, dbeta (fp32 parameter gradients; accumulate != 0 adds to what is there) and the coefficients of the apply pass:
//   coef[0][c] = gamma*invstd, coef[1][c] = dbeta/N, coef[2][c] = dgamma/N
__global__ void __launch_bounds__(BN_FIN_THREADS) bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, long long P, int C, const float* __restrict__ gamma,
                                       const float* __restrict__ invstd, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       int accumulate, float* __restrict__ coef) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  double s1, s2;
  bn_sum_partials(partial, nblocks, C, c, s1, s2);
  if (c >= C || threadIdx.x >= 32) return;
  const float g = gamma ? gamma[c] : 1.f;
  if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)s1;
  if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)s2;
  coef[c] = g * invstd[c];
  coef[C + c] = (float)(s1 / (double)P);
  coef[2 * C + c] = (float)(s2 / (double)P);
}

// dy = coef0 * (d' - coef1 - xhat*coef2) = k0*d' + k1*y + k2 per channel; optionally also writes d' (the gradient that flows
// into the shortcut branch).  The thread's 8 channels are loop-invariant (see bn_apply_fwd_kernel): 24 coefficients in registers.
template <bool BITS>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dout,
                                                                 const void* __restrict__ out_mask,
                                                                 const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                                                                 const float* __restrict__ invstd, const float* __restrict__ coef,
                                                                 __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dmasked,
                                                                 long long nvec, int C) {
  const int c0 = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) % (C >> 3)) * 8;
  float k0[8], k1[8], k2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    const float a = __ldg(coef + c), is = __ldg(invstd + c), mu = __ldg(mean + c);
    k0[j] = a;
    k1[j] = -a * __ldg(coef + 2 * C + c) * is;
    k2[j] = -a * __ldg(coef + C + c) - k1[j] * mu;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    Vec8 d = load8(dout + i * 8);
    if (out_mask) {
      if constexpr (BITS) {
        const uint32_t m = static_cast<const uint8_t*>(out_mask)[i];
#pragma unroll
        for (int j = 0; j < 8; ++j) d.v[j] = (m >> j) & 1u ? d.v[j] : 0.f;
      } else {
        const Vec8 m = load8(static_cast<const __nv_bfloat16*>(out_mask) + i * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) d.v[j] = m.v[j] > 0.f ? d.v[j] : 0.f;
      }
    }
    if (dmasked) store8(dmasked + i * 8, d.v);
    const Vec8 v = load8(y + i * 8);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(k0[j], d.v[j], fmaf(k1[j], v.v[j], k2[j]));
    store8(dy + i * 8, o);
  }
}

static int bn_grid_rows(long long P, int C) {
  const int rows = BN_THREADS / (C >> 3);
  long long blocks = ceil_div_ll(P, (long long)rows * 16);  // >= 16 row iterations per block: few partials to finalize
  if (blocks > BN_MAX_BLOCKS) blocks = BN_MAX_BLOCKS;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}
static int bn_grid_elems(long long nvec) {
  long long blocks = ceil_div_ll(nvec, BN_THREADS);
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  return (int)blocks;
}
static bool bn_c_ok(int C) { return C >= 8 && C <= 2048 && (C & 7) == 0 && (BN_THREADS % (C >> 3)) == 0; }

}  // namespace hk

extern "C" {

size_t hk_bn_workspace_bytes(int C) { return (size_t)hk::BN_MAX_BLOCKS * 2 * (size_t)C * sizeof(float); }

int hk_bn_train_stats(const void* y, long long P, int C, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, float momentum, float eps, float* mean_out, float* invstd_out, float* scale_out,
                      float* shift_out, void* ws, size_t ws_bytes, void* stream) {
  using namespace hk;
  HK_REQUIRE(y && mean_out && invstd_out && scale_out && shift_out && ws, "hk_bn_train_stats: null pointer");
  HK_REQUIRE(P > 0 && bn_c_ok(C), "hk_bn_train_stats: unsupported shape P=%lld C=%d (C must be a power-of-two multiple of 8 up to 2048)", P, C);
  HK_REQUIRE(ws_bytes >= hk_bn_workspace_bytes(C), "hk_bn_train_stats: workspace too small");
  HK_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0, "hk_bn_train_stats: y must be 16-byte aligned");
  const int blocks = bn_grid_rows(P, C);
  bn_stats_partial_kernel<<<blocks, BN_THREADS, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(y), P, C, static_cast<float*>(ws));
  int rc = check_launch("bn_stats_partial_kernel");
  if (rc) return rc;
  bn_stats_finalize_kernel<<<ceil_div(C, 32), BN_FIN_THREADS, 0, as_stream(stream)>>>(static_cast<const float*>(ws), blocks, P, C, gamma, beta,
                                                                           running_mean, running_var, momentum, eps, mean_out,
                                                                           invstd_out, scale_out, shift_out);
  return check_launch("bn_stats_finalize_kernel");
}

int hk_bn_apply_fwd(const void* y, const float* scale, const float* shift, const void* residual_or_null, int relu, void* out,
                    void* relu_bits_or_null, long long P, int C, void* stream) {
  using namespace hk;
  HK_REQUIRE(y && scale && shift && out, "hk_bn_apply_fwd: null pointer");
  HK_REQUIRE(P > 0 && C >= 8 && (C & 7) == 0, "hk_bn_apply_fwd: bad shape");
  const long long nvec = P * (C >> 3);
  bn_apply_fwd_kernel<<<bn_grid_elems(nvec), BN_THREADS, 0, as_stream(stream)>>>(
      static_cast<const __nv_bfloat16*>(y), scale, shift, static_cast<const __nv_bfloat16*>(residual_or_null), relu,
      static_cast<__nv_bfloat16*>(out), static_cast<uint8_t*>(relu_bits_or_null), nvec, C);
  return check_launch("bn_apply_fwd_kernel");
}

int hk_bn_train_bwd(const void* dout, const void* out_mask_or_null, int mask_is_bits, const void* y, const float* mean, const float* invstd,
                    const float* gamma, long long P, int C, float* dgamma, float* dbeta, int accumulate, void* dy,
                    void* dmasked_or_null, void* ws, size_t ws_bytes, void* stream) {
  using namespace hk;
  HK_REQUIRE(dout && y && mean && invstd && dy && ws, "hk_bn_train_bwd: null pointer");
  HK_REQUIRE(P > 0 && bn_c_ok(C), "hk_bn_train_bwd: unsupported shape P=%lld C=%d", P, C);
  HK_REQUIRE(ws_bytes >= hk_bn_workspace_bytes(C) + 3 * (size_t)C * sizeof(float), "hk_bn_train_bwd: workspace too small");
  const int blocks = bn_grid_rows(P, C);
  float* partial = static_cast<float*>(ws);
  float* coef = partial + (size_t)BN_MAX_BLOCKS * 2 * C;
  const __nv_bfloat16* d = static_cast<const __nv_bfloat16*>(dout);
  const void* m = out_mask_or_null;
  const __nv_bfloat16* yy = static_cast<const __nv_bfloat16*>(y);
  if (mask_is_bits) bn_bwd_partial_kernel<true><<<blocks, BN_THREADS, 0, as_stream(stream)>>>(d, m, yy, mean, invstd, P, C, partial);
  else bn_bwd_partial_kernel<false><<<blocks, BN_THREADS, 0, as_stream(stream)>>>(d, m, yy, mean, invstd, P, C, partial);
  int rc = check_launch("bn_bwd_partial_kernel");
  if (rc) return rc;
  bn_bwd_finalize_kernel<<<ceil_div(C, 32), BN_FIN_THREADS, 0, as_stream(stream)>>>(partial, blocks, P, C, gamma, invstd, dgamma, dbeta, accumulate, coef);
  rc = check_launch("bn_bwd_finalize_kernel");
  if (rc) return rc;
  const long long nvec = P * (C >> 3);
  if (mask_is_bits)
    bn_bwd_apply_kernel<true><<<bn_grid_elems(nvec), BN_THREADS, 0, as_stream(stream)>>>(d, m, yy, mean, invstd, coef, static_cast<__nv_bfloat16*>(dy),
                                                                                       static_cast<__nv_bfloat16*>(dmasked_or_null), nvec, C);
  else
    bn_bwd_apply_kernel<false><<<bn_grid_elems(nvec), BN_THREADS, 0, as_stream(stream)>>>(d, m, yy, mean, invstd, coef, static_cast<__nv_bfloat16*>(dy),
                                                                                        static_cast<__nv_bfloat16*>(dmasked_or_null), nvec, C);
  return check_launch("bn_bwd_apply_kernel");
}

}  // extern "C"
