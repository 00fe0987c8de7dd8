// Fused Adam over flat fp32 buffers.  Replaces the ~220 small launches per step of
// torch.optim.Adam(lr=1e-4, weight_decay=1e-4) at reference train.py:79,36 with one HBM-bound pass:
// 16 B read + 12 B written per parameter (p, g, m, v in; p, m, v out), 21.8 M parameters -> 0.61 GB per step.
// Arithmetic follows torch's single-tensor Adam (coupled L2, no amsgrad) operation by operation:
//   g += wd*p;  m += (g-m)*(1-b1);  v = v*b2 + (1-b2)*g*g;  p -= (lr/(1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
#include "hk_common.cuh"

namespace hk {

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n4,
            long long n, float lr_over_bc1, float b1, float b2, float eps, float wd, float inv_sqrt_bc2) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pa = reinterpret_cast<float*>(&pp);
    const float* ga = reinterpret_cast<const float*>(&gg);
    float* ma = reinterpret_cast<float*>(&mm);
    float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float grad = fmaf(wd, pa[j], ga[j]);
      ma[j] = fmaf(grad - ma[j], 1.0f - b1, ma[j]);
      va[j] = fmaf(va[j], b2, (1.0f - b2) * grad * grad);
      const float denom = fmaf(sqrtf(va[j]), inv_sqrt_bc2, eps);
      pa[j] = pa[j] - lr_over_bc1 * (ma[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // scalar tail (n % 4)
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    const float grad = fmaf(wd, p[i], g[i]);
    const float mi = fmaf(grad - m[i], 1.0f - b1, m[i]);
    const float vi = fmaf(v[i], b2, (1.0f - b2) * grad * grad);
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - lr_over_bc1 * (mi / fmaf(sqrtf(vi), inv_sqrt_bc2, eps));
  }
}

}  // namespace hk

extern "C" int hk_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                            float beta1, float beta2, float eps, float weight_decay, int step, void* stream) {
  using namespace hk;
  HK_REQUIRE(params && grads && exp_avg && exp_avg_sq, "hk_adam_step: null pointer");
  HK_REQUIRE(n > 0 && step >= 1, "hk_adam_step: n and step must be positive");
  HK_REQUIRE(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
               reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0, "hk_adam_step: buffers must be 16-byte aligned");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float lr_over_bc1 = (float)((double)lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const long long n4 = n >> 2;
  long long blocks = ceil_div_ll(n4 > 0 ? n4 : 1, 256);
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  adam_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(params, grads, exp_avg, exp_avg_sq, n4, n, lr_over_bc1, beta1, beta2, eps,
                                                          weight_decay, inv_sqrt_bc2);
  return check_launch("adam_kernel");
}
