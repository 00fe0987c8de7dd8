// Weight repack: OIHW fp32 state_dict tensor -> O(HW)I ("KRSC") fp32/bf16 GEMM-B operand, and
// BatchNorm(eval) -> per-channel (scale, bias) applied in the conv epilogue.
//
// The reference keeps nn.Conv2d weights as (cout, cin, kh, kw) and applies nn.BatchNorm2d as a separate
// op (src/resnet.py:46,49,139,187).  In eval() that op is y = (x - mean) / sqrt(var + eps) * gamma + beta,
// i.e. y = x * scale + bias with scale = gamma / sqrt(var + eps), bias = beta - mean * scale.  The scale is
// kept OUT of the weights (applied to the fp32 accumulator) so bf16 weight rounding is not perturbed.
#include "hk_common.cuh"

namespace hk {

template <typename OutT>
__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ w_oihw, OutT* __restrict__ w_out, int cout, int cin, int khw) {
  const long long total = (long long)cout * cin * khw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // destination index: ((o * khw + t) * cin + c)
    const int c = (int)(i % cin);
    const long long r = i / cin;
    const int t = (int)(r % khw);
    const int o = (int)(r / khw);
    const float v = __ldg(w_oihw + ((size_t)o * cin + c) * khw + t);
    store_from_float<OutT>(w_out + i, v);
  }
}

// Data-gradient weights: dX = conv(dY, W') with W'[ci, r', s', co] = W[co, ci, kh-1-r', kw-1-s'] (taps flipped, channel roles
// swapped); same pad/dilation as the forward conv when pad == dil*(k-1)/2 (every conv of this network).
__global__ void __launch_bounds__(256)
pack_weights_dgrad_kernel(const float* __restrict__ w_oihw, __nv_bfloat16* __restrict__ w_out, int cout, int cin, int kh, int kw) {
  const int khw = kh * kw;
  const long long total = (long long)cout * cin * khw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // destination index: ((ci * khw + t') * cout + co)
    const int co = (int)(i % cout);
    const long long r = i / cout;
    const int t = (int)(r % khw);
    const int ci = (int)(r / khw);
    w_out[i] = __float2bfloat16_rn(__ldg(w_oihw + ((size_t)co * cin + ci) * khw + (khw - 1 - t)));
  }
}

// All convolutions of the network in ONE launch (training repacks the weights every step): blockIdx.y = item, the x index space of an
// item is [0, n) for the forward layout (cout, khw, cin) followed by [n, 2n) for the data-gradient layout (cin, khw flipped, cout).
__global__ void __launch_bounds__(256) pack_weights_many_kernel(const HkPackItem* __restrict__ items) {
  const HkPackItem it = items[blockIdx.y];
  const long long n = (long long)it.cout * it.cin * it.khw;
  const long long total = it.w_dgrad ? 2 * n : n;
  const float* __restrict__ w = it.w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (i < n) {
      const int c = (int)(i % it.cin);
      const long long r = i / it.cin;
      const int t = (int)(r % it.khw);
      const int o = (int)(r / it.khw);
      static_cast<__nv_bfloat16*>(it.w_fwd)[i] = __float2bfloat16_rn(__ldg(w + ((size_t)o * it.cin + c) * it.khw + t));
    } else {
      const long long j = i - n;
      const int co = (int)(j % it.cout);
      const long long r = j / it.cout;
      const int t = (int)(r % it.khw);
      const int ci = (int)(r / it.khw);
      static_cast<__nv_bfloat16*>(it.w_dgrad)[j] = __float2bfloat16_rn(__ldg(w + ((size_t)co * it.cin + ci) * it.khw + (it.khw - 1 - t)));
    }
  }
}

__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                               const float* __restrict__ var, float eps, int cout, float* __restrict__ scale,
                               float* __restrict__ bias) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout) return;
  if (gamma) {
    // same op order as ATen's eval batch_norm: invstd = 1/sqrt(var+eps); w*invstd; b - mean*w*invstd
    const float invstd = 1.0f / sqrtf(var[i] + eps);
    const float s = gamma[i] * invstd;
    scale[i] = s;
    bias[i] = beta[i] - mean[i] * s;
  } else {
    scale[i] = 1.0f;
    bias[i] = 0.0f;
  }
}

}  // namespace hk

extern "C" int hk_pack_conv_weights(const float* w_oihw, const float* bn_gamma, const float* bn_beta, const float* bn_mean,
                                    const float* bn_var, float bn_eps, int cout, int cin, int kh, int kw, int w_dtype,
                                    void* w_out, float* scale_out, float* bias_out, void* stream) {
  using namespace hk;
  HK_REQUIRE(w_oihw && w_out && scale_out && bias_out, "hk_pack_conv_weights: null pointer");
  HK_REQUIRE(cout > 0 && cin > 0 && kh > 0 && kw > 0, "hk_pack_conv_weights: bad shape");
  const bool all = bn_gamma && bn_beta && bn_mean && bn_var;
  const bool none = !bn_gamma && !bn_beta && !bn_mean && !bn_var;
  HK_REQUIRE(all || none, "hk_pack_conv_weights: pass all four BatchNorm vectors or none");
  HK_REQUIRE(w_dtype == HK_BF16 || w_dtype == HK_F32, "hk_pack_conv_weights: w_dtype must be bf16 or f32");
  cudaStream_t s = as_stream(stream);
  const long long total = (long long)cout * cin * kh * kw;
  int blocks = (int)ceil_div_ll(total, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (w_dtype == HK_BF16)
    pack_weights_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(w_oihw, static_cast<__nv_bfloat16*>(w_out), cout, cin, kh * kw);
  else
    pack_weights_kernel<float><<<blocks, 256, 0, s>>>(w_oihw, static_cast<float*>(w_out), cout, cin, kh * kw);
  int rc = check_launch("pack_weights_kernel");
  if (rc) return rc;
  bn_fold_kernel<<<ceil_div(cout, 128), 128, 0, s>>>(bn_gamma, bn_beta, bn_mean, bn_var, bn_eps, cout, scale_out, bias_out);
  return check_launch("bn_fold_kernel");
}

extern "C" int hk_pack_conv_weights_dgrad(const float* w_oihw, int cout, int cin, int kh, int kw, void* w_out, void* stream) {
  using namespace hk;
  HK_REQUIRE(w_oihw && w_out, "hk_pack_conv_weights_dgrad: null pointer");
  HK_REQUIRE(cout > 0 && cin > 0 && kh > 0 && kw > 0, "hk_pack_conv_weights_dgrad: bad shape");
  const long long total = (long long)cout * cin * kh * kw;
  int blocks = (int)ceil_div_ll(total, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  pack_weights_dgrad_kernel<<<blocks, 256, 0, as_stream(stream)>>>(w_oihw, static_cast<__nv_bfloat16*>(w_out), cout, cin, kh, kw);
  return check_launch("pack_weights_dgrad_kernel");
}

extern "C" int hk_pack_conv_weights_many(const HkPackItem* items_dev, int n_items, long long max_elems, void* stream) {
  using namespace hk;
  HK_REQUIRE(items_dev && n_items > 0 && max_elems > 0, "hk_pack_conv_weights_many: bad argument");
  long long bx = ceil_div_ll(2 * max_elems, 256 * 8);  // 8 elements per thread for the largest item
  if (bx > 1024) bx = 1024;
  if (bx < 1) bx = 1;
  pack_weights_many_kernel<<<dim3((unsigned)bx, (unsigned)n_items), 256, 0, as_stream(stream)>>>(items_dev);
  return check_launch("pack_weights_many_kernel");
}
