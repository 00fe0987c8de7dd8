// fp32 CUDA-core implicit-GEMM convolution + fused affine/residual/ReLU epilogue, and NHWC max-pool.
//
// This is the fp32 CORRECTNESS mode of the backbone (tcgen05 has no fp32 MMA kind and
// TF32 cannot hold the 1e-4 bar), and the stem conv (Cin=3, K=147 is too thin for TMA im2col) in both
// modes.  GEMM view of nn.Conv2d (src/resnet.py:36,137,185): M = B*Ho*Wo output pixels, N = Cout,
// K = kh*kw*Cin with the channel index fastest (weights repacked to O(HW)I by hk_pack_conv_weights).
//   CTA tile 64(M) x 64(N) x 16(K), 256 threads, 4x4 register micro-tile, register-staged prefetch of
//   the next K slab while the current one is consumed from shared memory.
#include "hk_common.cuh"

namespace hk {

struct ConvFfmaArgs {
  const void* x;
  const float* w;  // (Cout, K) fp32
  const float* scale;
  const float* bias;
  const void* residual;  // NHWC, OutT, or null
  void* y;               // NHWC, OutT
  int B, H, W, Cin, Ho, Wo, Cout, kh, kw, stride, pad, dil, relu;
  int M, Ktot;
};

constexpr int FT_M = 64, FT_N = 64, FT_K = 16, FT_PAD = 4, FT_THREADS = 256;

template <typename InT, typename OutT, bool NCHW, bool VEC>
__global__ void __launch_bounds__(FT_THREADS)
conv_ffma_kernel(const ConvFfmaArgs a) {
  __shared__ __align__(16) float As[FT_K][FT_M + FT_PAD];
  __shared__ __align__(16) float Bs[FT_K][FT_N + FT_PAD];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * FT_M, n0 = blockIdx.y * FT_N;
  const InT* __restrict__ x = static_cast<const InT*>(a.x);

  // ---- loader role: one (row, 4 consecutive k) strip of A and one of B per thread ----
  const int lrow = tid >> 2, lk = (tid & 3) << 2;
  const int m = m0 + lrow;
  const bool mvalid = m < a.M;
  int pb = 0, iy0 = 0, ix0 = 0;
  if (mvalid) {
    pb = m / (a.Ho * a.Wo);
    const int rem = m - pb * a.Ho * a.Wo;
    const int oy = rem / a.Wo, ox = rem - oy * a.Wo;
    iy0 = oy * a.stride - a.pad;
    ix0 = ox * a.stride - a.pad;
  }
  const int nrow = n0 + lrow;
  const bool nvalid = nrow < a.Cout;
  const float* __restrict__ wrow = a.w + (size_t)(nvalid ? nrow : 0) * a.Ktot;

  float ra[4], rb[4];
  auto fetch = [&](int kt) {
    const int k = kt * FT_K + lk;
#pragma unroll
    for (int j = 0; j < 4; ++j) { ra[j] = 0.f; rb[j] = 0.f; }
    if (VEC) {
      // Cin % 4 == 0: the four k share one tap and are channel-contiguous
      if (k < a.Ktot) {
        const int tap = k / a.Cin, c = k - tap * a.Cin;
        const int r = tap / a.kw, s = tap - r * a.kw;
        const int iy = iy0 + r * a.dil, ix = ix0 + s * a.dil;
        if (mvalid && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) {
          const InT* p = x + (((size_t)pb * a.H + iy) * a.W + ix) * a.Cin + c;
          if constexpr (sizeof(InT) == 4) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(p));
            ra[0] = q.x; ra[1] = q.y; ra[2] = q.z; ra[3] = q.w;
          } else {
            const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
            unpack_bf16x2(q.x, ra[0], ra[1]);
            unpack_bf16x2(q.y, ra[2], ra[3]);
          }
        }
        if (nvalid) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(wrow + k));
          rb[0] = q.x; rb[1] = q.y; rb[2] = q.z; rb[3] = q.w;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = k + j;
        if (kk < a.Ktot) {
          const int tap = kk / a.Cin, c = kk - tap * a.Cin;
          const int r = tap / a.kw, s = tap - r * a.kw;
          const int iy = iy0 + r * a.dil, ix = ix0 + s * a.dil;
          if (mvalid && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) {
            const size_t idx = NCHW ? (((size_t)pb * a.Cin + c) * a.H + iy) * a.W + ix
                                    : (((size_t)pb * a.H + iy) * a.W + ix) * a.Cin + c;
            ra[j] = load_as_float<InT>(x + idx);
          }
          if (nvalid) rb[j] = __ldg(wrow + kk);
        }
      }
    }
  };

  // ---- compute role ----
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ktiles = (a.Ktot + FT_K - 1) / FT_K;
  fetch(0);
  for (int kt = 0; kt < ktiles; ++kt) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[lk + j][lrow] = ra[j];
      Bs[lk + j][lrow] = rb[j];
    }
    __syncthreads();
    if (kt + 1 < ktiles) fetch(kt + 1);
#pragma unroll
    for (int k = 0; k < FT_K; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue: y = act(acc * scale + bias + residual) ----
  const int nb = n0 + tx * 4;
  if (nb >= a.Cout) return;  // Cout % 4 == 0 (host-checked)
  const float4 sc = __ldg(reinterpret_cast<const float4*>(a.scale + nb));
  const float4 bi = __ldg(reinterpret_cast<const float4*>(a.bias + nb));
  const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, biv[4] = {bi.x, bi.y, bi.z, bi.w};
  OutT* __restrict__ y = static_cast<OutT*>(a.y);
  const OutT* __restrict__ res = static_cast<const OutT*>(a.residual);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int mm = m0 + ty * 4 + i;
    if (mm >= a.M) continue;
    const size_t off = (size_t)mm * a.Cout + nb;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = fmaf(acc[i][j], scv[j], biv[j]);
    if (res) {
      if constexpr (sizeof(OutT) == 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(res + off));
        v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w;
      } else {
        const uint2 q = __ldg(reinterpret_cast<const uint2*>(res + off));
        float r0, r1, r2, r3;
        unpack_bf16x2(q.x, r0, r1);
        unpack_bf16x2(q.y, r2, r3);
        v[0] += r0; v[1] += r1; v[2] += r2; v[3] += r3;
      }
    }
    if (a.relu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if constexpr (sizeof(OutT) == 4) {
      *reinterpret_cast<float4*>(y + off) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      *reinterpret_cast<uint2*>(y + off) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
    }
  }
}

template <typename InT, typename OutT>
static void launch_ffma(const ConvFfmaArgs& a, bool nchw, cudaStream_t s) {
  dim3 grid(ceil_div(a.M, FT_M), ceil_div(a.Cout, FT_N));
  const bool vec = !nchw && (a.Cin % 4 == 0);
  if (nchw)
    conv_ffma_kernel<InT, OutT, true, false><<<grid, FT_THREADS, 0, s>>>(a);
  else if (vec)
    conv_ffma_kernel<InT, OutT, false, true><<<grid, FT_THREADS, 0, s>>>(a);
  else
    conv_ffma_kernel<InT, OutT, false, false><<<grid, FT_THREADS, 0, s>>>(a);
}

int conv_ffma_launch(const HkConvDesc& d, const void* x, const void* w, const float* scale, const float* bias,
                     const void* residual, void* y, cudaStream_t s) {
  HK_REQUIRE(d.out_c % 4 == 0, "conv(FFMA): out_c=%d must be a multiple of 4", d.out_c);
  HK_REQUIRE((long long)d.batch * d.out_h * d.out_w < 0x7fffffffLL, "conv(FFMA): too many output pixels");
  HK_REQUIRE(!(d.in_is_nchw && d.in_dtype != HK_F32), "conv(FFMA): NCHW input must be fp32");
  ConvFfmaArgs a;
  a.x = x; a.w = static_cast<const float*>(w); a.scale = scale; a.bias = bias; a.residual = residual; a.y = y;
  a.B = d.batch; a.H = d.in_h; a.W = d.in_w; a.Cin = d.in_c; a.Ho = d.out_h; a.Wo = d.out_w; a.Cout = d.out_c;
  a.kh = d.kh; a.kw = d.kw; a.stride = d.stride; a.pad = d.pad; a.dil = d.dil; a.relu = d.relu;
  a.M = d.batch * d.out_h * d.out_w;
  a.Ktot = d.kh * d.kw * d.in_c;
  const bool nchw = d.in_is_nchw != 0;
  if (d.in_dtype == HK_F32 && d.out_dtype == HK_F32) launch_ffma<float, float>(a, nchw, s);
  else if (d.in_dtype == HK_F32 && d.out_dtype == HK_BF16) launch_ffma<float, __nv_bfloat16>(a, nchw, s);
  else if (d.in_dtype == HK_BF16 && d.out_dtype == HK_BF16) launch_ffma<__nv_bfloat16, __nv_bfloat16>(a, nchw, s);
  else if (d.in_dtype == HK_BF16 && d.out_dtype == HK_F32) launch_ffma<__nv_bfloat16, float>(a, nchw, s);
  else return fail(HK_ERR_BAD_ARG, "conv(FFMA): unsupported dtype combination in=%d out=%d", d.in_dtype, d.out_dtype);
  return check_launch("conv_ffma_kernel");
}

// ---- MaxPool2d(3, stride 2, pad 1), NHWC; padding behaves as -inf (torch semantics) ----
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C, int Ho, int Wo) {
  const unsigned cg = (unsigned)(C / VEC);
  const unsigned total = (unsigned)B * Ho * Wo * cg;  // host guarantees < 2^32
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const unsigned g = t % cg;
    unsigned r = t / cg;
    const int ox = (int)(r % (unsigned)Wo); r /= (unsigned)Wo;
    const int oy = (int)(r % (unsigned)Ho);
    const int b = (int)(r / (unsigned)Ho);
    float best[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) best[j] = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int iy = oy * 2 - 1 + dy;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int ix = ox * 2 - 1 + dx;
        if (ix < 0 || ix >= W) continue;
        const T* p = x + (((size_t)b * H + iy) * W + ix) * C + g * VEC;
        if constexpr (sizeof(T) == 4) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(p));
          best[0] = fmaxf(best[0], q.x); best[1] = fmaxf(best[1], q.y);
          best[2] = fmaxf(best[2], q.z); best[3] = fmaxf(best[3], q.w);
        } else {
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
          const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float lo, hi;
            unpack_bf16x2(qq[j], lo, hi);
            best[2 * j] = fmaxf(best[2 * j], lo);
            best[2 * j + 1] = fmaxf(best[2 * j + 1], hi);
          }
        }
      }
    }
    T* q = y + (((size_t)b * Ho + oy) * Wo + ox) * C + g * VEC;
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(q) = make_float4(best[0], best[1], best[2], best[3]);
    } else {
      *reinterpret_cast<uint4*>(q) = make_uint4(pack_bf16x2(best[0], best[1]), pack_bf16x2(best[2], best[3]),
                                                pack_bf16x2(best[4], best[5]), pack_bf16x2(best[6], best[7]));
    }
  }
}

}  // namespace hk

extern "C" int hk_maxpool3x3s2_fwd(const void* x, void* y, int dtype, int batch, int in_h, int in_w, int c, int out_h,
                                   int out_w, void* stream) {
  using namespace hk;
  HK_REQUIRE(x && y, "hk_maxpool3x3s2_fwd: null pointer");
  HK_REQUIRE(batch > 0 && in_h > 0 && in_w > 0 && c > 0, "hk_maxpool3x3s2_fwd: bad shape");
  HK_REQUIRE(out_h == (in_h + 2 - 3) / 2 + 1 && out_w == (in_w + 2 - 3) / 2 + 1, "hk_maxpool3x3s2_fwd: out shape mismatch");
  HK_REQUIRE(dtype == HK_F32 || dtype == HK_BF16, "hk_maxpool3x3s2_fwd: dtype must be f32 or bf16");
  const int vec = dtype == HK_F32 ? 4 : 8;
  HK_REQUIRE(c % vec == 0, "hk_maxpool3x3s2_fwd: C=%d must be a multiple of %d", c, vec);
  const long long total = (long long)batch * out_h * out_w * (c / vec);
  HK_REQUIRE(total < 0xffffffffLL, "hk_maxpool3x3s2_fwd: tensor too large for 32-bit indexing");
  int blocks = (int)ceil_div_ll(total, 256);
  const int cap = sm_count() * 32;
  if (blocks > cap) blocks = cap;
  if (dtype == HK_F32)
    maxpool3x3s2_kernel<float, 4><<<blocks, 256, 0, as_stream(stream)>>>(static_cast<const float*>(x), static_cast<float*>(y),
                                                                        batch, in_h, in_w, c, out_h, out_w);
  else
    maxpool3x3s2_kernel<__nv_bfloat16, 8><<<blocks, 256, 0, as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), batch, in_h, in_w, c, out_h, out_w);
  return check_launch("maxpool3x3s2_kernel");
}
