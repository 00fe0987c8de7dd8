// Exact, order-independent per-channel accumulators of the train-mode BatchNorm reductions, and the conv-epilogue helper that feeds
// them (reference: nn.BatchNorm2d under model.train(), src/resnet.py:46,49,139,187).
//
// BnAcc: a 128-bit two's-complement fixed-point integer (unit 2^-50) updated with 64-bit integer atomics.  Integer addition is
// associative, so the total is the exact sum of the fp32 partial sums added to it whatever order the blocks retire in:
// deterministic like a fixed-order two-level reduction, but without the per-block scratch and without a finalize launch.
#pragma once
#include "hk_common.cuh"

namespace hk {

struct alignas(32) BnAcc {
  unsigned long long lo, hi;   // two's-complement 128-bit integer, unit 2^-50
  unsigned long long poison;   // != 0: a partial sum was Inf/NaN or beyond the accumulator range -> the total reads as NaN
  unsigned long long pad;
};
constexpr int BN_ACC_FRAC_BITS = 50;

// p as a 128-bit fixed-point integer (lo, hi).  Returns false when there is nothing to add (zero, below resolution, or poisoned:
// Inf/NaN/out-of-range values mark the accumulator instead).
__device__ __forceinline__ bool bn_acc_split(BnAcc* a, float p, unsigned long long& lo, unsigned long long& hi) {
  const uint32_t bits = __float_as_uint(p);
  const uint32_t ex = (bits >> 23) & 0xffu;
  uint32_t man = bits & 0x7fffffu;
  if (ex == 0xffu) { atomicOr(&a->poison, 1ull); return false; }
  if (ex == 0u && man == 0u) return false;
  const int e = ex ? (int)ex : 1;
  if (ex) man |= 0x800000u;
  const int shift = e - 150 + BN_ACC_FRAC_BITS;     // p = man * 2^(e-150)
  unsigned __int128 v;
  if (shift >= 0) {
    if (shift > 100) { atomicOr(&a->poison, 2ull); return false; }   // |p| >= 2^74: far outside anything a finite BatchNorm produces
    v = static_cast<unsigned __int128>(man) << shift;
  } else {
    if (shift <= -24) return false;                  // below the accumulator's resolution (2^-50)
    v = man >> (-shift);
  }
  if (bits >> 31) v = static_cast<unsigned __int128>(0) - v;
  lo = static_cast<unsigned long long>(v);
  hi = static_cast<unsigned long long>(v >> 64);
  return true;
}

__device__ __forceinline__ void bn_acc_add(BnAcc* a, float p) {
  unsigned long long lo, hi;
  if (!bn_acc_split(a, p, lo, hi)) return;
  const unsigned long long old = atomicAdd(&a->lo, lo);
  const unsigned long long hi2 = hi + ((old + lo) < lo ? 1ull : 0ull);   // carry out of the low word: exact whatever the order
  if (hi2) atomicAdd(&a->hi, hi2);
}

__device__ __forceinline__ double bn_acc_read(const BnAcc* a) {
  const unsigned long long lo = a->lo, hi = a->hi;
  if (a->poison) return __longlong_as_double(0x7ff8000000000000ll);
  unsigned __int128 v = (static_cast<unsigned __int128>(hi) << 64) | lo;
  const bool neg = (hi >> 63) != 0;
  if (neg) v = static_cast<unsigned __int128>(0) - v;
  const double d = (static_cast<double>(static_cast<unsigned long long>(v >> 64)) * 18446744073709551616.0 +
                    static_cast<double>(static_cast<unsigned long long>(v))) * (1.0 / 1125899906842624.0);   // 2^-50
  return neg ? -d : d;
}

// ---- conv epilogue: sum y and sum y^2 per output channel of the bf16 values the epilogue stores ----
// The tcgen05 conv kernels stage every 128-pixel x 64-channel output chunk in shared memory as bf16 rows of 128 bytes (SWIZZLE_128B)
// for the TMA store.  With 256 epilogue threads, thread t sums channel pair (t & 31) over the 16 rows of row group (t >> 5) --
// a warp reads the 32 words of one row: conflict-free -- and leaves its four partial sums in a double-buffered scratch; the eight
// row groups are added in a fixed order by 128 of the threads after the NEXT chunk's first barrier (no extra barrier per chunk) into
// per-CTA per-channel totals, which go to the BnAcc accumulators once, at the end of the kernel.  The CTA's tile order is fixed, so
// the totals are deterministic.  Rows outside the image (ragged tiles) are excluded through `nvalid`.
constexpr int EPI_STATS_SCRATCH_FLOATS = 2 * 8 * 32 * 4;   // 8 KB
inline int epi_stats_smem_bytes(int cout) { return EPI_STATS_SCRATCH_FLOATS * 4 + cout * 2 * 4; }

struct EpiStats {
  float* scratch;   // [2][8 row groups][32 channel pairs][4]: s(c0), q(c0), s(c1), q(c1)
  float* stat;      // [Cout][2]
  int prev_base;    // first channel of the chunk whose scratch still awaits reduction; -1: none
  uint32_t k;       // chunks processed (scratch parity)
};

__device__ __forceinline__ void epi_stats_init(EpiStats& st, float* region, int cout, int t) {
  st.scratch = region;
  st.stat = region + EPI_STATS_SCRATCH_FLOATS;
  st.prev_base = -1;
  st.k = 0;
  for (int i = t; i < 2 * cout; i += 256) st.stat[i] = 0.f;   // visible to everybody after the first barrier of the first chunk
}

// after the first barrier of a chunk (all threads have finished the previous chunk, scratch writes included)
__device__ __forceinline__ void epi_stats_reduce_prev(EpiStats& st, int t) {
  if (st.prev_base >= 0 && t < 128) {
    const float* p = st.scratch + ((st.k & 1u) ^ 1u) * (8 * 32 * 4) + t;   // t = channel-in-chunk * 2 + (0: sum, 1: sum of squares)
    float acc = 0.f;
#pragma unroll
    for (int rg = 0; rg < 8; ++rg) acc += p[rg * 128];
    st.stat[st.prev_base * 2 + t] += acc;
  }
}

// after the second barrier of a chunk (the staged bf16 tile is complete): rows [rg*16, rg*16 + nvalid) of this thread's row group
__device__ __forceinline__ void epi_stats_chunk(EpiStats& st, const uint8_t* stage_buf, int t, int nvalid, int chan_base) {
  const int cp = t & 31, rg = t >> 5;
  float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
  const uint8_t* base = stage_buf + rg * 2048 + (cp & 3) * 4;
  const int g = cp >> 2;
  if (nvalid == 16) {   // the common case (warp-uniform: a warp is one row group): no per-row predicate
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = *reinterpret_cast<const uint32_t*>(base + i * 128 + (((g ^ i) & 7) << 4));   // (rg*16 + i) & 7 == i & 7
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float a, b;
      unpack_bf16x2(w[i], a, b);
      s0 += a; q0 = fmaf(a, a, q0);
      s1 += b; q1 = fmaf(b, b, q1);
    }
  } else {
    for (int i = 0; i < nvalid; ++i) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(base + i * 128 + (((g ^ i) & 7) << 4));
      float a, b;
      unpack_bf16x2(w, a, b);
      s0 += a; q0 = fmaf(a, a, q0);
      s1 += b; q1 = fmaf(b, b, q1);
    }
  }
  *reinterpret_cast<float4*>(st.scratch + (st.k & 1u) * (8 * 32 * 4) + (rg * 32 + cp) * 4) = make_float4(s0, q0, s1, q1);
  st.prev_base = chan_base;
  ++st.k;
}

// after the tile loop: bar() is the epilogue's named barrier
template <class Bar>
__device__ __forceinline__ void epi_stats_flush(EpiStats& st, int t, int cout, BnAcc* acc, Bar bar) {
  bar();
  epi_stats_reduce_prev(st, t);
  bar();
  // up to four values per thread per round with all their atomics in flight together: the returning low-word adds are independent, the
  // high-word adds need no return value -- one contended round trip per round instead of one per value
  for (int i0 = t; i0 < 2 * cout; i0 += 4 * 256) {
    unsigned long long lo[4], hi[4], old[4];
    bool live[4];
    BnAcc* dst[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * 256;
      live[u] = i < 2 * cout;
      dst[u] = acc + (live[u] ? (i & 1) * cout + (i >> 1) : 0);
      live[u] = live[u] && bn_acc_split(dst[u], live[u] ? st.stat[i] : 0.f, lo[u], hi[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) old[u] = live[u] ? atomicAdd(&dst[u]->lo, lo[u]) : 0ull;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (live[u]) {
        const unsigned long long hi2 = hi[u] + ((old[u] + lo[u]) < lo[u] ? 1ull : 0ull);
        if (hi2) atomicAdd(&dst[u]->hi, hi2);
      }
    }
  }
}

}  // namespace hk
