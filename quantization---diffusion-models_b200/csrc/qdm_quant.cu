// (b) Fused RTN quantise / dequantise / pack kernels.  HBM-bound: one 16-byte load per thread,
// group statistics by warp shuffles, every torch op of the reference replayed with its dtype
// rounding (see rnd<T>) so codes, scales and zeros are bit-exact with the reference:
//   AwqQuantizer.pseudo_quantize_tensor      quantize/quantizer.py:163-198
//   quantize_weight_absmax                    quantize/fake_quant.py:21-84
//   quantize_weight_per_channel_absmax        quantize/fake_quant.py:86-93
//   quantize_weight_per_tensor_absmax         quantize/fake_quant.py:97-105
//   quantize_activation_per_token_absmax      quantize/fake_quant.py:109-118
//   AWQ int4 GEMM layout                      utils/packing_utils.py:4-102, utils/quant_utils.py:10-39
#include "qdm_common.cuh"

int qdm_absmax_impl(const void* x, int dtype, int64_t numel, void* out, void* workspace,
                    size_t workspace_bytes, cudaStream_t st);

namespace {

enum QMode { Q_ZP = 0, Q_SYM = 1, Q_SYM_NOCLAMP = 2 };

// scale / zero of one group from its (max, min) [Q_ZP] or |.|max [symmetric].
template <typename T, int MODE>
__device__ __forceinline__ void group_params(float mx, float mn, float max_int, float& s, float& z) {
  const float floor_c = rnd<T>(1e-5f);  // python scalar 1e-5 seen through the tensor dtype
  if (MODE == Q_ZP) {
    float d = rnd<T>(__fsub_rn(mx, mn));            // max_val - min_val
    d = fmaxf(d, floor_c);                          // .clamp(min=1e-5)
    s = rnd<T>(__fdiv_rn(d, max_int));              // / max_int
    const float r = rintf(rnd<T>(__fdiv_rn(mn, s)));  // torch.round(min_val / scales)
    z = fminf(fmaxf(-r, 0.f), max_int);             // (-...).clamp_(0, max_int)
  } else {
    const float a = fmaxf(mx, floor_c);             // .clamp(min=1e-5)
    s = rnd<T>(__fdiv_rn(a, max_int));              // / q_max
    z = 0.f;
  }
}

// codes are exact small integers.  Zero-point codes are unsigned bytes (0..255); symmetric
// codes are two's complement and saturate to [-128, 127] (only reachable with QDM_Q_NO_CLAMP on
// bf16 inputs, where the reference itself produces +-128, see DESIGN.md "8-bit caveat").
template <int MODE>
__device__ __forceinline__ int8_t code_to_i8(float q) {
  const float lo = (MODE == Q_ZP) ? 0.f : -128.f, hi = (MODE == Q_ZP) ? 255.f : 127.f;
  int v = __float2int_rn(fminf(fmaxf(q, lo), hi));
  return (int8_t)(v & 0xff);
}

constexpr int kQThreads = 256;

// One 16-byte vector of a group through the reference's op chain.  x[] already holds the prepared
// (pre-multiplied / clipped) weights, (s, z) the group's parameters.  Writes the fake-quantised values
// to o and the integer codes (as floats) to cq.  FAST selects the reciprocal division (see qdm_common.cuh).
//   zero point: q = round(w/s); c = clamp(q + z, 0, max); dq = (c - z) * s      quantizer.py:177-182
//     (q + z and c - z are sums of small integers, exact in every dtype; when q is too large for that
//      both the rounded and the unrounded sum are far above max and clamp to it)
//   symmetric:  q = round(w/s); c = clamp(q, min, max) [not for NO_CLAMP]; dq = c * s   quantizer.py:186-190
template <typename T, int MODE, bool FAST, bool POST, bool POST_FAST>
__device__ __forceinline__ void rtn_vec(const float (&x)[ElemTraits<T>::kVec], float s, float z,
                                        float min_int, float max_int,
                                        const float (&pd)[ElemTraits<T>::kVec], const float (&rpd)[ElemTraits<T>::kVec],
                                        Vec16<T>& o, float (&cq)[ElemTraits<T>::kVec]) {
  constexpr int V = ElemTraits<T>::kVec;
  const float r = FAST ? rcp_approx(s) : 0.f;
  const float lo = (MODE == Q_ZP) ? __fsub_rn(0.f, z) : min_int, hi = (MODE == Q_ZP) ? __fsub_rn(max_int, z) : max_int;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    float q = rint_T<T>(rnd<T>(div_T<T, FAST, false>(x[j], s, r)));
    if (MODE != Q_ZP) q = copysignf(q, x[j]);    // torch.round keeps the sign of a zero result; s > 0
    if (MODE != Q_SYM_NOCLAMP) q = fminf(fmaxf(q, lo), hi);
    cq[j] = q;                                   // code - z (zero point) or the code itself
    if (POST) {
      const float v = rnd<T>(__fmul_rn(q, s));
      o.v[j] = ElemTraits<T>::from_f(div_T<T, POST_FAST, MODE != Q_ZP>(v, pd[j], rpd[j]));
    } else {
      o.v[j] = ElemTraits<T>::from_f(__fmul_rn(q, s));
    }
  }
}

// ---------------------------------------------------------------- fp16 (the reference dtype) packed path
// Everything except the division itself runs on half2 pairs (two elements per lane-instruction):
//   * x*pre_mul, clamp, min/max and q*s round exactly like "fp32 op, round to fp16" because the fp32
//     results are exact (22-bit products, comparisons), so HMUL2/HMNMX2 give the same bits;
//   * round-half-even is (q + 1536) - 1536 in fp16 (ulp 1 in [1024, 2048)): exact for |q| < 512, and above
//     that still far outside every clamp range (unclamped codes never exceed 128);
//   * the quotient is formed in fp32 (div_by_rcp) and packed back with one F2FP per pair.
// About 7 instructions per element instead of 13; that is what moves these kernels from issue-bound to HBM-bound.
struct H2x4 { __half2 h[4]; };

__device__ __forceinline__ H2x4 as_h2x4(const Vec16<__half>& v) {
  H2x4 r;
  *reinterpret_cast<uint4*>(&r) = *reinterpret_cast<const uint4*>(&v);
  return r;
}
__device__ __forceinline__ uint32_t h2_bits(__half2 v) { return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ __half2 bits_h2(uint32_t v) { return *reinterpret_cast<__half2*>(&v); }

// per-lane packed (max, min) [zero point] or (|.|max, -) [symmetric] of one vector
template <int MODE>
__device__ __forceinline__ void fold_h(const H2x4& x, __half2& m2, __half2& n2) {
  if (MODE == Q_ZP) {
    m2 = __hmax2(__hmax2(x.h[0], x.h[1]), __hmax2(x.h[2], x.h[3]));
    n2 = __hmin2(__hmin2(x.h[0], x.h[1]), __hmin2(x.h[2], x.h[3]));
  } else {
    m2 = __hmax2(__hmax2(__habs2(x.h[0]), __habs2(x.h[1])), __hmax2(__habs2(x.h[2]), __habs2(x.h[3])));
    n2 = m2;
  }
}
// fold the two halves, then ONE butterfly over `lpg` adjacent lanes on the pair (max, -min): max of negated
// minima is the negated minimum (negation is exact), so a single SHFL + HMNMX2 per step serves both.
template <int MODE>
__device__ __forceinline__ void reduce_h(__half2 m2, __half2 n2, int lpg, float& mx, float& mn) {
  __half2 p;
  if (MODE == Q_ZP) p = __halves2half2(__hmax(__low2half(m2), __high2half(m2)), __hneg(__hmin(__low2half(n2), __high2half(n2))));
  else p = __hmax2(m2, __lowhigh2highlow(m2));
  switch (lpg) {                                   // xor-butterfly steps commute: fall through from the widest
    case 32: p = __hmax2(p, __shfl_xor_sync(0xffffffffu, p, 16));
    case 16: p = __hmax2(p, __shfl_xor_sync(0xffffffffu, p, 8));
    case 8: p = __hmax2(p, __shfl_xor_sync(0xffffffffu, p, 4));
    case 4: p = __hmax2(p, __shfl_xor_sync(0xffffffffu, p, 2));
    case 2: p = __hmax2(p, __shfl_xor_sync(0xffffffffu, p, 1));
    default: break;
  }
  mx = __low2float(p);
  mn = (MODE == Q_ZP) ? -__high2float(p) : 0.f;
}
template <int MODE>
__device__ __forceinline__ void minmax_h(const H2x4& x, int lpg, float& mx, float& mn) {
  __half2 m2, n2;
  fold_h<MODE>(x, m2, n2);
  reduce_h<MODE>(m2, n2, lpg, mx, mn);
}

// group_params for fp16 with the reciprocal division (valid for every fp16 pair); also returns r ~ 1/s
template <int MODE>
__device__ __forceinline__ void group_params_h(float mx, float mn, float max_int, float r_max_int,
                                               float& s, float& z, float& r) {
  const float floor_c = rnd<__half>(1e-5f);
  if (MODE == Q_ZP) {
    const float d = fmaxf(rnd<__half>(__fsub_rn(mx, mn)), floor_c);
    s = rnd<__half>(div_by_rcp<false>(d, max_int, r_max_int));
    r = rcp_approx(s);
    const float t = rint_T<__half>(rnd<__half>(div_by_rcp<false>(mn, s, r)));
    z = fminf(fmaxf(-t, 0.f), max_int);
  } else {
    s = rnd<__half>(div_by_rcp<false>(fmaxf(mx, floor_c), max_int, r_max_int));
    r = rcp_approx(s);
    z = 0.f;
  }
}

// the RTN chain on a packed vector; cq receives (code - z) [zero point] or the code, as exact fp16 integers
template <int MODE, bool POST>
__device__ __forceinline__ void rtn_vec_h(const H2x4& x, float s, float r, float z, float min_int, float max_int,
                                          const float (&pd)[8], const float (&rpd)[8], H2x4& o, H2x4& cq) {
  const __half2 s2 = __float2half2_rn(s);                      // s is an fp16 value: exact
  const __half2 magic = __float2half2_rn(1536.f);
  const __half2 lo2 = __float2half2_rn(MODE == Q_ZP ? __fsub_rn(0.f, z) : min_int);   // +0 when z == 0
  const __half2 hi2 = __float2half2_rn(MODE == Q_ZP ? __fsub_rn(max_int, z) : max_int);
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const float2 xf = __half22float2(x.h[p]);
    __half2 q = __floats2half2_rn(div_by_rcp<false>(xf.x, s, r), div_by_rcp<false>(xf.y, s, r));
    q = __hsub2(__hadd2(q, magic), magic);
    if (MODE != Q_ZP) q = bits_h2((h2_bits(q) & 0x7fff7fffu) | (h2_bits(x.h[p]) & 0x80008000u));  // sign of a zero result
    if (MODE != Q_SYM_NOCLAMP) q = __hmin2(__hmax2(q, lo2), hi2);
    cq.h[p] = q;
    const __half2 v = __hmul2(q, s2);
    if (POST) {
      const float2 vf = __half22float2(v);
      o.h[p] = __floats2half2_rn(div_by_rcp<MODE != Q_ZP>(vf.x, pd[2 * p], rpd[2 * p]),
                                 div_by_rcp<MODE != Q_ZP>(vf.y, pd[2 * p + 1], rpd[2 * p + 1]));
    } else {
      o.h[p] = v;
    }
  }
}

// 8 code bytes from cq (adds z back for zero point; saturates symmetric codes to int8)
template <int MODE>
__device__ __forceinline__ uint2 code_bytes_h(const H2x4& cq, float z) {
  const __half2 z2 = __float2half2_rn(z), magic = __float2half2_rn(1536.f);
  uint32_t t[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    __half2 c = (MODE == Q_ZP) ? __hadd2(cq.h[p], z2)
                               : __hmin2(__hmax2(cq.h[p], __float2half2_rn(-128.f)), __float2half2_rn(127.f));
    t[p] = h2_bits(__hadd2(c, magic));                         // mantissa = 512 + c: low byte = c mod 256
  }
  return make_uint2(__byte_perm(t[0], t[1], 0x6420), __byte_perm(t[2], t[3], 0x6420));
}

// ---------------------------------------------------------------- power-of-two groups, vector path
// W[n_rows, k_cols] is a sequence of contiguous groups; a group is `lpg` adjacent lanes (one 16-byte
// vector per lane), so group min/max is an xor-shuffle butterfly.  Thread t owns vector column
// t % vpr for rows t / vpr, + rows_per_pass, ...: the per-column vectors of the AWQ search
// (pre_mul, post_div and its reciprocal) are loaded once per thread, and the next row's vector is
// requested before the current one is processed (two 16-byte loads in flight per thread).
template <typename T, int MODE, bool EXTRAS>
__global__ void __launch_bounds__(kQThreads)
quant_group_kernel(const T* __restrict__ w, int n_rows, int vpr, int lpg_shift, int rows_per_pass,
                   float max_int, float min_int,
                   const T* __restrict__ pre_mul, const T* __restrict__ clip_max,
                   const T* __restrict__ post_div,
                   T* __restrict__ dq, int8_t* __restrict__ codes,
                   T* __restrict__ scales, T* __restrict__ zeros) {
  constexpr int V = ElemTraits<T>::kVec;
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * kQThreads + threadIdx.x;
  const bool in_range = t < rows_per_pass * vpr;
  const int cv = in_range ? t % vpr : 0;
  const int row0 = in_range ? t / vpr : n_rows;
  const int lpg = 1 << lpg_shift;
  const int gpr = vpr >> lpg_shift;                // groups per row
  float pm[V], pd[V], rpd[V];
  H2x4 pmh;                                        // fp16 path keeps pre_mul packed
  const float r_max_int = rcp_approx(max_int);
  bool post_fast = true;
#pragma unroll
  for (int j = 0; j < V; ++j) { pm[j] = 1.f; pd[j] = 1.f; rpd[j] = 1.f; }
#pragma unroll
  for (int p = 0; p < 4; ++p) pmh.h[p] = __float2half2_rn(1.f);
  if (EXTRAS) {
    if (pre_mul) {
      Vec16<T> a = ld_vec16(pre_mul + int64_t(cv) * V);
#pragma unroll
      for (int j = 0; j < V; ++j) pm[j] = ElemTraits<T>::to_f(a.v[j]);
      if constexpr (std::is_same<T, __half>::value) pmh = as_h2x4(a);
    }
    if (post_div) {
      Vec16<T> a = ld_vec16(post_div + int64_t(cv) * V);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        pd[j] = ElemTraits<T>::to_f(a.v[j]);
        rpd[j] = rcp_approx(pd[j]);
        post_fast = post_fast && fastdiv_ok<T>(pd[j], 0.f);
      }
    }
  }
  Vec16<T> cur;
  if (row0 < n_rows) cur = ld_vec16_stream(w + (int64_t(row0) * vpr + cv) * V);
  for (int row = row0, base = 0; base < n_rows; base += rows_per_pass, row += rows_per_pass) {
    const bool active = row < n_rows;               // `base` keeps the trip count warp-uniform for the shuffles
    const int64_t i = int64_t(row) * vpr + cv;
    Vec16<T> nxt;
    if (row + rows_per_pass < n_rows && in_range) nxt = ld_vec16_stream(w + (i + int64_t(rows_per_pass) * vpr) * V);
    if constexpr (std::is_same<T, __half>::value) {
      H2x4 x;
      if (active) {
        x = as_h2x4(cur);
        if (EXTRAS) {
          if (pre_mul) {
#pragma unroll
            for (int p = 0; p < 4; ++p) x.h[p] = __hmul2(x.h[p], pmh.h[p]);
          }
          if (clip_max) {
            const __half2 c2 = __half2half2(clip_max[int64_t(row) * gpr + (cv >> lpg_shift)]);
#pragma unroll
            for (int p = 0; p < 4; ++p) x.h[p] = __hmin2(__hmax2(x.h[p], __hneg2(c2)), c2);
          }
        }
      } else {
#pragma unroll
        for (int p = 0; p < 4; ++p) x.h[p] = __float2half2_rn(0.f);
      }
      float mx, mn;
      minmax_h<MODE>(x, lpg, mx, mn);
      if (active) {
        float s, z, r;
        group_params_h<MODE>(mx, mn, max_int, r_max_int, s, z, r);
        if ((lane & (lpg - 1)) == 0) {
          const int64_t g = int64_t(row) * gpr + (cv >> lpg_shift);
          if (scales) scales[g] = __float2half_rn(s);
          if (zeros && MODE == Q_ZP) zeros[g] = __float2half_rn(z);
        }
        H2x4 o, cq;
        if (EXTRAS && post_div) rtn_vec_h<MODE, true>(x, s, r, z, min_int, max_int, pd, rpd, o, cq);
        else rtn_vec_h<MODE, false>(x, s, r, z, min_int, max_int, pd, rpd, o, cq);
        if (dq) *reinterpret_cast<uint4*>(dq + i * V) = *reinterpret_cast<const uint4*>(&o);
        if (codes) *reinterpret_cast<uint2*>(codes + i * V) = code_bytes_h<MODE>(cq, z);
      }
    } else {
    float x[V];
    if (active) {
#pragma unroll
      for (int j = 0; j < V; ++j) x[j] = ElemTraits<T>::to_f(cur.v[j]);
      if (EXTRAS) {
        if (pre_mul) {
#pragma unroll
          for (int j = 0; j < V; ++j) x[j] = rnd<T>(__fmul_rn(x[j], pm[j]));
        }
        if (clip_max) {
          const float c = ElemTraits<T>::to_f(clip_max[int64_t(row) * gpr + (cv >> lpg_shift)]);
#pragma unroll
          for (int j = 0; j < V; ++j) x[j] = fminf(fmaxf(x[j], -c), c);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) x[j] = 0.f;
    }
    float mx, mn;
    if (MODE == Q_ZP) {
      mx = x[0]; mn = x[0];
#pragma unroll
      for (int j = 1; j < V; ++j) { mx = fmaxf(mx, x[j]); mn = fminf(mn, x[j]); }
    } else {
      mx = 0.f; mn = 0.f;
#pragma unroll
      for (int j = 0; j < V; ++j) mx = fmaxf(mx, fabsf(x[j]));
    }
    for (int o = 1; o < lpg; o <<= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (MODE == Q_ZP) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if (active) {
      float s, z;
      group_params<T, MODE>(mx, mn, max_int, s, z);
      if ((lane & (lpg - 1)) == 0) {
        const int64_t g = int64_t(row) * gpr + (cv >> lpg_shift);
        if (scales) scales[g] = ElemTraits<T>::from_f(s);
        if (zeros && MODE == Q_ZP) zeros[g] = ElemTraits<T>::from_f(z);
      }
      float cq[V];
      Vec16<T> o;
      const float amax = (MODE == Q_ZP) ? fmaxf(fabsf(mx), fabsf(mn)) : mx;
      const bool fast = fastdiv_ok<T>(s, amax);
      if (EXTRAS && post_div) {
        if (fast && post_fast) rtn_vec<T, MODE, true, true, true>(x, s, z, min_int, max_int, pd, rpd, o, cq);
        else rtn_vec<T, MODE, false, true, false>(x, s, z, min_int, max_int, pd, rpd, o, cq);
      } else {
        if (fast) rtn_vec<T, MODE, true, false, false>(x, s, z, min_int, max_int, pd, rpd, o, cq);
        else rtn_vec<T, MODE, false, false, false>(x, s, z, min_int, max_int, pd, rpd, o, cq);
      }
      if (dq) st_vec16(dq + i * V, o);
      if (codes) {
        uint32_t pk[V / 4];
#pragma unroll
        for (int j = 0; j < V; ++j) {
          float c = (MODE == Q_ZP) ? __fadd_rn(cq[j], z) : fminf(fmaxf(cq[j], -128.f), 127.f);
          const uint32_t b = int_byte(c);
          if ((j & 3) == 0) pk[j / 4] = b; else pk[j / 4] |= b << (8 * (j & 3));
        }
        if (V == 8) *reinterpret_cast<uint2*>(codes + i * V) = make_uint2(pk[0], pk[V / 4 - 1]);
        else *reinterpret_cast<uint32_t*>(codes + i * V) = pk[0];
      }
    }
    }
    cur = nxt;
  }
}

// ---------------------------------------------------------------- arbitrary row length
// One warp per row (row = group of `cols` contiguous elements); two passes, the second one
// re-reads the row from L1/L2.  k_period is the K extent used to index pre_mul/post_div.
template <typename T, int MODE>
__global__ void __launch_bounds__(kQThreads)
quant_rows_warp_kernel(const T* __restrict__ x, int64_t rows, int64_t cols, int64_t k_period,
                       float max_int, float min_int,
                       const T* __restrict__ pre_mul, const T* __restrict__ clip_max,
                       const T* __restrict__ post_div,
                       T* __restrict__ dq, int8_t* __restrict__ codes,
                       T* __restrict__ scales, T* __restrict__ zeros, int vec_ok) {
  constexpr int V = ElemTraits<T>::kVec;
  const int lane = threadIdx.x & 31;
  const int wpb = kQThreads / 32;
  for (int64_t row = int64_t(blockIdx.x) * wpb + (threadIdx.x >> 5); row < rows;
       row += int64_t(gridDim.x) * wpb) {
    const T* p = x + row * cols;
    const float c = clip_max ? ElemTraits<T>::to_f(clip_max[row]) : 0.f;
    auto prep = [&](float v, int64_t e) {
      if (pre_mul) v = rnd<T>(__fmul_rn(v, ElemTraits<T>::to_f(pre_mul[(row * cols + e) % k_period])));
      if (clip_max) v = fminf(fmaxf(v, -c), c);
      return v;
    };
    float mx = (MODE == Q_ZP) ? -INFINITY : 0.f, mn = INFINITY;
    if (vec_ok) {
      // four 16-byte loads per lane in flight (rows longer than the register kernel's 4096 elements, e.g. K = 5120
      // activations, were latency-bound at one load per lane per trip: 1.7 TB/s)
      constexpr int U = 4;
      for (int64_t e0 = int64_t(lane) * V; e0 < cols; e0 += int64_t(U) * 32 * V) {
        Vec16<T> v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (e0 + int64_t(u) * 32 * V < cols) v[u] = ld_vec16(p + e0 + int64_t(u) * 32 * V);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t e = e0 + int64_t(u) * 32 * V;
          if (e < cols) {
#pragma unroll
            for (int j = 0; j < V; ++j) {
              float f = prep(ElemTraits<T>::to_f(v[u].v[j]), e + j);
              if (MODE == Q_ZP) { mx = fmaxf(mx, f); mn = fminf(mn, f); } else mx = fmaxf(mx, fabsf(f));
            }
          }
        }
      }
    } else {
      for (int64_t e = lane; e < cols; e += 32) {
        float f = prep(ElemTraits<T>::to_f(p[e]), e);
        if (MODE == Q_ZP) { mx = fmaxf(mx, f); mn = fminf(mn, f); } else mx = fmaxf(mx, fabsf(f));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (MODE == Q_ZP) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    float s, z;
    group_params<T, MODE>(mx, mn, max_int, s, z);
    if (lane == 0) {
      if (scales) scales[row] = ElemTraits<T>::from_f(s);
      if (zeros && MODE == Q_ZP) zeros[row] = ElemTraits<T>::from_f(z);
    }
    if (!dq && !codes) continue;
    auto finish = [&](float f, int64_t e, float& code) {
      float d = rtn_elem<T>(f, s, z, min_int, max_int, MODE == Q_ZP, MODE != Q_SYM_NOCLAMP, code);
      if (post_div) d = rnd<T>(__fdiv_rn(d, ElemTraits<T>::to_f(post_div[(row * cols + e) % k_period])));
      return d;
    };
    if (vec_ok) {
      constexpr int U = 4;
      for (int64_t e0 = int64_t(lane) * V; e0 < cols; e0 += int64_t(U) * 32 * V) {
        Vec16<T> vv[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (e0 + int64_t(u) * 32 * V < cols) vv[u] = ld_vec16(p + e0 + int64_t(u) * 32 * V);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t e = e0 + int64_t(u) * 32 * V;
          if (e >= cols) continue;
          Vec16<T> o;
          uint32_t pk[V / 4];
#pragma unroll
          for (int j = 0; j < V; ++j) {
            float code;
            o.v[j] = ElemTraits<T>::from_f(finish(prep(ElemTraits<T>::to_f(vv[u].v[j]), e + j), e + j, code));
            const uint32_t bb = uint32_t(uint8_t(code_to_i8<MODE>(code)));
            if ((j & 3) == 0) pk[j / 4] = bb; else pk[j / 4] |= bb << (8 * (j & 3));
          }
          if (dq) st_vec16(dq + row * cols + e, o);
          if (codes) {   // V consecutive codes as one 4- / 8-byte store (row * cols + e is a multiple of V: vec_ok)
            if (V == 8) *reinterpret_cast<uint2*>(codes + row * cols + e) = make_uint2(pk[0], pk[V / 4 - 1]);
            else *reinterpret_cast<uint32_t*>(codes + row * cols + e) = pk[0];
          }
        }
      }
    } else {
      for (int64_t e = lane; e < cols; e += 32) {
        float code;
        float d = finish(prep(ElemTraits<T>::to_f(p[e]), e), e, code);
        if (dq) dq[row * cols + e] = ElemTraits<T>::from_f(d);
        if (codes) codes[row * cols + e] = code_to_i8<MODE>(code);
      }
    }
  }
}

// Rows of up to 32*V*NV elements held in registers: one warp per row, every lane issues all of its
// 16-byte loads up front (NV in flight), the row statistics come from a shuffle butterfly and the row is
// quantised from registers -- one pass over HBM (elem read + outputs), unlike the two-pass kernel above.
// I8OUT: emit int8 codes + fp32 scale (the A8 of W8A8) instead of the fake-quantised row.
// LPR = lanes per row (32, 16 or 8): short rows (K = 320 / 640 activations: 40 / 80 vectors) share a warp, 32 / LPR rows at a
// time, so that every lane still has ~5 loads in flight and the butterfly has log2(LPR) steps -- with one warp per 640-byte
// row the per-token quantiser ran at 0.44 of the HBM peak (issue-bound on the per-row overhead), against 0.86 at K = 1280.
template <typename T, int MODE, int NV, bool I8OUT, int LPR = 32>
__global__ void __launch_bounds__(kQThreads)
quant_rows_reg_kernel(const T* __restrict__ x, int64_t rows, int cols, float max_int, float min_int,
                      T* __restrict__ dq, int8_t* __restrict__ codes,
                      T* __restrict__ scales, T* __restrict__ zeros, float* __restrict__ sx) {
  constexpr int V = ElemTraits<T>::kVec;
  constexpr int RPW = 32 / LPR;                          // rows per warp
  const int lane = threadIdx.x & (LPR - 1);              // lane inside the row's group
  const int sub = (threadIdx.x & 31) / LPR;              // which of the warp's rows
  const int wpb = kQThreads / 32;
  const float dummy[V] = {};
  const float dummy8[8] = {};
  if (I8OUT) {
    // the A8 quantiser sits between two GEMMs of a W8A8 model: launched with programmatic stream serialisation
    // (launch_rows_reg), it lets the GEMM that follows set up while it runs and itself becomes resident while the GEMM
    // before it drains; nothing of an earlier kernel is touched before the wait.  Both are no-ops in a plain launch.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  for (int64_t row0 = (int64_t(blockIdx.x) * wpb + (threadIdx.x >> 5)) * RPW; row0 < rows;
       row0 += int64_t(gridDim.x) * wpb * RPW) {
    // a group past the last row works on the last row again and skips every store (the shuffles below need all lanes)
    const bool row_on = row0 + sub < rows;
    const int64_t row = row_on ? row0 + sub : rows - 1;
    const T* p = x + row * cols;
    Vec16<T> raw[NV];
#pragma unroll
    for (int t = 0; t < NV; ++t) {
      const int e = (t * LPR + lane) * V;
      if (e < cols) raw[t] = ld_vec16_stream(p + e);
    }
    if constexpr (std::is_same<T, __half>::value) {
      __half2 m2 = __float2half2_rn((MODE == Q_ZP) ? -INFINITY : 0.f), n2 = __float2half2_rn(INFINITY);
#pragma unroll
      for (int t = 0; t < NV; ++t) {
        if ((t * LPR + lane) * V < cols) {
          __half2 a, b;
          fold_h<MODE>(as_h2x4(raw[t]), a, b);
          m2 = __hmax2(m2, a);
          n2 = __hmin2(n2, b);
        }
      }
      float mx, mn, s, z, r;
      reduce_h<MODE>(m2, n2, LPR, mx, mn);
      group_params_h<MODE>(mx, mn, max_int, rcp_approx(max_int), s, z, r);
      if (lane == 0 && row_on) {
        if (I8OUT) sx[row] = s;
        if (scales) scales[row] = __float2half_rn(s);
        if (zeros && MODE == Q_ZP) zeros[row] = __float2half_rn(z);
      }
      if ((!dq && !codes) || !row_on) continue;
#pragma unroll
      for (int t = 0; t < NV; ++t) {
        const int e = (t * LPR + lane) * V;
        if (e < cols) {
          H2x4 o, cq;
          rtn_vec_h<MODE, false>(as_h2x4(raw[t]), s, r, z, min_int, max_int, dummy8, dummy8, o, cq);
          if (dq) *reinterpret_cast<uint4*>(dq + row * cols + e) = *reinterpret_cast<const uint4*>(&o);
          if (codes) *reinterpret_cast<uint2*>(codes + row * cols + e) = code_bytes_h<MODE>(cq, z);
        }
      }
    } else {
      float mx = (MODE == Q_ZP) ? -INFINITY : 0.f, mn = INFINITY;
  #pragma unroll
      for (int t = 0; t < NV; ++t) {
        const int e = (t * LPR + lane) * V;
        if (e < cols) {
  #pragma unroll
          for (int j = 0; j < V; ++j) {
            const float f = ElemTraits<T>::to_f(raw[t].v[j]);
            if (MODE == Q_ZP) { mx = fmaxf(mx, f); mn = fminf(mn, f); } else mx = fmaxf(mx, fabsf(f));
          }
        }
      }
  #pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (MODE == Q_ZP) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      }
      float s, z;
      group_params<T, MODE>(mx, mn, max_int, s, z);
      if (lane == 0 && row_on) {
        if (I8OUT) sx[row] = s;
        if (scales) scales[row] = ElemTraits<T>::from_f(s);
        if (zeros && MODE == Q_ZP) zeros[row] = ElemTraits<T>::from_f(z);
      }
      if ((!dq && !codes) || !row_on) continue;
      const float amax = (MODE == Q_ZP) ? fmaxf(fabsf(mx), fabsf(mn)) : mx;
      const bool fast = fastdiv_ok<T>(s, amax);
  #pragma unroll
      for (int t = 0; t < NV; ++t) {
        const int e = (t * LPR + lane) * V;
        if (e < cols) {
          float xv[V], cq[V];
          Vec16<T> o;
  #pragma unroll
          for (int j = 0; j < V; ++j) xv[j] = ElemTraits<T>::to_f(raw[t].v[j]);
          if (fast) rtn_vec<T, MODE, true, false, false>(xv, s, z, min_int, max_int, dummy, dummy, o, cq);
          else rtn_vec<T, MODE, false, false, false>(xv, s, z, min_int, max_int, dummy, dummy, o, cq);
          if (dq) st_vec16(dq + row * cols + e, o);
          if (codes) {
            uint32_t pk[V / 4];
  #pragma unroll
            for (int j = 0; j < V; ++j) {
              float c = (MODE == Q_ZP) ? __fadd_rn(cq[j], z) : fminf(fmaxf(cq[j], -128.f), 127.f);
              const uint32_t b = int_byte(c);
              if ((j & 3) == 0) pk[j / 4] = b; else pk[j / 4] |= b << (8 * (j & 3));
            }
            if (V == 8) *reinterpret_cast<uint2*>(codes + row * cols + e) = make_uint2(pk[0], pk[V / 4 - 1]);
            else *reinterpret_cast<uint32_t*>(codes + row * cols + e) = pk[0];
          }
        }
      }
    }
  }
}

// One thread per row for tiny rows (conv weights viewed as [..., kw]: rows of 1..16 taps).
template <typename T, int MODE>
__global__ void __launch_bounds__(kQThreads)
quant_rows_thread_kernel(const T* __restrict__ x, int64_t rows, int cols, float max_int, float min_int,
                         T* __restrict__ dq, int8_t* __restrict__ codes,
                         T* __restrict__ scales, T* __restrict__ zeros) {
  constexpr int kMaxCols = 16;
  for (int64_t row = int64_t(blockIdx.x) * kQThreads + threadIdx.x; row < rows;
       row += int64_t(gridDim.x) * kQThreads) {
    const T* p = x + row * cols;
    float v[kMaxCols];
    float mx = (MODE == Q_ZP) ? -INFINITY : 0.f, mn = INFINITY;
#pragma unroll
    for (int e = 0; e < kMaxCols; ++e) {
      if (e < cols) {
        v[e] = ElemTraits<T>::to_f(p[e]);
        if (MODE == Q_ZP) { mx = fmaxf(mx, v[e]); mn = fminf(mn, v[e]); } else mx = fmaxf(mx, fabsf(v[e]));
      }
    }
    float s, z;
    group_params<T, MODE>(mx, mn, max_int, s, z);
    if (scales) scales[row] = ElemTraits<T>::from_f(s);
    if (zeros && MODE == Q_ZP) zeros[row] = ElemTraits<T>::from_f(z);
    const float amax = (MODE == Q_ZP) ? fmaxf(fabsf(mx), fabsf(mn)) : mx;
    const bool fast = fastdiv_ok<T>(s, amax);              // reciprocal division, see qdm_common.cuh
    const float r = rcp_approx(s);
    const float lo = (MODE == Q_ZP) ? __fsub_rn(0.f, z) : min_int, hi = (MODE == Q_ZP) ? __fsub_rn(max_int, z) : max_int;
#pragma unroll
    for (int e = 0; e < kMaxCols; ++e) {
      if (e < cols) {
        float q = rint_T<T>(rnd<T>(fast ? div_by_rcp<false>(v[e], s, r) : __fdiv_rn(v[e], s)));
        if (MODE != Q_ZP) q = copysignf(q, v[e]);
        if (MODE != Q_SYM_NOCLAMP) q = fminf(fmaxf(q, lo), hi);
        if (dq) dq[row * cols + e] = ElemTraits<T>::from_f(__fmul_rn(q, s));
        if (codes) codes[row * cols + e] = code_to_i8<MODE>(MODE == Q_ZP ? __fadd_rn(q, z) : q);
      }
    }
  }
}

// Tiny rows with vector traffic (conv weights viewed as [..., kw]: kw = 3 for the 3x3 convolutions that hold 70 % of
// the SD1.5 UNet's weights, kw = 1 for the 1x1 ones).  A thread takes the smallest run of whole rows that is a whole
// number of 16-byte vectors (COLS = 3: 8 rows = 3 vectors), so loads and stores are coalesced 16-byte accesses instead
// of 2-byte ones.  Symmetric modes only (the per-channel path of fake_quant.py:86-93); outputs: dq and / or scales.
template <typename T, int MODE, int COLS>
__global__ void __launch_bounds__(kQThreads)
quant_rows_tiny_kernel(const T* __restrict__ x, int64_t n_chunks, float max_int, float min_int,
                       T* __restrict__ dq, T* __restrict__ scales) {
  constexpr int V = ElemTraits<T>::kVec;
  constexpr int G = (COLS % 8 == 0) ? 8 : (COLS % 4 == 0) ? 4 : (COLS % 2 == 0) ? 2 : 1;   // gcd(COLS, 8)
  constexpr int GV = (V == 8) ? G : ((COLS % 4 == 0) ? 4 : (COLS % 2 == 0) ? 2 : 1);      // gcd(COLS, V)
  constexpr int NVEC = COLS / GV;             // vectors per chunk
  constexpr int R = V / GV;                   // rows per chunk
  static_assert(NVEC * V == R * COLS, "chunk = whole rows = whole vectors");
  (void)G;
  const float floor_c = rnd<T>(1e-5f);
  const float r_max_int = rcp_approx(max_int);
  for (int64_t c = int64_t(blockIdx.x) * kQThreads + threadIdx.x; c < n_chunks; c += int64_t(gridDim.x) * kQThreads) {
    float v[NVEC * V];
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {
      Vec16<T> a = ld_vec16_stream(x + (c * NVEC + i) * V);
#pragma unroll
      for (int j = 0; j < V; ++j) v[i * V + j] = ElemTraits<T>::to_f(a.v[j]);
    }
    float o[NVEC * V];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float mx = 0.f;
#pragma unroll
      for (int e = 0; e < COLS; ++e) mx = fmaxf(mx, fabsf(v[r * COLS + e]));
      const float a = fmaxf(mx, floor_c);
      const bool fast = fastdiv_ok<T>(max_int, a);
      const float s = rnd<T>(fast ? div_by_rcp<false>(a, max_int, r_max_int) : __fdiv_rn(a, max_int));
      if (scales) scales[c * R + r] = ElemTraits<T>::from_f(s);
      const bool fast2 = fastdiv_ok<T>(s, mx);
      const float rs = rcp_approx(s);
#pragma unroll
      for (int e = 0; e < COLS; ++e) {
        const float w = v[r * COLS + e];
        float q = rint_T<T>(rnd<T>(fast2 ? div_by_rcp<false>(w, s, rs) : __fdiv_rn(w, s)));
        q = copysignf(q, w);
        if (MODE != Q_SYM_NOCLAMP) q = fminf(fmaxf(q, min_int), max_int);
        o[r * COLS + e] = __fmul_rn(q, s);
      }
    }
    if (dq) {
#pragma unroll
      for (int i = 0; i < NVEC; ++i) {
        Vec16<T> a;
#pragma unroll
        for (int j = 0; j < V; ++j) a.v[j] = ElemTraits<T>::from_f(o[i * V + j]);
        st_vec16(dq + (c * NVEC + i) * V, a);
      }
    }
  }
}

// ---------------------------------------------------------------- whole-tensor scale
template <typename T>
__global__ void __launch_bounds__(kQThreads)
quant_flat_kernel(const T* __restrict__ x, int64_t numel, const T* __restrict__ absmax_dev, float max_int,
                  T* __restrict__ dq, int8_t* __restrict__ codes, T* __restrict__ scale_out, int vec_ok) {
  constexpr int V = ElemTraits<T>::kVec;
  float s, z;
  group_params<T, Q_SYM_NOCLAMP>(ElemTraits<T>::to_f(absmax_dev[0]), 0.f, max_int, s, z);
  if (scale_out && blockIdx.x == 0 && threadIdx.x == 0) scale_out[0] = ElemTraits<T>::from_f(s);
  const int64_t tid = int64_t(blockIdx.x) * kQThreads + threadIdx.x;
  const int64_t nthreads = int64_t(gridDim.x) * kQThreads;
  const int64_t nvec = vec_ok ? numel / V : 0;
  // more than 8 bits (a_bit = 16, q_max = 32767): plain IEEE division, outside what the fast path was verified for
  const bool fast = max_int <= 255.f && fastdiv_ok<T>(s, ElemTraits<T>::to_f(absmax_dev[0]));
  const float dummy[V] = {};
  for (int64_t i = tid; i < nvec; i += nthreads) {
    Vec16<T> v = ld_vec16_stream(x + i * V);
    float xv[V], cq[V];
    Vec16<T> o;
#pragma unroll
    for (int j = 0; j < V; ++j) xv[j] = ElemTraits<T>::to_f(v.v[j]);
    if (fast) rtn_vec<T, Q_SYM_NOCLAMP, true, false, false>(xv, s, 0.f, 0.f, 0.f, dummy, dummy, o, cq);
    else rtn_vec<T, Q_SYM_NOCLAMP, false, false, false>(xv, s, 0.f, 0.f, 0.f, dummy, dummy, o, cq);
    if (codes) {
#pragma unroll
      for (int j = 0; j < V; ++j) codes[i * V + j] = code_to_i8<Q_SYM_NOCLAMP>(cq[j]);
    }
    if (dq) st_vec16(dq + i * V, o);
  }
  for (int64_t i = nvec * V + tid; i < numel; i += nthreads) {
    float code;
    float d = rtn_elem<T>(ElemTraits<T>::to_f(x[i]), s, 0.f, 0.f, 0.f, false, false, code);
    if (dq) dq[i] = ElemTraits<T>::from_f(d);
    if (codes) codes[i] = code_to_i8<Q_SYM_NOCLAMP>(code);
  }
}

// ---------------------------------------------------------------- per-token int8 activation codes
// The A8 of W8A8 (quantize_activation_per_token_absmax, fake_quant.py:109-118) emitting the int8
// codes and the per-token scale instead of the fake-quantised tensor.  Optional `smooth[K]`
// divides the activation first (SmoothQuant's x / s when s cannot be folded into a previous op).
template <typename T>
__global__ void __launch_bounds__(kQThreads)
actquant_token_i8_kernel(const T* __restrict__ x, int64_t rows, int64_t cols, const T* __restrict__ smooth,
                         int8_t* __restrict__ xq, float* __restrict__ sx, int vec_ok) {
  constexpr int V = ElemTraits<T>::kVec;
  const int lane = threadIdx.x & 31;
  const int wpb = kQThreads / 32;
  for (int64_t row = int64_t(blockIdx.x) * wpb + (threadIdx.x >> 5); row < rows;
       row += int64_t(gridDim.x) * wpb) {
    const T* p = x + row * cols;
    auto prep = [&](float v, int64_t e) {
      if (smooth) v = rnd<T>(__fdiv_rn(v, ElemTraits<T>::to_f(smooth[e])));
      return v;
    };
    float mx = 0.f;
    constexpr int U = 4;   // 16-byte loads per lane in flight (this kernel serves rows longer than 4096 elements, e.g. K = 5120)
    if (vec_ok) {
      for (int64_t e0 = int64_t(lane) * V; e0 < cols; e0 += int64_t(U) * 32 * V) {
        Vec16<T> v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (e0 + int64_t(u) * 32 * V < cols) v[u] = ld_vec16(p + e0 + int64_t(u) * 32 * V);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t e = e0 + int64_t(u) * 32 * V;
          if (e < cols) {
#pragma unroll
            for (int j = 0; j < V; ++j) mx = fmaxf(mx, fabsf(prep(ElemTraits<T>::to_f(v[u].v[j]), e + j)));
          }
        }
      }
    } else {
      for (int64_t e = lane; e < cols; e += 32) mx = fmaxf(mx, fabsf(prep(ElemTraits<T>::to_f(p[e]), e)));
    }
    mx = warp_max(mx);
    float s, z;
    group_params<T, Q_SYM_NOCLAMP>(mx, 0.f, 127.f, s, z);
    if (lane == 0) sx[row] = s;
    if (vec_ok) {
      for (int64_t e0 = int64_t(lane) * V; e0 < cols; e0 += int64_t(U) * 32 * V) {
        Vec16<T> v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (e0 + int64_t(u) * 32 * V < cols) v[u] = ld_vec16(p + e0 + int64_t(u) * 32 * V);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t e = e0 + int64_t(u) * 32 * V;
          if (e >= cols) continue;
          int8_t cb[V];
#pragma unroll
          for (int j = 0; j < V; ++j) {
            const float q = rintf(rnd<T>(__fdiv_rn(prep(ElemTraits<T>::to_f(v[u].v[j]), e + j), s)));
            cb[j] = code_to_i8<Q_SYM_NOCLAMP>(q);
          }
          if (V == 8) *reinterpret_cast<uint2*>(xq + row * cols + e) = *reinterpret_cast<const uint2*>(cb);
          else *reinterpret_cast<uint32_t*>(xq + row * cols + e) = *reinterpret_cast<const uint32_t*>(cb);
        }
      }
    } else {
      for (int64_t e = lane; e < cols; e += 32) {
        const float q = rintf(rnd<T>(__fdiv_rn(prep(ElemTraits<T>::to_f(p[e]), e), s)));
        xq[row * cols + e] = code_to_i8<Q_SYM_NOCLAMP>(q);
      }
    }
  }
}

// ---------------------------------------------------------------- AWQ int4 layout
__device__ __constant__ int kAwqOrder[8] = {0, 2, 4, 6, 1, 3, 5, 7};

constexpr int kPackTileN = 64;  // 8 packed words
constexpr int kPackTileK = 64;

// codes_nk int8 [N, K] -> qweight int32 [K, N/8]; transposes through shared memory.
__global__ void __launch_bounds__(256)
pack_awq_kernel(const int8_t* __restrict__ codes, int64_t n_rows, int64_t k_cols, int32_t* __restrict__ qweight) {
  __shared__ uint8_t tile[kPackTileN][kPackTileK + 4];
  const int64_t n0 = int64_t(blockIdx.y) * kPackTileN, k0 = int64_t(blockIdx.x) * kPackTileK;
  for (int idx = threadIdx.x; idx < kPackTileN * kPackTileK; idx += 256) {
    const int r = idx / kPackTileK, c = idx % kPackTileK;
    const int64_t n = n0 + r, k = k0 + c;
    tile[r][c] = (n < n_rows && k < k_cols) ? (uint8_t)codes[n * k_cols + k] : 0;
  }
  __syncthreads();
  const int64_t words_per_row = n_rows / 8;
  for (int idx = threadIdx.x; idx < kPackTileK * (kPackTileN / 8); idx += 256) {
    const int k = idx / (kPackTileN / 8), c = idx % (kPackTileN / 8);
    if (k0 + k >= k_cols || n0 + 8 * c >= n_rows) continue;
    uint32_t word = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) word |= uint32_t(tile[8 * c + kAwqOrder[i]][k] & 0xF) << (4 * i);
    qweight[(k0 + k) * words_per_row + n0 / 8 + c] = (int32_t)word;
  }
}

// qweight [K, N/8] -> codes_kn int8 [K, N] in natural column order
__global__ void __launch_bounds__(256)
unpack_awq_kernel(const int32_t* __restrict__ qweight, int64_t n_words, int8_t* __restrict__ codes) {
  for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < n_words; i += int64_t(gridDim.x) * 256) {
    const uint32_t w = (uint32_t)qweight[i];
    uint8_t out[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) out[kAwqOrder[j]] = (w >> (4 * j)) & 0xF;
    *reinterpret_cast<uint2*>(codes + i * 8) = *reinterpret_cast<const uint2*>(out);
  }
}

// Fused zero-point RTN + AWQ pack.  Block tile: kFuseTileN out-rows x one group of K.
// Phase 1: every thread issues all of its 16-byte loads, group statistics by lanes-per-group butterflies,
// the 8 k-consecutive 4-bit codes of a lane are packed into one word of shared memory.
// Phase 2: a thread takes the 8 words of output word-column c (rows 8c + AWQ order) for one k-octet,
// transposes the 8x8 nibble matrix in registers (3 butterfly stages) and stores 8 qweight words;
// scales / zero points leave through shared memory as contiguous rows of the [G, N] outputs.
constexpr int kFuseTileN = 128;

__device__ __forceinline__ void nib_swap(uint32_t& a, uint32_t& b, int sh, uint32_t mask) {
  const uint32_t t = ((a >> sh) ^ b) & mask;   // a = even member (keeps low blocks), b = odd member
  b ^= t;
  a ^= t << sh;
}

template <typename T, int LPG>                               // LPG = lanes (16-byte vectors) per group row
__global__ void __launch_bounds__(256)
quant_pack_awq_kernel(const T* __restrict__ w, int64_t n_rows, int64_t k_cols, float max_int,
                      int32_t* __restrict__ qweight, int32_t* __restrict__ qzeros,
                      T* __restrict__ scales_t, T* __restrict__ dq) {
  constexpr int V = ElemTraits<T>::kVec;
  constexpr int kGroup = LPG * V;
  constexpr int kRowsPerPass = 256 / LPG;
  constexpr int kPasses = kFuseTileN / kRowsPerPass;
  constexpr int kPitch = LPG + 1;
  constexpr int kOctets = kGroup / 8;                        // k-octets (packed words per row) in the tile
  __shared__ uint32_t tile[kFuseTileN * kPitch];             // nibble-packed codes, one word per k-octet
  __shared__ T ssm[kFuseTileN];                              // scales
  __shared__ uint8_t zsm[kFuseTileN];                        // zero points
  const int64_t n0 = int64_t(blockIdx.y) * kFuseTileN;
  const int64_t gi = blockIdx.x;                             // group index along K
  const int64_t k0 = gi * kGroup;
  const int sub = threadIdx.x % LPG, r0 = threadIdx.x / LPG;
  const int rows_here = (n_rows - n0) < kFuseTileN ? int(n_rows - n0) : kFuseTileN;
  const float r_max_int = rcp_approx(max_int);
  const T* src = w + (n0 + r0) * k_cols + k0 + sub * V;
  T* dst_dq = dq ? dq + (n0 + r0) * k_cols + k0 + sub * V : nullptr;
  const int64_t pass_stride = int64_t(kRowsPerPass) * k_cols;
  Vec16<T> raw[kPasses];
#pragma unroll
  for (int p = 0; p < kPasses; ++p)
    if (r0 + p * kRowsPerPass < rows_here) raw[p] = ld_vec16_stream(src + p * pass_stride);
#pragma unroll
  for (int p = 0; p < kPasses; ++p) {
    const int r = r0 + p * kRowsPerPass;
    const bool valid = r < rows_here;
    uint32_t pk = 0;
    if constexpr (std::is_same<T, __half>::value) {
      H2x4 x;
      if (valid) x = as_h2x4(raw[p]);
      else {
#pragma unroll
        for (int q = 0; q < 4; ++q) x.h[q] = __float2half2_rn(0.f);
      }
      float mx, mn, s, z, r_s;
      minmax_h<Q_ZP>(x, LPG, mx, mn);
      group_params_h<Q_ZP>(mx, mn, max_int, r_max_int, s, z, r_s);
      if (sub == 0) {
        ssm[r] = __float2half_rn(s);
        zsm[r] = (uint8_t)int_byte(z);
      }
      H2x4 o, cq;
      const float dummy8[8] = {};
      rtn_vec_h<Q_ZP, false>(x, s, r_s, z, 0.f, max_int, dummy8, dummy8, o, cq);
      // nibble j of the word = code of element j: the codes sit in the mantissas of (c + 1536)
      const __half2 zm = __float2half2_rn(__fadd_rn(z, 1536.f));
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t u = h2_bits(__hadd2(cq.h[q], zm)) & 0x000F000Fu;
        pk |= ((u | (u >> 12)) & 0xFFu) << (8 * q);
      }
      if (dst_dq && valid) *reinterpret_cast<uint4*>(dst_dq + p * pass_stride) = *reinterpret_cast<const uint4*>(&o);
    } else {
      float x[V];
#pragma unroll
      for (int j = 0; j < V; ++j) x[j] = valid ? ElemTraits<T>::to_f(raw[p].v[j]) : 0.f;
      float mx = x[0], mn = x[0];
#pragma unroll
      for (int j = 1; j < V; ++j) { mx = fmaxf(mx, x[j]); mn = fminf(mn, x[j]); }
#pragma unroll
      for (int o = 1; o < LPG; o <<= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      }
      float s, z;
      group_params<T, Q_ZP>(mx, mn, max_int, s, z);
      if (sub == 0) {
        ssm[r] = ElemTraits<T>::from_f(s);
        zsm[r] = (uint8_t)int_byte(z);
      }
      float cq[V];
      Vec16<T> o;
      const float dummy[V] = {};
      if (fastdiv_ok<T>(s, fmaxf(fabsf(mx), fabsf(mn)))) rtn_vec<T, Q_ZP, true, false, false>(x, s, z, 0.f, max_int, dummy, dummy, o, cq);
      else rtn_vec<T, Q_ZP, false, false, false>(x, s, z, 0.f, max_int, dummy, dummy, o, cq);
#pragma unroll
      for (int j = 0; j < V; ++j) pk |= (int_byte(__fadd_rn(cq[j], z)) & 0xFu) << (4 * j);
      if (dst_dq && valid) st_vec16(dst_dq + p * pass_stride, o);
    }
    if (V == 8) {
      tile[r * kPitch + sub] = pk;
    } else {                                                 // fp32: two lanes share one k-octet word
      const uint32_t other = __shfl_xor_sync(0xffffffffu, pk, 1);
      if ((sub & 1) == 0) tile[r * kPitch + (sub >> 1)] = pk | (other << 16);
    }
  }
  __syncthreads();
  const int64_t words_per_row = n_rows / 8;
  constexpr int kWordCols = kFuseTileN / 8;                  // 16 packed words across the tile
  for (int idx = threadIdx.x; idx < kOctets * kWordCols; idx += 256) {
    const int kb = idx / kWordCols, c = idx % kWordCols;
    if (8 * c >= rows_here) continue;
    uint32_t m[8];                                           // m[i] = row 8c + order[i]; nibble t = code at k = 8kb + t
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = tile[(8 * c + 2 * (i & 3) + (i >> 2)) * kPitch + kb];
    // 8x8 nibble transpose: afterwards m[t] nibble i = old m[i] nibble t
    nib_swap(m[0], m[1], 4, 0x0F0F0F0Fu); nib_swap(m[2], m[3], 4, 0x0F0F0F0Fu);
    nib_swap(m[4], m[5], 4, 0x0F0F0F0Fu); nib_swap(m[6], m[7], 4, 0x0F0F0F0Fu);
    nib_swap(m[0], m[2], 8, 0x00FF00FFu); nib_swap(m[1], m[3], 8, 0x00FF00FFu);
    nib_swap(m[4], m[6], 8, 0x00FF00FFu); nib_swap(m[5], m[7], 8, 0x00FF00FFu);
    nib_swap(m[0], m[4], 16, 0x0000FFFFu); nib_swap(m[1], m[5], 16, 0x0000FFFFu);
    nib_swap(m[2], m[6], 16, 0x0000FFFFu); nib_swap(m[3], m[7], 16, 0x0000FFFFu);
    int32_t* dst = qweight + (k0 + 8 * kb) * words_per_row + n0 / 8 + c;
#pragma unroll
    for (int t = 0; t < 8; ++t) dst[t * words_per_row] = (int32_t)m[t];
  }
  if (threadIdx.x < kWordCols && 8 * int(threadIdx.x) < rows_here) {
    const int c = threadIdx.x;
    uint32_t word = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) word |= uint32_t(zsm[8 * c + 2 * (i & 3) + (i >> 2)] & 0xF) << (4 * i);
    qzeros[gi * words_per_row + n0 / 8 + c] = (int32_t)word;
  }
  for (int r = threadIdx.x; r < rows_here; r += 256) scales_t[gi * n_rows + n0 + r] = ssm[r];
}

// W_kn[k, 8c + j] = (q - z) * s, one packed word (8 outputs, 16-byte scale load and store for 2-byte T)
// per thread.  q - z is formed exactly in fp32 from the nibble without an I2F: the nibble is or-ed into the
// mantissa of 2^23 and (2^23 + z) subtracted.
template <typename T>
__global__ void __launch_bounds__(256)
dequant_awq_kernel(const int32_t* __restrict__ qweight, const int32_t* __restrict__ qzeros,
                   const T* __restrict__ scales, uint32_t total_words, uint32_t n_words, uint32_t group,
                   T* __restrict__ out) {
  const uint32_t stride = gridDim.x * 256u;
  for (uint32_t i0 = blockIdx.x * 256u + threadIdx.x; i0 < total_words; i0 += 2 * stride) {
    // two words per thread-step: all six loads are issued before the first unpack
    uint32_t qw[2], zw[2];
    uint4 sv[2];
    const T* sp[2];
    bool on[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t i = i0 + u * stride;
      on[u] = i < total_words;
      const uint32_t ii = on[u] ? i : i0;
      const uint32_t k = ii / n_words, c = ii - k * n_words;
      const uint32_t g = k / group;
      qw[u] = (uint32_t)__ldg(qweight + ii);
      zw[u] = (uint32_t)__ldg(qzeros + g * n_words + c);
      sp[u] = scales + (int64_t(g) * n_words + c) * 8;
      if constexpr (sizeof(T) == 2) sv[u] = __ldg(reinterpret_cast<const uint4*>(sp[u]));
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!on[u]) continue;
      T* dst = out + int64_t(i0 + u * stride) * 8;
      if constexpr (sizeof(T) == 2) {
        // 16-bit types, packed: nibbles i and i + 4 of a word are the ADJACENT columns 2i, 2i + 1 (AWQ order), so
        // (w >> 4i) & 0x000F000F or-ed into the mantissas of (1024 | 1024) [fp16] / (128 | 128) [bf16] is the pair
        // (magic + q_2i, magic + q_2i+1); the packed subtract of the zero-point pair is exact and the packed multiply by
        // the scale pair rounds once -- the same value as the fp32 form below, 6 instructions per two outputs, not ~16
        constexpr uint32_t MAGIC = std::is_same<T, __half>::value ? 0x64006400u : 0x43004300u;
        const uint32_t s2[4] = {sv[u].x, sv[u].y, sv[u].z, sv[u].w};
        uint32_t o2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t q2 = ((qw[u] >> (4 * i)) & 0x000F000Fu) | MAGIC;
          const uint32_t z2 = ((zw[u] >> (4 * i)) & 0x000F000Fu) | MAGIC;
          if constexpr (std::is_same<T, __half>::value) {
            const __half2 r = __hmul2(__hsub2(*reinterpret_cast<const __half2*>(&q2), *reinterpret_cast<const __half2*>(&z2)),
                                      *reinterpret_cast<const __half2*>(&s2[i]));
            o2[i] = *reinterpret_cast<const uint32_t*>(&r);
          } else {
            const __nv_bfloat162 r = __hmul2(__hsub2(*reinterpret_cast<const __nv_bfloat162*>(&q2), *reinterpret_cast<const __nv_bfloat162*>(&z2)),
                                             *reinterpret_cast<const __nv_bfloat162*>(&s2[i]));
            o2[i] = *reinterpret_cast<const uint32_t*>(&r);
          }
        }
        asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(o2[0]), "r"(o2[1]), "r"(o2[2]), "r"(o2[3]) : "memory");
      } else {
        T o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float qf = __uint_as_float(((qw[u] >> (4 * j)) & 0xFu) | 0x4B000000u);
          const float zf = __uint_as_float(((zw[u] >> (4 * j)) & 0xFu) | 0x4B000000u);
          const int col = 2 * (j & 3) + (j >> 2);                // AWQ order {0,2,4,6,1,3,5,7}[j]
          o[col] = ElemTraits<T>::from_f(__fmul_rn(__fsub_rn(qf, zf), ElemTraits<T>::to_f(sp[u][col])));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = o[j];
      }
    }
  }
}

int grid_for(int64_t work_items, int per_block) {
  int64_t b = (work_items + per_block - 1) / per_block;
  const int64_t cap = int64_t(QDM_NUM_SMS) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return int(b);
}

// kQThreads-wide launch, with programmatic stream serialisation when PDL (the kernel then starts with
// griddepcontrol.launch_dependents + griddepcontrol.wait); QDM_NO_PDL=1 launches plainly (A/B switch, read once)
template <bool PDL, typename Kern, typename... Args>
cudaError_t launch_maybe_pdl(Kern kern, unsigned grid, cudaStream_t st, Args... args) {
  static const bool no_pdl = getenv("QDM_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kQThreads);
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = (PDL && !no_pdl) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

template <typename T, int MODE, bool I8OUT>
int launch_rows_reg(const T* x, int64_t rows, int64_t cols, float max_int, float min_int,
                    T* dq, int8_t* codes, T* scales, T* zeros, float* sx, cudaStream_t st) {
  constexpr int V = ElemTraits<T>::kVec;
  const int64_t nv = (cols + 32 * V - 1) / (32 * V);
  // persistent grid: the CTAs the device holds at once (equal work per CTA, no partial last wave)
#define QDM_ROWS_REG(NV)                                                                              \
  do {                                                                                                \
    static int occ = 0;                                                                               \
    if (occ == 0) {                                                                                   \
      QDM_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, quant_rows_reg_kernel<T, MODE, NV, I8OUT>, kQThreads, 0)); \
      if (occ < 1) occ = 1;                                                                           \
    }                                                                                                 \
    int64_t grid = (rows + kQThreads / 32 - 1) / (kQThreads / 32);                                    \
    if (grid > int64_t(QDM_NUM_SMS) * occ) grid = int64_t(QDM_NUM_SMS) * occ;                         \
    QDM_CUDA_OK(launch_maybe_pdl<I8OUT>(quant_rows_reg_kernel<T, MODE, NV, I8OUT>, (unsigned)grid, st, x, rows, int(cols), max_int, \
                                        min_int, dq, codes, scales, zeros, sx));                          \
  } while (0)
  // short rows share a warp (<= 8 vectors per lane at 8 / 16 lanes per row)
#define QDM_ROWS_SUB(LPR)                                                                             \
  do {                                                                                                \
    static int occ = 0;                                                                               \
    if (occ == 0) {                                                                                   \
      QDM_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, quant_rows_reg_kernel<T, MODE, 8, I8OUT, LPR>, kQThreads, 0)); \
      if (occ < 1) occ = 1;                                                                           \
    }                                                                                                 \
    const int64_t rpc = int64_t(kQThreads / 32) * (32 / LPR);                                         \
    int64_t grid = (rows + rpc - 1) / rpc;                                                            \
    if (grid > int64_t(QDM_NUM_SMS) * occ) grid = int64_t(QDM_NUM_SMS) * occ;                         \
    QDM_CUDA_OK(launch_maybe_pdl<I8OUT>(quant_rows_reg_kernel<T, MODE, 8, I8OUT, LPR>, (unsigned)grid, st, x, rows, int(cols),   \
                                        max_int, min_int, dq, codes, scales, zeros, sx));                 \
  } while (0)
  const int64_t vecs = (cols + V - 1) / V;
  if (vecs > 32 && vecs <= 64 && rows >= 1024) QDM_ROWS_SUB(8);
  else if (vecs > 64 && vecs <= 128 && rows >= 1024) QDM_ROWS_SUB(16);
  else if (nv <= 1) QDM_ROWS_REG(1);
  else if (nv <= 2) QDM_ROWS_REG(2);
  else if (nv <= 4) QDM_ROWS_REG(4);
  else if (nv <= 8) QDM_ROWS_REG(8);
  else QDM_ROWS_REG(16);
#undef QDM_ROWS_REG
#undef QDM_ROWS_SUB
  QDM_LAUNCH_CHECK();
  return QDM_OK;
}

template <typename T, int MODE>
int launch_quant(const T* w, int64_t n_groups, int64_t group, int64_t k_period, int n_bits,
                 const T* pre_mul, const T* clip_max, const T* post_div,
                 T* dq, int8_t* codes, T* scales, T* zeros, cudaStream_t st) {
  constexpr int V = ElemTraits<T>::kVec;
  const float max_int = (MODE == Q_ZP) ? float((1 << n_bits) - 1) : float((1 << (n_bits - 1)) - 1);
  const float min_int = (MODE == Q_ZP) ? 0.f : -float(1 << (n_bits - 1));
  const int64_t numel = n_groups * group;
  const bool aligned = qdm_aligned16(w) && (!dq || qdm_aligned16(dq)) && (!pre_mul || qdm_aligned16(pre_mul)) &&
                       (!post_div || qdm_aligned16(post_div)) &&
                       (!codes || (reinterpret_cast<uintptr_t>(codes) & 7u) == 0);
  const int64_t lpg = group / V;
  const int64_t n_rows = numel / k_period;
  const bool extras = pre_mul || clip_max || post_div;
  // more than 8 bits (row-wise a_bit = 16, q_max = 32767): the packed fp16 chain's magic-number rounding and the reciprocal
  // division were verified for small codes only, so these go through the op-by-op IEEE kernel at the end
  const bool wide = max_int > 255.f;
  if (!wide && aligned && group % V == 0 && lpg <= 32 && (lpg & (lpg - 1)) == 0 && n_rows < (int64_t(1) << 31) &&
      k_period / V < (int64_t(1) << 20)) {
    const int64_t vpr = k_period / V;
    int lpg_shift = 0;
    while ((int64_t(1) << lpg_shift) < lpg) ++lpg_shift;
    // persistent grid = exactly the CTAs the device holds at once (no partial second wave), and whole
    // passes so that every pass covers the same number of rows
    static int occ_plain = 0, occ_extras = 0;
    int& occ = extras ? occ_extras : occ_plain;
    if (occ == 0) {
      if (extras) QDM_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, quant_group_kernel<T, MODE, true>, kQThreads, 0));
      else QDM_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, quant_group_kernel<T, MODE, false>, kQThreads, 0));
      if (occ < 1) occ = 1;
    }
    const int64_t target_threads = int64_t(QDM_NUM_SMS) * occ * kQThreads;
    int64_t r_max = target_threads / vpr;
    if (r_max < 1) r_max = 1;
    const int64_t passes = (n_rows + r_max - 1) / r_max;
    const int64_t rpp = (n_rows + passes - 1) / passes;
    const unsigned grid = (unsigned)((rpp * vpr + kQThreads - 1) / kQThreads);
    if (extras)
      quant_group_kernel<T, MODE, true><<<grid, kQThreads, 0, st>>>(
          w, int(n_rows), int(vpr), lpg_shift, int(rpp), max_int, min_int, pre_mul, clip_max, post_div, dq, codes, scales, zeros);
    else
      quant_group_kernel<T, MODE, false><<<grid, kQThreads, 0, st>>>(
          w, int(n_rows), int(vpr), lpg_shift, int(rpp), max_int, min_int, nullptr, nullptr, nullptr, dq, codes, scales, zeros);
  } else if (!wide && group <= 16 && !extras) {
    // whole chunks of rows through the vector kernel (symmetric modes, no int8 codes), the remaining rows one per thread
    int64_t done_rows = 0;
    if (MODE != Q_ZP && !codes && aligned && group < V) {
      const int gv = (group % 8 == 0 && V == 8) ? 8 : (group % 4 == 0) ? 4 : (group % 2 == 0) ? 2 : 1;
      const int64_t rows_per_chunk = V / gv, n_chunks = n_groups / rows_per_chunk;
      if (n_chunks > 0) {
        const int grid = grid_for(n_chunks, kQThreads);
        switch (group) {
#define QDM_TINY(C) case C: quant_rows_tiny_kernel<T, MODE, C><<<grid, kQThreads, 0, st>>>(w, n_chunks, max_int, min_int, dq, scales); break;
          QDM_TINY(1) QDM_TINY(2) QDM_TINY(3) QDM_TINY(4) QDM_TINY(5) QDM_TINY(6) QDM_TINY(7)
#undef QDM_TINY
          default: break;
        }
        QDM_LAUNCH_CHECK();
        done_rows = n_chunks * rows_per_chunk;
      }
    }
    if (done_rows < n_groups) {
      const int64_t rest = n_groups - done_rows, off = done_rows * group;
      quant_rows_thread_kernel<T, MODE><<<grid_for(rest, kQThreads), kQThreads, 0, st>>>(
          w + off, rest, int(group), max_int, min_int, dq ? dq + off : nullptr, codes ? codes + off : nullptr,
          scales ? scales + done_rows : nullptr, zeros ? zeros + done_rows : nullptr);
    } else {
      return QDM_OK;
    }
  } else if (!wide && aligned && !extras && group % V == 0 && group <= 32 * V * 16) {
    return launch_rows_reg<T, MODE, false>(w, n_groups, group, max_int, min_int, dq, codes, scales, zeros, nullptr, st);
  } else {
    const int vec_ok = aligned && group % V == 0;
    quant_rows_warp_kernel<T, MODE><<<grid_for(n_groups, kQThreads / 32), kQThreads, 0, st>>>(
        w, n_groups, group, k_period, max_int, min_int, pre_mul, clip_max, post_div, dq, codes, scales, zeros, vec_ok);
  }
  QDM_LAUNCH_CHECK();
  return QDM_OK;
}


// ---------------------------------------------------------------- AWQ clip search (quantize/quantizer.py:805-863)
// For every (out-row, group) the reference tries 10 shrink levels max_i = org_max * (1 - i/20), quantises the clamped
// group and keeps the level with the smallest  err_i = mean_tok ( sum_k x[t,k] q_i[k]  -  sum_k x[t,k] w[k] )^2 .
// It materialises co_b x n_tok x K products eleven times per batch of rows.  By linearity
//     err_i = d_i^T C d_i ,   d_i = q_i - w ,   C = X_g^T X_g / n_tok   (the group's g x g Gram matrix, fp32),
// and C does not depend on the out-row: one small pass over the sampled activations builds all G Gram matrices
// (group_gram_kernel), and the search itself (awq_clip_kernel) then costs g^2 FMAs per (row, group, level) instead of
// n_tok * g -- 4x fewer for n_tok = 512, g = 128 -- with no temporaries at all.  Q_i is the bit-exact RTN chain of this
// file; the quadratic form is fp32 (the reference rounds products and sums to fp16, so near-ties may pick the
// neighbouring level: tests bound the agreement and the loss ratio).
//
// awq_clip_kernel: CTA = one group (its Gram matrix in shared memory, transposed access C[k][j] is conflict free since C is
// symmetric) x a range of out-rows; warp = one row at a time, lane = E = g / 32 consecutive k.  Phase 1 computes the ten
// d_i (kept in registers and written k-major to a per-warp shared tile), phase 2 accumulates y_i = C d_i for all ten
// levels at once (per k: one 16-byte load of C, three broadcast loads of d, 10 E FMAs), phase 3 reduces d_i . y_i.
constexpr int kClipLevels = 10;
constexpr int kClipDPad = 12;        // floats per k row of the per-warp d tile (10 levels + pad: 16-byte aligned rows)
constexpr int kClipWarps = 8;

template <typename T, int GS>
__global__ void __launch_bounds__(256)
group_gram_kernel(const T* __restrict__ x, int64_t n_tok, int64_t ld, float* __restrict__ gram) {
  // grid (G, GS / 16): block = group g, 16 rows a of its Gram matrix; thread = (a, 8 columns b)
  constexpr int CH = 32;                             // tokens staged per step
  __shared__ float xs[CH][GS + 1];
  const int g = blockIdx.x, a = blockIdx.y * 16 + (threadIdx.x >> 4), b0 = (threadIdx.x & 15) * (GS / 16);
  float acc[GS / 16];
#pragma unroll
  for (int j = 0; j < GS / 16; ++j) acc[j] = 0.f;
  for (int64_t t0 = 0; t0 < n_tok; t0 += CH) {
    for (int i = threadIdx.x; i < CH * GS; i += 256) {
      const int tt = i / GS, c = i % GS;
      xs[tt][c] = (t0 + tt < n_tok) ? ElemTraits<T>::to_f(x[(t0 + tt) * ld + int64_t(g) * GS + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int tt = 0; tt < CH; ++tt) {
      const float xa = xs[tt][a];
#pragma unroll
      for (int j = 0; j < GS / 16; ++j) acc[j] = __fmaf_rn(xa, xs[tt][b0 + j], acc[j]);
    }
    __syncthreads();
  }
  const float inv = 1.f / float(n_tok);
#pragma unroll
  for (int j = 0; j < GS / 16; ++j) gram[(int64_t(g) * GS + a) * GS + b0 + j] = acc[j] * inv;
}

template <typename T, int GS, bool ZP>
__global__ void __launch_bounds__(32 * kClipWarps)
awq_clip_kernel(const T* __restrict__ w, int64_t co, int64_t ci, const float* __restrict__ gram, float max_int, float min_int,
                int n_grid, int n_levels, int rows_per_cta, T* __restrict__ best_max) {
  constexpr int E = GS / 32;
  extern __shared__ float clip_smem[];
  float* C = clip_smem;                                       // [GS][GS]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* dT = clip_smem + GS * GS + warp * (GS * kClipDPad);  // [GS][kClipDPad], this warp's
  const int g = blockIdx.x, G = int(ci / GS);
  for (int i = threadIdx.x; i < GS * GS / 4; i += blockDim.x)
    reinterpret_cast<float4*>(C)[i] = reinterpret_cast<const float4*>(gram + int64_t(g) * GS * GS)[i];
  __syncthreads();
  const int64_t r0 = int64_t(blockIdx.y) * rows_per_cta, r1 = min(co, r0 + rows_per_cta);
  for (int64_t r = r0 + warp; r < r1; r += kClipWarps) {
    float wv[E];
    {
      const T* p = w + r * ci + int64_t(g) * GS + lane * E;
#pragma unroll
      for (int e = 0; e < E; ++e) wv[e] = ElemTraits<T>::to_f(p[e]);
    }
    float amax = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) amax = fmaxf(amax, fabsf(wv[e]));
    amax = warp_max(amax);                                    // org_max_val (quantizer.py:833), exact
    float d_own[kClipLevels][E], maxv[kClipLevels];
#pragma unroll
    for (int i = 0; i < kClipLevels; ++i) {
      // max_val = org_max_val * (1 - i_s / n_grid): dtype tensor x python float = fp32 product rounded to the dtype
      const float mv = rnd<T>(__fmul_rn(amax, float(1.0 - double(i) / double(n_grid))));
      maxv[i] = mv;
      float c[E], mx = ZP ? -INFINITY : 0.f, mn = INFINITY;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        c[e] = fminf(fmaxf(wv[e], -mv), mv);                  // torch.clamp(w, min_val, max_val)
        if (ZP) { mx = fmaxf(mx, c[e]); mn = fminf(mn, c[e]); } else mx = fmaxf(mx, fabsf(c[e]));
      }
      mx = warp_max(mx);
      if (ZP) mn = -warp_max(-mn);
      float s, z;
      group_params<T, ZP ? Q_ZP : Q_SYM>(mx, mn, max_int, s, z);
#pragma unroll
      for (int e = 0; e < E; ++e) {
        float code;
        const float q = rtn_elem<T>(c[e], s, z, ZP ? 0.f : min_int, max_int, ZP, true, code);
        d_own[i][e] = (i < n_levels) ? q - wv[e] : 0.f;
        dT[(lane * E + e) * kClipDPad + i] = d_own[i][e];
      }
    }
    __syncwarp();
    float y[kClipLevels][E];
#pragma unroll
    for (int i = 0; i < kClipLevels; ++i)
#pragma unroll
      for (int e = 0; e < E; ++e) y[i][e] = 0.f;
#pragma unroll 2
    for (int k = 0; k < GS; ++k) {
      float cv[E];
      if (E == 4) {
        const float4 t = *reinterpret_cast<const float4*>(C + k * GS + lane * 4);
        cv[0] = t.x; cv[1] = t.y; cv[2] = t.z; cv[E - 1] = t.w;
      } else {
        const float2 t = *reinterpret_cast<const float2*>(C + k * GS + lane * 2);
        cv[0] = t.x; cv[E - 1] = t.y;
      }
      const float4 d0 = *reinterpret_cast<const float4*>(dT + k * kClipDPad);
      const float4 d1 = *reinterpret_cast<const float4*>(dT + k * kClipDPad + 4);
      const float2 d2 = *reinterpret_cast<const float2*>(dT + k * kClipDPad + 8);
      const float dk[kClipLevels] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w, d2.x, d2.y};
#pragma unroll
      for (int i = 0; i < kClipLevels; ++i)
#pragma unroll
        for (int e = 0; e < E; ++e) y[i][e] = __fmaf_rn(cv[e], dk[i], y[i][e]);
    }
    float best = amax, min_err = INFINITY;                    // min_errs = ones * 1e9 is +inf in fp16 (quantizer.py:836)
#pragma unroll
    for (int i = 0; i < kClipLevels; ++i) {
      float err = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) err = __fmaf_rn(y[i][e], d_own[i][e], err);
      err = warp_sum(err);
      if (i < n_levels && err < min_err) { min_err = err; best = maxv[i]; }   // strict <: the first minimum (quantizer.py:851)
    }
    if (lane == 0) best_max[r * G + g] = ElemTraits<T>::from_f(best);
    __syncwarp();   // the d tile is rewritten by the next row
  }
}

template <typename T>
int dispatch_mode(unsigned flags, const T* w, int64_t n_groups, int64_t group, int64_t k_period, int n_bits,
                  const T* pre_mul, const T* clip_max, const T* post_div,
                  T* dq, int8_t* codes, T* scales, T* zeros, cudaStream_t st) {
  if (flags & QDM_Q_ZERO_POINT)
    return launch_quant<T, Q_ZP>(w, n_groups, group, k_period, n_bits, pre_mul, clip_max, post_div, dq, codes, scales, zeros, st);
  if (flags & QDM_Q_NO_CLAMP)
    return launch_quant<T, Q_SYM_NOCLAMP>(w, n_groups, group, k_period, n_bits, pre_mul, clip_max, post_div, dq, codes, scales, zeros, st);
  return launch_quant<T, Q_SYM>(w, n_groups, group, k_period, n_bits, pre_mul, clip_max, post_div, dq, codes, scales, zeros, st);
}

}  // namespace

#define QDM_DISPATCH_DTYPE(dtype, ...)                                  \
  switch (dtype) {                                                      \
    case QDM_F16: { using T = __half; __VA_ARGS__; } break;             \
    case QDM_BF16: { using T = __nv_bfloat16; __VA_ARGS__; } break;     \
    case QDM_F32: { using T = float; __VA_ARGS__; } break;              \
    default: qdm_set_error("unknown dtype %d", dtype); return QDM_ERR_INVALID; \
  }

extern "C" int qdm_quant_group(const void* w, int dtype, int64_t n_rows, int64_t k_cols, int group,
                               int n_bits, unsigned flags,
                               const void* pre_mul, const void* clip_max, const void* post_div,
                               void* dq, int8_t* codes, void* scales, void* zeros, void* stream) {
  QDM_REQUIRE(w, "qdm_quant_group: null weight");
  QDM_REQUIRE(n_rows > 0 && k_cols > 0, "qdm_quant_group: empty tensor [%lld, %lld]", (long long)n_rows, (long long)k_cols);
  QDM_REQUIRE(group > 0 && k_cols % group == 0, "qdm_quant_group: group %d must divide k_cols %lld", group, (long long)k_cols);
  QDM_REQUIRE(n_bits >= 2 && n_bits <= 8, "qdm_quant_group: n_bits %d outside [2, 8]", n_bits);
  QDM_REQUIRE((flags & ~(QDM_Q_ZERO_POINT | QDM_Q_NO_CLAMP)) == 0 &&
              (flags & (QDM_Q_ZERO_POINT | QDM_Q_NO_CLAMP)) != (QDM_Q_ZERO_POINT | QDM_Q_NO_CLAMP),
              "qdm_quant_group: bad flags 0x%x", flags);
  QDM_DEVICE_GATE();
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_groups = n_rows * (k_cols / group);
  QDM_DISPATCH_DTYPE(dtype, return (dispatch_mode<T>(flags, (const T*)w, n_groups, group, k_cols, n_bits,
                                                     (const T*)pre_mul, (const T*)clip_max, (const T*)post_div,
                                                     (T*)dq, codes, (T*)scales, (T*)zeros, st)));
  return QDM_OK;
}

extern "C" int qdm_quant_rowwise(const void* x, int dtype, int64_t rows, int64_t cols, int n_bits, unsigned flags,
                                 void* dq, int8_t* codes, void* scales, void* zeros, void* stream) {
  QDM_REQUIRE(x, "qdm_quant_rowwise: null input");
  QDM_REQUIRE(rows > 0 && cols > 0, "qdm_quant_rowwise: empty tensor [%lld, %lld]", (long long)rows, (long long)cols);
  // up to 16 bits for the fake-quant output (the reference's default a_bit = 16, q_max = 32767: fake_quant.py:112);
  // integer codes are one byte, so they exist for <= 8 bits only
  QDM_REQUIRE(n_bits >= 2 && n_bits <= 16, "qdm_quant_rowwise: n_bits %d outside [2, 16]", n_bits);
  QDM_REQUIRE(n_bits <= 8 || !codes, "qdm_quant_rowwise: int8 codes need n_bits <= 8 (got %d)", n_bits);
  QDM_REQUIRE((flags & ~(QDM_Q_ZERO_POINT | QDM_Q_NO_CLAMP)) == 0 &&
              (flags & (QDM_Q_ZERO_POINT | QDM_Q_NO_CLAMP)) != (QDM_Q_ZERO_POINT | QDM_Q_NO_CLAMP),
              "qdm_quant_rowwise: bad flags 0x%x", flags);
  QDM_DEVICE_GATE();
  cudaStream_t st = (cudaStream_t)stream;
  QDM_DISPATCH_DTYPE(dtype, return (dispatch_mode<T>(flags, (const T*)x, rows, cols, cols, n_bits, nullptr, nullptr, nullptr,
                                                     (T*)dq, codes, (T*)scales, (T*)zeros, st)));
  return QDM_OK;
}

extern "C" int qdm_actquant_token_i8(const void* x, int dtype, int64_t rows, int64_t cols, const void* smooth,
                                     int8_t* xq, float* sx, void* stream) {
  QDM_REQUIRE(x && xq && sx, "qdm_actquant_token_i8: null pointer");
  QDM_REQUIRE(rows > 0 && cols > 0, "qdm_actquant_token_i8: empty tensor [%lld, %lld]", (long long)rows, (long long)cols);
  QDM_DEVICE_GATE();
  cudaStream_t st = (cudaStream_t)stream;
  QDM_DISPATCH_DTYPE(dtype, {
    constexpr int V = ElemTraits<T>::kVec;
    const int vec_ok = qdm_aligned16(x) && cols % V == 0 && (reinterpret_cast<uintptr_t>(xq) & 7u) == 0;
    if (vec_ok && !smooth && cols <= 32 * V * 16)
      return (launch_rows_reg<T, Q_SYM_NOCLAMP, true>((const T*)x, rows, cols, 127.f, -128.f, nullptr, xq, nullptr,
                                                      nullptr, sx, st));
    actquant_token_i8_kernel<T><<<grid_for(rows, kQThreads / 32), kQThreads, 0, st>>>(
        (const T*)x, rows, cols, (const T*)smooth, xq, sx, vec_ok);
    QDM_LAUNCH_CHECK();
  });
  return QDM_OK;
}

extern "C" int qdm_quant_tensor(const void* x, int dtype, int64_t numel, int n_bits,
                                void* dq, int8_t* codes, void* scale_out,
                                void* workspace, size_t workspace_bytes, void* stream) {
  QDM_REQUIRE(x && workspace, "qdm_quant_tensor: null pointer");
  QDM_REQUIRE(numel > 0, "qdm_quant_tensor: empty tensor");
  QDM_REQUIRE(n_bits >= 2 && n_bits <= 16, "qdm_quant_tensor: n_bits %d outside [2, 16]", n_bits);
  QDM_REQUIRE(n_bits <= 8 || !codes, "qdm_quant_tensor: int8 codes need n_bits <= 8 (got %d)", n_bits);
  const size_t head = 256;  // absmax scalar lives at the start of the workspace
  QDM_REQUIRE(workspace_bytes >= head + qdm_absmax_workspace_bytes(numel), "qdm_quant_tensor: workspace too small");
  QDM_DEVICE_GATE();
  cudaStream_t st = (cudaStream_t)stream;
  int rc = qdm_absmax_impl(x, dtype, numel, workspace, (char*)workspace + head, workspace_bytes - head, st);
  if (rc != QDM_OK) return rc;
  const float max_int = float((1 << (n_bits - 1)) - 1);
  QDM_DISPATCH_DTYPE(dtype, {
    constexpr int V = ElemTraits<T>::kVec;
    const int vec_ok = qdm_aligned16(x) && (!dq || qdm_aligned16(dq));
    quant_flat_kernel<T><<<grid_for(numel / V + 1, kQThreads), kQThreads, 0, st>>>(
        (const T*)x, numel, (const T*)workspace, max_int, (T*)dq, codes, (T*)scale_out, vec_ok);
    QDM_LAUNCH_CHECK();
  });
  return QDM_OK;
}

extern "C" size_t qdm_quant_tensor_workspace_bytes(int64_t numel) {
  return 256 + qdm_absmax_workspace_bytes(numel);
}

extern "C" int qdm_pack_awq(const int8_t* codes_nk, int64_t n_rows, int64_t k_cols, int32_t* qweight, void* stream) {
  QDM_REQUIRE(codes_nk && qweight, "qdm_pack_awq: null pointer");
  QDM_REQUIRE(n_rows > 0 && k_cols > 0 && n_rows % 8 == 0, "qdm_pack_awq: n_rows %lld must be a positive multiple of 8",
              (long long)n_rows);
  QDM_DEVICE_GATE();
  dim3 grid((unsigned)((k_cols + kPackTileK - 1) / kPackTileK), (unsigned)((n_rows + kPackTileN - 1) / kPackTileN));
  pack_awq_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(codes_nk, n_rows, k_cols, qweight);
  QDM_LAUNCH_CHECK();
  return QDM_OK;
}

extern "C" int qdm_unpack_awq(const int32_t* qweight, int64_t k_rows, int64_t n_cols, int8_t* codes_kn, void* stream) {
  QDM_REQUIRE(qweight && codes_kn, "qdm_unpack_awq: null pointer");
  QDM_REQUIRE(k_rows > 0 && n_cols > 0 && n_cols % 8 == 0, "qdm_unpack_awq: n_cols %lld must be a positive multiple of 8",
              (long long)n_cols);
  QDM_REQUIRE((reinterpret_cast<uintptr_t>(codes_kn) & 7u) == 0, "qdm_unpack_awq: codes must be 8-byte aligned");
  QDM_DEVICE_GATE();
  const int64_t n_words = k_rows * (n_cols / 8);
  unpack_awq_kernel<<<grid_for(n_words, 256), 256, 0, (cudaStream_t)stream>>>(qweight, n_words, codes_kn);
  QDM_LAUNCH_CHECK();
  return QDM_OK;
}

namespace {
template <typename T>
void launch_pack(int lpg, dim3 grid, const T* w, int64_t n_rows, int64_t k_cols, int32_t* qweight, int32_t* qzeros,
                 T* scales_t, T* dq, cudaStream_t st) {
  switch (lpg) {
    case 4: quant_pack_awq_kernel<T, 4><<<grid, 256, 0, st>>>(w, n_rows, k_cols, 15.f, qweight, qzeros, scales_t, dq); break;
    case 8: quant_pack_awq_kernel<T, 8><<<grid, 256, 0, st>>>(w, n_rows, k_cols, 15.f, qweight, qzeros, scales_t, dq); break;
    case 16: quant_pack_awq_kernel<T, 16><<<grid, 256, 0, st>>>(w, n_rows, k_cols, 15.f, qweight, qzeros, scales_t, dq); break;
    default: quant_pack_awq_kernel<T, 32><<<grid, 256, 0, st>>>(w, n_rows, k_cols, 15.f, qweight, qzeros, scales_t, dq); break;
  }
}
}  // namespace

extern "C" int qdm_quant_pack_awq(const void* w, int dtype, int64_t n_rows, int64_t k_cols, int group,
                                  int32_t* qweight, int32_t* qzeros, void* scales_t, void* dq, void* stream) {
  QDM_REQUIRE(w && qweight && qzeros && scales_t, "qdm_quant_pack_awq: null pointer");
  QDM_REQUIRE(n_rows > 0 && k_cols > 0 && group > 0 && k_cols % group == 0,
              "qdm_quant_pack_awq: group %d must divide k_cols %lld", group, (long long)k_cols);
  QDM_REQUIRE(n_rows % 8 == 0, "qdm_quant_pack_awq: n_rows %lld must be a multiple of 8", (long long)n_rows);
  QDM_REQUIRE(qdm_aligned16(w) && (!dq || qdm_aligned16(dq)), "qdm_quant_pack_awq: tensors must be 16-byte aligned");
  QDM_DEVICE_GATE();
  cudaStream_t st = (cudaStream_t)stream;
  QDM_DISPATCH_DTYPE(dtype, {
    constexpr int V = ElemTraits<T>::kVec;
    const int lpg = group / V;
    QDM_UNSUPPORTED(group % V == 0 && group >= 32 && lpg >= 4 && lpg <= 32 && (lpg & (lpg - 1)) == 0 && group <= 256,
                    "qdm_quant_pack_awq: group %d unsupported for this dtype", group);
    dim3 grid((unsigned)(k_cols / group), (unsigned)((n_rows + kFuseTileN - 1) / kFuseTileN));
    launch_pack<T>(lpg, grid, (const T*)w, n_rows, k_cols, qweight, qzeros, (T*)scales_t, (T*)dq, st);
    QDM_LAUNCH_CHECK();
  });
  return QDM_OK;
}

extern "C" int qdm_dequant_awq(const int32_t* qweight, const int32_t* qzeros, const void* scales_t, int dtype,
                               int64_t k_rows, int64_t n_cols, int group, void* out_kn, void* stream) {
  QDM_REQUIRE(qweight && qzeros && scales_t && out_kn, "qdm_dequant_awq: null pointer");
  QDM_REQUIRE(k_rows > 0 && n_cols > 0 && n_cols % 8 == 0 && group > 0 && k_rows % group == 0,
              "qdm_dequant_awq: bad shape K=%lld N=%lld group=%d", (long long)k_rows, (long long)n_cols, group);
  QDM_REQUIRE(qdm_aligned16(scales_t) && qdm_aligned16(out_kn), "qdm_dequant_awq: scales / out must be 16-byte aligned");
  const int64_t total = k_rows * (n_cols / 8);
  QDM_UNSUPPORTED(total < (int64_t(1) << 32) - 256 * int64_t(QDM_NUM_SMS) * 16, "qdm_dequant_awq: more than 2^32 packed words");
  QDM_DEVICE_GATE();
  cudaStream_t st = (cudaStream_t)stream;
  QDM_DISPATCH_DTYPE(dtype, {
    dequant_awq_kernel<T><<<grid_for(total, 256), 256, 0, st>>>(qweight, qzeros, (const T*)scales_t, (uint32_t)total,
                                                               (uint32_t)(n_cols / 8), (uint32_t)group, (T*)out_kn);
    QDM_LAUNCH_CHECK();
  });
  return QDM_OK;
}

// ---------------------------------------------------------------- self test of the reciprocal division
namespace {
template <typename T> struct BitsOf;
template <> struct BitsOf<__half> {
  static __device__ float f(uint16_t b) { return __half2float(__ushort_as_half(b)); }
  static constexpr int kMaxFinite = 0x7BFF;
};
template <> struct BitsOf<__nv_bfloat16> {
  static __device__ float f(uint16_t b) { return __uint_as_float(uint32_t(b) << 16); }
  static constexpr int kMaxFinite = 0x7F7F;
};
// block = one positive finite divisor (by bit pattern), threads sweep all 65536 dividend patterns
template <typename T>
__global__ void fastdiv_selftest_kernel(unsigned long long* out) {
  const float s = BitsOf<T>::f(uint16_t(blockIdx.x + 1));
  const float r = rcp_approx(s);
  unsigned long long tested = 0, bad = 0;
  for (int wb = threadIdx.x; wb < 65536; wb += blockDim.x) {
    const float w = BitsOf<T>::f(uint16_t(wb));
    if (!(fabsf(w) <= 3.4e38f)) continue;                      // inf / nan patterns
    if (!fastdiv_ok<T>(s, fabsf(w))) continue;                 // outside the window the kernels use __fdiv_rn
    ++tested;
    const float ref = rnd<T>(__fdiv_rn(w, s)), got = rnd<T>(div_by_rcp<true>(w, s, r));
    if (__float_as_uint(ref) != __float_as_uint(got)) ++bad;
  }
  atomicAdd(&out[0], tested);
  if (bad) atomicAdd(&out[1], bad);
}
}  // namespace

extern "C" int qdm_selftest_fastdiv(int dtype, uint64_t* out_host) {
  QDM_REQUIRE(out_host, "qdm_selftest_fastdiv: null pointer");
  QDM_REQUIRE(dtype == QDM_F16 || dtype == QDM_BF16, "qdm_selftest_fastdiv: dtype must be f16 or bf16");
  QDM_DEVICE_GATE();
  unsigned long long* d = nullptr;
  QDM_CUDA_OK(cudaMalloc(&d, 2 * sizeof(unsigned long long)));
  cudaMemset(d, 0, 2 * sizeof(unsigned long long));
  if (dtype == QDM_F16) fastdiv_selftest_kernel<__half><<<BitsOf<__half>::kMaxFinite, 256>>>(d);
  else fastdiv_selftest_kernel<__nv_bfloat16><<<BitsOf<__nv_bfloat16>::kMaxFinite, 256>>>(d);
  cudaError_t e = cudaMemcpy(out_host, d, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) { qdm_set_error("qdm_selftest_fastdiv: %s", cudaGetErrorString(e)); return QDM_ERR_CUDA; }
  qdm_count_launch();
  return QDM_OK;
}

extern "C" size_t qdm_awq_clip_workspace_bytes(int64_t ci, int group) {
  if (ci <= 0 || group <= 0 || ci % group) return 0;
  return size_t(ci / group) * size_t(group) * size_t(group) * sizeof(float);
}

extern "C" int qdm_awq_clip_search(const void* w, int dtype, int64_t co, int64_t ci, int group, int n_bits, unsigned flags,
                                   const void* x, int64_t n_tok, int64_t ld_x, int n_grid, float max_shrink,
                                   void* best_max, void* workspace, size_t workspace_bytes, void* stream) {
  QDM_REQUIRE(w && x && best_max && workspace, "qdm_awq_clip_search: null pointer");
  QDM_REQUIRE(co > 0 && ci > 0 && n_tok > 0 && ld_x >= ci, "qdm_awq_clip_search: empty problem");
  QDM_REQUIRE(dtype == QDM_F16 || dtype == QDM_BF16, "qdm_awq_clip_search: dtype must be f16 or bf16");
  QDM_REQUIRE((group == 64 || group == 128) && ci % group == 0, "qdm_awq_clip_search: group %d must be 64 or 128 and divide ci=%lld",
              group, (long long)ci);
  QDM_REQUIRE(n_bits >= 2 && n_bits <= 8, "qdm_awq_clip_search: n_bits %d outside [2, 8]", n_bits);
  QDM_REQUIRE((flags & ~QDM_Q_ZERO_POINT) == 0, "qdm_awq_clip_search: bad flags 0x%x", flags);
  const int n_levels = int(max_shrink * float(n_grid));
  QDM_REQUIRE(n_grid > 0 && n_levels >= 1 && n_levels <= kClipLevels, "qdm_awq_clip_search: %d shrink levels (1..%d supported)", n_levels, kClipLevels);
  QDM_REQUIRE(workspace_bytes >= qdm_awq_clip_workspace_bytes(ci, group) && qdm_aligned16(workspace), "qdm_awq_clip_search: workspace too small");
  QDM_DEVICE_GATE();
  cudaStream_t st = (cudaStream_t)stream;
  const bool zp = (flags & QDM_Q_ZERO_POINT) != 0;
  const float max_int = zp ? float((1 << n_bits) - 1) : float((1 << (n_bits - 1)) - 1);
  const float min_int = zp ? 0.f : -float(1 << (n_bits - 1));
  const int G = int(ci / group);
  float* gram = static_cast<float*>(workspace);
  // rows per CTA: enough CTAs to fill the device (2 per SM), whole multiples of the warp count
  int64_t rows_per_cta = (co * G + 2 * QDM_NUM_SMS - 1) / (2 * QDM_NUM_SMS);
  rows_per_cta = (rows_per_cta + kClipWarps - 1) / kClipWarps * kClipWarps;
  if (rows_per_cta < kClipWarps) rows_per_cta = kClipWarps;
  if (rows_per_cta > co) rows_per_cta = (co + kClipWarps - 1) / kClipWarps * kClipWarps;
  const dim3 grid(unsigned(G), unsigned((co + rows_per_cta - 1) / rows_per_cta));
#define QDM_CLIP(TT, GSZ, ZPB)                                                                                         \
  do {                                                                                                                 \
    group_gram_kernel<TT, GSZ><<<dim3(unsigned(G), GSZ / 16), 256, 0, st>>>((const TT*)x, n_tok, ld_x, gram);          \
    QDM_LAUNCH_CHECK();                                                                                                \
    const size_t smem = size_t(GSZ) * GSZ * 4 + size_t(kClipWarps) * GSZ * kClipDPad * 4;                              \
    static bool attr_set = false;                                                                                      \
    if (!attr_set) {                                                                                                   \
      QDM_CUDA_OK(cudaFuncSetAttribute(awq_clip_kernel<TT, GSZ, ZPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      attr_set = true;                                                                                                 \
    }                                                                                                                  \
    awq_clip_kernel<TT, GSZ, ZPB><<<grid, 32 * kClipWarps, smem, st>>>((const TT*)w, co, ci, gram, max_int, min_int, n_grid, \
                                                                      n_levels, int(rows_per_cta), (TT*)best_max);    \
    QDM_LAUNCH_CHECK();                                                                                                \
  } while (0)
  if (dtype == QDM_F16) {
    if (group == 128) { if (zp) QDM_CLIP(__half, 128, true); else QDM_CLIP(__half, 128, false); }
    else { if (zp) QDM_CLIP(__half, 64, true); else QDM_CLIP(__half, 64, false); }
  } else {
    if (group == 128) { if (zp) QDM_CLIP(__nv_bfloat16, 128, true); else QDM_CLIP(__nv_bfloat16, 128, false); }
    else { if (zp) QDM_CLIP(__nv_bfloat16, 64, true); else QDM_CLIP(__nv_bfloat16, 64, false); }
  }
#undef QDM_CLIP
  return QDM_OK;
}
