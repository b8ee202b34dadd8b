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

// ---------------------------------------------------------------- power-of-two groups, vector path
// The flattened tensor is a sequence of contiguous groups; a group is `lpg` adjacent lanes
// (one 16-byte vector per lane), so group min/max is an xor-shuffle butterfly.
template <typename T, int MODE>
__global__ void __launch_bounds__(kQThreads)
quant_group_kernel(const T* __restrict__ w, int64_t n_vec, int64_t k_cols, int lpg,
                   float max_int, float min_int,
                   const T* __restrict__ pre_mul, const T* __restrict__ clip_max,
                   const T* __restrict__ post_div,
                   T* __restrict__ dq, int8_t* __restrict__ codes,
                   T* __restrict__ scales, T* __restrict__ zeros) {
  constexpr int V = ElemTraits<T>::kVec;
  const int lane = threadIdx.x & 31;
  const int64_t stride = int64_t(gridDim.x) * kQThreads;
  for (int64_t base = int64_t(blockIdx.x) * kQThreads; base < n_vec; base += stride) {
    const int64_t i = base + threadIdx.x;
    const bool active = i < n_vec;
    float x[V];
    int64_t col = 0;
    if (active) {
      Vec16<T> v = ld_vec16_stream(w + i * V);
#pragma unroll
      for (int j = 0; j < V; ++j) x[j] = ElemTraits<T>::to_f(v.v[j]);
      col = (i * V) % k_cols;
      if (pre_mul) {
        Vec16<T> pm = ld_vec16(pre_mul + col);
#pragma unroll
        for (int j = 0; j < V; ++j) x[j] = rnd<T>(__fmul_rn(x[j], ElemTraits<T>::to_f(pm.v[j])));
      }
      if (clip_max) {
        const float c = ElemTraits<T>::to_f(clip_max[i / lpg]);
#pragma unroll
        for (int j = 0; j < V; ++j) x[j] = fminf(fmaxf(x[j], -c), c);
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
    if (!active) continue;
    float s, z;
    group_params<T, MODE>(mx, mn, max_int, s, z);
    if ((lane & (lpg - 1)) == 0) {
      const int64_t g = i / lpg;
      if (scales) scales[g] = ElemTraits<T>::from_f(s);
      if (zeros && MODE == Q_ZP) zeros[g] = ElemTraits<T>::from_f(z);
    }
    Vec16<T> o;
    int8_t cb[V];
    Vec16<T> pd;
    if (post_div) pd = ld_vec16(post_div + col);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float code;
      float d = rtn_elem<T>(x[j], s, z, min_int, max_int, MODE == Q_ZP, MODE != Q_SYM_NOCLAMP, code);
      if (post_div) d = rnd<T>(__fdiv_rn(d, ElemTraits<T>::to_f(pd.v[j])));
      o.v[j] = ElemTraits<T>::from_f(d);
      cb[j] = code_to_i8<MODE>(code);
    }
    if (dq) st_vec16(dq + i * V, o);
    if (codes) {
      if (V == 8) {
        *reinterpret_cast<uint2*>(codes + i * V) = *reinterpret_cast<const uint2*>(cb);
      } else {
        *reinterpret_cast<uint32_t*>(codes + i * V) = *reinterpret_cast<const uint32_t*>(cb);
      }
    }
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
      for (int64_t e = int64_t(lane) * V; e < cols; e += 32 * V) {
        Vec16<T> v = ld_vec16(p + e);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          float f = prep(ElemTraits<T>::to_f(v.v[j]), e + j);
          if (MODE == Q_ZP) { mx = fmaxf(mx, f); mn = fminf(mn, f); } else mx = fmaxf(mx, fabsf(f));
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
      for (int64_t e = int64_t(lane) * V; e < cols; e += 32 * V) {
        Vec16<T> v = ld_vec16(p + e);
        Vec16<T> o;
        int8_t cb[V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
          float code;
          o.v[j] = ElemTraits<T>::from_f(finish(prep(ElemTraits<T>::to_f(v.v[j]), e + j), e + j, code));
          cb[j] = code_to_i8<MODE>(code);
        }
        if (dq) st_vec16(dq + row * cols + e, o);
        if (codes) {
#pragma unroll
          for (int j = 0; j < V; ++j) codes[row * cols + e + j] = cb[j];
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
#pragma unroll
    for (int e = 0; e < kMaxCols; ++e) {
      if (e < cols) {
        float code;
        float d = rtn_elem<T>(v[e], s, z, min_int, max_int, MODE == Q_ZP, MODE != Q_SYM_NOCLAMP, code);
        if (dq) dq[row * cols + e] = ElemTraits<T>::from_f(d);
        if (codes) codes[row * cols + e] = code_to_i8<MODE>(code);
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
  for (int64_t i = tid; i < nvec; i += nthreads) {
    Vec16<T> v = ld_vec16_stream(x + i * V);
    Vec16<T> o;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float code;
      o.v[j] = ElemTraits<T>::from_f(rtn_elem<T>(ElemTraits<T>::to_f(v.v[j]), s, 0.f, 0.f, 0.f, false, false, code));
      if (codes) codes[i * V + j] = code_to_i8<Q_SYM_NOCLAMP>(code);
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
    if (vec_ok) {
      for (int64_t e = int64_t(lane) * V; e < cols; e += 32 * V) {
        Vec16<T> v = ld_vec16(p + e);
#pragma unroll
        for (int j = 0; j < V; ++j) mx = fmaxf(mx, fabsf(prep(ElemTraits<T>::to_f(v.v[j]), e + j)));
      }
    } else {
      for (int64_t e = lane; e < cols; e += 32) mx = fmaxf(mx, fabsf(prep(ElemTraits<T>::to_f(p[e]), e)));
    }
    mx = warp_max(mx);
    float s, z;
    group_params<T, Q_SYM_NOCLAMP>(mx, 0.f, 127.f, s, z);
    if (lane == 0) sx[row] = s;
    if (vec_ok) {
      for (int64_t e = int64_t(lane) * V; e < cols; e += 32 * V) {
        Vec16<T> v = ld_vec16(p + e);
        int8_t cb[V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float q = rintf(rnd<T>(__fdiv_rn(prep(ElemTraits<T>::to_f(v.v[j]), e + j), s)));
          cb[j] = code_to_i8<Q_SYM_NOCLAMP>(q);
        }
        if (V == 8) *reinterpret_cast<uint2*>(xq + row * cols + e) = *reinterpret_cast<const uint2*>(cb);
        else *reinterpret_cast<uint32_t*>(xq + row * cols + e) = *reinterpret_cast<const uint32_t*>(cb);
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

// Fused zero-point RTN + AWQ pack.  Block tile: 64 out-rows x one group of K.
// Phase 1 (lanes-per-group butterflies) leaves codes/scale/zero in shared memory,
// phase 2 assembles the transposed int32 words.
template <typename T>
__global__ void __launch_bounds__(256)
quant_pack_awq_kernel(const T* __restrict__ w, int64_t n_rows, int64_t k_cols, int group, float max_int,
                      int32_t* __restrict__ qweight, int32_t* __restrict__ qzeros,
                      T* __restrict__ scales_t, T* __restrict__ dq) {
  constexpr int V = ElemTraits<T>::kVec;
  extern __shared__ uint8_t smem_raw[];
  const int pitch = group + 4;
  uint8_t* tile = smem_raw;                                  // [64][group + 4] codes
  uint8_t* zsm = smem_raw + kPackTileN * pitch;              // [64] zero points
  const int lpg = group / V;                                 // lanes per row of the tile
  const int rows_per_pass = 256 / lpg;
  const int64_t n0 = int64_t(blockIdx.y) * kPackTileN;
  const int64_t gi = blockIdx.x;                             // group index along K
  const int64_t k0 = gi * group;
  const int sub = threadIdx.x % lpg;
  for (int r = threadIdx.x / lpg; r < kPackTileN; r += rows_per_pass) {
    const int64_t n = n0 + r;                                // n_rows % 64 == 0: always valid
    Vec16<T> v = ld_vec16_stream(w + n * k_cols + k0 + sub * V);
    float x[V];
#pragma unroll
    for (int j = 0; j < V; ++j) x[j] = ElemTraits<T>::to_f(v.v[j]);
    float mx = x[0], mn = x[0];
#pragma unroll
    for (int j = 1; j < V; ++j) { mx = fmaxf(mx, x[j]); mn = fminf(mn, x[j]); }
    for (int o = 1; o < lpg; o <<= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    float s, z;
    group_params<T, Q_ZP>(mx, mn, max_int, s, z);
    if (sub == 0) {
      scales_t[gi * n_rows + n] = ElemTraits<T>::from_f(s);
      zsm[r] = (uint8_t)__float2int_rn(z);
    }
    Vec16<T> o;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float code;
      o.v[j] = ElemTraits<T>::from_f(rtn_elem<T>(x[j], s, z, 0.f, max_int, true, true, code));
      tile[r * pitch + sub * V + j] = (uint8_t)__float2int_rn(code);
    }
    if (dq) st_vec16(dq + n * k_cols + k0 + sub * V, o);
  }
  __syncthreads();
  const int64_t words_per_row = n_rows / 8;
  for (int idx = threadIdx.x; idx < group * 8; idx += 256) {
    const int k = idx >> 3, c = idx & 7;
    uint32_t word = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) word |= uint32_t(tile[(8 * c + kAwqOrder[i]) * pitch + k] & 0xF) << (4 * i);
    qweight[(k0 + k) * words_per_row + n0 / 8 + c] = (int32_t)word;
  }
  if (threadIdx.x < 8) {
    const int c = threadIdx.x;
    uint32_t word = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) word |= uint32_t(zsm[8 * c + kAwqOrder[i]] & 0xF) << (4 * i);
    qzeros[gi * words_per_row + n0 / 8 + c] = (int32_t)word;
  }
}

// W_kn[k, 8c + j] = (q - z) * s, one packed word (8 outputs, one 16-byte store for 2-byte T) per thread
template <typename T>
__global__ void __launch_bounds__(256)
dequant_awq_kernel(const int32_t* __restrict__ qweight, const int32_t* __restrict__ qzeros,
                   const T* __restrict__ scales, int64_t k_rows, int64_t n_words, int group,
                   T* __restrict__ out) {
  const int64_t total = k_rows * n_words;
  for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < total; i += int64_t(gridDim.x) * 256) {
    const int64_t k = i / n_words, c = i % n_words;
    const uint32_t qw = (uint32_t)qweight[i];
    const uint32_t zw = (uint32_t)qzeros[(k / group) * n_words + c];
    const T* sp = scales + (k / group) * n_words * 8 + c * 8;
    T o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int q = (qw >> (4 * j)) & 0xF, z = (zw >> (4 * j)) & 0xF;
      const int col = kAwqOrder[j];
      o[col] = ElemTraits<T>::from_f(__fmul_rn(float(q - z), ElemTraits<T>::to_f(sp[col])));
    }
    T* dst = out + k * n_words * 8 + c * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = o[j];
  }
}

int grid_for(int64_t work_items, int per_block) {
  int64_t b = (work_items + per_block - 1) / per_block;
  const int64_t cap = int64_t(QDM_NUM_SMS) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return int(b);
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
  if (aligned && group % V == 0 && lpg <= 32 && (lpg & (lpg - 1)) == 0 && k_period % V == 0) {
    const int64_t n_vec = numel / V;
    quant_group_kernel<T, MODE><<<grid_for(n_vec, kQThreads), kQThreads, 0, st>>>(
        w, n_vec, k_period, int(lpg), max_int, min_int, pre_mul, clip_max, post_div, dq, codes, scales, zeros);
  } else if (group <= 16 && !pre_mul && !clip_max && !post_div) {
    quant_rows_thread_kernel<T, MODE><<<grid_for(n_groups, kQThreads), kQThreads, 0, st>>>(
        w, n_groups, int(group), max_int, min_int, dq, codes, scales, zeros);
  } else {
    const int vec_ok = aligned && group % V == 0;
    quant_rows_warp_kernel<T, MODE><<<grid_for(n_groups, kQThreads / 32), kQThreads, 0, st>>>(
        w, n_groups, group, k_period, max_int, min_int, pre_mul, clip_max, post_div, dq, codes, scales, zeros, vec_ok);
  }
  QDM_LAUNCH_CHECK();
  return QDM_OK;
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
  QDM_REQUIRE(n_bits >= 2 && n_bits <= 8, "qdm_quant_rowwise: n_bits %d outside [2, 8]", n_bits);
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
  QDM_REQUIRE(n_bits >= 2 && n_bits <= 8, "qdm_quant_tensor: n_bits %d outside [2, 8]", n_bits);
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

extern "C" int qdm_quant_pack_awq(const void* w, int dtype, int64_t n_rows, int64_t k_cols, int group,
                                  int32_t* qweight, int32_t* qzeros, void* scales_t, void* dq, void* stream) {
  QDM_REQUIRE(w && qweight && qzeros && scales_t, "qdm_quant_pack_awq: null pointer");
  QDM_REQUIRE(n_rows > 0 && k_cols > 0 && group > 0 && k_cols % group == 0,
              "qdm_quant_pack_awq: group %d must divide k_cols %lld", group, (long long)k_cols);
  QDM_REQUIRE(n_rows % kPackTileN == 0, "qdm_quant_pack_awq: n_rows %lld must be a multiple of 64", (long long)n_rows);
  QDM_REQUIRE(qdm_aligned16(w) && (!dq || qdm_aligned16(dq)), "qdm_quant_pack_awq: tensors must be 16-byte aligned");
  QDM_DEVICE_GATE();
  cudaStream_t st = (cudaStream_t)stream;
  QDM_DISPATCH_DTYPE(dtype, {
    constexpr int V = ElemTraits<T>::kVec;
    const int lpg = group / V;
    QDM_UNSUPPORTED(group % V == 0 && lpg >= 1 && lpg <= 32 && (lpg & (lpg - 1)) == 0 && group <= 256,
                    "qdm_quant_pack_awq: group %d unsupported for this dtype", group);
    dim3 grid((unsigned)(k_cols / group), (unsigned)(n_rows / kPackTileN));
    const size_t smem = size_t(kPackTileN) * (group + 4) + kPackTileN;
    quant_pack_awq_kernel<T><<<grid, 256, smem, st>>>((const T*)w, n_rows, k_cols, group, 15.f, qweight, qzeros,
                                                      (T*)scales_t, (T*)dq);
    QDM_LAUNCH_CHECK();
  });
  return QDM_OK;
}

extern "C" int qdm_dequant_awq(const int32_t* qweight, const int32_t* qzeros, const void* scales_t, int dtype,
                               int64_t k_rows, int64_t n_cols, int group, void* out_kn, void* stream) {
  QDM_REQUIRE(qweight && qzeros && scales_t && out_kn, "qdm_dequant_awq: null pointer");
  QDM_REQUIRE(k_rows > 0 && n_cols > 0 && n_cols % 8 == 0 && group > 0 && k_rows % group == 0,
              "qdm_dequant_awq: bad shape K=%lld N=%lld group=%d", (long long)k_rows, (long long)n_cols, group);
  QDM_DEVICE_GATE();
  cudaStream_t st = (cudaStream_t)stream;
  QDM_DISPATCH_DTYPE(dtype, {
    const int64_t total = k_rows * (n_cols / 8);
    dequant_awq_kernel<T><<<grid_for(total, 256), 256, 0, st>>>(qweight, qzeros, (const T*)scales_t, k_rows,
                                                               n_cols / 8, group, (T*)out_kn);
    QDM_LAUNCH_CHECK();
  });
  return QDM_OK;
}
