// Shared helpers for libqdm.so kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <type_traits>
#include "../../include/qdm.h"

// ---------------------------------------------------------------- error plumbing
void qdm_set_error(const char* fmt, ...);
void qdm_count_launch(int n = 1);
int qdm_require_device();  // QDM_OK only on a cc-10.0 device (cached per device)

#define QDM_DEVICE_GATE()                              \
  do {                                                 \
    int _g = qdm_require_device();                     \
    if (_g != QDM_OK) return _g;                       \
  } while (0)

#define QDM_REQUIRE(cond, ...)                         \
  do {                                                 \
    if (!(cond)) {                                     \
      qdm_set_error(__VA_ARGS__);                      \
      return QDM_ERR_INVALID;                          \
    }                                                  \
  } while (0)

#define QDM_UNSUPPORTED(cond, ...)                     \
  do {                                                 \
    if (!(cond)) {                                     \
      qdm_set_error(__VA_ARGS__);                      \
      return QDM_ERR_UNSUPPORTED;                      \
    }                                                  \
  } while (0)

#define QDM_CUDA_OK(expr)                                                        \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) {                                                     \
      qdm_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                    __FILE__, __LINE__);                                         \
      return QDM_ERR_CUDA;                                                       \
    }                                                                            \
  } while (0)

#define QDM_LAUNCH_CHECK()                                                       \
  do {                                                                           \
    cudaError_t _e = cudaGetLastError();                                         \
    if (_e != cudaSuccess) {                                                     \
      qdm_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),  \
                    __FILE__, __LINE__);                                         \
      return QDM_ERR_CUDA;                                                       \
    }                                                                            \
    qdm_count_launch();                                                          \
  } while (0)

static inline int qdm_dtype_size(int dtype) {
  return dtype == QDM_F32 ? 4 : 2;
}
static inline bool qdm_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int QDM_NUM_SMS = 148;

// ---------------------------------------------------------------- element types
// Every torch op of the reference is replayed as: compute in fp32, round to the tensor
// dtype (RN-even).  `rnd<T>(x)` is that rounding step; for fp32 it is the identity.
template <typename T> struct ElemTraits;
template <> struct ElemTraits<__half> {
  static constexpr int kVec = 8;  // elements per 16-byte vector
  __device__ __forceinline__ static float to_f(__half v) { return __half2float(v); }
  __device__ __forceinline__ static __half from_f(float v) { return __float2half_rn(v); }
};
template <> struct ElemTraits<__nv_bfloat16> {
  static constexpr int kVec = 8;
  __device__ __forceinline__ static float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ __forceinline__ static __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};
template <> struct ElemTraits<float> {
  static constexpr int kVec = 4;
  __device__ __forceinline__ static float to_f(float v) { return v; }
  __device__ __forceinline__ static float from_f(float v) { return v; }
};

template <typename T>
__device__ __forceinline__ float rnd(float v) {
  return ElemTraits<T>::to_f(ElemTraits<T>::from_f(v));
}
template <>
__device__ __forceinline__ float rnd<float>(float v) { return v; }

// 16-byte vector of T
template <typename T>
struct __align__(16) Vec16 {
  T v[ElemTraits<T>::kVec];
};

template <typename T>
__device__ __forceinline__ Vec16<T> ld_vec16(const T* p) {
  Vec16<T> r;
  *reinterpret_cast<uint4*>(&r) = *reinterpret_cast<const uint4*>(p);
  return r;
}
// streaming (read-once) load: bypass L1 allocation
template <typename T>
__device__ __forceinline__ Vec16<T> ld_vec16_stream(const T* p) {
  Vec16<T> r;
  uint4 u;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
               : "l"(p));
  *reinterpret_cast<uint4*>(&r) = u;
  return r;
}
template <typename T>
__device__ __forceinline__ void st_vec16(T* p, const Vec16<T>& r) {
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&r);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------- exact division without the IEEE slow path
// torch divides fp16/bf16 tensors as rnd<T>(fp32 w / fp32 s) with a correctly rounded fp32 quotient.
// `__fdiv_rn` per element costs a MUFU, ~8 FFMA, an FCHK and a divergent slow-path call, which makes the
// quantise kernels issue-bound instead of HBM-bound.  With one reciprocal estimate r ~ 1/s per divisor,
//   q0 = w * r;  e = fma(-q0, s, w) (exact remainder);  q = fma(e, r, q0)
// gives rnd<T>(q) == rnd<T>(__fdiv_rn(w, s)) for EVERY pair of fp16 values, and for every pair of bf16
// values with s in [2^-60, 2^60] and |w| <= 2^64 -- proven by exhaustion over all 2^31 pairs on the device
// (qdm_selftest_fastdiv, run by tests/test_gpu_quant.py).  Outside that window (bf16 only: quotient
// overflow, subnormal divisors) and for fp32 tensors the kernels take `__fdiv_rn`.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// KEEP_ZERO_SIGN: a zero dividend keeps the sign IEEE division gives it (fma(+0, r, -0) would lose it);
// callers that fix the sign themselves or feed an integer add skip the two extra instructions.
template <bool KEEP_ZERO_SIGN>
__device__ __forceinline__ float div_by_rcp(float w, float s, float r) {
  const float q0 = __fmul_rn(w, r);
  const float e = __fmaf_rn(-q0, s, w);
  const float q1 = __fmaf_rn(e, r, q0);
  return (KEEP_ZERO_SIGN && w == 0.f) ? q0 : q1;
}
// may the divisor s (> 0) be used with div_by_rcp for dividends of magnitude <= amax?
template <typename T>
__device__ __forceinline__ bool fastdiv_ok(float s, float amax) {
  if (std::is_same<T, __half>::value) return true;
  if (std::is_same<T, float>::value) return false;
  return s >= 0x1p-60f && s <= 0x1p50f && amax <= 0x1p50f;   // |q * s| <= 2^8 * 2^50 keeps post-division inside too
}
template <typename T, bool FAST, bool KEEP_ZERO_SIGN>
__device__ __forceinline__ float div_T(float w, float s, float r) {
  return FAST ? div_by_rcp<KEEP_ZERO_SIGN>(w, s, r) : __fdiv_rn(w, s);
}
// round-half-even to an integer.  16-bit dtypes: magic-number add (two FADD on the FMA pipe instead of an
// FRND on the quarter-rate XU pipe), exact for |q| < 2^22 and order/sign/inf-preserving above, where every
// caller clamps.  A zero result comes out as +0; callers that need torch.round's -0 copy the dividend's sign.
template <typename T>
__device__ __forceinline__ float rint_T(float q) {
  if (std::is_same<T, float>::value) return rintf(q);
  return __fadd_rn(__fadd_rn(q, 12582912.f), -12582912.f);
}
// low byte of a small integer held in a float (two's complement), without an F2I
__device__ __forceinline__ uint32_t int_byte(float c) {
  return __float_as_uint(__fadd_rn(c, 12582912.f)) & 0xffu;
}

// The RTN chain shared by every quantiser of the reference.  One call = the torch ops
//   q = round(w / s) [+ z, clamp(lo,hi), - z] ; dq = q * s
// with a dtype rounding after each op.  Returns the integer code (before "- z") in `code`.
template <typename T>
__device__ __forceinline__ float rtn_elem(float w, float s, float z, float lo, float hi,
                                          bool use_zero, bool do_clamp, float& code) {
  float q = rnd<T>(__fdiv_rn(w, s));
  q = rintf(q);  // round-half-even, exact in every dtype
  if (use_zero) q = rnd<T>(__fadd_rn(q, z));
  if (do_clamp) q = fminf(fmaxf(q, lo), hi);
  code = q;
  if (use_zero) q = rnd<T>(__fsub_rn(q, z));
  return rnd<T>(__fmul_rn(q, s));
}
