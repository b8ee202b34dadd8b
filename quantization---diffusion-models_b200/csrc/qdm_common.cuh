// Shared helpers for libqdm.so kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
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
