// GEGLU between the two packed Linears of a diffusers FeedForward (ff.net.0.proj -> ff.net.2):
//   y[m, f] = x[m, f] * gelu(x[m, F + f])          x = ff.net.0.proj output [M, 2F], y = ff.net.2 input [M, F]
// HBM-bound elementwise work (4 B read + 2 B written per output element), one 16-byte vector of each half per
// thread-step, read-once loads that do not allocate in L1.  Arithmetic replays the two torch ops of
// `h * F.gelu(gate)` (erf form, ATen ActivationGeluKernel.cu: x * 0.5 * (1 + erf(x * M_SQRT1_2)) in fp32, rounded to the
// tensor dtype; then the product in fp32, rounded once more), so that the packed path and the plain torch skeleton differ
// by at most the last bit of erff.
#include "qdm_common.cuh"

namespace {

constexpr int GEGLU_THREADS = 256;
constexpr int GEGLU_UNROLL = 2;   // two row-vectors per thread in flight (4 x 16-byte loads)

template <typename T>
__global__ void __launch_bounds__(GEGLU_THREADS)
geglu_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t M, int64_t F) {
  constexpr int V = ElemTraits<T>::kVec;
  const int64_t vec_per_row = F / V, total = M * vec_per_row;
  const int64_t stride = int64_t(gridDim.x) * GEGLU_THREADS;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");   // x is the previous kernel's output
  for (int64_t i0 = int64_t(blockIdx.x) * GEGLU_THREADS + threadIdx.x; i0 < total; i0 += GEGLU_UNROLL * stride) {
    Vec16<T> h[GEGLU_UNROLL], g[GEGLU_UNROLL];
    int64_t off_y[GEGLU_UNROLL];
#pragma unroll
    for (int u = 0; u < GEGLU_UNROLL; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < total) {
        const int64_t row = i / vec_per_row, c = (i - row * vec_per_row) * V;
        const T* px = x + row * (2 * F) + c;
        h[u] = ld_vec16_stream(px);
        g[u] = ld_vec16_stream(px + F);
        off_y[u] = row * F + c;
      }
    }
#pragma unroll
    for (int u = 0; u < GEGLU_UNROLL; ++u) {
      if (i0 + u * stride < total) {
        Vec16<T> o;
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const float gv = ElemTraits<T>::to_f(g[u].v[e]);
          const float ge = rnd<T>(__fmul_rn(__fmul_rn(gv, 0.5f), __fadd_rn(1.0f, erff(__fmul_rn(gv, 0.70710678118654752440f)))));
          o.v[e] = ElemTraits<T>::from_f(__fmul_rn(ElemTraits<T>::to_f(h[u].v[e]), ge));
        }
        st_vec16(y + off_y[u], o);
      }
    }
  }
}

template <typename T>
int geglu_launch(const void* x, void* y, int64_t M, int64_t F, cudaStream_t st) {
  const int64_t total = M * (F / ElemTraits<T>::kVec);
  const int64_t want = (total + GEGLU_THREADS * GEGLU_UNROLL - 1) / (GEGLU_THREADS * GEGLU_UNROLL);
  const int64_t cap = int64_t(QDM_NUM_SMS) * 8;   // 8 resident CTAs per SM, grid-stride beyond
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(want < cap ? (want < 1 ? 1 : want) : cap));
  cfg.blockDim = dim3(GEGLU_THREADS);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  QDM_CUDA_OK(cudaLaunchKernelEx(&cfg, geglu_kernel<T>, (const T*)x, (T*)y, M, F));
  QDM_LAUNCH_CHECK();
  return QDM_OK;
}

}  // namespace

extern "C" int qdm_geglu(const void* x, int dtype, int64_t M, int64_t F, void* y, void* stream) {
  QDM_REQUIRE(x && y, "qdm_geglu: null pointer");
  QDM_REQUIRE(dtype == QDM_F16 || dtype == QDM_BF16, "qdm_geglu: dtype must be f16 or bf16");
  QDM_REQUIRE(M >= 0 && F > 0 && F % 8 == 0, "qdm_geglu: M=%lld, F=%lld (F must be a positive multiple of 8)", (long long)M, (long long)F);
  QDM_REQUIRE(qdm_aligned16(x) && qdm_aligned16(y), "qdm_geglu: x / y must be 16-byte aligned");
  QDM_DEVICE_GATE();
  if (M == 0) return QDM_OK;
  return dtype == QDM_BF16 ? geglu_launch<__nv_bfloat16>(x, y, M, F, (cudaStream_t)stream)
                           : geglu_launch<__half>(x, y, M, F, (cudaStream_t)stream);
}
