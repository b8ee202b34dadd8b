// (a) Calibration-statistic reductions: column |x| max / sum, row |x| max, tensor |x| max,
// AWQ normalised-weight column sum, squared-difference sum.  All HBM-bound; 16-byte
// coalesced loads, warp shuffles + shared-memory staging, two-stage fixed-tree reductions
// (no float atomics) so every result is deterministic run to run.
#include "qdm_common.cuh"

namespace {

constexpr int kColThreads = 256;   // 8 warps: warp = row lane, lane = 16-byte column vector
constexpr int kColWarps = kColThreads / 32;

enum ColOp { COL_ABSMAX = 0, COL_ABSSUM = 1 };

// Stage 1 of a column reduction over x[rows, cols] (row stride ld).  Block (bx, by) owns
// columns [bx*32*V, (bx+1)*32*V) and rows by, by+gridDim.y, ... in units of 8-row slabs.
// Writes partial[by][col] (fp32).
template <typename T, int OP>
__global__ void __launch_bounds__(kColThreads)
col_reduce_stage1(const T* __restrict__ x, int64_t rows, int64_t cols, int64_t ld,
                  float* __restrict__ partial) {
  constexpr int V = ElemTraits<T>::kVec;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t c0 = (int64_t(blockIdx.x) * 32 + lane) * V;
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  if (c0 < cols) {
    const int64_t row_step = int64_t(gridDim.y) * kColWarps;
    int64_t r = int64_t(blockIdx.y) * kColWarps + warp;
    // two loads in flight per thread
    for (; r + row_step < rows; r += 2 * row_step) {
      Vec16<T> a = ld_vec16_stream(x + r * ld + c0);
      Vec16<T> b = ld_vec16_stream(x + (r + row_step) * ld + c0);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float fa = fabsf(ElemTraits<T>::to_f(a.v[i])), fb = fabsf(ElemTraits<T>::to_f(b.v[i]));
        if (OP == COL_ABSMAX) acc[i] = fmaxf(acc[i], fmaxf(fa, fb));
        else acc[i] += fa + fb;
      }
    }
    if (r < rows) {
      Vec16<T> a = ld_vec16_stream(x + r * ld + c0);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float fa = fabsf(ElemTraits<T>::to_f(a.v[i]));
        if (OP == COL_ABSMAX) acc[i] = fmaxf(acc[i], fa);
        else acc[i] += fa;
      }
    }
  }
  __shared__ float sm[kColWarps][32 * V + 1];
#pragma unroll
  for (int i = 0; i < V; ++i) sm[warp][lane * V + i] = acc[i];
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * V; c += kColThreads) {
    float v = sm[0][c];
#pragma unroll
    for (int w = 1; w < kColWarps; ++w) v = (OP == COL_ABSMAX) ? fmaxf(v, sm[w][c]) : v + sm[w][c];
    const int64_t col = int64_t(blockIdx.x) * 32 * V + c;
    if (col < cols) partial[int64_t(blockIdx.y) * cols + col] = v;
  }
}

// scalar fallback (cols or ld not a multiple of the vector width, or unaligned base)
template <typename T, int OP>
__global__ void __launch_bounds__(kColThreads)
col_reduce_stage1_scalar(const T* __restrict__ x, int64_t rows, int64_t cols, int64_t ld,
                         float* __restrict__ partial) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t c = int64_t(blockIdx.x) * 32 + lane;
  float acc = 0.f;
  if (c < cols) {
    for (int64_t r = int64_t(blockIdx.y) * kColWarps + warp; r < rows; r += int64_t(gridDim.y) * kColWarps) {
      float f = fabsf(ElemTraits<T>::to_f(x[r * ld + c]));
      acc = (OP == COL_ABSMAX) ? fmaxf(acc, f) : acc + f;
    }
  }
  __shared__ float sm[kColWarps][33];
  sm[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && c < cols) {
    float v = sm[0][lane];
#pragma unroll
    for (int w = 1; w < kColWarps; ++w) v = (OP == COL_ABSMAX) ? fmaxf(v, sm[w][lane]) : v + sm[w][lane];
    partial[int64_t(blockIdx.y) * cols + c] = v;
  }
}

// Stage 2: fold `splits` partial rows in fixed order.  mode 1 = running max against out.
template <typename TOut, int OP>
__global__ void col_reduce_stage2(const float* __restrict__ partial, int splits, int64_t cols,
                                  TOut* __restrict__ out, int mode) {
  const int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float v = partial[c];
  for (int s = 1; s < splits; ++s) {
    float p = partial[int64_t(s) * cols + c];
    v = (OP == COL_ABSMAX) ? fmaxf(v, p) : v + p;
  }
  if (mode == 1) v = fmaxf(v, ElemTraits<TOut>::to_f(out[c]));
  out[c] = ElemTraits<TOut>::from_f(v);  // exact for absmax: v is a value of the input dtype
}

int col_splits(int64_t rows, int64_t col_blocks, int ctas_per_sm = 4) {
  // ~ctas_per_sm CTAs per SM in flight, at least 8 rows per CTA, at most 256 splits
  int64_t want = (int64_t(QDM_NUM_SMS) * ctas_per_sm) / col_blocks;
  if (want < 1) want = 1;
  int64_t max_by_rows = (rows + kColWarps - 1) / kColWarps;
  int64_t s = want < max_by_rows ? want : max_by_rows;
  if (s < 1) s = 1;
  if (s > 256) s = 256;
  return int(s);
}

template <typename T, int OP, typename TOut>
int launch_col_reduce(const T* x, int64_t rows, int64_t cols, int64_t ld, TOut* out, int mode,
                      float* ws, size_t ws_bytes, cudaStream_t st) {
  constexpr int V = ElemTraits<T>::kVec;
  const bool vec_ok = (cols % V == 0) && (ld % V == 0) && qdm_aligned16(x);
  const int64_t col_blocks = vec_ok ? (cols + 32 * V - 1) / (32 * V) : (cols + 31) / 32;
  const int splits = col_splits(rows, col_blocks, OP == COL_ABSMAX ? 8 : 4);
  QDM_REQUIRE(ws_bytes >= size_t(splits) * cols * sizeof(float),
              "column reduction workspace too small: %zu < %zu", ws_bytes,
              size_t(splits) * cols * sizeof(float));
  dim3 grid((unsigned)col_blocks, (unsigned)splits);
  if (vec_ok)
    col_reduce_stage1<T, OP><<<grid, kColThreads, 0, st>>>(x, rows, cols, ld, ws);
  else
    col_reduce_stage1_scalar<T, OP><<<grid, kColThreads, 0, st>>>(x, rows, cols, ld, ws);
  QDM_LAUNCH_CHECK();
  col_reduce_stage2<TOut, OP><<<(unsigned)((cols + 255) / 256), 256, 0, st>>>(ws, splits, cols, out, mode);
  QDM_LAUNCH_CHECK();
  return QDM_OK;
}

// ---------------------------------------------------------------------------------------------
// One-pass calibration hook statistic (utils/calib_data.py:105-124):
// column |x| max AND column |x| sum from a single read of x, folded in place into the caller's
// running accumulators by the second stage -- nothing per call is retained on the host side.
// Stage 1 is col_reduce_stage1 with both accumulators live (VEC = false: scalar columns).
template <typename T, bool VEC>
__global__ void __launch_bounds__(kColThreads)
col_stats_stage1(const T* __restrict__ x, int64_t rows, int64_t cols, int64_t ld,
                 float* __restrict__ pmax, float* __restrict__ psum) {
  constexpr int V = VEC ? ElemTraits<T>::kVec : 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t c0 = (int64_t(blockIdx.x) * 32 + lane) * V;
  float amax[V], asum[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { amax[i] = 0.f; asum[i] = 0.f; }
  if (c0 < cols) {
    const int64_t row_step = int64_t(gridDim.y) * kColWarps;
    int64_t r = int64_t(blockIdx.y) * kColWarps + warp;
    if constexpr (VEC) {
      for (; r + row_step < rows; r += 2 * row_step) {   // two loads in flight per thread
        Vec16<T> a = ld_vec16_stream(x + r * ld + c0);
        Vec16<T> b = ld_vec16_stream(x + (r + row_step) * ld + c0);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float fa = fabsf(ElemTraits<T>::to_f(a.v[i])), fb = fabsf(ElemTraits<T>::to_f(b.v[i]));
          amax[i] = fmaxf(amax[i], fmaxf(fa, fb));
          asum[i] += fa + fb;
        }
      }
      if (r < rows) {
        Vec16<T> a = ld_vec16_stream(x + r * ld + c0);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float fa = fabsf(ElemTraits<T>::to_f(a.v[i]));
          amax[i] = fmaxf(amax[i], fa);
          asum[i] += fa;
        }
      }
    } else {
      for (; r < rows; r += row_step) {
        float fa = fabsf(ElemTraits<T>::to_f(x[r * ld + c0]));
        amax[0] = fmaxf(amax[0], fa);
        asum[0] += fa;
      }
    }
  }
  __shared__ float smx[kColWarps][32 * V + 1];
  __shared__ float sms[kColWarps][32 * V + 1];
#pragma unroll
  for (int i = 0; i < V; ++i) { smx[warp][lane * V + i] = amax[i]; sms[warp][lane * V + i] = asum[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * V; c += kColThreads) {
    float m = smx[0][c], s = sms[0][c];
#pragma unroll
    for (int w = 1; w < kColWarps; ++w) { m = fmaxf(m, smx[w][c]); s += sms[w][c]; }
    const int64_t col = int64_t(blockIdx.x) * 32 * V + c;
    if (col < cols) {
      pmax[int64_t(blockIdx.y) * cols + col] = m;
      psum[int64_t(blockIdx.y) * cols + col] = s;
    }
  }
}

// Stage 2: fixed-order fold of the partials, then the in-place updates (one thread per column, launches are
// stream-ordered, so the accumulators are deterministic): out_max = colmax | max(out_max, colmax);
// acc_maxsum += colmax (the numerator of "mean over calls of the per-call max", StableDiffusion1_x.py:104-112);
// acc_abssum += sum_r |x| (the numerator of x_mean, quantizer.py:642-659).  fp64 accumulators: a sum of fp16
// values (multiples of 2^-24 below 2^16) is exact in fp64 up to 2^13 calls x 2^16, hence order-free.
template <typename T>
__global__ void col_stats_stage2(const float* __restrict__ pmax, const float* __restrict__ psum, int splits,
                                 int64_t cols, T* __restrict__ out_max, int max_mode,
                                 double* __restrict__ acc_maxsum, double* __restrict__ acc_abssum) {
  const int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float m = pmax[c], s = psum[c];
  for (int k = 1; k < splits; ++k) {
    m = fmaxf(m, pmax[int64_t(k) * cols + c]);
    s += psum[int64_t(k) * cols + c];
  }
  if (acc_maxsum) acc_maxsum[c] += double(m);
  if (acc_abssum) acc_abssum[c] += double(s);
  if (out_max) {
    if (max_mode == 1) m = fmaxf(m, ElemTraits<T>::to_f(out_max[c]));
    out_max[c] = ElemTraits<T>::from_f(m);
  }
}

template <typename T>
int launch_col_stats(const T* x, int64_t rows, int64_t cols, int64_t ld, T* out_max, int max_mode,
                     double* acc_maxsum, double* acc_abssum, float* ws, size_t ws_bytes, cudaStream_t st) {
  constexpr int V = ElemTraits<T>::kVec;
  const bool vec_ok = (cols % V == 0) && (ld % V == 0) && qdm_aligned16(x);
  const int64_t col_blocks = vec_ok ? (cols + 32 * V - 1) / (32 * V) : (cols + 31) / 32;
  const int splits = col_splits(rows, col_blocks, 4);
  QDM_REQUIRE(ws_bytes >= 2 * size_t(splits) * cols * sizeof(float),
              "qdm_colstats workspace too small: %zu < %zu", ws_bytes, 2 * size_t(splits) * cols * sizeof(float));
  float* pmax = ws;
  float* psum = ws + size_t(splits) * cols;
  dim3 grid((unsigned)col_blocks, (unsigned)splits);
  if (vec_ok)
    col_stats_stage1<T, true><<<grid, kColThreads, 0, st>>>(x, rows, cols, ld, pmax, psum);
  else
    col_stats_stage1<T, false><<<grid, kColThreads, 0, st>>>(x, rows, cols, ld, pmax, psum);
  QDM_LAUNCH_CHECK();
  col_stats_stage2<T><<<(unsigned)((cols + 255) / 256), 256, 0, st>>>(pmax, psum, splits, cols, out_max, max_mode,
                                                                      acc_maxsum, acc_abssum);
  QDM_LAUNCH_CHECK();
  return QDM_OK;
}

// ------------------------------------------------------------------ row |x| max
// warp per row for long rows
template <typename T>
__global__ void __launch_bounds__(256)
row_absmax_warp(const T* __restrict__ x, int64_t rows, int64_t cols, T* __restrict__ out, bool vec_ok) {
  constexpr int V = ElemTraits<T>::kVec;
  const int lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* p = x + row * cols;
  float m = 0.f;
  if (vec_ok) {
    for (int64_t c = int64_t(lane) * V; c < cols; c += 32 * V) {
      Vec16<T> a = ld_vec16_stream(p + c);
#pragma unroll
      for (int i = 0; i < V; ++i) m = fmaxf(m, fabsf(ElemTraits<T>::to_f(a.v[i])));
    }
  } else {
    for (int64_t c = lane; c < cols; c += 32) m = fmaxf(m, fabsf(ElemTraits<T>::to_f(p[c])));
  }
  m = warp_max(m);
  if (lane == 0) out[row] = ElemTraits<T>::from_f(m);
}
// thread per row for tiny rows (conv taps)
template <typename T>
__global__ void __launch_bounds__(256)
row_absmax_thread(const T* __restrict__ x, int64_t rows, int cols, T* __restrict__ out) {
  const int64_t row = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const T* p = x + row * cols;
  float m = 0.f;
  for (int c = 0; c < cols; ++c) m = fmaxf(m, fabsf(ElemTraits<T>::to_f(p[c])));
  out[row] = ElemTraits<T>::from_f(m);
}

// ------------------------------------------------------------------ tensor |x| max, sq-diff sum
constexpr int kFlatThreads = 256;
template <typename T>
__global__ void __launch_bounds__(kFlatThreads)
absmax_stage1(const T* __restrict__ x, int64_t numel, float* __restrict__ partial, bool vec_ok) {
  constexpr int V = ElemTraits<T>::kVec;
  float m = 0.f;
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthreads = int64_t(gridDim.x) * blockDim.x;
  if (vec_ok) {
    const int64_t nvec = numel / V;
    for (int64_t i = tid; i < nvec; i += nthreads) {
      Vec16<T> a = ld_vec16_stream(x + i * V);
#pragma unroll
      for (int j = 0; j < V; ++j) m = fmaxf(m, fabsf(ElemTraits<T>::to_f(a.v[j])));
    }
    for (int64_t i = nvec * V + tid; i < numel; i += nthreads) m = fmaxf(m, fabsf(ElemTraits<T>::to_f(x[i])));
  } else {
    for (int64_t i = tid; i < numel; i += nthreads) m = fmaxf(m, fabsf(ElemTraits<T>::to_f(x[i])));
  }
  m = warp_max(m);
  __shared__ float sm[kFlatThreads / 32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = sm[0];
    for (int w = 1; w < kFlatThreads / 32; ++w) v = fmaxf(v, sm[w]);
    partial[blockIdx.x] = v;
  }
}
template <typename T>
__global__ void absmax_stage2(const float* __restrict__ partial, int n, T* __restrict__ out) {
  float m = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) m = fmaxf(m, partial[i]);
  m = warp_max(m);
  if (threadIdx.x == 0) out[0] = ElemTraits<T>::from_f(m);
}

template <typename T>
__global__ void __launch_bounds__(kFlatThreads)
sqdiff_stage1(const T* __restrict__ a, const T* __restrict__ b, int64_t numel,
              double* __restrict__ partial, bool vec_ok) {
  constexpr int V = ElemTraits<T>::kVec;
  float acc = 0.f;  // per-thread fp32 like torch's float sum; block/grid folding in fp64
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthreads = int64_t(gridDim.x) * blockDim.x;
  auto term = [](T x, T y) {
    float d = rnd<T>(__fsub_rn(ElemTraits<T>::to_f(x), ElemTraits<T>::to_f(y)));  // (a - b) in dtype
    return __fmul_rn(d, d);                                                          // .float().pow(2)
  };
  if (vec_ok) {
    const int64_t nvec = numel / V;
    int64_t i = tid;
    for (; i + nthreads < nvec; i += 2 * nthreads) {           // four 16-byte loads in flight per thread
      Vec16<T> va = ld_vec16_stream(a + i * V), vb = ld_vec16_stream(b + i * V);
      Vec16<T> vc = ld_vec16_stream(a + (i + nthreads) * V), vd = ld_vec16_stream(b + (i + nthreads) * V);
#pragma unroll
      for (int j = 0; j < V; ++j) acc += term(va.v[j], vb.v[j]);
#pragma unroll
      for (int j = 0; j < V; ++j) acc += term(vc.v[j], vd.v[j]);
    }
    if (i < nvec) {
      Vec16<T> va = ld_vec16_stream(a + i * V), vb = ld_vec16_stream(b + i * V);
#pragma unroll
      for (int j = 0; j < V; ++j) acc += term(va.v[j], vb.v[j]);
    }
    for (int64_t i2 = nvec * V + tid; i2 < numel; i2 += nthreads) acc += term(a[i2], b[i2]);
  } else {
    for (int64_t i = tid; i < numel; i += nthreads) acc += term(a[i], b[i]);
  }
  double d = (double)acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  __shared__ double sm[kFlatThreads / 32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = sm[0];
    for (int w = 1; w < kFlatThreads / 32; ++w) v += sm[w];
    partial[blockIdx.x] = v;
  }
}
__global__ void sqdiff_stage2(const double* __restrict__ partial, int n, double* __restrict__ out) {
  // one CTA, fixed order: thread t folds partial[t], partial[t + 256], ..., then a fixed shuffle / shared tree
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = sm[0];
    for (int w = 1; w < int(blockDim.x >> 5); ++w) t += sm[w];
    out[0] = t;
  }
}

int flat_blocks(int64_t numel, int vec) {
  int64_t b = (numel / vec + kFlatThreads - 1) / kFlatThreads;
  int64_t cap = int64_t(QDM_NUM_SMS) * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return int(b);
}

// ------------------------------------------------------------------ AWQ w_scale column sum
// Block: 8 warps; warp = row lane, lane = 16-byte vector; block covers 32*V columns.
// group max over LPG = group/V adjacent lanes by xor-shuffles.
template <typename T>
__global__ void __launch_bounds__(kColThreads)
awq_wsum_stage1(const T* __restrict__ w, int64_t n_rows, int64_t k_cols, int lanes_per_group,
                float* __restrict__ partial) {
  constexpr int V = ElemTraits<T>::kVec;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t c0 = (int64_t(blockIdx.x) * 32 + lane) * V;
  const bool active = c0 < k_cols;
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  // one row of the group statistics + normalised accumulation; `have` is warp-uniform per row
  auto row_step = [&](const Vec16<T>& v, bool have) {
    if constexpr (std::is_same<T, __half>::value) {
      // fp16: |w| and the group max stay packed (exact), the quotient is formed in fp32 and rounded once
      __half2 a2[4];
#pragma unroll
      for (int p = 0; p < 4; ++p)
        a2[p] = (have && active) ? __habs2(reinterpret_cast<const __half2*>(&v)[p]) : __float2half2_rn(0.f);
      __half2 m2 = __hmax2(__hmax2(a2[0], a2[1]), __hmax2(a2[2], a2[3]));
      m2 = __hmax2(m2, __lowhigh2highlow(m2));
      switch (lanes_per_group) {
        case 32: m2 = __hmax2(m2, __shfl_xor_sync(0xffffffffu, m2, 16));
        case 16: m2 = __hmax2(m2, __shfl_xor_sync(0xffffffffu, m2, 8));
        case 8: m2 = __hmax2(m2, __shfl_xor_sync(0xffffffffu, m2, 4));
        case 4: m2 = __hmax2(m2, __shfl_xor_sync(0xffffffffu, m2, 2));
        case 2: m2 = __hmax2(m2, __shfl_xor_sync(0xffffffffu, m2, 1));
        default: break;
      }
      const float denom = rnd<T>(__fadd_rn(__low2float(m2), 1e-6f));
      const float rd = rcp_approx(denom);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float2 af = __half22float2(a2[p]);
        const float2 qf = __half22float2(__floats2half2_rn(div_by_rcp<false>(af.x, denom, rd), div_by_rcp<false>(af.y, denom, rd)));
        acc[2 * p] += qf.x;
        acc[2 * p + 1] += qf.y;
      }
      return;
    }
    float a[V];
    float m = 0.f;
    if (have && active) {
#pragma unroll
      for (int i = 0; i < V; ++i) { a[i] = fabsf(ElemTraits<T>::to_f(v.v[i])); m = fmaxf(m, a[i]); }
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) a[i] = 0.f;
    }
    for (int o = 1; o < lanes_per_group; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float denom = rnd<T>(__fadd_rn(m, 1e-6f));  // amax + 1e-6 (python scalar), rounded to dtype
    if (fastdiv_ok<T>(denom, m)) {                    // reciprocal division, exact (qdm_common.cuh)
      const float rd = rcp_approx(denom);
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] += rnd<T>(div_by_rcp<false>(a[i], denom, rd));
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] += rnd<T>(__fdiv_rn(a[i], denom));
    }
  };
  const int64_t step = int64_t(gridDim.y) * kColWarps;
  if constexpr (std::is_same<T, __half>::value) {
    // fp16: rows in PAIRS -- the two rows' lane maxima travel through the group butterfly as the two halves of one half2
    // (half the shuffles per row), the quotients |w| / (max + 1e-6) are formed two at a time with packed fp32 FMAs
    // (fma.rn.f32x2: same rounding as the scalar form, half the issue slots); four 16-byte loads in flight per thread.
    // Rows are still accumulated in the order r, r + step, r + 2 step, ...: bit-identical sums to the one-row form.
    float2 acc2[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) acc2[p] = make_float2(0.f, 0.f);
    auto pair_step = [&](const Vec16<T>& va, const Vec16<T>& vb, bool have_b) {
      __half2 a2[4], b2[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        a2[p] = active ? __habs2(reinterpret_cast<const __half2*>(&va)[p]) : __float2half2_rn(0.f);
        b2[p] = (active && have_b) ? __habs2(reinterpret_cast<const __half2*>(&vb)[p]) : __float2half2_rn(0.f);
      }
      __half2 ma = __hmax2(__hmax2(a2[0], a2[1]), __hmax2(a2[2], a2[3]));
      __half2 mb = __hmax2(__hmax2(b2[0], b2[1]), __hmax2(b2[2], b2[3]));
      ma = __hmax2(ma, __lowhigh2highlow(ma));
      mb = __hmax2(mb, __lowhigh2highlow(mb));
      __half2 m2 = __lows2half2(ma, mb);   // (row a, row b)
      switch (lanes_per_group) {
        case 32: m2 = __hmax2(m2, __shfl_xor_sync(0xffffffffu, m2, 16));
        case 16: m2 = __hmax2(m2, __shfl_xor_sync(0xffffffffu, m2, 8));
        case 8: m2 = __hmax2(m2, __shfl_xor_sync(0xffffffffu, m2, 4));
        case 4: m2 = __hmax2(m2, __shfl_xor_sync(0xffffffffu, m2, 2));
        case 2: m2 = __hmax2(m2, __shfl_xor_sync(0xffffffffu, m2, 1));
        default: break;
      }
      const float da = rnd<T>(__fadd_rn(__low2float(m2), 1e-6f)), db = rnd<T>(__fadd_rn(__high2float(m2), 1e-6f));
      const float ra = rcp_approx(da), rb = rcp_approx(db);
      auto add_row = [&](const __half2 (&x2)[4], float d, float r) {
        const float2 rr = make_float2(r, r), nd = make_float2(-d, -d);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float2 af = __half22float2(x2[p]);
          const float2 q0 = __fmul2_rn(af, rr);
          const float2 e = __ffma2_rn(q0, nd, af);          // exact remainder a - q0 * d
          const float2 q1 = __ffma2_rn(e, rr, q0);
          acc2[p] = __fadd2_rn(acc2[p], __half22float2(__floats2half2_rn(q1.x, q1.y)));
        }
      };
      add_row(a2, da, ra);
      if (have_b) add_row(b2, db, rb);
    };
    for (int64_t r = int64_t(blockIdx.y) * kColWarps + warp; r < n_rows; r += 4 * step) {
      Vec16<T> v[4];
      bool have[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        have[u] = r + u * step < n_rows;
        if (active && have[u]) v[u] = ld_vec16_stream(w + (r + u * step) * k_cols + c0);
      }
      pair_step(v[0], v[1], have[1]);
      if (have[2]) pair_step(v[2], v[3], have[3]);
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) { acc[2 * p] = acc2[p].x; acc[2 * p + 1] = acc2[p].y; }
  } else {
    for (int64_t r = int64_t(blockIdx.y) * kColWarps + warp; r < n_rows; r += 2 * step) {   // two loads in flight
      const bool two = r + step < n_rows;
      Vec16<T> va, vb;
      if (active) va = ld_vec16_stream(w + r * k_cols + c0);
      if (active && two) vb = ld_vec16_stream(w + (r + step) * k_cols + c0);
      row_step(va, true);
      if (two) row_step(vb, true);
    }
  }
  __shared__ float sm[kColWarps][32 * V + 1];
#pragma unroll
  for (int i = 0; i < V; ++i) sm[warp][lane * V + i] = acc[i];
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * V; c += kColThreads) {
    float v = sm[0][c];
#pragma unroll
    for (int ww = 1; ww < kColWarps; ++ww) v += sm[ww][c];
    const int64_t col = int64_t(blockIdx.x) * 32 * V + c;
    if (col < k_cols) partial[int64_t(blockIdx.y) * k_cols + col] = v;
  }
}

}  // namespace

// ====================================================================== C ABI
extern "C" size_t qdm_colreduce_workspace_bytes(int64_t rows, int64_t cols) {
  (void)rows;
  return size_t(256) * size_t(cols > 0 ? cols : 1) * sizeof(float);
}

#define QDM_DISPATCH_DTYPE(dtype, ...)                                  \
  switch (dtype) {                                                      \
    case QDM_F16: { using T = __half; __VA_ARGS__; } break;             \
    case QDM_BF16: { using T = __nv_bfloat16; __VA_ARGS__; } break;     \
    case QDM_F32: { using T = float; __VA_ARGS__; } break;              \
    default: qdm_set_error("unknown dtype %d", dtype); return QDM_ERR_INVALID; \
  }

extern "C" int qdm_colabsmax(const void* x, int dtype, int64_t rows, int64_t cols, int64_t ld,
                             void* out, int mode, void* workspace, size_t workspace_bytes, void* stream) {
  QDM_REQUIRE(x && out && workspace, "qdm_colabsmax: null pointer");
  QDM_REQUIRE(rows > 0 && cols > 0 && ld >= cols, "qdm_colabsmax: bad shape rows=%lld cols=%lld ld=%lld",
              (long long)rows, (long long)cols, (long long)ld);
  QDM_REQUIRE(mode == 0 || mode == 1, "qdm_colabsmax: mode must be 0 or 1");
  QDM_DEVICE_GATE();
  cudaStream_t st = (cudaStream_t)stream;
  QDM_DISPATCH_DTYPE(dtype, return (launch_col_reduce<T, COL_ABSMAX, T>((const T*)x, rows, cols, ld, (T*)out, mode,
                                                                       (float*)workspace, workspace_bytes, st)));
  return QDM_OK;
}

extern "C" int qdm_colabssum(const void* x, int dtype, int64_t rows, int64_t cols, int64_t ld,
                             float* out_sum, void* workspace, size_t workspace_bytes, void* stream) {
  QDM_REQUIRE(x && out_sum && workspace, "qdm_colabssum: null pointer");
  QDM_REQUIRE(rows > 0 && cols > 0 && ld >= cols, "qdm_colabssum: bad shape");
  QDM_DEVICE_GATE();
  cudaStream_t st = (cudaStream_t)stream;
  QDM_DISPATCH_DTYPE(dtype, return (launch_col_reduce<T, COL_ABSSUM, float>((const T*)x, rows, cols, ld, out_sum, 0,
                                                                           (float*)workspace, workspace_bytes, st)));
  return QDM_OK;
}

extern "C" size_t qdm_colstats_workspace_bytes(int64_t rows, int64_t cols) {
  return 2 * qdm_colreduce_workspace_bytes(rows, cols);
}

extern "C" int qdm_colstats(const void* x, int dtype, int64_t rows, int64_t cols, int64_t ld, void* out_max,
                            int max_mode, double* acc_maxsum, double* acc_abssum, void* workspace,
                            size_t workspace_bytes, void* stream) {
  QDM_REQUIRE(x && workspace, "qdm_colstats: null pointer");
  QDM_REQUIRE(out_max || acc_maxsum || acc_abssum, "qdm_colstats: no output requested");
  QDM_REQUIRE(rows > 0 && cols > 0 && ld >= cols, "qdm_colstats: bad shape rows=%lld cols=%lld ld=%lld",
              (long long)rows, (long long)cols, (long long)ld);
  QDM_REQUIRE(max_mode == 0 || max_mode == 1, "qdm_colstats: max_mode must be 0 or 1");
  QDM_DEVICE_GATE();
  cudaStream_t st = (cudaStream_t)stream;
  QDM_DISPATCH_DTYPE(dtype, return (launch_col_stats<T>((const T*)x, rows, cols, ld, (T*)out_max, max_mode, acc_maxsum,
                                                       acc_abssum, (float*)workspace, workspace_bytes, st)));
  return QDM_OK;
}

extern "C" int qdm_rowabsmax(const void* x, int dtype, int64_t rows, int64_t cols, void* out, void* stream) {
  QDM_REQUIRE(x && out, "qdm_rowabsmax: null pointer");
  QDM_REQUIRE(rows > 0 && cols > 0, "qdm_rowabsmax: bad shape");
  QDM_DEVICE_GATE();
  cudaStream_t st = (cudaStream_t)stream;
  QDM_DISPATCH_DTYPE(dtype, {
    constexpr int V = ElemTraits<T>::kVec;
    if (cols < 64) {
      row_absmax_thread<T><<<(unsigned)((rows + 255) / 256), 256, 0, st>>>((const T*)x, rows, (int)cols, (T*)out);
    } else {
      const bool vec_ok = (cols % V == 0) && qdm_aligned16(x);
      row_absmax_warp<T><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>((const T*)x, rows, cols, (T*)out, vec_ok);
    }
    QDM_LAUNCH_CHECK();
  });
  return QDM_OK;
}

extern "C" size_t qdm_absmax_workspace_bytes(int64_t numel) {
  (void)numel;
  return size_t(QDM_NUM_SMS) * 8 * sizeof(float);
}

int qdm_absmax_impl(const void* x, int dtype, int64_t numel, void* out, void* workspace,
                    size_t workspace_bytes, cudaStream_t st) {
  QDM_DISPATCH_DTYPE(dtype, {
    constexpr int V = ElemTraits<T>::kVec;
    const int blocks = flat_blocks(numel, V);
    QDM_REQUIRE(workspace_bytes >= size_t(blocks) * sizeof(float), "qdm_absmax: workspace too small");
    absmax_stage1<T><<<blocks, kFlatThreads, 0, st>>>((const T*)x, numel, (float*)workspace, qdm_aligned16(x));
    QDM_LAUNCH_CHECK();
    absmax_stage2<T><<<1, 32, 0, st>>>((const float*)workspace, blocks, (T*)out);
    QDM_LAUNCH_CHECK();
  });
  return QDM_OK;
}

extern "C" int qdm_absmax(const void* x, int dtype, int64_t numel, void* out, void* workspace,
                          size_t workspace_bytes, void* stream) {
  QDM_REQUIRE(x && out && workspace, "qdm_absmax: null pointer");
  QDM_REQUIRE(numel > 0, "qdm_absmax: empty tensor");
  QDM_DEVICE_GATE();
  return qdm_absmax_impl(x, dtype, numel, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" size_t qdm_sqdiff_workspace_bytes(int64_t numel) {
  (void)numel;
  return size_t(QDM_NUM_SMS) * 8 * sizeof(double);
}

extern "C" int qdm_sqdiff_sum(const void* a, const void* b, int dtype, int64_t numel, double* out,
                              void* workspace, size_t workspace_bytes, void* stream) {
  QDM_REQUIRE(a && b && out && workspace, "qdm_sqdiff_sum: null pointer");
  QDM_REQUIRE(numel > 0, "qdm_sqdiff_sum: empty tensor");
  QDM_DEVICE_GATE();
  cudaStream_t st = (cudaStream_t)stream;
  QDM_DISPATCH_DTYPE(dtype, {
    constexpr int V = ElemTraits<T>::kVec;
    const int blocks = flat_blocks(numel, V);
    QDM_REQUIRE(workspace_bytes >= size_t(blocks) * sizeof(double), "qdm_sqdiff_sum: workspace too small");
    const bool vec_ok = qdm_aligned16(a) && qdm_aligned16(b);
    sqdiff_stage1<T><<<blocks, kFlatThreads, 0, st>>>((const T*)a, (const T*)b, numel, (double*)workspace, vec_ok);
    QDM_LAUNCH_CHECK();
    sqdiff_stage2<<<1, 256, 0, st>>>((const double*)workspace, blocks, out);
    QDM_LAUNCH_CHECK();
  });
  return QDM_OK;
}

extern "C" int qdm_awq_wsum(const void* w, int dtype, int64_t n_rows, int64_t k_cols, int group,
                            float* out_sum, void* workspace, size_t workspace_bytes, void* stream) {
  QDM_REQUIRE(w && out_sum && workspace, "qdm_awq_wsum: null pointer");
  QDM_REQUIRE(n_rows > 0 && k_cols > 0 && group > 0 && k_cols % group == 0,
              "qdm_awq_wsum: group %d must divide k_cols %lld", group, (long long)k_cols);
  QDM_REQUIRE(qdm_aligned16(w), "qdm_awq_wsum: weight must be 16-byte aligned");
  QDM_DEVICE_GATE();
  cudaStream_t st = (cudaStream_t)stream;
  QDM_DISPATCH_DTYPE(dtype, {
    constexpr int V = ElemTraits<T>::kVec;
    const int lpg = group / V;
    QDM_UNSUPPORTED(group % V == 0 && lpg >= 1 && lpg <= 32 && (lpg & (lpg - 1)) == 0,
                    "qdm_awq_wsum: group %d unsupported (need %d*2^j <= %d)", group, V, 32 * V);
    const int64_t col_blocks = (k_cols + 32 * V - 1) / (32 * V);
    const int splits = col_splits(n_rows, col_blocks, 8);
    QDM_REQUIRE(workspace_bytes >= size_t(splits) * k_cols * sizeof(float), "qdm_awq_wsum: workspace too small");
    dim3 grid((unsigned)col_blocks, (unsigned)splits);
    awq_wsum_stage1<T><<<grid, kColThreads, 0, st>>>((const T*)w, n_rows, k_cols, lpg, (float*)workspace);
    QDM_LAUNCH_CHECK();
    col_reduce_stage2<float, COL_ABSSUM><<<(unsigned)((k_cols + 255) / 256), 256, 0, st>>>(
        (const float*)workspace, splits, k_cols, out_sum, 0);
    QDM_LAUNCH_CHECK();
  });
  return QDM_OK;
}
