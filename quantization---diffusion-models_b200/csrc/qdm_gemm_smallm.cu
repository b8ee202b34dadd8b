// (c') W4A16 for M <= 64 rows: the time-embedding / AdaLN projections of the diffusion models (SD1.5: 24 of the 184
// Linear calls of a step with M = 16; SD3.5: norm1.linear [14592, 2432] with M = batch).  These are weight-bandwidth
// bound (0.5 B per weight, every weight used once per 16 rows) and latency bound: the persistent tcgen05 kernel needs
// K / 64 serial pipeline periods plus ~5 us of setup for them.  Here every CTA owns ONE packed word column (8 output
// columns), its 8 warps split K eight ways, x is staged once per CTA in shared memory, each warp runs mma.sync.m16n8k16 (fp32 accumulate) on weights it
// dequantises straight from the AWQ words in registers, and the four partial tiles are folded through shared memory.
// The B fragment of m16n8k16 wants, per lane, column n = lane / 4 and rows k = 2 (lane % 4) + {0, 1, 8, 9}: one nibble
// of four words of the column -- the [0,2,4,6,1,3,5,7] nibble order only changes the shift.
// (q - z) * s is formed exactly as utils/packing_utils.py:87-102 does (integer difference, one rounding in the
// tensor dtype), so the product differs from the large-M kernel only by accumulation order.
#include "qdm_common.cuh"
#include <stdlib.h>

namespace {

constexpr int SM_WARPS = 8;     // K is split eight ways inside the CTA
constexpr int SM_MAX_MT = 4;    // up to 64 rows
constexpr int SM_CHUNK = 8;     // k16 steps whose packed words are requested together (32 loads in flight per lane)

template <bool BF16>
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  if (BF16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
}

// (q - z) * s for two k-consecutive weights of one column -> packed pair in the tensor dtype
template <bool BF16>
__device__ __forceinline__ uint32_t dq_pair(uint32_t w0, uint32_t w1, int shift, int z, uint16_t s_bits) {
  const int q0 = int((w0 >> shift) & 0xFu) - z, q1 = int((w1 >> shift) & 0xFu) - z;
  if (BF16) {
    const __nv_bfloat16 s = __ushort_as_bfloat16(s_bits);
    __nv_bfloat162 r = __halves2bfloat162(__hmul(__int2bfloat16_rn(q0), s), __hmul(__int2bfloat16_rn(q1), s));
    return *reinterpret_cast<uint32_t*>(&r);
  } else {
    const __half s = __ushort_as_half(s_bits);
    __half2 r = __halves2half2(__hmul(__int2half_rn(q0), s), __hmul(__int2half_rn(q1), s));
    return *reinterpret_cast<uint32_t*>(&r);
  }
}

template <bool BF16, int MT>
__global__ void __launch_bounds__(SM_WARPS * 32)
w4a16_smallm_kernel(const uint16_t* __restrict__ x, const uint32_t* __restrict__ qweight, const uint32_t* __restrict__ qzeros,
                    const uint16_t* __restrict__ scales, const uint16_t* __restrict__ bias, uint16_t* __restrict__ y,
                    int M, int N, int K, int group) {
  extern __shared__ uint4 xs_raw[];                  // x staged once per CTA: [16 MT][K + 8] (pitch: conflict-free fragments)
  uint16_t* xs = reinterpret_cast<uint16_t*>(xs_raw);
  __shared__ float red[SM_WARPS][MT][4][32];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // programmatic dependent launch
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wc = blockIdx.x;                        // packed word column: output columns 8 wc .. 8 wc + 7
  const int words_per_row = N >> 3;
  const int pitch = K + 8;
  const int t = lane & 3, nl = lane >> 2;           // fragment coordinates: k pair index, column / row inside the tile
  const int shift = 4 * ((nl >> 1) + 4 * (nl & 1)); // nibble of column nl: AWQ order {0,2,4,6,1,3,5,7} inverted
  const int steps = K >> 4;                         // k16 steps
  const int spw = (steps + SM_WARPS - 1) / SM_WARPS;
  const int s_begin = warp * spw, s_end = min(steps, s_begin + spw);

  // the first chunk of packed words is requested before x is staged: both round trips overlap
  uint32_t w[SM_CHUNK][4], zw[SM_CHUNK];
  uint16_t sc[SM_CHUNK];
  auto load_chunk = [&](int s0) {
#pragma unroll
    for (int u = 0; u < SM_CHUNK; ++u) {
      const int st = s0 + u;
      const bool on = st < s_end;
      const int kk = st << 4;
      const uint32_t* wp = qweight + int64_t(kk + 2 * t) * words_per_row + wc;
#pragma unroll
      for (int i = 0; i < 4; ++i)   // rows 2t, 2t + 1, 2t + 8, 2t + 9
        w[u][i] = on ? __ldg(wp + int64_t((i & 1) + 8 * (i >> 1)) * words_per_row) : 0u;
      const int g = on ? kk / group : 0;             // group % 16 == 0: a k16 step never straddles groups
      zw[u] = on ? __ldg(qzeros + int64_t(g) * words_per_row + wc) : 0u;
      sc[u] = on ? __ldg(scales + int64_t(g) * N + 8 * wc + nl) : uint16_t(0);
    }
  };
  load_chunk(s_begin);   // constants (packed words, zeros, scales): fetched while the previous kernel drains
  asm volatile("griddepcontrol.wait;" ::: "memory");   // x and y belong to earlier kernels: nothing of theirs is touched before this
  {
    const int vec_per_row = K >> 3;
    for (int idx = threadIdx.x; idx < 16 * MT * vec_per_row; idx += SM_WARPS * 32) {
      const int row = idx / vec_per_row, c8 = idx - row * vec_per_row;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (row < M) v = __ldg(reinterpret_cast<const uint4*>(x + int64_t(row) * K) + c8);
      *reinterpret_cast<uint4*>(xs + row * pitch + c8 * 8) = v;
    }
  }
  __syncthreads();

  float acc[MT][4];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[m][i] = 0.f;
  for (int s0 = s_begin; s0 < s_end; s0 += SM_CHUNK) {
    if (s0 != s_begin) load_chunk(s0);
#pragma unroll
    for (int u = 0; u < SM_CHUNK; ++u) {
      const int st = s0 + u;
      if (st >= s_end) break;
      const int kk = st << 4;
      const int z = int((zw[u] >> shift) & 0xFu);
      uint32_t b[2];
      b[0] = dq_pair<BF16>(w[u][0], w[u][1], shift, z, sc[u]);
      b[1] = dq_pair<BF16>(w[u][2], w[u][3], shift, z, sc[u]);
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const uint16_t* xp = xs + (16 * m + nl) * pitch + kk + 2 * t;
        uint32_t a[4];
        a[0] = *reinterpret_cast<const uint32_t*>(xp);
        a[1] = *reinterpret_cast<const uint32_t*>(xp + 8 * pitch);
        a[2] = *reinterpret_cast<const uint32_t*>(xp + 8);
        a[3] = *reinterpret_cast<const uint32_t*>(xp + 8 * pitch + 8);
        mma_16816<BF16>(acc[m], a, b);
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int i = 0; i < 4; ++i) red[warp][m][i][lane] = acc[m][i];
  __syncthreads();
  // fold the K slices in a fixed order and store: warp w takes m-tiles w, w + 8, ...
  for (int m = warp; m < MT; m += SM_WARPS) {
    float c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v = red[0][m][i][lane];
#pragma unroll
      for (int ww = 1; ww < SM_WARPS; ++ww) v += red[ww][m][i][lane];
      c[i] = v;
    }
    const int col = 8 * wc + 2 * t;
    float b0 = 0.f, b1 = 0.f;
    if (bias) {
      if (BF16) { b0 = __bfloat162float(__ushort_as_bfloat16(bias[col])); b1 = __bfloat162float(__ushort_as_bfloat16(bias[col + 1])); }
      else { b0 = __half2float(__ushort_as_half(bias[col])); b1 = __half2float(__ushort_as_half(bias[col + 1])); }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = 16 * m + nl + 8 * h;
      if (row < M) {
        uint32_t o;
        if (BF16) { __nv_bfloat162 v = __floats2bfloat162_rn(c[2 * h] + b0, c[2 * h + 1] + b1); o = *reinterpret_cast<uint32_t*>(&v); }
        else { __half2 v = __floats2half2_rn(c[2 * h] + b0, c[2 * h + 1] + b1); o = *reinterpret_cast<uint32_t*>(&v); }
        *reinterpret_cast<uint32_t*>(y + int64_t(row) * N + col) = o;
      }
    }
  }
}

template <bool BF16, int MT>
int launch_one(dim3 grid, size_t smem, const void* x, const int32_t* qweight, const int32_t* qzeros, const void* scales,
               const void* bias, void* y, int M, int N, int K, int group, cudaStream_t st) {
  auto kern = w4a16_smallm_kernel<BF16, MT>;
  static size_t smem_set = 0;   // per instantiation; grows monotonically
  if (smem > 48 * 1024 && smem > smem_set) {
    QDM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    smem_set = 200 * 1024;
  }
  // programmatic stream serialisation, as for the tcgen05 kernels (the kernel waits before its first global access)
  static const bool no_pdl = getenv("QDM_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(SM_WARPS * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  QDM_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, (const uint16_t*)x, (const uint32_t*)qweight, (const uint32_t*)qzeros,
                                 (const uint16_t*)scales, (const uint16_t*)bias, (uint16_t*)y, M, N, K, group));
  QDM_LAUNCH_CHECK();
  return QDM_OK;
}

template <bool BF16>
int launch_mt(int mt, dim3 grid, size_t smem, const void* x, const int32_t* qweight, const int32_t* qzeros, const void* scales,
              const void* bias, void* y, int M, int N, int K, int group, cudaStream_t st) {
  switch (mt) {
    case 1: return launch_one<BF16, 1>(grid, smem, x, qweight, qzeros, scales, bias, y, M, N, K, group, st);
    case 2: return launch_one<BF16, 2>(grid, smem, x, qweight, qzeros, scales, bias, y, M, N, K, group, st);
    case 3: return launch_one<BF16, 3>(grid, smem, x, qweight, qzeros, scales, bias, y, M, N, K, group, st);
    default: return launch_one<BF16, 4>(grid, smem, x, qweight, qzeros, scales, bias, y, M, N, K, group, st);
  }
}

}  // namespace

// When the small-M kernel is used.  Measured (GPU-side, L2 flushed): it wins for M <= 32 and weights up to ~6 M
// elements (M = 16, 1280 x 1280: 9.0 vs 14.9 us); every CTA reads one 4-byte word column, i.e. an eighth of each 32-byte
// sector it touches, so for large weights (AdaLN 14592 x 2432) the L2 -> SM over-fetch makes it slower than the tcgen05
// kernel (78 vs 28 us) -- a sector-wide (8 word columns per CTA), cluster-split-K version is the next step there.
bool qdm_gemm_w4a16_smallm_fits(int64_t M, int64_t N, int64_t K) {
  const int64_t mt = (M + 15) / 16;
  return M <= 32 && N * K <= (int64_t(13) << 19) && mt * 16 * (K + 8) * 2 <= 190 * 1024;
}

// called by qdm_gemm_w4a16 for M <= 64 (arguments already validated there)
int qdm_gemm_w4a16_smallm(const void* x, const int32_t* qweight, const int32_t* qzeros, const void* scales, const void* bias,
                          void* y, int is_bf16, int64_t M, int64_t N, int64_t K, int group, cudaStream_t st) {
  const int mt = int((M + 15) / 16);
  const size_t smem = size_t(mt) * 16 * (K + 8) * 2;
  dim3 grid((unsigned)(N / 8));
  return is_bf16 ? launch_mt<true>(mt, grid, smem, x, qweight, qzeros, scales, bias, y, (int)M, (int)N, (int)K, group, st)
                 : launch_mt<false>(mt, grid, smem, x, qweight, qzeros, scales, bias, y, (int)M, (int)N, (int)K, group, st);
}
