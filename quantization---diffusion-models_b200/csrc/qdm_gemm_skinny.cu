// (c'') W4A16 for M <= 32 rows, any weight size: the AdaLN / time-embedding / pooled-text projections of the diffusion
// models (SD3.5: norm1.linear [14592, 2432] with M = batch, 75 calls per step; SDXL: M = 8; SD1.5: M = 16).  These are
// bound by the packed-weight bandwidth (0.5 B per weight, every weight used once) and by latency, so the kernel is
// built around the weight stream:
//   * a warp-step is a [16 k rows x 64 columns] block of packed words: lane (nl = lane / 4, t = lane % 4) loads the word
//     of column group nl (8 adjacent words = one 32-byte sector per row) from rows kk + 4 t + {0, 1, 2, 3} -- every
//     sector it touches is fully used.  RING build (N % 32 == 0, the 16-byte alignment of every row piece): each warp streams
//     its packed words through a private shared-memory ring filled by 16-byte cp.async (4 chunks of 64 k rows x 32 B:
//     6-8 KB per warp, ~100 KB per SM in flight -- the 4-byte register loads of the first version kept 32-64 KB per SM in
//     flight and reached 1.0-1.8 TB/s); rows are stored permuted so that the lanes' word reads are conflict-free.  The
//     zero points / scales of the NEXT quantisation group are prefetched into registers when a group is entered.
//     Fallback (N % 32 != 0): two register buffers keep 16-32 4-byte loads in flight per lane;
//   * the weights are the A operand of mma.sync.m16n8k16 (W^T [16 columns x 16 k] times x^T [16 k x 8 rows]): which
//     physical row / column plays which logical (n, k) of the fragment is free as long as A, B and the output agree, so
//     the lane's four words ARE four A fragments (column pairs (2j, 2j+1) of its word, j = 0..3) without any shuffle:
//     a[0] = column 2j, rows 4t, 4t+1;  a[1] = column 2j+1, same rows;  a[2], a[3] = rows 4t+2, 4t+3.  The x fragment is
//     one 8-byte shared-memory load (x[m = nl][kk + 4t .. 4t+3]);
//   * unpack: one byte permute pairs the same byte of two rows, one and-or drops the nibbles into the mantissa of 1024.0
//     (fp16) / 128.0 (bf16), the packed subtract of (magic + z) gives the exact integer q - z and the packed multiply by
//     the scale rounds once: bit-identical to dequantize_gemm (utils/packing_utils.py:87-102), ~1.1 instructions per weight;
//   * K is split over the 8 warps of a CTA and over the CTAs of a thread-block CLUSTER (1, 2, 4 or 8, chosen so that
//     every SM holds ~2 CTAs); partial sums are folded through shared memory inside the CTA and through distributed
//     shared memory inside the cluster, both in a fixed order: deterministic, no workspace, no atomics.
#include "qdm_common.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>

namespace cg = cooperative_groups;

namespace {

constexpr int SK_WARPS = 8;
constexpr int SK_U = 4;        // warp-steps whose packed words are requested together (16 loads in flight per lane)
constexpr int SK_MAX_MT = 4;   // up to 32 rows of x (m8 tiles)

template <bool BF16>
__device__ __forceinline__ void mma_w_x(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (BF16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}

// packed (v - zm) * s in the tensor dtype: v and zm hold magic + integer, so the difference is exact; one rounding
template <bool BF16>
__device__ __forceinline__ uint32_t sub_mul2(uint32_t v, uint32_t zm, uint32_t s) {
  if (BF16) {
    __nv_bfloat162 r = __hmul2(__hsub2(*reinterpret_cast<__nv_bfloat162*>(&v), *reinterpret_cast<__nv_bfloat162*>(&zm)),
                               *reinterpret_cast<__nv_bfloat162*>(&s));
    return *reinterpret_cast<uint32_t*>(&r);
  } else {
    __half2 r = __hmul2(__hsub2(*reinterpret_cast<__half2*>(&v), *reinterpret_cast<__half2*>(&zm)), *reinterpret_cast<__half2*>(&s));
    return *reinterpret_cast<uint32_t*>(&r);
  }
}

constexpr int SK_RING = 4;                        // cp.async ring depth (chunks of SK_U warp-steps) per warp
constexpr int SK_CHUNK_BYTES = SK_U * 16 * 32;    // 64 k rows x 32 B

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool on) {
  const int n = on ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// POST (M <= 8): zero point and scale are applied to the fp32 accumulator once per quantisation group instead of to every
// weight pair.  The tensor core is fed the raw codes -- fp16: the nibble IS the fp16 subnormal q * 2^-24 (`p & 0x000F000F`, one
// lop3 per weight pair; the high nibble of a byte stays in place as 16 q * 2^-24), bf16: (16 + q) / 16 -- the per-step sums of x
// come from shared memory, and at a group boundary  y += s_c * (F_c * acc_g - (off_c + z_c) * sum_k x)  with F_c = 2^24 / 2^20
// (fp16 low / high nibble) or 16 (bf16).  ~1.6 instead of ~3 issue slots per weight, which is what bounds this kernel.  The
// products q * x are exact in fp32 and the scale is applied in fp32: closer to the exact product than the per-weight form
// (which rounds every (q - z) * s to 16 bits first), equal to it within ~2^-11 relative -- the GEMM parity bar is 1e-2.
#ifndef SK_F16_OFF
#define SK_F16_OFF 0        // fp16 code offset: 0 = subnormal codes; 1024 = (1024 + q) * 2^-24 (normal numbers, A/B fallback)
#endif

template <bool BF16>
__device__ __forceinline__ float h16_to_f(uint32_t bits) {
  return BF16 ? __uint_as_float(bits << 16) : __half2float(__ushort_as_half((unsigned short)bits));
}

template <bool BF16, int MT, bool RING, bool POST>
__global__ void __launch_bounds__(SK_WARPS * 32)
w4a16_skinny_kernel(const uint16_t* __restrict__ x, const uint32_t* __restrict__ qweight, const uint32_t* __restrict__ qzeros,
                    const uint16_t* __restrict__ scales, const uint16_t* __restrict__ bias, uint16_t* __restrict__ y,
                    int M, int N, int K, int group, int steps_per_cta, int WN) {
  constexpr uint32_t MAGIC = BF16 ? 0x43004300u : 0x64006400u;   // 128.0 | 1024.0 in both halves
  extern __shared__ uint4 sk_smem_raw[];
  // [8 MT rows][k slice + 16] x slice, then the per-warp partial sums [warps][4 j][MT][4][32], then the CTA's sum
  uint16_t* xs = reinterpret_cast<uint16_t*>(sk_smem_raw);
  const int k_slice = steps_per_cta * 16;
  const int pitch = k_slice + 16;
  // per-warp area: the cp.async ring [SK_RING][64 rows x 32 B] during the main loop (RING), the warp's partial sums
  // [4 j][MT][4][32] after it (the same bytes: a warp writes its sums only when it is done with its ring)
  constexpr int WARP_AREA = RING ? SK_RING * SK_CHUNK_BYTES : 4 * MT * 4 * 32 * 4;
  static_assert(WARP_AREA >= 4 * MT * 4 * 32 * 4, "the partial sums fit the ring");
  uint8_t* area0 = reinterpret_cast<uint8_t*>(xs + M * pitch);   // only the M real rows of x are staged (rows >= M are zero registers)
  auto red_w = [&](int w) { return reinterpret_cast<float*>(area0 + w * WARP_AREA); };
  float* part = reinterpret_cast<float*>(area0 + SK_WARPS * WARP_AREA);   // [WN column groups][4 j][MT][4][32]
  float* xsum_s = part + WN * 4 * MT * 4 * 32;                            // POST: [k16 step of this CTA][8 MT rows] sums of x
  const uint32_t ring0 = uint32_t(__cvta_generic_to_shared(area0)) + (threadIdx.x >> 5) * WARP_AREA;

  cg::cluster_group cluster = cg::this_cluster();
  const int crank = int(cluster.block_rank()), csize = int(cluster.num_blocks());
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // programmatic dependent launch

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nl = lane >> 2, t = lane & 3;
  const int words_per_row = N >> 3;
  // the CTA's 8 warps are WN column groups (64 columns each: WN x 32 contiguous bytes of every k row, so that DRAM and L2
  // lines are used whole) times WK = 8 / WN slices of K
  const int wn = warp % WN, wk = warp / WN, WK = SK_WARPS / WN;
  const int wc = (int(blockIdx.x) * WN + wn) * 8 + nl;  // this lane's packed word column (8 output columns)
  const bool col_on = wc < words_per_row;
  const int steps_total = K >> 4;
  const int s_lo = crank * steps_per_cta;               // this CTA's k16 steps
  const int s_hi = min(steps_total, s_lo + steps_per_cta);
  const int n_steps = max(0, s_hi - s_lo);
  const int spw = (n_steps + WK - 1) / WK;              // contiguous chunk per warp: the group changes rarely
  const int w_lo = s_lo + wk * spw, w_hi = min(s_hi, w_lo + spw);

  // two register buffers of SK_U warp-steps each: the next chunk's words are requested before the current chunk is
  // unpacked, so a warp always has 16-32 loads in flight instead of alternating between waiting and computing
  uint32_t wa[SK_U][4], wb[SK_U][4];
  auto load_words = [&](uint32_t (&w)[SK_U][4], int s0) {
#pragma unroll
    for (int u = 0; u < SK_U; ++u) {
      const bool on = col_on && (s0 + u) < w_hi;
      const uint32_t* wp = qweight + int64_t(((s0 + u) << 4) + 4 * t) * words_per_row + wc;
#pragma unroll
      for (int i = 0; i < 4; ++i) w[u][i] = on ? __ldg(wp + int64_t(i) * words_per_row) : 0u;
    }
  };
  // first chunk of packed words and the first group's zero points / scales: in flight while x is staged
  const int gshift = 31 - __clz(group);   // group is a power of two (checked by the caller): no integer division in the loop
  const int g_pre = (w_lo << 4) >> gshift;
  uint32_t zw_pre = 0u;
  uint4 sv_pre = make_uint4(0, 0, 0, 0);
  // RING: lane L copies the 16-byte pieces L, L + 32, L + 64, L + 96 of a chunk; piece p = (k row p / 2, half p % 2) of the
  // warp's 32-byte column window; row r of a chunk is stored at row (r & ~15) | ((r & 3) << 2) | ((r >> 2) & 3), so that
  // the word reads of lanes (nl, t) -- rows 4 t + i -- fall into 32 different banks
  const int wc0 = (int(blockIdx.x) * WN + wn) * 8;
  auto ring_issue = [&](int chunk) {   // chunk index counted from w_lo in units of SK_U steps
    const int s0 = w_lo + chunk * SK_U;
    const uint32_t dst0 = ring0 + uint32_t(chunk % SK_RING) * SK_CHUNK_BYTES;
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const int pce = lane + 32 * q4, r = pce >> 1, h = pce & 1;
      const int rp = (r & ~15) | ((r & 3) << 2) | ((r >> 2) & 3);
      const bool on = (s0 + (r >> 4)) < w_hi && (wc0 + 4 * h) < words_per_row;
      const uint32_t* src = qweight + int64_t((s0 << 4) + r) * words_per_row + wc0 + 4 * h;
      cp_async16_zfill(dst0 + uint32_t(rp * 32 + h * 16), on ? src : qweight, on);
    }
  };
  const int n_chunks = (w_hi - w_lo + SK_U - 1) / SK_U;
  if (RING) {
#pragma unroll
    for (int c = 0; c < SK_RING - 1; ++c) {
      if (c < n_chunks) ring_issue(c);
      cp_async_commit();
    }
  }
  if (w_lo < w_hi) {
    if (!RING) load_words(wa, w_lo);
    if (col_on) {
      zw_pre = __ldg(qzeros + int64_t(g_pre) * words_per_row + wc);
      sv_pre = __ldg(reinterpret_cast<const uint4*>(scales + int64_t(g_pre) * N + 8 * wc));
    }
  }

  // the loads above are constants (packed words, zeros, scales): they fly while the previous kernel drains; x and y belong
  // to earlier kernels and are not touched before this wait
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // stage this CTA's k slice of the M rows of x
  {
    const int vec_per_row = k_slice >> 3;
    const int k0 = s_lo << 4;
    for (int idx = threadIdx.x; idx < M * vec_per_row; idx += SK_WARPS * 32) {
      const int row = idx / vec_per_row, c8 = idx - row * vec_per_row;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (k0 + c8 * 8 < K) v = __ldg(reinterpret_cast<const uint4*>(x + int64_t(row) * K + k0) + c8);
      *reinterpret_cast<uint4*>(xs + row * pitch + c8 * 8) = v;
    }
  }
  __syncthreads();
  if (POST) {   // sum of the 16 x values of every (k16 step, row): what the code offset / zero point multiplies
    for (int idx = threadIdx.x; idx < steps_per_cta * 8 * MT; idx += SK_WARPS * 32) {
      const int stp = idx / (8 * MT), row = idx - stp * (8 * MT);
      float sum = 0.f;
      if (row < M && s_lo + stp < steps_total) {
        const uint4 v0 = *reinterpret_cast<const uint4*>(xs + row * pitch + 16 * stp);
        const uint4 v1 = *reinterpret_cast<const uint4*>(xs + row * pitch + 16 * stp + 8);
        const uint32_t h[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) sum += h16_to_f<BF16>(h[e] & 0xFFFFu) + h16_to_f<BF16>(h[e] >> 16);
      }
      xsum_s[idx] = sum;
    }
    __syncthreads();
  }

  float acc[4][MT][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][m][i] = 0.f;

  int g_cur = -1;
  uint32_t zm[8], sc2[8];   // per column of this lane's word: (magic + z) and the scale, each duplicated into both halves
  // POST: the current group's partial sums of codes * x, the sums of x they belong to, and the group's fp32 parameters
  float acc_g[POST ? 4 : 1][POST ? MT : 1][4], xs_acc[POST ? MT : 1][2], sc_f[POST ? 8 : 1], kz_f[POST ? 8 : 1];
  if (POST) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc_g[POST ? j : 0][POST ? m : 0][i] = 0.f;
#pragma unroll
    for (int m = 0; m < MT; ++m) xs_acc[POST ? m : 0][0] = xs_acc[POST ? m : 0][1] = 0.f;
  }
  auto flush_group = [&]() {   // y += s_c * (F_c * acc_g - (off_c + z_c) * sum x), then the group accumulators restart
    if constexpr (POST) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = 2 * j + (i >> 1);
            const float fc = BF16 ? 16.f : ((c & 2) ? 1048576.f : 16777216.f);   // high nibbles of a byte carry 16 q
            const float v = __fmaf_rn(acc_g[j][m][i], fc, -kz_f[c] * xs_acc[m][i & 1]);
            acc[j][m][i] = __fmaf_rn(sc_f[c], v, acc[j][m][i]);
            acc_g[j][m][i] = 0.f;
          }
#pragma unroll
      for (int m = 0; m < MT; ++m) xs_acc[m][0] = xs_acc[m][1] = 0.f;
    }
  };
  auto do_chunk = [&](const uint32_t (&w)[SK_U][4], int s0) {
#pragma unroll
    for (int u = 0; u < SK_U; ++u) {
      const int st = s0 + u;
      if (st >= w_hi) break;
      const int kk = st << 4;
      const int g = kk >> gshift;                       // group = 64 * 2^j: a step never straddles two groups
      if (g != g_cur) {
        if (POST && g_cur >= 0) flush_group();
        g_cur = g;
        // zw_pre / sv_pre hold this group's parameters (requested when the previous group was entered, or before the x
        // staging for the first one); request the next group's now: they are needed group / 16 steps from here
        const uint32_t zw = zw_pre;
        const uint4 sv = sv_pre;
        if (col_on && (g + 1) * group < (w_hi << 4)) {
          zw_pre = __ldg(qzeros + int64_t(g + 1) * words_per_row + wc);
          sv_pre = __ldg(reinterpret_cast<const uint4*>(scales + int64_t(g + 1) * N + 8 * wc));
        }
        const uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int nib = (c >> 1) + 4 * (c & 1);       // column c of a word sits in nibble {0,4,1,5,2,6,3,7}[c]
          const uint32_t z = (zw >> (4 * nib)) & 0xFu;
          zm[c] = MAGIC | z | (z << 16);
          const uint32_t s16 = (c & 1) ? (sw[c >> 1] >> 16) : (sw[c >> 1] & 0xFFFFu);
          sc2[c] = s16 | (s16 << 16);
          if constexpr (POST) {
            sc_f[c] = h16_to_f<BF16>(s16);
            kz_f[c] = float(z) + (BF16 ? 16.f : float(SK_F16_OFF) / ((c & 2) ? 16.f : 1.f));
          }
        }
      }
      // x fragment: rows kk + 4t .. 4t + 3 of x row nl (+ 8 per m-tile)
      uint2 xb[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m)
        xb[m] = (8 * m + nl < M) ? *reinterpret_cast<const uint2*>(xs + (8 * m + nl) * pitch + (kk - (s_lo << 4)) + 4 * t) : make_uint2(0u, 0u);
      // unpack: P = (byte b of row r, byte b of row r + 1) -> nibbles 2b (low) and 2b + 1 (high) of both rows
      uint32_t a[4][4];
#pragma unroll
      for (int rp = 0; rp < 2; ++rp) {
        const uint32_t r0 = w[u][2 * rp], r1 = w[u][2 * rp + 1];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const uint32_t p = __byte_perm(r0, r1, 0x4400 + 0x1111 * b);   // bytes [r0.b, r0.b, r1.b, r1.b]
          const uint32_t lo = (p & 0x000F000Fu) | MAGIC;            // nibble 2b     of (row r, row r + 1)
          const uint32_t hi = ((p >> 4) & 0x000F000Fu) | MAGIC;     // nibble 2b + 1 of (row r, row r + 1)
          // nibble n holds column order[n], order = {0,2,4,6,1,3,5,7}
          const int c_lo = (2 * b < 4) ? 2 * (2 * b) : 2 * (2 * b - 4) + 1;
          const int c_hi = (2 * b + 1 < 4) ? 2 * (2 * b + 1) : 2 * (2 * b + 1 - 4) + 1;
          // column c = 2j (+1): fragment j = c / 2, register (c & 1) + 2 * rp
          if constexpr (POST) {
            constexpr uint32_t OFFB = SK_F16_OFF ? 0x04000400u : 0u;   // exponent field 1: 2^-24 * (1024 + mantissa)
            a[c_lo >> 1][(c_lo & 1) + 2 * rp] = BF16 ? (((p << 3) & 0x00780078u) | 0x3F803F80u) : ((p & 0x000F000Fu) | OFFB);
            a[c_hi >> 1][(c_hi & 1) + 2 * rp] = BF16 ? (((p >> 1) & 0x00780078u) | 0x3F803F80u) : ((p & 0x00F000F0u) | OFFB);
          } else {
            a[c_lo >> 1][(c_lo & 1) + 2 * rp] = sub_mul2<BF16>(lo, zm[c_lo], sc2[c_lo]);
            a[c_hi >> 1][(c_hi & 1) + 2 * rp] = sub_mul2<BF16>(hi, zm[c_hi], sc2[c_hi]);
          }
        }
      }
      if constexpr (POST) {
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const float2 s2 = *reinterpret_cast<const float2*>(xsum_s + (st - s_lo) * (8 * MT) + 8 * m + 2 * t);
          xs_acc[m][0] += s2.x;
          xs_acc[m][1] += s2.y;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int m = 0; m < MT; ++m) mma_w_x<BF16>(acc_g[j][m], a[j], xb[m].x, xb[m].y);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int m = 0; m < MT; ++m) mma_w_x<BF16>(acc[j][m], a[j], xb[m].x, xb[m].y);
      }
    }
  };
  if (RING) {
    for (int c = 0; c < n_chunks; ++c) {
      cp_async_wait<SK_RING - 2>();   // chunk c has landed (this lane's pieces) ...
      __syncwarp();                   // ... and every lane's; all lanes are also done reading chunk c - 1
      if (c + SK_RING - 1 < n_chunks) ring_issue(c + SK_RING - 1);   // into the slot of chunk c - 1
      cp_async_commit();
#pragma unroll
      for (int u = 0; u < SK_U; ++u)
#pragma unroll
        for (int i = 0; i < 4; ++i)   // row 16 u + 4 t + i is stored at row 16 u + 4 i + t
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wa[u][i]) : "r"(ring0 + uint32_t(c % SK_RING) * SK_CHUNK_BYTES + uint32_t((16 * u + 4 * i + t) * 32 + nl * 4)));
      do_chunk(wa, w_lo + c * SK_U);
    }
  } else {
    for (int s0 = w_lo; s0 < w_hi; s0 += 2 * SK_U) {        // wa holds chunk s0
      if (s0 + SK_U < w_hi) load_words(wb, s0 + SK_U);
      do_chunk(wa, s0);
      if (s0 + 2 * SK_U < w_hi) load_words(wa, s0 + 2 * SK_U);
      if (s0 + SK_U < w_hi) do_chunk(wb, s0 + SK_U);
    }
  }

  if (POST && g_cur >= 0) flush_group();   // the last group of this warp's k range
  // fold the warps' K chunks (fixed order), then the cluster's K slices (fixed order)
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int i = 0; i < 4; ++i) red_w(warp)[((j * MT + m) * 4 + i) * 32 + lane] = acc[j][m][i];
  __syncthreads();
  constexpr int PER_LANE = 4 * MT * 4;
  for (int it = warp; it < WN * PER_LANE; it += SK_WARPS) {   // item = (column group, element): sum its WK slices
    const int cg_ = it / PER_LANE, e = it - cg_ * PER_LANE;
    float v = red_w(cg_)[e * 32 + lane];                      // warp index = k * WN + cg_
    for (int k = 1; k < WK; ++k) v += red_w(k * WN + cg_)[e * 32 + lane];
    part[it * 32 + lane] = v;
  }
  cluster.sync();   // every CTA's `part` is complete and visible cluster-wide
  // Reduce-scatter over the cluster: element e is folded by CTA e % csize (and one of its warps), which reads the csize
  // partials through distributed shared memory -- all loads issued before the first add (a dependent chain of remote
  // loads costs ~0.3 us each) -- adds them in rank order, adds the bias and stores.  Element e = (j * MT + m) * 4 + i of
  // lane (nl, t) is y[8 m + 2 t + (i & 1)][8 wc + 2 j + (i >> 1)].
  for (int it = crank + csize * warp; it < WN * PER_LANE; it += csize * SK_WARPS) {
    float vals[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) vals[r] = (r < csize) ? cluster.map_shared_rank(part, r)[it * 32 + lane] : 0.f;
    float v = vals[0];
#pragma unroll
    for (int r = 1; r < 8; ++r) v += vals[r];
    const int cg_ = it / PER_LANE, e = it - cg_ * PER_LANE;
    const int i = e & 3, jm = e >> 2, m = jm % MT, j = jm / MT;
    const int wco = (int(blockIdx.x) * WN + cg_) * 8 + nl;   // the word column this ITEM belongs to (not this warp's)
    const int row = 8 * m + 2 * t + (i & 1), col = 8 * wco + 2 * j + (i >> 1);
    if (wco < words_per_row && row < M) {
      if (bias) v += BF16 ? __bfloat162float(__ushort_as_bfloat16(bias[col])) : __half2float(__ushort_as_half(bias[col]));
      y[int64_t(row) * N + col] = BF16 ? __bfloat16_as_ushort(__float2bfloat16_rn(v)) : __half_as_ushort(__float2half_rn(v));
    }
  }
  cluster.sync();   // every CTA's shared memory stays alive until its peers have read it
}

bool skinny_post_ok(int mt) {
  static const bool off = getenv("QDM_SKINNY_NO_POST") != nullptr;   // A/B switch, read once
  return mt == 1 && !off;   // M <= 8; at M = 16 the second accumulator set costs the second resident CTA (measured 16.2 vs 15.1 us)
}

size_t skinny_smem(int64_t M, int steps_per_cta, int wn, bool ring) {
  const int mt = int((M + 7) / 8);
  const size_t sums = size_t(4) * mt * 4 * 32 * sizeof(float);
  const size_t warp_area = ring ? size_t(SK_RING) * SK_CHUNK_BYTES : sums;
  const size_t xsum = skinny_post_ok(mt) ? size_t(steps_per_cta) * 8 * mt * sizeof(float) : 0;
  return size_t(M) * (steps_per_cta * 16 + 16) * 2 + SK_WARPS * warp_area + wn * sums + xsum;
}

template <bool BF16, int MT, bool RING, bool POST>
int skinny_launch(dim3 grid, int ks, size_t smem, const void* x, const int32_t* qweight, const int32_t* qzeros, const void* scales,
                  const void* bias, void* y, int M, int N, int K, int group, int steps_per_cta, int wn, cudaStream_t st) {
  auto kern = w4a16_skinny_kernel<BF16, MT, RING, POST>;
  static size_t smem_set = 0;   // per instantiation; grows monotonically
  if (smem > 48 * 1024 && smem > smem_set) {
    QDM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    smem_set = 200 * 1024;
  }
  static const bool no_pdl = getenv("QDM_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(SK_WARPS * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = (unsigned)ks; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = no_pdl ? 1 : 2;
  QDM_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, (const uint16_t*)x, (const uint32_t*)qweight, (const uint32_t*)qzeros,
                                 (const uint16_t*)scales, (const uint16_t*)bias, (uint16_t*)y, M, N, K, group, steps_per_cta, wn));
  QDM_LAUNCH_CHECK();
  return QDM_OK;
}

template <bool BF16, bool RING>
int skinny_mt(int mt, dim3 grid, int ks, size_t smem, const void* x, const int32_t* qweight, const int32_t* qzeros, const void* scales,
              const void* bias, void* y, int M, int N, int K, int group, int spc, int wn, cudaStream_t st) {
  const bool post = skinny_post_ok(mt);
  switch (mt) {
    case 1: return post ? skinny_launch<BF16, 1, RING, true>(grid, ks, smem, x, qweight, qzeros, scales, bias, y, M, N, K, group, spc, wn, st)
                        : skinny_launch<BF16, 1, RING, false>(grid, ks, smem, x, qweight, qzeros, scales, bias, y, M, N, K, group, spc, wn, st);
    case 2: return skinny_launch<BF16, 2, RING, false>(grid, ks, smem, x, qweight, qzeros, scales, bias, y, M, N, K, group, spc, wn, st);
    case 3: return skinny_launch<BF16, 3, RING, false>(grid, ks, smem, x, qweight, qzeros, scales, bias, y, M, N, K, group, spc, wn, st);
    default: return skinny_launch<BF16, 4, RING, false>(grid, ks, smem, x, qweight, qzeros, scales, bias, y, M, N, K, group, spc, wn, st);
  }
}

// the 16-byte cp.async pieces need every k row's column window 16-byte aligned
bool skinny_ring_ok(int64_t N) {
  static const bool off = getenv("QDM_SKINNY_NO_RING") != nullptr;   // A/B switch, read once
  return N % 32 == 0 && !off;
}

// Work split: WN = 4 column groups per CTA (128 contiguous bytes of every k row) when N is wide enough, K over the
// remaining warps and over the cluster: enough CTAs for ~2 per SM, at least two k16 steps per warp, shared memory in bounds
void skinny_plan(int64_t M, int64_t N, int64_t K, int* ks_out, int* spc_out, int* wn_out, size_t* smem_out) {
  const int mt = int((M + 7) / 8);
  const int wn = N >= 256 ? 4 : (N >= 128 ? 2 : 1);
  const int wk = SK_WARPS / wn;
  const int64_t col_blocks = (N + 64 * wn - 1) / (64 * wn), steps = K / 16;
  int ks = 1;
  // every CTA resident at once (2 per SM up to 16 rows, 1 above: 124-126 vs 160-168 registers): a second wave would
  // repeat the whole load / reduce latency chain
  const int64_t resident = int64_t(mt <= 2 ? 2 : 1) * QDM_NUM_SMS;
  while (ks < 8 && col_blocks * ks * 2 <= resident && steps / (2 * ks) >= 2 * wk) ks *= 2;
  int spc = int((steps + ks - 1) / ks);
  const bool ring = skinny_ring_ok(N);
  while (ks < 8 && skinny_smem(M, spc, wn, ring) > 160 * 1024) { ks *= 2; spc = int((steps + ks - 1) / ks); }
  *ks_out = ks;
  *spc_out = spc;
  *wn_out = wn;
  *smem_out = skinny_smem(M, spc, wn, ring);
}

}  // namespace

bool qdm_gemm_w4a16_skinny_fits(int64_t M, int64_t N, int64_t K) {
  if (M > 8 * SK_MAX_MT || K % 16 || N % 8) return false;
  int ks, spc, wn;
  size_t smem;
  skinny_plan(M, N, K, &ks, &spc, &wn, &smem);
  return smem <= 160 * 1024;
}

// called by qdm_gemm_w4a16 for M <= 32 (arguments already validated there)
int qdm_gemm_w4a16_skinny(const void* x, const int32_t* qweight, const int32_t* qzeros, const void* scales, const void* bias,
                          void* y, int is_bf16, int64_t M, int64_t N, int64_t K, int group, cudaStream_t st) {
  int ks, spc, wn;
  size_t smem;
  skinny_plan(M, N, K, &ks, &spc, &wn, &smem);
  const int mt = int((M + 7) / 8);
  dim3 grid((unsigned)((N + 64 * wn - 1) / (64 * wn)), (unsigned)ks);
  if (skinny_ring_ok(N))
    return is_bf16 ? skinny_mt<true, true>(mt, grid, ks, smem, x, qweight, qzeros, scales, bias, y, (int)M, (int)N, (int)K, group, spc, wn, st)
                   : skinny_mt<false, true>(mt, grid, ks, smem, x, qweight, qzeros, scales, bias, y, (int)M, (int)N, (int)K, group, spc, wn, st);
  return is_bf16 ? skinny_mt<true, false>(mt, grid, ks, smem, x, qweight, qzeros, scales, bias, y, (int)M, (int)N, (int)K, group, spc, wn, st)
                 : skinny_mt<false, false>(mt, grid, ks, smem, x, qweight, qzeros, scales, bias, y, (int)M, (int)N, (int)K, group, spc, wn, st);
}
