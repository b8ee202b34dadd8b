// (c)(d) Quantized-linear GEMMs on the 5th-gen tensor cores: tcgen05.mma with TMEM accumulators,
// TMA-fed operand tiles, mbarrier pipelines, persistent warp-specialised CTAs (one per SM).
//
//   G_F16     y = x[M,K] . w[N,K]^T            both operands by TMA, K-major SW128
//             = WxAxLinear.forward on fake-quant weights            quantize/fake_quant.py:223
//   G_F16_KN  y = x[M,K] . w_kn[K,N]           B by TMA in MN-major SW128 (dequantize_gemm output,
//                                              utils/packing_utils.py:87-102)
//   G_W4      y = x[M,K] . dequant(qweight,qzeros,scales)[K,N]
//             AWQ int4 layout consumed as stored (utils/packing_utils.py:4-27): 8 dequant warps
//             unpack nibble pairs (lop3 magic-number trick the [0,2,4,6,1,3,5,7] order exists for),
//             apply (q - z) * s in the tensor dtype (bit-exact with dequantize_gemm) and write the
//             MN-major SW128 B tile straight into the pipeline stage.
//   G_I8      y = (xq[M,K] . wq[N,K]^T) * sx[m] * sw[n]   int8 x int8 -> int32 (kind::i8), the W8A8 of
//             quantize_activation_per_token_absmax x quantize_weight_per_channel_absmax
//             (quantize/fake_quant.py:86-93,109-118) with the dequant scales in the epilogue.
//
// Warp roles (CTA of 256 threads, 512 for G_W4):
//   warp 0  lane 0 : TMA producer          warp 1 lane 0 : MMA issuer
//   warp 2         : TMEM alloc / dealloc  warp 3        : idle
//   warps 4-7      : epilogue (TMEM -> registers -> smem transpose -> coalesced 16-byte stores)
//   warps 8-15     : int4 dequant producers (G_W4 only)
#include "qdm_gemm_dev.cuh"
#include <mutex>
#include <stdlib.h>

bool qdm_gemm_w4a16_skinny_fits(int64_t M, int64_t N, int64_t K);
int qdm_gemm_w4a16_skinny(const void* x, const int32_t* qweight, const int32_t* qzeros, const void* scales, const void* bias,
                          void* y, int is_bf16, int64_t M, int64_t N, int64_t K, int group, cudaStream_t st);
int qdm_w4rp_gemm(const void* x, const void* blob, const void* bias, void* y, int is_bf16, int64_t M, int64_t N, int64_t K,
                  int subs, int sub_n, cudaStream_t st);
int qdm_w4ts_gemm(const void* x, const void* blob, const void* bias, void* y, int is_bf16, int64_t M, int64_t N, int64_t K, int tile_t,
                  cudaStream_t st);
bool qdm_gemm_w4a16_smallm_fits(int64_t M, int64_t N, int64_t K);
int qdm_gemm_w4a16_smallm(const void* x, const int32_t* qweight, const int32_t* qzeros, const void* scales, const void* bias,
                          void* y, int is_bf16, int64_t M, int64_t N, int64_t K, int group, cudaStream_t st);

using namespace qdmg;

namespace {

// int4 dequant producers.  The 8 producer warps form two groups of 128 threads that take alternate k-blocks
// (group g fills pipeline iterations g, g+2, ...): the per-k-block overhead of a warp (cursor arithmetic,
// barrier wait, proxy fence, arrive) is paid half as often and each warp has two MMA periods to hide the
// latency of its dependent instruction chain.  A thread owns one packed word column `wc` (8 output columns) and
// the k rows kr, kr+RPP, ... of the 64-row k-block, and writes the MN-major SW128 B tile of NLOC columns.
// Global loads run DIST of the group's k-blocks ahead of the shared-memory writes and do not stop at tile
// boundaries (the prefetch cursor walks the same (tile, kb) sequence as every other role).  `tile_n` is the tile
// extent in N, `col_off` this CTA's column offset inside the tile (0, or rank * NLOC in a CTA pair).  full
// barriers live at full_addr + 8*stage (a shared::cluster address when CLUSTER), empty barriers at
// empty_addr + 8*stage (always local).
constexpr int DQ_GROUPS = 2;
constexpr int DQ_GROUP_THREADS = 32 * NUM_DQ_WARPS / DQ_GROUPS;   // 128

template <int NLOC, bool BF16, int STAGES, int STAGE_BYTES, bool CLUSTER>
__device__ __forceinline__ void w4_producer_loop(const GemmParams& p, int dt, int lane, int first_tile, int tile_stride,
                                                 int num_tiles, int n_tiles, int num_kb, int tile_n, int col_off, int nloc,
                                                 uint32_t b_stage0, uint32_t empty_addr, uint32_t full_addr) {
  constexpr int WPR = NLOC / 8;                  // packed words per k row of this CTA's tile part
  constexpr int RPP = DQ_GROUP_THREADS / WPR;    // k rows covered by one pass of the group
  constexpr int PASSES = 64 / RPP;
  constexpr int DIST = 2;                        // ring depth - 1; loads run DIST + 1 group-k-blocks ahead
  static_assert(RPP >= 1 && PASSES >= 1 && 64 % RPP == 0, "bad producer tiling");
  const int grp = dt / DQ_GROUP_THREADS, tg = dt % DQ_GROUP_THREADS;
  const int wc = tg % WPR, kr = tg / WPR;
  const int words_per_row = p.N / 8;
  const bool in_tile = wc * 8 < nloc;            // `nloc` <= NLOC columns of the tile part are in use
  const int gshift = 31 - __clz(p.group >> 6);   // group / 64 is a power of two (host-checked)
  // MN-major SW128 tile: 64-column chunk (wc>>3), k row stride 128 B, 16-byte slot (wc&7) ^ (k&7)
  const uint32_t chunk_off = uint32_t(wc >> 3) * (64 * ROW_BYTES);
  auto row_off = [&](int ps) {
    const uint32_t k = uint32_t(kr + ps * RPP);
    return chunk_off + k * ROW_BYTES + ((uint32_t(wc & 7) ^ (k & 7)) << 4);
  };
  struct Pf {            // one prefetched k-block of this thread
    uint32_t w[PASSES];  // packed weight words
    uint32_t zw;         // packed zero points of the group
    uint4 sv;            // 8 scales
  };
  Pf ring[DIST + 1];
  // prefetch cursor of this group: k-block `pf_kb` of tile `pf_tile`, word column pf_wcol
  int pf_tile = first_tile, pf_kb = grp, pf_wcol = 0;
  bool pf_valid = false;
  auto enter_tile = [&]() {
    pf_wcol = ((pf_tile % n_tiles) * tile_n + col_off) / 8 + wc;
    pf_valid = in_tile && pf_tile < num_tiles && pf_wcol < words_per_row;
  };
  auto normalise = [&]() {   // carry pf_kb into pf_tile
    bool moved = false;
    while (pf_kb >= num_kb) { pf_kb -= num_kb; pf_tile += tile_stride; moved = true; }
    if (moved) enter_tile();
  };
  enter_tile();
  normalise();
  auto prefetch = [&](Pf& f) {
    f.zw = 0u;
    f.sv = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) f.w[ps] = 0u;
    if (pf_valid) {
      const int32_t* src = p.qweight + int64_t(pf_kb * 64 + kr) * words_per_row + pf_wcol;
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) f.w[ps] = (uint32_t)__ldg(src + int64_t(ps * RPP) * words_per_row);
      const int64_t g = pf_kb >> gshift;
      f.zw = (uint32_t)__ldg(p.qzeros + g * words_per_row + pf_wcol);
      f.sv = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.scales) + g * p.N + pf_wcol * 8));
    }
    pf_kb += DQ_GROUPS;
    normalise();
  };
#pragma unroll
  for (int d = 0; d < DIST + 1; ++d) prefetch(ring[d]);

  // constants kept in registers so that (x & mask) | magic is ONE lop3
  uint32_t mask_lo = 0x000F000Fu, mask_hi = 0x00F000F0u;
  uint32_t magic = BF16 ? 0x43004300u : 0x64006400u;  // 128.0 / 1024.0: the nibble lands in the low mantissa bits
  asm volatile("" : "+r"(mask_lo), "+r"(mask_hi), "+r"(magic));
  auto and_or = [](uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));  // (a & b) | c
    return d;
  };

  int stage = grp % STAGES;
  uint32_t phase = 0;
  auto process = [&](const Pf& f) {
    const uint32_t sp[4] = {f.sv.x, f.sv.y, f.sv.z, f.sv.w};
    // zero-point operands of the exact (q - z) step, per nibble pair
    uint32_t zsub[4];
    if (BF16) {
#pragma unroll
      for (int q = 0; q < 4; ++q) zsub[q] = and_or(f.zw >> (4 * q), mask_lo, magic);        // 128 + z
    } else {
      const uint32_t zs = f.zw >> 8;
      zsub[0] = and_or(f.zw, mask_lo, magic);                                                // 1024 + z
      zsub[2] = and_or(zs, mask_lo, magic);
      // high nibbles decode as 1024 + 16 z; (.)/16 = 64 + z exactly
      const __half2 sixteenth = __float2half2_rn(0.0625f);
      const uint32_t z1 = and_or(f.zw, mask_hi, magic), z3 = and_or(zs, mask_hi, magic);
      __half2 h1 = __hmul2(*reinterpret_cast<const __half2*>(&z1), sixteenth);
      __half2 h3 = __hmul2(*reinterpret_cast<const __half2*>(&z3), sixteenth);
      zsub[1] = *reinterpret_cast<uint32_t*>(&h1);
      zsub[3] = *reinterpret_cast<uint32_t*>(&h3);
    }
    mbar_wait(empty_addr + 8u * stage, phase ^ 1);
    const uint32_t b_dst = b_stage0 + stage * STAGE_BYTES;
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      if (!in_tile) break;   // columns past the tile are never read by the MMA
      const uint32_t w = f.w[ps];
      uint32_t o[4];
      if (BF16) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t t = and_or(w >> (4 * q), mask_lo, magic);  // {128 + q(col 2q), 128 + q(col 2q+1)}
          __nv_bfloat162 d = __hsub2(*reinterpret_cast<const __nv_bfloat162*>(&t), *reinterpret_cast<const __nv_bfloat162*>(&zsub[q]));
          d = __hmul2(d, *reinterpret_cast<const __nv_bfloat162*>(&sp[q]));
          o[q] = *reinterpret_cast<uint32_t*>(&d);
        }
      } else {
        const uint32_t ws = w >> 8;
        const __half2 sixteenth = __float2half2_rn(0.0625f);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t src = (q < 2) ? w : ws;
          __half2 d;
          if ((q & 1) == 0) {   // low nibble of each byte: 1024 + q, exact subtract
            const uint32_t t = and_or(src, mask_lo, magic);
            d = __hsub2(*reinterpret_cast<const __half2*>(&t), *reinterpret_cast<const __half2*>(&zsub[q]));
          } else {              // high nibble: 1024 + 16 q; fma(., 1/16, -(64 + z)) = q - z exactly
            const uint32_t t = and_or(src, mask_hi, magic);
            d = __hfma2(*reinterpret_cast<const __half2*>(&t), sixteenth, __hneg2(*reinterpret_cast<const __half2*>(&zsub[q])));
          }
          d = __hmul2(d, *reinterpret_cast<const __half2*>(&sp[q]));   // (q - z) * s, one rounding
          o[q] = *reinterpret_cast<uint32_t*>(&d);
        }
      }
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(b_dst + row_off(ps)), "r"(o[0]), "r"(o[1]), "r"(o[2]),
                   "r"(o[3])
                   : "memory");
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      if (CLUSTER) mbar_arrive_cluster(full_addr + 8u * stage);
      else mbar_arrive(full_addr + 8u * stage);
    }
    stage += DQ_GROUPS;
    if (stage >= STAGES) { stage -= STAGES; phase ^= 1; }
  };
  // The ring is indexed with compile-time constants (loop unrolled by DIST + 1): rotating it with register
  // copies would make every iteration wait for the loads it has just issued.
  const int my_tiles = first_tile < num_tiles ? (num_tiles - first_tile + tile_stride - 1) / tile_stride : 0;
  const int total = my_tiles * num_kb;                       // pipeline iterations of this CTA
  const int mine = (total - grp + DQ_GROUPS - 1) / DQ_GROUPS;  // ... of which this group fills `mine`
  for (int it = 0; it < mine; it += DIST + 1) {
#pragma unroll
    for (int u = 0; u < DIST + 1; ++u) {
      if (it + u < mine) {
        // loads are issued AFTER the proxy fence of this k-block: the fence waits for every outstanding memory
        // operation of the thread, prefetches included (measured as long-scoreboard stalls on FENCE.VIEW.ASYNC)
        process(ring[u]);
        prefetch(ring[u]);   // ring[u] is free again: refill it with k-block it + u + DIST + 1
      }
    }
  }
}

// int4 dequant from the TMA-staged raw ring (see RawCfg).  Same thread mapping and arithmetic as w4_producer_loop,
// but the packed words, scales and zero points come from shared memory: nothing global is outstanding at the proxy
// fence, so a k-block costs its own instructions instead of a memory round trip (measured with the loads in this
// loop: 0.63 us per k-block on latency-bound shapes).  `total` = pipeline iterations of this CTA.
template <int NLOC, bool BF16, int STAGES, int STAGE_BYTES, int RAW_N, int GROUPS, bool CLUSTER>
__device__ __forceinline__ void w4_dequant_loop(int dt, int lane, int first_tile, int tile_stride, int num_tiles, int n_tiles,
                                                int num_kb, int tile_n, int col_off, int nloc, int group, uint32_t b_stage0, uint32_t raw0,
                                                uint32_t empty_addr, uint32_t full_addr, uint32_t raw_full_addr,
                                                uint32_t raw_empty_addr, long long* trace, int sk_u0 = -1, int sk_u1 = 0) {
  // stream-K (sk_u0 >= 0): the CTA works on the k-block units [sk_u0, sk_u1) of the tile-major unit sequence, segment by
  // segment from the LAST tile to the first (see qdm_gemm2_sk_kernel); first_tile / tile_stride / num_tiles are unused.
  using R = RawCfg<NLOC>;
  TRC_DECL;
  constexpr int WPR = NLOC / 8;                  // packed words per k row of this CTA's tile part
  constexpr int GROUP_THREADS = 32 * NUM_DQ_WARPS / GROUPS;   // GROUPS groups take k-blocks round robin: their
                                                               // wait / fence / arrive latencies overlap GROUPS-fold
  constexpr int RPP = GROUP_THREADS / WPR;       // k rows covered by one pass of the group
  constexpr int PASSES = 64 / RPP;
  static_assert(RPP >= 1 && PASSES >= 1 && 64 % RPP == 0, "bad producer tiling");
  const int grp = dt / GROUP_THREADS, tg = dt % GROUP_THREADS;
  const int wc = tg % WPR, kr = tg / WPR;
  const bool in_tile = wc * 8 < nloc;            // `nloc` <= NLOC columns of the tile part are in use
  const uint32_t chunk_off = uint32_t(wc >> 3) * (64 * ROW_BYTES);
  auto row_off = [&](int ps) {
    const uint32_t k = uint32_t(kr + ps * RPP);
    return chunk_off + k * ROW_BYTES + ((uint32_t(wc & 7) ^ (k & 7)) << 4);
  };
  uint32_t mask_lo = 0x000F000Fu, mask_hi = 0x00F000F0u;
  uint32_t magic = BF16 ? 0x43004300u : 0x64006400u;  // 128.0 / 1024.0: the nibble lands in the low mantissa bits
  asm volatile("" : "+r"(mask_lo), "+r"(mask_hi), "+r"(magic));
  auto and_or = [](uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));  // (a & b) | c
    return d;
  };
  auto lds32 = [](uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
  };
  static_assert((RAW_N & (RAW_N - 1)) == 0, "raw ring depth must be a power of two");
  int stage = grp % STAGES;
  uint32_t phase = uint32_t(grp / STAGES) & 1u;                // more groups than stages: the first use may be a second lap
  const bool sk = sk_u0 >= 0;
  const int my_tiles = first_tile < num_tiles ? (num_tiles - first_tile + tile_stride - 1) / tile_stride : 0;
  const int total = sk ? sk_u1 - sk_u0 : my_tiles * num_kb;    // pipeline iterations of this CTA
  const int rs_per_tile = (num_kb + 1) >> 1;                   // raw stages (k-block pairs) per tile
  // stream-K geometry: first / last tile of the range, k-block bounds inside them, raw stages of the last tile's segment
  const int sk_t0 = sk ? sk_u0 / num_kb : 0, sk_ka0 = sk ? sk_u0 - sk_t0 * num_kb : 0;
  const int sk_t1 = sk ? (sk_u1 - 1) / num_kb : 0, sk_ke1 = sk ? (sk_u1 - 1) - sk_t1 * num_kb + 1 : 0;
  int tile = first_tile, kb = grp, tl = 0;                     // position of iteration `it`: tile, k-block, local tile count
  for (int it = grp; it < total; it += GROUPS) {
    int rseq;
    bool lone;
    if (sk) {
      // iteration `it` counts k-blocks over the segments in REVERSE tile order (k ascending inside a segment)
      // segment lengths: last tile first
      int rem = it, t = sk_t1, ka, ke, rbase = 0;
      for (;;) {
        ka = (t == sk_t0) ? sk_ka0 : 0;
        ke = (t == sk_t1) ? sk_ke1 : num_kb;
        if (rem < ke - ka) break;
        rem -= ke - ka;
        rbase += ((ke - 1) >> 1) - (ka >> 1) + 1;
        --t;
      }
      tile = t;
      kb = ka + rem;
      rseq = rbase + (kb >> 1) - (ka >> 1);
      const int partner = kb ^ 1;
      lone = !(partner >= ka && partner < ke);
    } else {
      while (kb >= num_kb) { kb -= num_kb; tile += tile_stride; ++tl; }
      rseq = tl * rs_per_tile + (kb >> 1);                     // raw stage sequence number of this CTA
      lone = (kb == num_kb - 1) && (kb & 1) == 0;              // odd K tail: this group is the stage's only consumer
    }
    // word offset of this tile part inside the (aligned-down) staged box
    const uint32_t shift = uint32_t(((tile % n_tiles) * tile_n + col_off) >> 3) & 3u;
    const uint32_t half = uint32_t(kb & 1);                    // which k-block of the raw stage
    const uint32_t srow = (group == 64) ? half : 0u;           // quantisation group row inside the stage
    const int rs = rseq & (RAW_N - 1);
    const uint32_t rphase = uint32_t(rseq / RAW_N) & 1u;
    const uint32_t w_off = (uint32_t((kr + half * 64) * R::WPRX + wc) + shift) * 4u;
    kb += GROUPS;
    const uint32_t raw = raw0 + uint32_t(rs) * R::BYTES;
    mbar_wait(raw_full_addr + 8u * rs, rphase);
    if (tg == 0) TRC(trace, 3 + grp, 1000000 + it);
    uint32_t w[PASSES];
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) w[ps] = lds32(raw + w_off + uint32_t(ps * RPP * R::WPRX) * 4u);
    const uint32_t zw = lds32(raw + R::QW_BYTES + R::SC_BYTES + srow * R::ZW_ROW_BYTES + (uint32_t(wc) + shift) * 4u);
    uint32_t sp[4];
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(sp[0]), "=r"(sp[1]), "=r"(sp[2]), "=r"(sp[3])
                 : "r"(raw + R::QW_BYTES + srow * R::SC_ROW_BYTES + uint32_t(wc) * 16u));
    uint32_t zsub[4];
    if (BF16) {
#pragma unroll
      for (int q = 0; q < 4; ++q) zsub[q] = and_or(zw >> (4 * q), mask_lo, magic);        // 128 + z
    } else {
      const uint32_t zs = zw >> 8;
      zsub[0] = and_or(zw, mask_lo, magic);                                                // 1024 + z
      zsub[2] = and_or(zs, mask_lo, magic);
      const __half2 sixteenth = __float2half2_rn(0.0625f);   // high nibbles decode as 1024 + 16 z; /16 = 64 + z exactly
      const uint32_t z1 = and_or(zw, mask_hi, magic), z3 = and_or(zs, mask_hi, magic);
      __half2 h1 = __hmul2(*reinterpret_cast<const __half2*>(&z1), sixteenth);
      __half2 h3 = __hmul2(*reinterpret_cast<const __half2*>(&z3), sixteenth);
      zsub[1] = *reinterpret_cast<uint32_t*>(&h1);
      zsub[3] = *reinterpret_cast<uint32_t*>(&h3);
    }
    mbar_wait(empty_addr + 8u * stage, phase ^ 1);
    if (tg == 0) TRC(trace, 3 + grp, 2000000 + it);
    const uint32_t b_dst = b_stage0 + stage * STAGE_BYTES;
    if (in_tile) {   // columns past the tile are never read by the MMA
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) {
        const uint32_t wv = w[ps];
        uint32_t o[4];
        if (BF16) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t t = and_or(wv >> (4 * q), mask_lo, magic);  // {128 + q(col 2q), 128 + q(col 2q+1)}
            __nv_bfloat162 d = __hsub2(*reinterpret_cast<const __nv_bfloat162*>(&t), *reinterpret_cast<const __nv_bfloat162*>(&zsub[q]));
            d = __hmul2(d, *reinterpret_cast<const __nv_bfloat162*>(&sp[q]));
            o[q] = *reinterpret_cast<uint32_t*>(&d);
          }
        } else {
          const uint32_t ws = wv >> 8;
          const __half2 sixteenth = __float2half2_rn(0.0625f);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t src = (q < 2) ? wv : ws;
            __half2 d;
            if ((q & 1) == 0) {   // low nibble of each byte: 1024 + q, exact subtract
              const uint32_t t = and_or(src, mask_lo, magic);
              d = __hsub2(*reinterpret_cast<const __half2*>(&t), *reinterpret_cast<const __half2*>(&zsub[q]));
            } else {              // high nibble: 1024 + 16 q; fma(., 1/16, -(64 + z)) = q - z exactly
              const uint32_t t = and_or(src, mask_hi, magic);
              d = __hfma2(*reinterpret_cast<const __half2*>(&t), sixteenth, __hneg2(*reinterpret_cast<const __half2*>(&zsub[q])));
            }
            d = __hmul2(d, *reinterpret_cast<const __half2*>(&sp[q]));   // (q - z) * s, one rounding
            o[q] = *reinterpret_cast<uint32_t*>(&d);
          }
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(b_dst + row_off(ps)), "r"(o[0]), "r"(o[1]), "r"(o[2]),
                     "r"(o[3])
                     : "memory");
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      if (CLUSTER) mbar_arrive_cluster(full_addr + 8u * stage);
      else mbar_arrive(full_addr + 8u * stage);
      mbar_arrive_n(raw_empty_addr + 8u * rs, lone ? 2u : 1u);   // stands in for the absent second k-block's warps
    }
    if (tg == 0) TRC(trace, 3 + grp, 3000000 + it);
    stage += GROUPS;
    while (stage >= STAGES) { stage -= STAGES; phase ^= 1; }
  }
}

// ---------------------------------------------------------------- the kernel
// RAWT (G_W4 only): packed words / scales / zero points arrive by TMA (map_b = qweight, map_s, map_z) in the raw
// ring; otherwise (N % 32 != 0: strides TMA cannot express) the dequant warps load them from global memory.
template <int BLOCK_N, int KIND, bool BF16, bool RAWT>
__global__ void __launch_bounds__(Cfg<BLOCK_N, KIND>::THREADS, 1)
qdm_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ CUtensorMap map_s, const __grid_constant__ CUtensorMap map_z,
                const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_y16, const GemmParams p) {
  using C = Cfg<BLOCK_N, KIND>;
  using R = RawCfg<BLOCK_N>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // layout: [stages x (A | B)] [epilogue: 4 store stagings (1024-aligned), 4 fp32 vectors] [raw int4 ring] [barriers] [tmem ptr]
  const uint32_t epi_base = smem_base + STAGES * C::STAGE_BYTES;
  const uint32_t raw_base = epi_base + C::EPI_BYTES;
  const uint32_t bar_base = raw_base + C::RAW_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  auto raw_full_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 4 + s); };
  auto raw_empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 4 + RAW_STAGES + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(
      smem_gen + STAGES * C::STAGE_BYTES + C::EPI_BYTES + C::RAW_BYTES + 8 * (2 * STAGES + 4 + 2 * RAW_STAGES));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int tile_n = p.tile_n;
  const int n_tiles = (p.N + tile_n - 1) / tile_n;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (p.K + C::K_PER_BLOCK - 1) / C::K_PER_BLOCK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    if (KIND != G_W4 || RAWT) tma_prefetch_desc(&map_b);
    if (KIND == G_W4 && RAWT) { tma_prefetch_desc(&map_s); tma_prefetch_desc(&map_z); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), (KIND == G_W4 && RAWT) ? C::FULL_COUNT_RAW : C::FULL_COUNT);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), 128);
    }
    for (int s = 0; s < RAW_STAGES; ++s) {
      mbar_init(raw_full_bar(s), 1);
      mbar_init(raw_empty_bar(s), C::RAW_EMPTY_COUNT);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(const_cast<uint32_t*>(tmem_ptr_smem))),
                 "n"(C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();   // everything above is on-chip set-up; from here on global memory of earlier kernels is read

  if (warp == 0) {
    // ===================================================== TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * BLOCK_M, n0 = (tile % n_tiles) * tile_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t a_dst = smem_base + stage * C::STAGE_BYTES;
          const uint32_t b_dst = a_dst + A_STAGE_BYTES;
          const int kc = kb * C::K_PER_BLOCK;
          int ka = kc, ma = m0;   // A coordinates: shifted rows / per-tap channel offset for the implicit-GEMM convolution
          int tap = 0;
          if (p.conv_cin) {
            tap = kc / p.conv_cin;
            ka = kc - tap * p.conv_cin;
            ma = m0 + p.conv_off[tap];
          }
          auto load_a = [&]() {
            if (p.conv_w) {
              const int hrow = m0 / p.conv_w;
              tma_load_4d(a_dst, &map_a, full_bar(stage), ka, tap % 3 - 1, (hrow % p.conv_h) * p.conv_stride + tap / 3 - 1, hrow / p.conv_h);
            } else {
              tma_load_2d(a_dst, &map_a, full_bar(stage), ka, ma);
            }
          };
          if (KIND == G_W4) {
            mbar_expect_tx(full_bar(stage), A_STAGE_BYTES);
            load_a();
          } else {
            // B bytes: tile_n rows of 128 B (K-major box), or ceil(tile_n / 64) boxes of 64 x 64 (MN-major)
            const int kn_chunks = (tile_n + 63) / 64;
            mbar_expect_tx(full_bar(stage), A_STAGE_BYTES + (KIND == G_F16_KN ? kn_chunks * 64 : tile_n) * ROW_BYTES);
            load_a();
            if (KIND == G_F16_KN) {
              for (int c = 0; c < kn_chunks; ++c)
                tma_load_2d(b_dst + c * (64 * ROW_BYTES), &map_b, full_bar(stage), n0 + c * 64, kc);
            } else {
              tma_load_2d(b_dst, &map_b, full_bar(stage), kc, n0);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (elect_one()) {
      constexpr bool B_MN = (KIND == G_F16_KN || KIND == G_W4);
      const uint32_t idesc = (KIND == G_I8) ? make_idesc(2, 1, 0, BLOCK_M, tile_n)
                                            : make_idesc(1, BF16 ? 1 : 0, B_MN ? 1 : 0, BLOCK_M, tile_n);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      uint32_t ready = 0;   // did the peek inside the previous k-block already see this stage full?
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (!ready) mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * C::STAGE_BYTES;
          const uint32_t b_addr = a_addr + A_STAGE_BYTES;
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == STAGES) { nstage = 0; nphase ^= 1; }
          const uint64_t da = make_smem_desc(a_addr, 16, 1024);   // 4 x 32 bytes of K per 128-byte row inside the block
          const uint64_t db = B_MN ? make_smem_desc(b_addr, 64 * ROW_BYTES, 1024) : make_smem_desc(b_addr, 16, 1024);
          ready = issue_kblock_ss<KIND, false>(tmem_c, da, db, B_MN ? 128u : 2u, idesc, uint32_t(kb != 0), full_bar(nstage), nphase);
          umma_commit(empty_bar(stage));
          if (kb == num_kb - 1) umma_commit(tmem_full_bar(acc));
          stage = nstage; phase = nphase;
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ===================================================== raw int4 producer: one stage = two k-blocks of packed B
    if (KIND == G_W4 && RAWT && elect_one()) {
      const int srows = p.group == 64 ? 2 : 1;                    // quantisation groups per 128 k rows
      const int gdiv = p.group < 128 ? 128 / p.group : 1, gmul = p.group > 128 ? p.group / 128 : 1;
      const uint32_t tx = R::tx_bytes(srows);
      int rs = 0;
      uint32_t rphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n0 = (tile % n_tiles) * tile_n;
        for (int j = 0; 2 * j < num_kb; ++j) {
          mbar_wait(raw_empty_bar(rs), rphase ^ 1);
          const uint32_t raw = raw_base + uint32_t(rs) * R::BYTES;
          const int grow = j * gdiv / gmul;                       // first group row of k rows [128 j, 128 j + 128)
          mbar_expect_tx(raw_full_bar(rs), tx);
          tma_load_2d(raw, &map_b, raw_full_bar(rs), (n0 >> 3) & ~3, j * 128);
          tma_load_2d(raw + R::QW_BYTES, &map_s, raw_full_bar(rs), n0, grow);
          tma_load_2d(raw + R::QW_BYTES + R::SC_BYTES, &map_z, raw_full_bar(rs), (n0 >> 3) & ~3, grow);
          if (++rs == C::RAW_N) { rs = 0; rphase ^= 1; }
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================================================== epilogue
    const int ew = warp - 4;  // TMEM lane quarter this warp may read (warp % 4)
    const uint32_t stg = epi_base + ew * EPI_STG_BYTES;
    float* vec_sm = reinterpret_cast<float*>(smem_gen + STAGES * C::STAGE_BYTES + 4 * EPI_STG_BYTES) + ew * (EPI_VEC_BYTES / 4) * (KIND == G_I8 ? 2 : 1);
    if (lane == 0) { tma_prefetch_desc(&map_y); tma_prefetch_desc(&map_y16); }
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * BLOCK_M, n0 = (tile % n_tiles) * tile_n;
      mbar_wait(tmem_full_bar(acc), acc_phase);
      tc_fence_after();
      // the CTA's last tile is shared with the (by then idle) dequant warps: chunks 0, 3, ... stay here
      const bool last = KIND == G_W4 && tile + int(gridDim.x) >= num_tiles;
      epilogue_drain<BLOCK_N, KIND, BF16>(p, &map_y, &map_y16, stg, vec_sm, tmem_base + (uint32_t(ew * 32) << 16) + acc * BLOCK_N,
                                          m0 + ew * 32, n0, lane, 0, last ? 3 : 1);
      tc_fence_before();
      mbar_arrive(tmem_empty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();
  } else if (KIND == G_W4 && warp >= 8) {
    // ===================================================== int4 dequant producers
    if (RAWT) {
      w4_dequant_loop<BLOCK_N, BF16, STAGES, C::STAGE_BYTES, C::RAW_N, C::RAW_GROUPS, false>(
          threadIdx.x - 256, lane, blockIdx.x, gridDim.x, num_tiles, n_tiles, num_kb, tile_n, 0, tile_n, p.group,
          smem_base + A_STAGE_BYTES, raw_base, empty_bar(0), full_bar(0), raw_full_bar(0), raw_empty_bar(0), p.trace);
    } else {
      w4_producer_loop<BLOCK_N, BF16, STAGES, C::STAGE_BYTES, false>(
          p, threadIdx.x - 256, lane, blockIdx.x, gridDim.x, num_tiles, n_tiles, num_kb, tile_n, 0, tile_n,
          smem_base + A_STAGE_BYTES, bar_base + 8u * STAGES, bar_base);
    }
    // ---- help drain the CTA's last tile: two more warp sets (TMEM lane quarter = warp % 4) take chunks 1, 4, ... and
    // 2, 5, ...; staging lives in the pipeline stages, which are idle once the last accumulator is complete
    const int my_tiles = int(blockIdx.x) < num_tiles ? (num_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x) : 0;
    if (my_tiles > 0) {
      const int lt = my_tiles - 1, tile = int(blockIdx.x) + lt * int(gridDim.x);
      const int acc = lt & 1, dw = warp - 8, ew = dw & 3, set = 1 + (dw >> 2);
      const int m0 = (tile / n_tiles) * BLOCK_M, n0 = (tile % n_tiles) * tile_n;
      mbar_wait(tmem_full_bar(acc), uint32_t(lt >> 1) & 1u);
      tc_fence_after();
      const uint32_t stg = smem_base + A_STAGE_BYTES + uint32_t(dw) * EPI_STG_BYTES;                    // B part of stage 0
      float* vec_sm = reinterpret_cast<float*>(smem_gen + C::STAGE_BYTES + A_STAGE_BYTES) + dw * 256;   // B part of stage 1
      epilogue_drain<BLOCK_N, KIND, BF16>(p, &map_y, &map_y16, stg, vec_sm, tmem_base + (uint32_t(ew * 32) << 16) + acc * BLOCK_N,
                                          m0 + ew * 32, n0, lane, set, 3);
      if (lane == 0) tma_store_wait_all();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- the CTA-pair kernel (cta_group::2)
// Two CTAs of one cluster (one TPC) work on a 256 x BLOCK_N tile: CTA r holds A rows [128 r, 128 r + 128) and
// the B columns [NLOC r, NLOC r + NLOC) (NLOC = BLOCK_N / 2) of every k-block in ITS shared memory; the leader
// (rank 0) issues tcgen05.mma.cta_group::2 with M = 256 and each CTA's tensor core accumulates its 128 rows x
// BLOCK_N columns in its own TMEM.  Per SM this halves the B operand -- and with it the int4 dequant work and
// the shared-memory traffic -- per flop.  Barriers: `full` lives in the leader (TMA bytes of both CTAs and
// the dequant warps of both CTAs arrive there), `empty`/`tmem_full` are signalled in both CTAs by multicast
// commits, `tmem_empty` lives in the leader and collects the epilogue warps of both CTAs.
template <int BLOCK_N, int KIND>
struct Cfg2 {
  static constexpr int NLOC = BLOCK_N / 2;
  static constexpr int B_STAGE_BYTES = NLOC * ROW_BYTES;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int EPI_BYTES = 4 * (EPI_STG_BYTES + EPI_VEC_BYTES * (KIND == G_I8 ? 2 : 1));
  static constexpr int RAW_N = 4;
  static constexpr int RAW_BYTES = (KIND == G_W4) ? RAW_N * RawCfg<NLOC>::BYTES : 0;
  static constexpr int STAGES = (227 * 1024 - 2048 - EPI_BYTES - RAW_BYTES) / STAGE_BYTES > 8 ? 8 : (227 * 1024 - 2048 - EPI_BYTES - RAW_BYTES) / STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + RAW_BYTES + EPI_BYTES + 1024 + 256;
  static constexpr int THREADS = (KIND == G_W4) ? 512 : 256;
  static constexpr int TMEM_COLS = (2 * BLOCK_N <= 256) ? 256 : 512;
  static constexpr int K_PER_BLOCK = (KIND == G_I8) ? 128 : 64;
  // leader's arrive.expect_tx (covers the TMA bytes of both CTAs) + the dequant warps of both CTAs
  static constexpr int FULL_COUNT = (KIND == G_W4) ? 1 + NUM_DQ_WARPS : 1;   // one producer group per CTA (global-load path)
  static constexpr int RAW_GROUPS = 4;
  static constexpr int FULL_COUNT_RAW = 1 + 2 * NUM_DQ_WARPS / RAW_GROUPS;    // one group per CTA, two CTAs
  static constexpr int RAW_EMPTY_COUNT = 2 * NUM_DQ_WARPS / RAW_GROUPS;
};

// QUAD: clusters of FOUR CTAs = two pairs that work on the same 256 rows and on adjacent n-tiles.  The A operand is the
// same for both pairs, so each CTA fetches only 64 of its 128 A rows and multicasts them to its counterpart in the other
// pair: half the TMA rows per k-block per SM (the TMA engine serves ~1 row of 128 B per 3.3 cycles and was the limiter
// of every W4 shape).  A stage may be refilled once BOTH pairs have consumed it (empty barriers count two commits).
template <int BLOCK_N, int KIND, bool BF16, bool RAWT, bool QUAD>
__global__ void __cluster_dims__(QUAD ? 4 : 2, 1, 1) __launch_bounds__(Cfg2<BLOCK_N, KIND>::THREADS, 1)
qdm_gemm2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_s, const __grid_constant__ CUtensorMap map_z,
                 const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_y16, const GemmParams p) {
  using C = Cfg2<BLOCK_N, KIND>;
  constexpr int STAGES = C::STAGES;
  constexpr int NLOC = C::NLOC;
  using R = RawCfg<NLOC>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t epi_base = smem_base + STAGES * C::STAGE_BYTES;
  const uint32_t raw_base = epi_base + C::EPI_BYTES;
  const uint32_t bar_base = raw_base + C::RAW_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  auto raw_full_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 4 + s); };
  auto raw_empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 4 + RAW_STAGES + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(
      smem_gen + STAGES * C::STAGE_BYTES + C::EPI_BYTES + C::RAW_BYTES + 8 * (2 * STAGES + 4 + 2 * RAW_STAGES));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();            // rank in the cluster: 0..1 (pair) or 0..3 (quad)
  const uint32_t rank = crank & 1u;                    // role inside the CTA pair: 0 = leader
  const int pgrp = QUAD ? int(crank >> 1) : 0;         // which pair of the quad
  const uint32_t leader_crank = crank & ~1u;
  const int m_tiles = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int tile_n = p.tile_n, nloc = p.tile_n / 2;   // this CTA holds `nloc` of the tile's B columns
  // tile sequence of this pair: tiles pair, pair + num_pairs, ... of an [m_tiles, n_tiles] grid, n fastest.  In a quad the
  // two pairs take the even / odd n-tiles of the same m-tile; n_tiles is rounded up to even (a tile past N loads zeros
  // and stores nothing).
  const int n_tiles_real = (p.N + tile_n - 1) / tile_n;
  const int n_tiles = QUAD ? (n_tiles_real + 1) & ~1 : n_tiles_real;
  const int num_tiles = m_tiles * n_tiles;
  const int cl = int(blockIdx.x) / (QUAD ? 4 : 2), num_cl = int(gridDim.x) / (QUAD ? 4 : 2);
  const int pair = QUAD ? 2 * cl + pgrp : cl, num_pairs = QUAD ? 2 * num_cl : num_cl;
  const int num_kb = (p.K + C::K_PER_BLOCK - 1) / C::K_PER_BLOCK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    if (KIND != G_W4 || RAWT) tma_prefetch_desc(&map_b);
    if (KIND == G_W4 && RAWT) { tma_prefetch_desc(&map_s); tma_prefetch_desc(&map_z); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), (KIND == G_W4 && RAWT) ? C::FULL_COUNT_RAW : C::FULL_COUNT);
      mbar_init(empty_bar(s), QUAD ? 2 : 1);   // quad: both pairs' MMA issuers commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), 8);   // 4 epilogue warps x 2 CTAs
    }
    for (int s = 0; s < RAW_STAGES; ++s) {
      mbar_init(raw_full_bar(s), 1);
      mbar_init(raw_empty_bar(s), C::RAW_EMPTY_COUNT);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(const_cast<uint32_t*>(tmem_ptr_smem))),
                 "n"(C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();   // barrier inits and TMEM allocation of BOTH CTAs are visible before anything remote happens
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t leader_full0 = mapa_shared(full_bar(0), leader_crank);
  const uint32_t leader_tmem_empty0 = mapa_shared(tmem_empty_bar(0), leader_crank);
  const uint16_t pair_mask = uint16_t(3u << (2 * pgrp)), all_mask = QUAD ? uint16_t(0xF) : uint16_t(0x3);

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs; each loads its own halves)
    if (elect_one()) {
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      TRC_DECL;
      int trc_it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m0 = (tile / n_tiles) * (2 * BLOCK_M) + int(rank) * BLOCK_M;
        const int n0 = (tile % n_tiles) * tile_n + int(rank) * nloc;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          TRC(p.trace, 0, 2000000 + trc_it);
          ++trc_it;
          const uint32_t a_dst = smem_base + stage * C::STAGE_BYTES;
          const uint32_t b_dst = a_dst + A_STAGE_BYTES;
          const uint32_t lf = leader_full0 + 8u * stage;
          const int kc = kb * C::K_PER_BLOCK;
          // the peer's TMA bytes land on the leader's barrier too; the peer itself does not arrive (its loads of
          // phase n+1 cannot start before its `empty` barrier says the leader consumed phase n)
          if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * (A_STAGE_BYTES + (KIND == G_W4 ? 0 : nloc * ROW_BYTES)));
          if (QUAD) {   // my 64 of the 128 rows, to me and to my counterpart in the other pair
            tma_load_2d_pair_mc(a_dst + uint32_t(pgrp) * (64 * ROW_BYTES), &map_a, lf, kc, m0 + pgrp * 64,
                                uint16_t((1u << rank) | (1u << (rank + 2))));
          } else if (p.conv_cin) {   // implicit-GEMM convolution: tap-shifted rows, channel offset inside the tap
            const int tap = kc / p.conv_cin;
            if (p.conv_w) {
              const int hrow = m0 / p.conv_w;
              tma_load_4d_pair(a_dst, &map_a, lf, kc - tap * p.conv_cin, tap % 3 - 1, (hrow % p.conv_h) * p.conv_stride + tap / 3 - 1, hrow / p.conv_h);
            } else {
              tma_load_2d_pair(a_dst, &map_a, lf, kc - tap * p.conv_cin, m0 + p.conv_off[tap]);
            }
          } else {
            tma_load_2d_pair(a_dst, &map_a, lf, kc, m0);
          }
          if (KIND != G_W4) tma_load_2d_pair(b_dst, &map_b, lf, kc, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    if (rank == 0 && elect_one()) {
      constexpr bool B_MN = (KIND == G_W4);
      const uint32_t idesc = (KIND == G_I8) ? make_idesc(2, 1, 0, 2 * BLOCK_M, tile_n)
                                            : make_idesc(1, BF16 ? 1 : 0, B_MN ? 1 : 0, 2 * BLOCK_M, tile_n);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      TRC_DECL;
      int trc_it = 0;
      uint32_t ready = 0;   // did the peek inside the previous k-block already see this stage full?
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1);
        TRC(p.trace, 1, 1000000 + trc_it);
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (!ready) mbar_wait(full_bar(stage), phase);
          TRC(p.trace, 1, (ready ? 2000000 : 9000000) + trc_it);
          ++trc_it;
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * C::STAGE_BYTES;
          const uint32_t b_addr = a_addr + A_STAGE_BYTES;
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == STAGES) { nstage = 0; nphase ^= 1; }
          const uint64_t da = make_smem_desc(a_addr, 16, 1024);
          const uint64_t db = B_MN ? make_smem_desc(b_addr, 64 * ROW_BYTES, 1024) : make_smem_desc(b_addr, 16, 1024);
          // the next stage's barrier is tested inside the block; its result is read after the fourth MMA
          ready = issue_kblock_ss<KIND, true>(tmem_c, da, db, B_MN ? 128u : 2u, idesc, uint32_t(kb != 0), full_bar(nstage), nphase);
          umma_commit_pair(empty_bar(stage), all_mask);
          if (kb == num_kb - 1) umma_commit_pair(tmem_full_bar(acc), pair_mask);
          stage = nstage; phase = nphase;
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ===================================================== raw int4 producer (each CTA: its own packed columns)
    if (KIND == G_W4 && RAWT && elect_one()) {
      const int srows = p.group == 64 ? 2 : 1;
      const int gdiv = p.group < 128 ? 128 / p.group : 1, gmul = p.group > 128 ? p.group / 128 : 1;
      const uint32_t tx = R::tx_bytes(srows);
      int rs = 0;
      uint32_t rphase = 0;
      TRC_DECL;
      int trc_it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int n0 = (tile % n_tiles) * tile_n + int(rank) * nloc;
        for (int j = 0; 2 * j < num_kb; ++j) {
          mbar_wait(raw_empty_bar(rs), rphase ^ 1);
          TRC(p.trace, 7, 1000000 + trc_it);
          ++trc_it;
          const uint32_t raw = raw_base + uint32_t(rs) * R::BYTES;
          const int grow = j * gdiv / gmul;
#ifdef QDM_EXP_NORAW   // timing experiment only (wrong results): what do the raw TMA loads cost?
          (void)raw; (void)grow; (void)tx;
          mbar_arrive(raw_full_bar(rs));
#else
          mbar_expect_tx(raw_full_bar(rs), tx);
          tma_load_2d(raw, &map_b, raw_full_bar(rs), (n0 >> 3) & ~3, j * 128);
          tma_load_2d(raw + R::QW_BYTES, &map_s, raw_full_bar(rs), n0, grow);
          tma_load_2d(raw + R::QW_BYTES + R::SC_BYTES, &map_z, raw_full_bar(rs), (n0 >> 3) & ~3, grow);
#endif
          if (++rs == C::RAW_N) { rs = 0; rphase ^= 1; }
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================================================== epilogue (each CTA drains its own 128 rows)
    pdl_wait();
    const int ew = warp - 4;
    const uint32_t stg = epi_base + ew * EPI_STG_BYTES;
    float* vec_sm = reinterpret_cast<float*>(smem_gen + STAGES * C::STAGE_BYTES + 4 * EPI_STG_BYTES) + ew * (EPI_VEC_BYTES / 4) * (KIND == G_I8 ? 2 : 1);
    if (lane == 0) { tma_prefetch_desc(&map_y); tma_prefetch_desc(&map_y16); }
    int acc = 0;
    uint32_t acc_phase = 0;
    TRC_DECL;
    int trc_it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m0 = (tile / n_tiles) * (2 * BLOCK_M) + int(rank) * BLOCK_M, n0 = (tile % n_tiles) * tile_n;
      mbar_wait(tmem_full_bar(acc), acc_phase);
      if (threadIdx.x == 128) TRC(p.trace, 2, 1000000 + trc_it);
      tc_fence_after();
      const bool last = KIND == G_W4 && tile + num_pairs >= num_tiles;   // shared with the dequant warps, see below
      epilogue_drain<BLOCK_N, KIND, BF16>(p, &map_y, &map_y16, stg, vec_sm, tmem_base + (uint32_t(ew * 32) << 16) + acc * BLOCK_N,
                                          m0 + ew * 32, n0, lane, 0, last ? 3 : 1);
      tc_fence_before();
      __syncwarp();
      if (threadIdx.x == 128) TRC(p.trace, 2, 2000000 + trc_it);
      ++trc_it;
      if (lane == 0) mbar_arrive_cluster(leader_tmem_empty0 + 8u * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();
  } else if (KIND == G_W4 && warp >= 8) {
    // ===================================================== int4 dequant producers (each CTA: its NLOC columns)
    if (RAWT) {
      w4_dequant_loop<NLOC, BF16, STAGES, C::STAGE_BYTES, C::RAW_N, C::RAW_GROUPS, true>(
          threadIdx.x - 256, lane, pair, num_pairs, num_tiles, n_tiles, num_kb, tile_n, int(rank) * nloc, nloc, p.group,
          smem_base + A_STAGE_BYTES, raw_base, empty_bar(0), leader_full0, raw_full_bar(0), raw_empty_bar(0), p.trace);
    } else {
      w4_producer_loop<NLOC, BF16, STAGES, C::STAGE_BYTES, true>(
          p, threadIdx.x - 256, lane, pair, num_pairs, num_tiles, n_tiles, num_kb, tile_n, int(rank) * nloc, nloc,
          smem_base + A_STAGE_BYTES, bar_base + 8u * STAGES, leader_full0);
    }
    // ---- help drain the pair's last tile (same scheme as the single-CTA kernel): the exposed tail of every launch
    const int my_tiles = pair < num_tiles ? (num_tiles - pair + num_pairs - 1) / num_pairs : 0;
    if (my_tiles > 0) {
      const int lt = my_tiles - 1, tile = pair + lt * num_pairs;
      const int acc = lt & 1, dw = warp - 8, ew = dw & 3, set = 1 + (dw >> 2);
      const int m0 = (tile / n_tiles) * (2 * BLOCK_M) + int(rank) * BLOCK_M, n0 = (tile % n_tiles) * tile_n;
      mbar_wait(tmem_full_bar(acc), uint32_t(lt >> 1) & 1u);
      tc_fence_after();
      // staging: B parts (16 KB each) of stages 0 and 1, fp32 vectors: B part of stage 2
      const uint32_t stg = smem_base + uint32_t(dw >> 2) * C::STAGE_BYTES + A_STAGE_BYTES + uint32_t(dw & 3) * EPI_STG_BYTES;
      float* vec_sm = reinterpret_cast<float*>(smem_gen + 2 * C::STAGE_BYTES + A_STAGE_BYTES) + dw * 256;
      epilogue_drain<BLOCK_N, KIND, BF16>(p, &map_y, &map_y16, stg, vec_sm, tmem_base + (uint32_t(ew * 32) << 16) + acc * BLOCK_N,
                                          m0 + ew * 32, n0, lane, set, 3);
      if (lane == 0) tma_store_wait_all();
    }
  }

  tc_fence_before();
  // execution-only rendezvous (the peer may still be reading this CTA's operands / signalling its barriers): a
  // relaxed arrive, so the epilogue warps do not sit in a MEMBAR.ALL.GPU behind their output stores
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- stream-K CTA-pair kernel (W4)
// The classic kernel gives every pair whole tiles: 80 tiles on 74 pairs cost two waves, and a 25-tile problem with a long
// K runs on 25 pairs.  Here the work is the tile-major sequence of k-block units (tiles x K/64) and pair p takes the units
// [p U / P, (p + 1) U / P): at most one tile is shared with the previous pair and one with the next.  A pair walks its
// segments from its LAST tile to its first:
//   * a segment that ends before the tile does (always the first one processed) leaves its fp32 accumulator in the pair's
//     workspace slot and raises a flag per epilogue warp;
//   * the segment that contains the tile's last k-block (processed last if it is a partial one) waits for the flags of the
//     pairs that hold the tile's earlier parts, adds their partials in slot order (deterministic) and runs the normal epilogue.
// Producers therefore publish early and consumers read late: nobody waits for long, and waits only go to lower pair
// indices, so there is no cycle.  The flags are reset by their reader, so the protocol also holds under CUDA-graph replay.
__device__ __forceinline__ void sk_wait_flag(const uint32_t* flag) {
  uint32_t v = 0, polls = 0;
  long long t0 = 0;
  for (;;) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v != 0) break;
    if (++polls == 4096) t0 = clock64();
    if (polls > 4096 && (polls & 1023) == 0 && clock64() - t0 > 6000000000LL) __trap();
  }
}

template <bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg2<256, G_W4>::THREADS, 1)
qdm_gemm2_sk_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_s, const __grid_constant__ CUtensorMap map_z,
                    const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_y16, const GemmParams p) {
  using C = Cfg2<256, G_W4>;
  constexpr int STAGES = C::STAGES;
  constexpr int NLOC = C::NLOC;
  constexpr int BLOCK_N = 256;
  using R = RawCfg<NLOC>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t epi_base = smem_base + STAGES * C::STAGE_BYTES;
  const uint32_t raw_base = epi_base + C::EPI_BYTES;
  const uint32_t bar_base = raw_base + C::RAW_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  auto raw_full_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 4 + s); };
  auto raw_empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 4 + RAW_STAGES + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(
      smem_gen + STAGES * C::STAGE_BYTES + C::EPI_BYTES + C::RAW_BYTES + 8 * (2 * STAGES + 4 + 2 * RAW_STAGES));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int m_tiles = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int tile_n = p.tile_n, nloc = p.tile_n / 2;
  const int n_tiles = (p.N + tile_n - 1) / tile_n;
  const int num_kb = p.K / 64;
  const long long units = (long long)m_tiles * n_tiles * num_kb;
  const int u0 = int(units * pair / num_pairs), u1 = int(units * (pair + 1) / num_pairs);
  // segments of [u0, u1) from the last tile to the first; f(tile, ka, ke): k-blocks [ka, ke) of `tile`
  auto for_each_seg = [&](auto&& f) {
    int u = u1;
    while (u > u0) {
      const int tile = (u - 1) / num_kb;
      const int ke = (u - 1) - tile * num_kb + 1;
      const int ka = (u - u0 >= ke) ? 0 : ke - (u - u0);
      f(tile, ka, ke);
      u -= ke - ka;
    }
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b); tma_prefetch_desc(&map_s); tma_prefetch_desc(&map_z);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), C::FULL_COUNT_RAW); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), 8); }
    for (int s = 0; s < RAW_STAGES; ++s) { mbar_init(raw_full_bar(s), 1); mbar_init(raw_empty_bar(s), C::RAW_EMPTY_COUNT); }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(const_cast<uint32_t*>(tmem_ptr_smem))),
                 "n"(C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  // only the roles that touch memory of EARLIER kernels wait (TMA producer: activations; epilogue: output buffer, W8A8 row
  // scales, stream-K partials); packed weights / scales / zeros are constants and are fetched while the previous kernel drains
  const uint32_t leader_full0 = mapa_shared(full_bar(0), 0);
  const uint32_t leader_tmem_empty0 = mapa_shared(tmem_empty_bar(0), 0);

  if (warp == 0) {
    // ===================================================== TMA producer: A
    if (elect_one()) {
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      for_each_seg([&](int tile, int ka, int ke) {
        const int m0 = (tile / n_tiles) * (2 * BLOCK_M) + int(rank) * BLOCK_M;
        for (int kb = ka; kb < ke; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * A_STAGE_BYTES);
          tma_load_2d_pair(smem_base + stage * C::STAGE_BYTES, &map_a, leader_full0 + 8u * stage, kb * 64, m0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      });
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = make_idesc(1, BF16 ? 1 : 0, 1, 2 * BLOCK_M, tile_n);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0, ready = 0;
      for_each_seg([&](int, int ka, int ke) {
        mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + acc * BLOCK_N;
        for (int kb = ka; kb < ke; ++kb) {
          if (!ready) mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * C::STAGE_BYTES;
          const uint32_t b_addr = a_addr + A_STAGE_BYTES;
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == STAGES) { nstage = 0; nphase ^= 1; }
          ready = issue_kblock_ss<G_W4, true>(tmem_c, make_smem_desc(a_addr, 16, 1024), make_smem_desc(b_addr, 64 * ROW_BYTES, 1024), 128u,
                                              idesc, uint32_t(kb != ka), full_bar(nstage), nphase);
          umma_commit_pair(empty_bar(stage), 3);
          if (kb == ke - 1) umma_commit_pair(tmem_full_bar(acc), 3);
          stage = nstage; phase = nphase;
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      });
    }
  } else if (warp == 3) {
    // ===================================================== raw int4 producer
    if (elect_one()) {
      const int srows = p.group == 64 ? 2 : 1;
      const int gdiv = p.group < 128 ? 128 / p.group : 1, gmul = p.group > 128 ? p.group / 128 : 1;
      const uint32_t tx = R::tx_bytes(srows);
      int rs = 0;
      uint32_t rphase = 0;
      for_each_seg([&](int tile, int ka, int ke) {
        const int n0 = (tile % n_tiles) * tile_n + int(rank) * nloc;
        for (int j = ka >> 1; j <= (ke - 1) >> 1; ++j) {
          mbar_wait(raw_empty_bar(rs), rphase ^ 1);
          const uint32_t raw = raw_base + uint32_t(rs) * R::BYTES;
          const int grow = j * gdiv / gmul;
          mbar_expect_tx(raw_full_bar(rs), tx);
          tma_load_2d(raw, &map_b, raw_full_bar(rs), (n0 >> 3) & ~3, j * 128);
          tma_load_2d(raw + R::QW_BYTES, &map_s, raw_full_bar(rs), n0, grow);
          tma_load_2d(raw + R::QW_BYTES + R::SC_BYTES, &map_z, raw_full_bar(rs), (n0 >> 3) & ~3, grow);
          if (++rs == C::RAW_N) { rs = 0; rphase ^= 1; }
        }
      });
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================================================== epilogue
    pdl_wait();
    const int ew = warp - 4;
    const uint32_t stg = epi_base + ew * EPI_STG_BYTES;
    float* vec_sm = reinterpret_cast<float*>(smem_gen + STAGES * C::STAGE_BYTES + 4 * EPI_STG_BYTES) + ew * (EPI_VEC_BYTES / 4);
    if (lane == 0) { tma_prefetch_desc(&map_y); tma_prefetch_desc(&map_y16); }
    constexpr int64_t kSlot = 2 * 128 * 256;                 // floats per pair slot
    const int64_t my_row = int64_t(rank) * (128 * 64) + ew * 32 + lane;   // in 16-byte units: [rank][h][i][row]
    int acc = 0;
    uint32_t acc_phase = 0;
    for_each_seg([&](int tile, int ka, int ke) {
      const int m0 = (tile / n_tiles) * (2 * BLOCK_M) + int(rank) * BLOCK_M, n0 = (tile % n_tiles) * tile_n;
      const uint32_t taddr0 = tmem_base + (uint32_t(ew * 32) << 16) + acc * BLOCK_N;
      mbar_wait(tmem_full_bar(acc), acc_phase);
      tc_fence_after();
#ifdef QDM_EXP_SK_NOFIX   // timing experiment only (wrong results): stream-K without parking / adding partials
      if (false) {
#else
      if (ke < num_kb) {
#endif
        // ---- not the tile's last part: park the fp32 accumulator in this pair's slot and raise the flag
        // slot layout [rank][half h][16-byte chunk i][128 rows]: a warp's store / load of chunk i covers 512 contiguous bytes
        uint4* dst = reinterpret_cast<uint4*>(p.sk_data + int64_t(pair) * kSlot) + my_row;
        const int n_halves = (min(tile_n, p.N - n0) + 31) / 32;
        for (int h = 0; h < n_halves; ++h) {
          uint32_t v[32];
          tmem_ld32(taddr0 + h * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            __stcg(dst + (h * 8 + i) * 128, make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.sk_flags + (pair * 2 + int(rank)) * 4 + ew), "r"(1u) : "memory");
      } else {
        // ---- the tile's last part: earlier parts (if any) sit in the slots of the pairs pa .. pair - 1
        int n_parts = 0;
#if defined(QDM_EXP_SK_NOFIX) || defined(QDM_EXP_SK_NOREDUCE)
        if (false) {
#else
        if (ka > 0) {
#endif
          const long long x0 = (long long)tile * num_kb;       // first unit of the tile; its owner is pa
          int pa = int(x0 * num_pairs / units);
          while (units * (pa + 1) / num_pairs <= x0) ++pa;
          while (units * pa / num_pairs > x0) --pa;
          n_parts = pair - pa;
          if (lane == 0)
            for (int q = pa; q < pair; ++q) sk_wait_flag(p.sk_flags + (q * 2 + int(rank)) * 4 + ew);
          __syncwarp();
        }
        epilogue_drain<BLOCK_N, G_W4, BF16>(p, &map_y, &map_y16, stg, vec_sm, taddr0, m0 + ew * 32, n0, lane, 0, 1,
                                            p.sk_data + int64_t(pair - n_parts) * kSlot + my_row * 4, n_parts, kSlot);
        if (n_parts > 0) {   // every flag has exactly one reader warp: hand it back empty for the next launch / replay
          __syncwarp();
          if (lane == 0)
            for (int q = pair - n_parts; q < pair; ++q) p.sk_flags[(q * 2 + int(rank)) * 4 + ew] = 0u;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tmem_empty0 + 8u * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    });
    if (lane == 0) tma_store_wait_all();
  } else if (warp >= 8) {
    // ===================================================== int4 dequant
    w4_dequant_loop<NLOC, BF16, STAGES, C::STAGE_BYTES, C::RAW_N, C::RAW_GROUPS, true>(
        threadIdx.x - 256, lane, 0, 1, 0, n_tiles, num_kb, tile_n, int(rank) * nloc, nloc, p.group,
        smem_base + A_STAGE_BYTES, raw_base, empty_bar(0), leader_full0, raw_full_bar(0), raw_empty_bar(0), p.trace, u0, u1);
  }

  tc_fence_before();
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- B-stationary CTA-pair kernel (W4, K <= 320)
// For the K = 320 layers of the UNet (M = 65536 tokens: 40 of the step's 184 launches, 37 % of its time) a pair's B tile
// part is tiny: K x nloc fp16 = 80 KB.  The generic kernel re-fetches and re-dequantises it for every one of the pair's
// ~7-35 tiles and pays packed-operand TMA rows, dequant latency and a B stage per k-block for it.  Here every pair keeps
// ONE n-tile (grid = n_tiles x floor(74 / n_tiles) pairs, so tile % n_tiles is constant per pair), dequantises its
// K x nloc part once into shared memory, and then only streams A: 5 stages of 16 KB, one TMA box per k-block.  The
// dequant warps become a second epilogue warp set after that (chunks alternate between the two sets), because with 5
// k-blocks per tile the epilogue is as long as the main loop.
struct CfgBS {
  static constexpr int NLOC = 128;
  // 6 A stages + 5 resident k-blocks (K <= 320: every small-K layer of the three denoisers) instead of 5 + 6 (K <= 384): the
  // kernel is bound by the bytes of A in flight per SM x the L2 latency, measured 65536 x 2560 x 320 94.9 -> 90.8 us,
  // 65536 x 960 x 320 43.2 -> 41.7, 65536 x 320 x 320 24.1 -> 23.1
#ifndef QDM_BS_ASTAGES
#define QDM_BS_ASTAGES 6
#endif
#ifndef QDM_BS_KB
#define QDM_BS_KB 5
#endif
  static constexpr int A_STAGES = QDM_BS_ASTAGES;
  static constexpr int BST_KB = QDM_BS_KB;                // resident k-blocks: K <= 320
  static constexpr int B_KB_BYTES = NLOC * ROW_BYTES;     // 16 KB per k-block
  static constexpr int RAW_N = 2;                         // raw ring, used once; afterwards staging of epilogue set 1
  static constexpr int RAW_BYTES = RAW_N * RawCfg<NLOC>::BYTES;
  static constexpr int EPI_BYTES = 4 * (EPI_STG_BYTES + EPI_VEC_BYTES);
  static_assert(RAW_BYTES >= EPI_BYTES, "epilogue set 1 stages in the raw ring");
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_BYTES = A_STAGES * A_STAGE_BYTES + BST_KB * B_KB_BYTES + EPI_BYTES + RAW_BYTES + 1024 + BAR_BYTES;
  static constexpr int THREADS = 512;
  static constexpr int TMEM_COLS = 512;
  static constexpr int GROUPS = 4;
};

template <bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CfgBS::THREADS, 1)
qdm_gemm2_bstat_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const __grid_constant__ CUtensorMap map_s, const __grid_constant__ CUtensorMap map_z,
                       const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_y16, const GemmParams p) {
  using C = CfgBS;
  using R = RawCfg<C::NLOC>;
  constexpr int SA = C::A_STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // layout: [A stages] [resident B: BST_KB x 16 KB] [epilogue set 0] [raw ring / epilogue set 1] [barriers] [tmem ptr]
  const uint32_t bst_base = smem_base + SA * A_STAGE_BYTES;
  const uint32_t epi_base = bst_base + C::BST_KB * C::B_KB_BYTES;
  const uint32_t raw_base = epi_base + C::EPI_BYTES;
  const uint32_t bar_base = raw_base + C::RAW_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (SA + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * SA + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * SA + 2 + a); };
  auto raw_full_bar = [&](int s) { return bar_base + 8u * (2 * SA + 4 + s); };
  auto raw_empty_bar = [&](int s) { return bar_base + 8u * (2 * SA + 4 + RAW_STAGES + s); };
  auto bst_full_bar = [&](int kb) { return bar_base + 8u * (2 * SA + 4 + 2 * RAW_STAGES + kb); };
  auto bst_free_bar = [&](int kb) { return bar_base + 8u * (2 * SA + 4 + 2 * RAW_STAGES + C::BST_KB + kb); };   // never completes
  constexpr int kNumBars = 2 * SA + 4 + 2 * RAW_STAGES + 2 * C::BST_KB;
  static_assert(8 * kNumBars + 8 <= C::BAR_BYTES, "barrier area");
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + (bar_base - smem_base) + 8 * kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;   // num_pairs % n_tiles == 0 (host): fixed n-tile per pair
  const int m_tiles = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int tile_n = p.tile_n, nloc = p.tile_n / 2;
  const int n_tiles = (p.N + tile_n - 1) / tile_n;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = p.K / 64;
  const int n_idx = pair % n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b); tma_prefetch_desc(&map_s); tma_prefetch_desc(&map_z);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), 16); }   // 2 sets x 4 warps x 2 CTAs
    for (int s = 0; s < RAW_STAGES; ++s) { mbar_init(raw_full_bar(s), 1); mbar_init(raw_empty_bar(s), 2 * NUM_DQ_WARPS / C::GROUPS); }
    for (int kb = 0; kb < C::BST_KB; ++kb) { mbar_init(bst_full_bar(kb), 2 * NUM_DQ_WARPS / C::GROUPS); mbar_init(bst_free_bar(kb), 1); }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(const_cast<uint32_t*>(tmem_ptr_smem))),
                 "n"(C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  // only the roles that touch memory of EARLIER kernels wait (TMA producer: activations; epilogue: output buffer, W8A8 row
  // scales, stream-K partials); packed weights / scales / zeros are constants and are fetched while the previous kernel drains
  const uint32_t leader_full0 = mapa_shared(full_bar(0), 0);
  const uint32_t leader_tmem_empty0 = mapa_shared(tmem_empty_bar(0), 0);
  const uint32_t leader_bst_full0 = mapa_shared(bst_full_bar(0), 0);

  // one epilogue warp set: TMEM lane quarter = warp % 4, chunks `set`, set + 2, ...
  auto epilogue_role = [&](int set, int ew, uint32_t stg, float* vec_sm) {
    pdl_wait();
    if (lane == 0) { tma_prefetch_desc(&map_y); tma_prefetch_desc(&map_y16); }
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m0 = (tile / n_tiles) * (2 * BLOCK_M) + int(rank) * BLOCK_M, n0 = (tile % n_tiles) * tile_n;
      mbar_wait(tmem_full_bar(acc), acc_phase);
      tc_fence_after();
      epilogue_drain<256, G_W4, BF16>(p, &map_y, &map_y16, stg, vec_sm, tmem_base + (uint32_t(ew * 32) << 16) + acc * 256,
                                      m0 + ew * 32, n0, lane, set, 2);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tmem_empty0 + 8u * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();
  };

  if (warp == 0) {
    // ===================================================== TMA producer: A only
    if (elect_one()) {
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m0 = (tile / n_tiles) * (2 * BLOCK_M) + int(rank) * BLOCK_M;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * A_STAGE_BYTES);
          tma_load_2d_pair(smem_base + stage * A_STAGE_BYTES, &map_a, leader_full0 + 8u * stage, kb * 64, m0);
          if (++stage == SA) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = make_idesc(1, BF16 ? 1 : 0, 1, 2 * BLOCK_M, tile_n);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0, ready = 0;
      bool first = true;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + acc * 256;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (first) mbar_wait(bst_full_bar(kb), 0);   // the resident B part of this k-block has been written (both CTAs)
          if (!ready) mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * A_STAGE_BYTES;
          const uint32_t b_addr = bst_base + kb * C::B_KB_BYTES;
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == SA) { nstage = 0; nphase ^= 1; }
          ready = issue_kblock_ss<G_W4, true>(tmem_c, make_smem_desc(a_addr, 16, 1024), make_smem_desc(b_addr, 64 * ROW_BYTES, 1024), 128u,
                                              idesc, uint32_t(kb != 0), full_bar(nstage), nphase);
          umma_commit_pair(empty_bar(stage), 3);
          if (kb == num_kb - 1) umma_commit_pair(tmem_full_bar(acc), 3);
          stage = nstage; phase = nphase;
        }
        first = false;
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ===================================================== raw int4 producer: this pair's n-tile, once
    if (pair < num_tiles && elect_one()) {
      const int srows = p.group == 64 ? 2 : 1;
      const int gdiv = p.group < 128 ? 128 / p.group : 1, gmul = p.group > 128 ? p.group / 128 : 1;
      const uint32_t tx = R::tx_bytes(srows);
      const int n0 = n_idx * tile_n + int(rank) * nloc;
      int rs = 0;
      uint32_t rphase = 0;
      for (int j = 0; 2 * j < num_kb; ++j) {
        mbar_wait(raw_empty_bar(rs), rphase ^ 1);
        const uint32_t raw = raw_base + uint32_t(rs) * R::BYTES;
        const int grow = j * gdiv / gmul;
        mbar_expect_tx(raw_full_bar(rs), tx);
        tma_load_2d(raw, &map_b, raw_full_bar(rs), (n0 >> 3) & ~3, j * 128);
        tma_load_2d(raw + R::QW_BYTES, &map_s, raw_full_bar(rs), n0, grow);
        tma_load_2d(raw + R::QW_BYTES + R::SC_BYTES, &map_z, raw_full_bar(rs), (n0 >> 3) & ~3, grow);
        if (++rs == C::RAW_N) { rs = 0; rphase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================================================== epilogue set 0
    const int ew = warp - 4;
    epilogue_role(0, ew, epi_base + ew * EPI_STG_BYTES,
                  reinterpret_cast<float*>(smem_gen + (epi_base - smem_base) + 4 * EPI_STG_BYTES) + ew * (EPI_VEC_BYTES / 4));
  } else if (warp >= 8) {
    // ===================================================== dequant once, then epilogue set 1 (warps 8-11)
    if (pair < num_tiles) {
      // one "tile" whose n index is n_idx: stage index == k-block, nothing ever has to be waited for on the B side
      w4_dequant_loop<C::NLOC, BF16, C::BST_KB, C::B_KB_BYTES, C::RAW_N, C::GROUPS, true>(
          threadIdx.x - 256, lane, n_idx, 1 << 20, n_idx + 1, 1 << 20, num_kb, tile_n, int(rank) * nloc, nloc, p.group,
          bst_base, raw_base, bst_free_bar(0), leader_bst_full0, raw_full_bar(0), raw_empty_bar(0), p.trace);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");   // all dequant warps are done with the raw ring: it becomes staging
    if (warp < 12) {
      const int ew = warp - 8;
      epilogue_role(1, ew, raw_base + ew * EPI_STG_BYTES,
                    reinterpret_cast<float*>(smem_gen + (raw_base - smem_base) + 4 * EPI_STG_BYTES) + ew * (EPI_VEC_BYTES / 4));
    }
  }

  tc_fence_before();
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- host side
}  // namespace

namespace qdmg {
PFN_cuTensorMapEncodeTiled g_encode = nullptr;
std::mutex g_encode_mu;

int get_encode_fn() {
  std::lock_guard<std::mutex> lk(g_encode_mu);
  if (g_encode) return QDM_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    qdm_set_error("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
    return QDM_ERR_CUDA;
  }
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
  return QDM_OK;
}

// 2-D row-major tensor [rows, cols] of `elem_bytes` elements; box = {box_cols, box_rows}, 128-byte swizzle.
int make_map(CUtensorMap* map, const void* ptr, int elem_bytes, int64_t rows, int64_t cols, int box_cols, int box_rows,
             bool swizzle128) {
  const CUtensorMapDataType dt = elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                               : elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    qdm_set_error("cuTensorMapEncodeTiled failed (%d) for [%lld, %lld] x %d B", (int)r, (long long)rows, (long long)cols, elem_bytes);
    return QDM_ERR_CUDA;
  }
  return QDM_OK;
}

}  // namespace qdmg

namespace {

// Tile width.  The kernels are built for up to 256 columns per tile; the width actually used is chosen per problem
// so that the persistent grid is not left with a nearly empty last wave (80 tiles on 74 CTA pairs cost two full
// waves).  Model per (tile, k-block), in cycles: tensor pipe 2*tn, shared-memory traffic (A written + read,
// B written + read) 256 + tn for a CTA pair and 256 + 2*tn for a single CTA, plus a fixed per-tile cost.
int g_quad_units = 32;   // resident quad clusters (a cluster must fit in one GPC); refined by the occupancy query at first launch

// mode: 1 single CTA, 2 CTA pair, 4 quad cluster (two pairs on adjacent n-tiles sharing the A loads)
int choose_tile_n(int64_t M, int64_t N, int mode, double* cost_out = nullptr) {
  const bool pair = mode >= 2;
  const int64_t rows = pair ? 2 * BLOCK_M : BLOCK_M;
  const int64_t units = mode == 4 ? g_quad_units : pair ? QDM_NUM_SMS / 2 : QDM_NUM_SMS;
  const int64_t m_tiles = (M + rows - 1) / rows;
  const int n_cap = int((N + 15) / 16 * 16);
  int best = 256;
  double best_cost = 1e300;
  for (int tn = 256; tn >= 32; tn -= 16) {
    if (tn > n_cap && tn != 256) continue;
    const int t = tn > n_cap ? n_cap : tn;
    int64_t n_tiles = (N + t - 1) / t;
    if (mode == 4) n_tiles = (n_tiles + 1) / 2;          // a quad takes two n-tiles at a time
    const int64_t waves = (m_tiles * n_tiles + units - 1) / units;
    // cycles per k-block: tensor pipe ~2 t; operand delivery (TMA rows: A 128 or 64 per CTA + packed rows) -- measured
    // ~600 + t for a pair, ~350 + t expected for a quad
    const double mma = 2.0 * t, feed = mode == 4 ? 350.0 + t : pair ? 256.0 + t : 256.0 + 2.0 * t;
    const double cost = double(waves) * ((mma > feed ? mma : feed) + 64.0);
    if (cost < best_cost * 0.999) { best_cost = cost; best = t; }
  }
  if (cost_out) *cost_out = best_cost;
  return best;
}

// Launch with programmatic stream serialisation (see pdl_wait in the kernels); QDM_NO_PDL=1 launches plainly.
template <typename Kern>
cudaError_t launch_pdl(Kern kern, unsigned grid, unsigned threads, size_t smem, cudaStream_t st, const CUtensorMap& a,
                       const CUtensorMap& b, const CUtensorMap& s, const CUtensorMap& z, const CUtensorMap& y,
                       const CUtensorMap& y16, const struct GemmParams& p) {
  static const bool no_pdl = getenv("QDM_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kern, a, b, s, z, y, y16, p);
}

struct Maps {
  CUtensorMap a, b, s, z;   // b: B operand (or packed qweight), s / z: W4 scales / zero points (raw TMA path)
  CUtensorMap y, y16;       // output [M, N]: 32 x 64 boxes (SWIZZLE_128B) and 32 x 16 slices (dense) for the epilogue's TMA stores
  bool raw = false;
  bool quad = false;        // W4 raw path on quad clusters: the A map has 64-row boxes
};

template <int BLOCK_N, int KIND, bool BF16, bool RAWT>
int launch_gemm(const Maps& m, const GemmParams& p, cudaStream_t st) {
  using C = Cfg<BLOCK_N, KIND>;
  auto kern = qdm_gemm_kernel<BLOCK_N, KIND, BF16, RAWT>;
  static bool attr_set = false;  // per instantiation; benign race (idempotent)
  if (!attr_set) {
    QDM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M, n_tiles = (p.N + p.tile_n - 1) / p.tile_n;
  const int tiles = m_tiles * n_tiles;
  const int grid = tiles < QDM_NUM_SMS ? tiles : QDM_NUM_SMS;
  QDM_CUDA_OK(launch_pdl(kern, (unsigned)grid, C::THREADS, C::SMEM_BYTES, st, m.a, m.b, m.s, m.z, m.y, m.y16, p));
  QDM_LAUNCH_CHECK();
  return QDM_OK;
}

template <int BLOCK_N, int KIND, bool BF16, bool RAWT, bool QUAD>
int launch_gemm2(const Maps& m, const GemmParams& p, cudaStream_t st) {
  using C = Cfg2<BLOCK_N, KIND>;
  auto kern = qdm_gemm2_kernel<BLOCK_N, KIND, BF16, RAWT, QUAD>;
  static bool attr_set = false;
  if (!attr_set) {
    QDM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int m_tiles = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M), n_tiles = (p.N + p.tile_n - 1) / p.tile_n;
  int pairs;
  if (QUAD) {   // whole clusters of two pairs; a cluster must fit inside one GPC, so fewer than 148 / 4 may be resident
    static int max_clusters = 0;
    if (max_clusters == 0) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(QDM_NUM_SMS / 4 * 4);
      cfg.blockDim = dim3(C::THREADS);
      cfg.dynamicSmemBytes = C::SMEM_BYTES;
      cudaLaunchAttribute attr;
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = 4; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = QDM_NUM_SMS / 4 - 5; }
      max_clusters = n < QDM_NUM_SMS / 4 ? n : QDM_NUM_SMS / 4;
      g_quad_units = max_clusters;
      if (getenv("QDM_DEBUG")) fprintf(stderr, "qdm: %d quad clusters resident\n", max_clusters);
    }
    const int st_tiles = m_tiles * ((n_tiles + 1) / 2);
    pairs = 2 * (st_tiles < max_clusters ? st_tiles : max_clusters);
  } else {
    const int tiles = m_tiles * n_tiles;
    pairs = tiles < QDM_NUM_SMS / 2 ? tiles : QDM_NUM_SMS / 2;
  }
#ifdef QDM_TRACE
  if (getenv("QDM_TRACE")) {
    static long long* tbuf = nullptr;
    if (!tbuf) cudaMalloc(&tbuf, 16 * 2048 * sizeof(long long));
    cudaMemset(tbuf, 0, 16 * 2048 * sizeof(long long));
    GemmParams pt = p;
    pt.trace = tbuf;
    kern<<<2 * pairs, C::THREADS, C::SMEM_BYTES, st>>>(m.a, m.b, m.s, m.z, m.y, m.y16, pt);
    cudaDeviceSynchronize();
    static long long host[16 * 2048];
    cudaMemcpy(host, tbuf, sizeof(host), cudaMemcpyDeviceToHost);
    fprintf(stderr, "QDMTRACE begin M=%d N=%d K=%d tile_n=%d\n", p.M, p.N, p.K, p.tile_n);
    for (int r = 0; r < 16; ++r)
      for (int i = 0; i < 1000 && host[r * 2048 + 2 * i]; ++i)
        fprintf(stderr, "QDMTRACE %d %lld %lld\n", r, host[r * 2048 + 2 * i], host[r * 2048 + 2 * i + 1]);
    QDM_LAUNCH_CHECK();
    return QDM_OK;
  }
#endif
  QDM_CUDA_OK(launch_pdl(kern, (unsigned)(2 * pairs), C::THREADS, C::SMEM_BYTES, st, m.a, m.b, m.s, m.z, m.y, m.y16, p));
  QDM_LAUNCH_CHECK();
  return QDM_OK;
}

// stream-K workspace (per device, set by qdm_gemm_set_workspace): [flags: 4 KB][74 pair slots x 2 x 128 x 256 fp32]
constexpr size_t SK_FLAG_BYTES = 4096;
constexpr size_t SK_SLOT_FLOATS = 2 * 128 * 256;
constexpr size_t SK_WS_BYTES = SK_FLAG_BYTES + size_t(QDM_NUM_SMS / 2) * SK_SLOT_FLOATS * sizeof(float);
void* g_sk_ws[64] = {};

template <bool BF16>
int launch_sk(const Maps& m, const GemmParams& p, int pairs, cudaStream_t st) {
  using C = Cfg2<256, G_W4>;
  auto kern = qdm_gemm2_sk_kernel<BF16>;
  static bool attr_set = false;
  if (!attr_set) {
    QDM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  QDM_CUDA_OK(launch_pdl(kern, (unsigned)(2 * pairs), C::THREADS, C::SMEM_BYTES, st, m.a, m.b, m.s, m.z, m.y, m.y16, p));
  QDM_LAUNCH_CHECK();
  return QDM_OK;
}

template <bool BF16>
int launch_bstat(const Maps& m, const GemmParams& p, int pairs, cudaStream_t st) {
  auto kern = qdm_gemm2_bstat_kernel<BF16>;
  static bool attr_set = false;
  if (!attr_set) {
    QDM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgBS::SMEM_BYTES));
    attr_set = true;
  }
  QDM_CUDA_OK(launch_pdl(kern, (unsigned)(2 * pairs), CfgBS::THREADS, CfgBS::SMEM_BYTES, st, m.a, m.b, m.s, m.z, m.y, m.y16, p));
  QDM_LAUNCH_CHECK();
  return QDM_OK;
}

// Bring-up / A-B switches.  The environment is read ONCE (first GEMM call of the process), never in the per-call path.
struct W4Opts {
  bool no_smallm, no_skinny, no_tma, no_bstat, no_sk, no_rp;
  int ts_waves;           // A/B knob of the TS dispatch rule (QDM_W4_TS_WAVES: largest tile count in waves), read once
  W4Opts()
      : no_smallm(getenv("QDM_W4_NO_SMALLM") != nullptr), no_skinny(getenv("QDM_W4_NO_SKINNY") != nullptr),
        no_tma(getenv("QDM_W4_NO_TMA") != nullptr), no_bstat(getenv("QDM_W4_NO_BSTAT") != nullptr),
        no_sk(getenv("QDM_W4_NO_SK") != nullptr), no_rp(getenv("QDM_W4_NO_RP") != nullptr),
        ts_waves(getenv("QDM_W4_TS_WAVES") ? atoi(getenv("QDM_W4_TS_WAVES")) : 1000000) {}
};
const W4Opts& w4_opts() {
  static const W4Opts o;
  return o;
}
// qdm_set_w4_disable: a process-wide OR-mask on top of the environment switches, for tests and A/B timing that must flip a
// kernel family off and on inside one process (the environment is only read once)
int g_w4_disable = 0;
W4Opts w4_opts_now() {
  W4Opts o = w4_opts();
  const int d = g_w4_disable;
  o.no_smallm |= (d & QDM_W4_NO_SMALLM) != 0; o.no_skinny |= (d & QDM_W4_NO_SKINNY) != 0; o.no_tma |= (d & QDM_W4_NO_TMA) != 0;
  o.no_bstat |= (d & QDM_W4_NO_BSTAT) != 0; o.no_sk |= (d & QDM_W4_NO_SK) != 0; o.no_rp |= (d & QDM_W4_NO_RP) != 0;
  return o;
}

// which kernel the last GEMM call of this thread launched (tests assert the dispatch; qdm_gemm_last_variant)
thread_local int g_last_variant = 0, g_last_tile = 0;
inline void note_variant(int v, int tile) { g_last_variant = v; g_last_tile = tile; }

// 0: heuristic, 1: force the single-CTA kernel, 2: force the CTA-pair kernel, 4: force quad clusters where the kernel
// supports them, 8: stream-K, 16 / 32: the repacked-weight kernel with one / two sub-tiles, 64: never the repacked-weight
// kernel (bring-up / A-B timing; test-only, set through qdm_set_gemm_mode)
int g_force_ctas = 0;
int g_force_tile = 0;   // with 16 / 32: sub-tile width override (multiple of 32), 0 = cost model
bool g_no_rp = false;   // mode 64: the heuristics of the non-repacked kernels, as if no blob had been passed

// Repacked-weight kernel (qdm_gemm_w4rp.cu): when is it taken, and with which tile?  Measured (profiles/README.md, round 2):
// reading the packed operand as one bulk copy per tile part instead of three tensor-map loads does not change the
// k-block period (the pipeline is bound by shared-memory traffic: A written + read, B written by the dequant warps + read),
// so the one-sub-tile form ties with the AWQ-tensor kernel; the two-sub-tile form (256 x up to 512 columns, all of TMEM
// as ONE accumulator set) wins exactly when it turns a two-wave problem into ONE wave -- 4096 x 1280 x 1280: 144 tiles of
// 256 x 144 on 74 pairs -> 64 tiles of 256 x 320, 23.2 -> 20.5 us; 4096 x 1280 x 5120: 60.8 (stream-K) -> 53.9 us -- and
// loses when two or more waves remain (its epilogue does not overlap the next tile's main loop).  Forced modes (16 / 32)
// keep a cost model over all widths for A/B timing.
bool choose_rp(int64_t M, int64_t N, int64_t K, int* subs_out, int* sub_n_out) {
  const int64_t P = QDM_NUM_SMS / 2, m_tiles = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M), num_kb = K / 64;
  if (g_force_ctas == 0) {
    const int tn_old = choose_tile_n(M, N, 2);
    if (m_tiles * ((N + tn_old - 1) / tn_old) <= P) return false;          // already one wave of whole tiles
    for (int sn = 32; sn <= 256; sn += 32) {                                // narrowest two-sub-tile width that is one wave
      const int W = 2 * sn;
      if (W - sn >= N) break;
      if (m_tiles * ((N + W - 1) / W) <= P) { *subs_out = 2; *sub_n_out = sn; return true; }
    }
    return false;
  }
  double best = 1e300;
  int bs = 1, bn = 256;
  for (int subs = 1; subs <= 2; ++subs) {
    if (g_force_ctas == 16 && subs != 1) continue;
    if (g_force_ctas == 32 && subs != 2) continue;
    for (int sn = 32; sn <= 256; sn += 32) {
      if (g_force_tile && sn != g_force_tile) continue;
      const int W = subs * sn;
      if (subs == 2 && W - sn >= N) continue;             // the second sub-tile would never hold a column
      const int64_t n_sup = (N + W - 1) / W, tiles = m_tiles * n_sup, waves = (tiles + P - 1) / P;
      const double mma = 2.0 * W, feed = 430.0 + 0.5 * W;
      const double perkb = (mma > feed ? mma : feed) + 60.0;
      const double epi = 500.0 + 4.5 * W;
      const double cost = double(waves) * (double(num_kb) * perkb + (subs == 2 ? epi : 0.0)) + (subs == 1 ? epi : 0.0);
      if (cost < best * 0.999) { best = cost; bs = subs; bn = sn; }
    }
  }
  *subs_out = bs; *sub_n_out = bn;
  return true;
}

bool use_pair(const GemmParams& p) {
  if (g_force_ctas == 1) return false;
  if (g_force_ctas >= 2 && g_force_ctas <= 8) return true;
  if (g_force_ctas == 128) return p.M > BLOCK_M;
  return p.M > BLOCK_M;   // a second 128-row half exists
}

template <int KIND, bool RAWT>
int dispatch_gemm_r(const Maps& m, const GemmParams& p, bool pair, cudaStream_t st) {
  if (pair && m.quad) {
    if (KIND == G_W4 && RAWT)   // the only instantiation built for quad clusters
      return p.is_bf16 ? launch_gemm2<256, KIND, true, RAWT, KIND == G_W4 && RAWT>(m, p, st)
                       : launch_gemm2<256, KIND, false, RAWT, KIND == G_W4 && RAWT>(m, p, st);
  }
  if (pair) return p.is_bf16 ? launch_gemm2<256, KIND, true, RAWT, false>(m, p, st) : launch_gemm2<256, KIND, false, RAWT, false>(m, p, st);
  return p.is_bf16 ? launch_gemm<256, KIND, true, RAWT>(m, p, st) : launch_gemm<256, KIND, false, RAWT>(m, p, st);
}
template <int KIND>
int dispatch_gemm(const Maps& m, const GemmParams& p, bool pair, cudaStream_t st) {
  if (KIND == G_W4 && m.raw) return dispatch_gemm_r<KIND, KIND == G_W4>(m, p, pair, st);
  return dispatch_gemm_r<KIND, false>(m, p, pair, st);
}

int check_common(const char* fn, const void* x, const void* w, void* y, int dtype, int64_t M, int64_t N, int64_t K) {
  QDM_REQUIRE(x && w && y, "%s: null pointer", fn);
  QDM_REQUIRE(M > 0 && N > 0 && K > 0, "%s: empty problem M=%lld N=%lld K=%lld", fn, (long long)M, (long long)N, (long long)K);
  QDM_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), "%s: dimension too large", fn);
  QDM_REQUIRE(dtype == QDM_F16 || dtype == QDM_BF16, "%s: dtype must be f16 or bf16", fn);
  QDM_REQUIRE(N % 8 == 0, "%s: N=%lld must be a multiple of 8", fn, (long long)N);
  QDM_REQUIRE(qdm_aligned16(x) && qdm_aligned16(w) && qdm_aligned16(y), "%s: operands must be 16-byte aligned", fn);
  return QDM_OK;
}

// 3x3 / stride 1 / pad 1 convolution on a zero-padded NHWC grid [B, H+2, W+2, C]: the output pixel stored at padded
// position r = (b*(H+2) + h+1)*(W+2) + w+1 is sum over taps (dy, dx) of x_pad[r + (dy-1)*(W+2) + (dx-1), :] @ W[:, dy, dx, :]^T,
// i.e. one GEMM with M = B*(H+2)*(W+2) rows, K = 9*C, whose A rows are shifted by a constant per tap.  Border rows of
// the output grid are computed too (from wrapped-around neighbours) and are ignored by the caller; rows shifted out
// of [0, M) are zero-filled by TMA.
struct ConvGeom {
  int64_t B, H, W, C;   // INPUT grid
  bool direct;          // unpadded NHWC input behind a 4-D tensor map (see GemmParams::conv_w)
  int stride = 1;       // 1, or 2 (direct form only): output grid H/2 x W/2, output pixel (ho, wo) reads input rows 2*ho + dy - 1
  int64_t Ho() const { return H / stride; }
  int64_t Wo() const { return W / stride; }
};
// a 128-row tile must be a box of whole (output) image rows: W divides 128 and the box is either a divisor of H image
// rows or a whole number of images
bool conv_direct_ok(int64_t H, int64_t W) {
  if (W <= 0 || H <= 0 || W > 128 || 128 % W) return false;
  const int64_t rows = 128 / W;
  return rows <= H ? H % rows == 0 : rows % H == 0;
}
// Stride 2: the box spans stride * (output pixels) input coordinates and the map's element strides (traversal strides)
// make the TMA unit fetch every other pixel of it, so the tile in shared memory is the same 128 rows x 64 channels as at
// stride 1; the box start (2*ho0 + dy - 1, dx - 1) may be -1, which is the zero padding.
int make_map_nhwc(CUtensorMap* map, const void* ptr, const ConvGeom& g) {
  const int64_t st = g.stride, ho = g.Ho(), wo = g.Wo();
  const int64_t rows = 128 / wo;
  const int64_t bh = rows <= ho ? rows : ho, bb = rows <= ho ? 1 : rows / ho;
  cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.B};
  cuuint64_t strides[3] = {(cuuint64_t)g.C * 2, (cuuint64_t)g.W * g.C * 2, (cuuint64_t)g.H * g.W * g.C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(wo * st), (cuuint32_t)(bh * st), (cuuint32_t)bb};
  cuuint32_t estr[4] = {1, (cuuint32_t)st, (cuuint32_t)st, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    qdm_set_error("cuTensorMapEncodeTiled failed (%d) for NHWC [%lld, %lld, %lld, %lld] stride %d", (int)r, (long long)g.B,
                  (long long)g.H, (long long)g.W, (long long)g.C, g.stride);
    return QDM_ERR_CUDA;
  }
  return QDM_OK;
}
int conv_check(const char* fn, const ConvGeom& g, int64_t N) {
  QDM_REQUIRE(g.B > 0 && g.H > 0 && g.W > 0 && g.C > 0 && N > 0, "%s: empty problem", fn);
  QDM_REQUIRE(g.C % 64 == 0, "%s: C=%lld must be a multiple of 64 (one k-block never straddles two taps)", fn, (long long)g.C);
  QDM_REQUIRE(g.B * (g.H + 2) * (g.W + 2) < (1LL << 31) && 9 * g.C < (1LL << 31), "%s: dimension too large", fn);
  QDM_REQUIRE(g.stride == 1 || (g.stride == 2 && g.direct && g.H % 2 == 0 && g.W % 2 == 0),
              "%s: stride 2 needs even H=%lld, W=%lld", fn, (long long)g.H, (long long)g.W);
  if (g.direct && !conv_direct_ok(g.Ho(), g.Wo())) {
    qdm_set_error("%s: output grid %lld x %lld cannot be tiled by whole image rows (its width must divide 128)%s",
                  fn, (long long)g.Ho(), (long long)g.Wo(), g.stride == 1 ? "; use the padded-grid entry" : "");
    return QDM_ERR_UNSUPPORTED;
  }
  return QDM_OK;
}
int64_t conv_rows(const ConvGeom& g) { return g.direct ? g.B * g.Ho() * g.Wo() : g.B * (g.H + 2) * (g.W + 2); }
void conv_fill(GemmParams* p, const ConvGeom& g) {
  p->conv_cin = int(g.C);
  p->conv_stride = g.stride;
  if (g.direct) { p->conv_w = int(g.Wo()); p->conv_h = int(g.Ho()); }
  for (int dy = 0; dy < 3; ++dy)
    for (int dx = 0; dx < 3; ++dx) p->conv_off[dy * 3 + dx] = int((dy - 1) * (g.W + 2) + (dx - 1));
}

int gemm_f16_impl(const char* fn, const void* x, const void* w, const void* bias, void* y, int dtype, int64_t M, int64_t N,
                  int64_t K, const ConvGeom* conv, cudaStream_t stream) {
  int rc = check_common(fn, x, w, y, dtype, M, N, K);
  if (rc) return rc;
  QDM_REQUIRE(K % 8 == 0, "%s: K=%lld must be a multiple of 8", fn, (long long)K);
  QDM_REQUIRE(!bias || qdm_aligned16(bias), "%s: bias must be 16-byte aligned", fn);
  QDM_DEVICE_GATE();
  if ((rc = get_encode_fn())) return rc;
  Maps m;
  if (conv && conv->direct) rc = make_map_nhwc(&m.a, x, *conv);
  else rc = make_map(&m.a, x, 2, M, conv ? conv->C : K, 64, BLOCK_M);
  if (rc) return rc;
  if ((rc = make_map(&m.y, y, 2, M, N, EPI_COLS, 32))) return rc;
  if ((rc = make_map(&m.y16, y, 2, M, N, 16, 32, false))) return rc;
  GemmParams p{};
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.bias = bias; p.y = y; p.is_bf16 = dtype == QDM_BF16;
  if (conv) conv_fill(&p, *conv);
  const bool pair = use_pair(p);
  p.tile_n = choose_tile_n(M, N, pair ? 2 : 1);
  if ((rc = make_map(&m.b, w, 2, N, K, 64, pair ? p.tile_n / 2 : p.tile_n))) return rc;
  m.s = m.a; m.z = m.a;
  return dispatch_gemm<G_F16>(m, p, pair, stream);
}

}  // namespace

extern "C" size_t qdm_gemm_workspace_bytes(void) { return SK_WS_BYTES; }

extern "C" int qdm_gemm_set_workspace(void* workspace, size_t bytes, void* stream) {
  int dev = 0;
  QDM_CUDA_OK(cudaGetDevice(&dev));
  QDM_REQUIRE(dev >= 0 && dev < 64, "qdm_gemm_set_workspace: device index %d", dev);
  if (!workspace) { g_sk_ws[dev] = nullptr; return QDM_OK; }
  QDM_REQUIRE(bytes >= SK_WS_BYTES && qdm_aligned16(workspace), "qdm_gemm_set_workspace: need %zu bytes, 16-byte aligned", SK_WS_BYTES);
  QDM_CUDA_OK(cudaMemsetAsync(workspace, 0, SK_FLAG_BYTES, (cudaStream_t)stream));   // all flags empty
  g_sk_ws[dev] = workspace;
  return QDM_OK;
}

extern "C" int qdm_set_gemm_mode(int ctas) {
  const int mode = ctas & 0xff, tile = ctas >> 8;   // bits 8.. : sub-tile width override of the repacked-weight kernel
  QDM_REQUIRE(mode == 0 || mode == 1 || mode == 2 || mode == 4 || mode == 8 || mode == 16 || mode == 32 || mode == 64 || mode == 128,
              "qdm_set_gemm_mode: 0 (auto), 1 (single CTA), 2 (CTA pair), 4 (quad cluster), 8 (stream-K), 16 / 32 (repacked weights, "
              "one / two sub-tiles) or 64 (no repacked weights)");
  QDM_REQUIRE(tile == 0 || (tile % 16 == 0 && tile <= 384), "qdm_set_gemm_mode: tile width %d must be a multiple of 16 <= 384", tile);
  g_force_ctas = mode == 64 ? 0 : mode;
  g_no_rp = mode == 64;
  g_force_tile = tile;
  return QDM_OK;
}

extern "C" int qdm_set_w4_disable(int mask) {
  const int all = QDM_W4_NO_SMALLM | QDM_W4_NO_SKINNY | QDM_W4_NO_TMA | QDM_W4_NO_BSTAT | QDM_W4_NO_SK | QDM_W4_NO_RP;
  QDM_REQUIRE((mask & ~all) == 0, "qdm_set_w4_disable: unknown bits in mask 0x%x", mask);
  g_w4_disable = mask;
  return QDM_OK;
}

extern "C" int qdm_gemm_f16(const void* x, const void* w, const void* bias, void* y, int dtype,
                            int64_t M, int64_t N, int64_t K, void* stream) {
  return gemm_f16_impl("qdm_gemm_f16", x, w, bias, y, dtype, M, N, K, nullptr, (cudaStream_t)stream);
}

extern "C" int qdm_conv3x3_f16(const void* x_pad, const void* w_tap, const void* bias, void* y_pad, int dtype,
                               int64_t B, int64_t H, int64_t W, int64_t C, int64_t N, void* stream) {
  const ConvGeom g{B, H, W, C, false};
  int rc = conv_check("qdm_conv3x3_f16", g, N);
  if (rc) return rc;
  return gemm_f16_impl("qdm_conv3x3_f16", x_pad, w_tap, bias, y_pad, dtype, conv_rows(g), N, 9 * C, &g, (cudaStream_t)stream);
}

extern "C" int qdm_conv3x3_direct_ok(int64_t H, int64_t W) { return conv_direct_ok(H, W) ? 1 : 0; }

extern "C" int qdm_conv3x3_nhwc_f16(const void* x, const void* w_tap, const void* bias, void* y, int dtype,
                                    int64_t B, int64_t H, int64_t W, int64_t C, int64_t N, void* stream) {
  const ConvGeom g{B, H, W, C, true};
  int rc = conv_check("qdm_conv3x3_nhwc_f16", g, N);
  if (rc) return rc;
  return gemm_f16_impl("qdm_conv3x3_nhwc_f16", x, w_tap, bias, y, dtype, conv_rows(g), N, 9 * C, &g, (cudaStream_t)stream);
}

extern "C" int qdm_conv3x3s2_nhwc_f16(const void* x, const void* w_tap, const void* bias, void* y, int dtype,
                                      int64_t B, int64_t H, int64_t W, int64_t C, int64_t N, void* stream) {
  const ConvGeom g{B, H, W, C, true, 2};
  int rc = conv_check("qdm_conv3x3s2_nhwc_f16", g, N);
  if (rc) return rc;
  return gemm_f16_impl("qdm_conv3x3s2_nhwc_f16", x, w_tap, bias, y, dtype, conv_rows(g), N, 9 * C, &g, (cudaStream_t)stream);
}

extern "C" int qdm_gemm_f16_kn(const void* x, const void* w_kn, const void* bias, void* y, int dtype,
                               int64_t M, int64_t N, int64_t K, void* stream) {
  int rc = check_common("qdm_gemm_f16_kn", x, w_kn, y, dtype, M, N, K);
  if (rc) return rc;
  QDM_REQUIRE(K % 8 == 0, "qdm_gemm_f16_kn: K=%lld must be a multiple of 8", (long long)K);
  QDM_REQUIRE(!bias || qdm_aligned16(bias), "qdm_gemm_f16_kn: bias must be 16-byte aligned");
  QDM_DEVICE_GATE();
  if ((rc = get_encode_fn())) return rc;
  Maps m;
  if ((rc = make_map(&m.a, x, 2, M, K, 64, BLOCK_M))) return rc;
  if ((rc = make_map(&m.y, y, 2, M, N, EPI_COLS, 32))) return rc;
  if ((rc = make_map(&m.y16, y, 2, M, N, 16, 32, false))) return rc;
  if ((rc = make_map(&m.b, w_kn, 2, K, N, 64, 64))) return rc;
  m.s = m.a; m.z = m.a;
  GemmParams p{};
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.bias = bias; p.y = y; p.is_bf16 = dtype == QDM_BF16;
  p.tile_n = choose_tile_n(M, N, 1);
  return dispatch_gemm<G_F16_KN>(m, p, false, (cudaStream_t)stream);
}

// Token-tile width of the TS kernel (qdm_gemm_w4ts.cu).  Cost model fitted to the per-shape measurements of
// profiles/ts_models_wide_r02.txt (cycles at the ~1.65 GHz the SMs run at inside these kernels; it reproduces the measured
// shapes to ~8 %): a K = 128 stage costs 450 + 2.25 T cycles with two alternating tiles of T <= 192 tokens (883 at T = 192
// against 768 cycles of tensor-pipe work; 600 per tile for fill / drain), and 8 Ts + 60 with ONE wide tile of two
// sub-tiles of Ts = T / 2 tokens (sixteen MMAs per stage: tensor-pipe-bound), whose epilogue (1800 + 14 T cycles) is not
// overlapped.  Wide tiles win where they remove a wave on a long main loop (4096 x 1280 x 5120: 41.1 vs 47.6 us); a 5 %
// margin keeps the overlapped form on ties.
int choose_ts_tile(int64_t M, int64_t N, int64_t K, double* cost_out) {
  const int64_t P = QDM_NUM_SMS / 2, n_blks = (N + 255) / 256, num_st = (K / 64 + 1) / 2;
  int best = 192;
  double best_cost = 1e300;
  for (int t = 384; t >= 32; t -= 32) {   // multiples of 32: the epilogue stores 32-token boxes; > 192: two sub-tiles (WIDE)
    if (g_force_tile && t != g_force_tile) continue;
    const bool wide = t > 192;
    if (wide && t % 64 != 0) continue;
    const int64_t tiles = n_blks * ((M + t - 1) / t), waves = (tiles + P - 1) / P;
    double cost = wide ? double(waves) * (double(num_st) * (8.0 * (t / 2) + 60.0) + 1800.0 + 14.0 * t) / 0.95
                       : double(waves) * (double(num_st) * (450.0 + 2.25 * t) + 600.0);
    if (cost < best_cost * 0.999) { best_cost = cost; best = t; }
  }
  if (cost_out) *cost_out = best_cost;
  return best;
}

static int gemm_w4a16_impl(const void* x, const int32_t* qweight, const int32_t* qzeros, const void* scales,
                           const void* bias, void* y, int dtype, int64_t M, int64_t N, int64_t K, int group,
                           const ConvGeom* conv, void* stream, const void* blob = nullptr, const void* blob_ts = nullptr) {
  int rc = check_common("qdm_gemm_w4a16", x, qweight, y, dtype, M, N, K);
  if (rc) return rc;
  QDM_REQUIRE(qzeros && scales, "qdm_gemm_w4a16: null qzeros/scales");
  QDM_REQUIRE(K % 64 == 0, "qdm_gemm_w4a16: K=%lld must be a multiple of 64", (long long)K);
  QDM_REQUIRE(group > 0 && group % 64 == 0 && K % group == 0 && ((group / 64) & (group / 64 - 1)) == 0,
              "qdm_gemm_w4a16: group=%d must be 64 * 2^j and divide K=%lld", group, (long long)K);
  QDM_REQUIRE(qdm_aligned16(scales) && (!bias || qdm_aligned16(bias)), "qdm_gemm_w4a16: scales/bias must be 16-byte aligned");
  QDM_DEVICE_GATE();
  // M <= 32 and small weights: latency bound -> the mma.sync kernel of qdm_gemm_smallm.cu (g_force_ctas keeps the
  // tcgen05 path reachable for A/B timing)
  // M <= 32: weight-bandwidth / latency bound.  Measured (profiles/README.md): the one-word-column kernel of
  // qdm_gemm_smallm.cu is as fast or faster up to ~2 M weights (16 x 1280 x 320: 4.7 vs 8.1 us, 16 x 1280 x 1280: 9.5 vs
  // 9.8 us); above that its sector over-fetch dominates and the sector-wide cluster-split-K kernel of qdm_gemm_skinny.cu
  // wins (1 x 2432 x 2432: 10 vs 21.6 us; 1 x 14592 x 2432: 18.1 vs 32.6 us on the tcgen05 kernel).
  const W4Opts opt = w4_opts_now();
  const bool small_w = N * K <= (int64_t(1) << 21) && qdm_gemm_w4a16_smallm_fits(M, N, K) && !opt.no_smallm;
  if (!conv && !small_w && qdm_gemm_w4a16_skinny_fits(M, N, K) && g_force_ctas == 0 && !opt.no_skinny) {
    note_variant(QDM_GEMM_SKINNY, 0);
    return qdm_gemm_w4a16_skinny(x, qweight, qzeros, scales, bias, y, dtype == QDM_BF16, M, N, K, group, (cudaStream_t)stream);
  }
  if (!conv && qdm_gemm_w4a16_smallm_fits(M, N, K) && g_force_ctas == 0 && !opt.no_smallm) {
    note_variant(QDM_GEMM_SMALLM, 0);
    return qdm_gemm_w4a16_smallm(x, qweight, qzeros, scales, bias, y, dtype == QDM_BF16, M, N, K, group, (cudaStream_t)stream);
  }
  // Weights as the TMEM A operand (qdm_gemm_w4ts.cu).  Measured against the kernels below over every Linear shape of the
  // three denoisers (profiles/ts_models_r02.txt): it wins by 10-45 % on 27 of 31 shapes with M > 32 -- 4096 x 10240 x 1280:
  // 79.4 vs 89.2 us (cuBLAS f16 79.4), 4096 x 2432 x 2432: 41.5 vs 56.3, 8192 x 1280 x 1280: 26.1 vs 33.7, the text-token
  // shapes 333 x 2432 x 9728: 44.6 vs 63.9 and 1232 x 1280 x 768: 9.0 vs 11.3 (any multiple of 32 tokens per tile, a hidden
  // epilogue, no shared-memory round trip of the weights, one barrier test + one commit per K = 128 on the issuing thread).
  // It loses where K <= 320 with many tiles (the B-stationary kernel dequantises each weight once per CTA, not once per
  // tile: 65536 x 320 x 320 25.2 vs 31.8) and where N = 320 leaves most of a 256-channel block empty.
  if (blob_ts && !conv && K >= 128 && M > 32 && (g_force_ctas == 128 || (g_force_ctas == 0 && !g_no_rp && !opt.no_rp))) {
    bool take = g_force_ctas == 128;
    const int t = choose_ts_tile(M, N, K, nullptr);
    if (!take) {
      const int64_t tiles = ((N + 255) / 256) * ((M + t - 1) / t);
      bool bstat = false;
      if (K <= 64 * CfgBS::BST_KB && N % 32 == 0 && M > BLOCK_M && !opt.no_bstat && !opt.no_tma) {
        const int tn = choose_tile_n(M, N, 2);
        const int64_t m_tiles = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M), n_tiles = (N + tn - 1) / tn, max_pairs = QDM_NUM_SMS / 2;
        bstat = n_tiles <= max_pairs && m_tiles * n_tiles >= 3 * max_pairs;
      }
      // 256-channel blocks: N = 320 idles 3/8 and N = 640 1/6 of the MMA rows and dequant warps of the last block; with many
      // token tiles (M >= 16384) the AWQ-tensor kernel's 160- / 224-wide tiles win there (65536 x 320 x 1280: 60.0 vs 74.2 us,
      // 16384 x 640 x 2560: 49.9 vs 54.6, 32768 x 640 x 2560: 89.8 vs 102.5; 16384 x 640 x 640 ties)
      const bool wasteful = ((N + 255) / 256) * 256 * 100 > int64_t(N) * 115 && M >= 16384;
      take = !bstat && !wasteful && tiles <= int64_t(opt.ts_waves) * (QDM_NUM_SMS / 2);
    }
    if (take) {
      note_variant(QDM_GEMM_TS, t);
      return qdm_w4ts_gemm(x, blob_ts, bias, y, dtype == QDM_BF16, M, N, K, t, (cudaStream_t)stream);
    }
  }
  // Repacked weights: every CTA-pair problem except the small-K / many-tile ones the B-stationary kernel keeps
  // (K >= 192: the last-tile helpers' barrier-parity argument needs >= 3 k-blocks per tile, see qdm_gemm_w4rp.cu)
  if (blob && !conv && M > BLOCK_M && K >= 192 && !opt.no_rp && !g_no_rp && (g_force_ctas == 0 || g_force_ctas == 16 || g_force_ctas == 32)) {
    bool bstat = false;
    if (g_force_ctas == 0 && K <= 64 * CfgBS::BST_KB && N % 32 == 0 && !opt.no_bstat && !opt.no_tma) {
      const int tn = choose_tile_n(M, N, 2);
      const int64_t m_tiles = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M), n_tiles = (N + tn - 1) / tn, max_pairs = QDM_NUM_SMS / 2;
      bstat = n_tiles <= max_pairs && m_tiles * n_tiles >= 3 * max_pairs;
    }
    int subs = 1, sub_n = 256;
    if (!bstat && choose_rp(M, N, K, &subs, &sub_n)) {
      note_variant(subs == 2 ? QDM_GEMM_RP2 : QDM_GEMM_RP1, subs * sub_n);
      return qdm_w4rp_gemm(x, blob, bias, y, dtype == QDM_BF16, M, N, K, subs, sub_n, (cudaStream_t)stream);
    }
  }
  if ((rc = get_encode_fn())) return rc;
  Maps m;
  if ((rc = make_map(&m.y, y, 2, M, N, EPI_COLS, 32))) return rc;
  if ((rc = make_map(&m.y16, y, 2, M, N, 16, 32, false))) return rc;
  GemmParams p{};
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.group = group;
  p.qweight = qweight; p.qzeros = qzeros; p.scales = scales; p.bias = bias; p.y = y; p.is_bf16 = dtype == QDM_BF16;
  if (conv) conv_fill(&p, *conv);   // only the whole-tile kernels below know about tap-shifted A rows
  const bool pair = use_pair(p);
  // packed operands by TMA when their row strides are multiples of 16 bytes (N % 32 == 0): boxes are always the
  // full tile part the kernel was built for (columns past the tile are loaded and ignored, past N zero-filled)
  m.raw = (N % 32 == 0) && qdm_aligned16(qzeros) && !opt.no_tma;   // A/B switch for bring-up
  double cost2 = 0, cost4 = 0;
  p.tile_n = choose_tile_n(M, N, pair ? 2 : 1, &cost2);
  if (pair && m.raw && g_force_ctas != 2 && !conv) {   // quad clusters when the cost model prefers them (or when forced)
    const int t4 = choose_tile_n(M, N, 4, &cost4);
    (void)cost4;   // measured: no gain from the shared A loads (profiles/README.md), so quads run only when forced
    if (g_force_ctas == 4) { m.quad = true; p.tile_n = t4; }
  }
  if (conv && conv->direct) rc = make_map_nhwc(&m.a, x, *conv);
  else rc = make_map(&m.a, x, 2, M, conv ? conv->C : K, 64, m.quad ? BLOCK_M / 2 : BLOCK_M);
  if (rc) return rc;
  m.b = m.a; m.s = m.a; m.z = m.a;
  if (m.raw) {
    const int nloc_max = pair ? 128 : 256, G = int(K / group);
    const int srows = group == 64 ? 2 : 1;   // quantisation groups per raw stage (128 k rows)
    if ((rc = make_map(&m.b, qweight, 4, K, N / 8, nloc_max / 8 + 4, 128, false))) return rc;
    if ((rc = make_map(&m.s, scales, 2, G, N, nloc_max, srows, false))) return rc;
    if ((rc = make_map(&m.z, qzeros, 4, G, N / 8, nloc_max / 8 + 4, srows, false))) return rc;
  }
  // B-stationary pair kernel: small K (resident B fits), enough tiles per pair for the one-time dequant to pay off
  if (!conv && pair && m.raw && !m.quad && g_force_ctas == 0 && K <= 64 * CfgBS::BST_KB && !opt.no_bstat) {
    const int64_t m_tiles = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M), n_tiles = (N + p.tile_n - 1) / p.tile_n;
    const int64_t max_pairs = QDM_NUM_SMS / 2;
    if (n_tiles <= max_pairs && m_tiles * n_tiles >= 3 * max_pairs) {
      const int pairs = int(n_tiles * (max_pairs / n_tiles));   // multiple of n_tiles: every pair keeps one n-tile
      note_variant(QDM_GEMM_BSTAT, p.tile_n);
      return p.is_bf16 ? launch_bstat<true>(m, p, pairs, (cudaStream_t)stream) : launch_bstat<false>(m, p, pairs, (cudaStream_t)stream);
    }
  }
  // stream-K pair kernel: when whole-tile waves leave pairs idle (few tiles with a long K, or a nearly empty last wave).
  // Cost model in cycles per k-block of a pair, measured: ~600 + tile_n; parking one partial accumulator and adding one
  // back costs ~15000 cycles (8 us) per launch as measured, so stream-K wins for long-K problems with an awkward tile count
  // (4096 x 1280 x 5120: 57 instead of 71 us) and loses for short ones (4096 x 1280 x 1280: 28.6 vs 23.1 us).
  int dev = 0;
  if (!conv && pair && m.raw && !m.quad && (g_force_ctas == 0 || g_force_ctas == 8) && !opt.no_sk &&
      cudaGetDevice(&dev) == cudaSuccess && dev < 64 && g_sk_ws[dev]) {
    const int64_t P = QDM_NUM_SMS / 2, m_tiles = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M), num_kb = K / 64;
    const int64_t n_t = (N + 255) / 256;
    const int tn = int(((N + n_t - 1) / n_t + 15) / 16 * 16);
    const int64_t units = m_tiles * n_t * num_kb;
    const int64_t pairs = units < P ? units : P;
    const double sk_cost = double((units + pairs - 1) / pairs) * (600.0 + tn) + 15000.0;
    const int64_t c_tiles = m_tiles * ((N + p.tile_n - 1) / p.tile_n);
    const double classic_cost = double((c_tiles + P - 1) / P) * double(num_kb) * (600.0 + p.tile_n);
    // a tile shared by three or more pairs would make its last part wait for several partials: only when every pair has
    // at least one tile's worth of k-blocks
    if (g_force_ctas == 8 || (units >= P * num_kb && sk_cost < 0.92 * classic_cost)) {
      GemmParams ps = p;
      ps.tile_n = tn;
      ps.sk_flags = reinterpret_cast<uint32_t*>(g_sk_ws[dev]);
      ps.sk_data = reinterpret_cast<float*>(reinterpret_cast<char*>(g_sk_ws[dev]) + SK_FLAG_BYTES);
      note_variant(QDM_GEMM_STREAMK, tn);
      return p.is_bf16 ? launch_sk<true>(m, ps, int(pairs), (cudaStream_t)stream) : launch_sk<false>(m, ps, int(pairs), (cudaStream_t)stream);
    }
  }
  note_variant(m.quad ? QDM_GEMM_QUAD : pair ? QDM_GEMM_PAIR : QDM_GEMM_SINGLE, p.tile_n);
  return dispatch_gemm<G_W4>(m, p, pair, (cudaStream_t)stream);
}

extern "C" int qdm_gemm_w4a16(const void* x, const int32_t* qweight, const int32_t* qzeros, const void* scales,
                              const void* bias, void* y, int dtype, int64_t M, int64_t N, int64_t K, int group,
                              void* stream) {
  return gemm_w4a16_impl(x, qweight, qzeros, scales, bias, y, dtype, M, N, K, group, nullptr, stream);
}

extern "C" int qdm_gemm_w4a16_rp(const void* x, const int32_t* qweight, const int32_t* qzeros, const void* scales,
                                 const void* blob, const void* bias, void* y, int dtype, int64_t M, int64_t N, int64_t K,
                                 int group, void* stream) {
  return gemm_w4a16_impl(x, qweight, qzeros, scales, bias, y, dtype, M, N, K, group, nullptr, stream, blob);
}

extern "C" int qdm_gemm_w4a16_plan(const void* x, const int32_t* qweight, const int32_t* qzeros, const void* scales,
                                   const void* blob_rp, const void* blob_ts, const void* bias, void* y, int dtype, int64_t M,
                                   int64_t N, int64_t K, int group, void* stream) {
  return gemm_w4a16_impl(x, qweight, qzeros, scales, bias, y, dtype, M, N, K, group, nullptr, stream, blob_rp, blob_ts);
}

extern "C" int qdm_gemm_last_variant(int* tile_n) {
  if (tile_n) *tile_n = g_last_tile;
  return g_last_variant;
}

extern "C" int qdm_conv3x3_w4a16(const void* x_pad, const int32_t* qweight, const int32_t* qzeros, const void* scales,
                                 const void* bias, void* y_pad, int dtype, int64_t B, int64_t H, int64_t W, int64_t C,
                                 int64_t N, int group, void* stream) {
  const ConvGeom g{B, H, W, C, false};
  int rc = conv_check("qdm_conv3x3_w4a16", g, N);
  if (rc) return rc;
  return gemm_w4a16_impl(x_pad, qweight, qzeros, scales, bias, y_pad, dtype, conv_rows(g), N, 9 * C, group, &g, stream);
}

extern "C" int qdm_conv3x3_nhwc_w4a16(const void* x, const int32_t* qweight, const int32_t* qzeros, const void* scales,
                                      const void* bias, void* y, int dtype, int64_t B, int64_t H, int64_t W, int64_t C,
                                      int64_t N, int group, void* stream) {
  const ConvGeom g{B, H, W, C, true};
  int rc = conv_check("qdm_conv3x3_nhwc_w4a16", g, N);
  if (rc) return rc;
  return gemm_w4a16_impl(x, qweight, qzeros, scales, bias, y, dtype, conv_rows(g), N, 9 * C, group, &g, stream);
}

extern "C" int qdm_conv3x3s2_nhwc_w4a16(const void* x, const int32_t* qweight, const int32_t* qzeros, const void* scales,
                                        const void* bias, void* y, int dtype, int64_t B, int64_t H, int64_t W, int64_t C,
                                        int64_t N, int group, void* stream) {
  const ConvGeom g{B, H, W, C, true, 2};
  int rc = conv_check("qdm_conv3x3s2_nhwc_w4a16", g, N);
  if (rc) return rc;
  return gemm_w4a16_impl(x, qweight, qzeros, scales, bias, y, dtype, conv_rows(g), N, 9 * C, group, &g, stream);
}

extern "C" int qdm_gemm_w8a8(const int8_t* xq, const float* sx, const int8_t* wq, const float* sw,
                             const void* bias, void* y, int out_dtype, int64_t M, int64_t N, int64_t K,
                             void* stream) {
  int rc = check_common("qdm_gemm_w8a8", xq, wq, y, out_dtype, M, N, K);
  if (rc) return rc;
  QDM_REQUIRE(sx && sw, "qdm_gemm_w8a8: null scales");
  QDM_REQUIRE(K % 16 == 0, "qdm_gemm_w8a8: K=%lld must be a multiple of 16", (long long)K);
  QDM_REQUIRE(qdm_aligned16(sw) && (!bias || qdm_aligned16(bias)), "qdm_gemm_w8a8: sw/bias must be 16-byte aligned");
  QDM_DEVICE_GATE();
  if ((rc = get_encode_fn())) return rc;
  Maps m;
  if ((rc = make_map(&m.a, xq, 1, M, K, 128, BLOCK_M))) return rc;
  if ((rc = make_map(&m.y, y, 2, M, N, EPI_COLS, 32))) return rc;
  if ((rc = make_map(&m.y16, y, 2, M, N, 16, 32, false))) return rc;
  GemmParams p{};
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.sx = sx; p.sw = sw; p.bias = bias; p.y = y; p.is_bf16 = out_dtype == QDM_BF16;
  const bool pair = use_pair(p);
  p.tile_n = choose_tile_n(M, N, pair ? 2 : 1);
  if ((rc = make_map(&m.b, wq, 1, N, K, 128, pair ? p.tile_n / 2 : p.tile_n))) return rc;
  m.s = m.a; m.z = m.a;
  return dispatch_gemm<G_I8>(m, p, pair, (cudaStream_t)stream);
}

