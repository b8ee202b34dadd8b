// (c) W4A16 GEMM on REPACKED int4 weights (the "RP" path): y = x[M,K] . dequant(qweight,qzeros,scales)[K,N] + bias.
//
// Why a second storage form.  The AWQ layout (utils/packing_utils.py:4-27: qweight [K, N/8] int32) is what a checkpoint
// holds and what WQLinear_GEMM exposes; it stays untouched.  For the tensor-core kernel it is an awkward source: the
// packed words of one tile part are 64-128 short rows of 40-80 bytes per k-block, and the TMA unit (and the L2 behind
// it) is charged per ROW (~3.3 cycles each, profiles/README.md), so the int4 operand cost as much fabric time as half
// an fp16 operand.  qdm_w4a16_repack() (run once per weight, at module load) rewrites words, scales and zero points as
//     blob[K/128][ceil(N/16)] blocks of 1104 bytes:
//         [128 k rows][2 words]  packed words of 16 output columns        (1024 B)
//         [2][16] dtype          scales of k rows [0,64) and [64,128)      (  64 B)
//         [2][2]  int32          packed zero points, same two halves       (  16 B)
// so that everything a CTA needs for 128 k rows of its tile part is ONE contiguous range: one `cp.async.bulk` per raw
// stage instead of three tensor-map loads, whole 128-byte lines, no per-row overhead.  The block stride (276 words)
// spreads the 8 blocks of a tile part over distinct shared-memory banks for the dequant warps' 4-byte reads.
//
// Why wide tiles.  A CTA pair normally owns a 256 x tile_n (<= 256) tile with double-buffered TMEM accumulators.  The
// A operand (16 KB per k-block per CTA) is the dominant fabric load and is amortised over tile_n columns only; and
// mid-sized layers (4096 x 1280 x 1280: 144 tiles of 256 x 144 on 74 pairs) need two waves.  SUBS = 2 gives a pair TWO
// adjacent sub-tiles (256 x 2*sub_n <= 512 columns, all 512 TMEM columns as one single-buffered accumulator set): one
// A stage feeds 2 x 4 tcgen05.mma per k-block, the per-flop A traffic (L2 -> SM and shared-memory writes) halves, and
// 4096 x 1280 x 1280 becomes 64 tiles = one wave.  The price: the epilogue no longer overlaps the next tile's main
// loop, so the host's cost model only picks SUBS = 2 where that is cheaper (qdm_gemm.cu: choose_rp).
//
// Roles per CTA (16 warps) as in qdm_gemm2_kernel: warp 0 TMA producer of A, warp 1 MMA issuer (leader CTA), warp 2
// TMEM allocator, warp 3 raw producer (bulk copies of blob ranges), warps 4-7 epilogue, warps 8-15 dequant (4 groups
// of 2 warps taking k-blocks round robin).  Dequant arithmetic is the one of qdm_gemm.cu: (q - z) exact in the 16-bit
// type via the magic-number trick, * s with one rounding -- bit-identical to dequantize_gemm (packing_utils.py:87-102).
#include "qdm_gemm_dev.cuh"

using namespace qdmg;

namespace {

constexpr int RP_BLK_COLS = 16;
constexpr int RP_SC_OFF = 1024;
constexpr int RP_ZW_OFF = 1088;
constexpr int RP_BLK_BYTES = 1104;
constexpr int RP_NB_PER_CTA = 8;   // repack kernel: blocks per CTA along N (64 contiguous bytes of every qweight row)

// ---------------------------------------------------------------- one-time repack
__global__ void __launch_bounds__(128) w4rp_repack_kernel(const int32_t* __restrict__ qweight, const int32_t* __restrict__ qzeros,
                                                          const uint16_t* __restrict__ scales, int N, int K, int group, int nb_total,
                                                          uint8_t* __restrict__ blob) {
  const int nb_groups = (nb_total + RP_NB_PER_CTA - 1) / RP_NB_PER_CTA;
  const int kg = blockIdx.x / nb_groups, nb0 = (blockIdx.x % nb_groups) * RP_NB_PER_CTA;
  const int t = threadIdx.x;   // k row inside the 128-row group
  const int words_per_row = N / 8;
  const int k = kg * 128 + t;
  for (int i = 0; i < RP_NB_PER_CTA; ++i) {
    const int nb = nb0 + i;
    if (nb >= nb_total) break;
    uint8_t* out = blob + (size_t(kg) * nb_total + nb) * RP_BLK_BYTES;
    uint2 w = make_uint2(0u, 0u);
    const int wc = nb * 2;
    if (k < K) {
      const int32_t* src = qweight + int64_t(k) * words_per_row + wc;
      if (wc < words_per_row) w.x = uint32_t(src[0]);
      if (wc + 1 < words_per_row) w.y = uint32_t(src[1]);
    }
    *reinterpret_cast<uint2*>(out + t * 8) = w;
    if (t < 32) {   // scales of the two 64-row halves
      const int h = t >> 4, c = t & 15;
      const int kk = kg * 128 + 64 * h, col = nb * RP_BLK_COLS + c;
      uint16_t v = 0;
      if (kk < K && col < N) v = scales[int64_t(kk / group) * N + col];
      reinterpret_cast<uint16_t*>(out + RP_SC_OFF)[h * 16 + c] = v;
    } else if (t < 36) {   // packed zero points of the two halves
      const int j = t - 32, h = j >> 1, w2 = j & 1;
      const int kk = kg * 128 + 64 * h;
      uint32_t v = 0;
      if (kk < K && wc + w2 < words_per_row) v = uint32_t(qzeros[int64_t(kk / group) * words_per_row + wc + w2]);
      reinterpret_cast<uint32_t*>(out + RP_ZW_OFF)[h * 2 + w2] = v;
    }
  }
}

// ---------------------------------------------------------------- kernel configuration
template <int SUBS>
struct CfgRP {
  static constexpr int NLOC = 128;                                  // columns per CTA per sub-tile (sub_n <= 256)
  static constexpr int B_SUB_BYTES = NLOC * ROW_BYTES;              // 16 KB
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + SUBS * B_SUB_BYTES;
  static constexpr int RAW_SUB_BYTES = (NLOC / RP_BLK_COLS) * RP_BLK_BYTES;   // 8 blocks = 8832 B (= 69 * 128)
  static constexpr int RAW_STAGE_BYTES = SUBS * RAW_SUB_BYTES;      // one raw stage = 128 k rows of every sub-tile part
  static constexpr int RAW_N = SUBS == 1 ? 4 : 3;
  static constexpr int RAW_BYTES = RAW_N * RAW_STAGE_BYTES;
  static constexpr int EPI_BYTES = 4 * (EPI_STG_BYTES + EPI_VEC_BYTES);
  static constexpr int STAGES_RAW = (227 * 1024 - 2048 - EPI_BYTES - RAW_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + RAW_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int THREADS = 512;
  static constexpr int TMEM_COLS = 512;
  static constexpr int ACCS = SUBS == 1 ? 2 : 1;                    // accumulator sets: double buffered only for one sub-tile
  // Dequant groups take k-blocks round robin.  A group only ever waits on the barriers of ITS k-blocks, so it skips the
  // other groups' phases of a stage; an mbarrier parity wait can tell adjacent phases apart only, which is safe as long
  // as a group cannot run two laps ahead of the consumer: STAGES >= GROUPS for the pipeline stages (arriving at k-block
  // kb, the MMA has consumed kb - GROUPS - STAGES; the phase that could be mistaken is kb - 2 STAGES), and every group
  // a consumer of every raw stage (two k-blocks) or of every RAW_N-th one.  One sub-tile: 5 stages, 4 groups of 2 warps,
  // raw ring of 4 (groups 0/1 and 2/3 alternate stages).  Two sub-tiles: 3 stages -> 2 groups of 4 warps (even / odd
  // k-blocks), so that both groups read every raw stage.
  static constexpr int GROUPS = SUBS == 1 ? 4 : 2;
  static constexpr int FULL_COUNT = 1 + 2 * NUM_DQ_WARPS / GROUPS;  // leader's expect_tx + one group per CTA, two CTAs
  static constexpr int RAW_EMPTY_COUNT = 2 * NUM_DQ_WARPS / GROUPS; // the warps of the stage's two k-blocks
  static_assert(STAGES >= 3, "the last-tile helpers stage in the B parts of pipeline stages 0..2");
  static_assert(STAGES >= GROUPS && (GROUPS == 2 || RAW_N % 2 == 0), "parity waits: see GROUPS");
  static_assert(RAW_SUB_BYTES % 128 == 0, "raw sub-stages stay 128-byte aligned");
};

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

// int4 dequant from the blob ranges staged by the raw producer.  Thread mapping: a group is 64 (128) threads; thread tg
// owns word column wc = tg % 16 (8 output columns) of the CTA's 128-column part and the k rows kr + RPP ps (kr = tg / 16,
// RPP = 4 or 8) of the 64-row k-block; per sub-tile it reads 16 (8) words + 1 zero-point word + 8 scales and writes as
// many 16-byte slots of the MN-major SW128 B tile.  Bank check for the word reads of one warp (wc = lane % 16, two adjacent k rows): address / 4 =
// 276 b + 2 k + (wc & 1) with b = wc / 2; 276 = 20 (mod 32) and {20 b} = {0, 20, 8, 28, 16, 4, 24, 12}: 32 distinct banks.
template <int SUBS, bool BF16, int STAGES, int STAGE_BYTES, int RAW_N>
__device__ __forceinline__ void w4rp_dequant_loop(int dt, int lane, int first_tile, int tile_stride, int num_tiles, int num_kb,
                                                  int nloc, uint32_t b_stage0, uint32_t raw0, uint32_t empty_addr,
                                                  uint32_t full_addr, uint32_t raw_full_addr, uint32_t raw_empty_addr) {
  using C = CfgRP<SUBS>;
  constexpr int GROUPS = C::GROUPS;
  constexpr int GROUP_THREADS = 32 * NUM_DQ_WARPS / GROUPS;   // 64
  constexpr int WPR = C::NLOC / 8;                            // 16 word columns
  constexpr int RPP = GROUP_THREADS / WPR;                    // 4 k rows per pass
  constexpr int PASSES = 64 / RPP;                            // 16
  const int grp = dt / GROUP_THREADS, tg = dt % GROUP_THREADS;
  const int wc = tg % WPR, kr = tg / WPR;
  const bool in_tile = wc * 8 < nloc;
  const uint32_t chunk_off = uint32_t(wc >> 3) * (64 * ROW_BYTES);
  auto row_off = [&](int ps) {
    const uint32_t k = uint32_t(kr + ps * RPP);
    return chunk_off + k * ROW_BYTES + ((uint32_t(wc & 7) ^ (k & 7)) << 4);
  };
  uint32_t mask_lo = 0x000F000Fu, mask_hi = 0x00F000F0u;
  uint32_t magic = BF16 ? 0x43004300u : 0x64006400u;   // 128.0 / 1024.0: the nibble lands in the low mantissa bits
  asm volatile("" : "+r"(mask_lo), "+r"(mask_hi), "+r"(magic));
  auto and_or = [](uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));   // (a & b) | c
    return d;
  };
  auto lds32 = [](uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
  };
  const uint32_t blk_off = uint32_t(wc >> 1) * RP_BLK_BYTES + uint32_t(wc & 1) * 4u;   // word (k = 0) of this thread's column
  int stage = grp % STAGES;
  uint32_t phase = uint32_t(grp / STAGES) & 1u;   // more groups than stages: the first use may be a second lap
  const int my_tiles = first_tile < num_tiles ? (num_tiles - first_tile + tile_stride - 1) / tile_stride : 0;
  const int total = my_tiles * num_kb;
  const int rs_per_tile = (num_kb + 1) >> 1;
  int kb = grp, tl = 0;
  for (int it = grp; it < total; it += GROUPS) {
    while (kb >= num_kb) { kb -= num_kb; ++tl; }
    const int rseq = tl * rs_per_tile + (kb >> 1);            // raw stage sequence number of this CTA
    const bool lone = (kb == num_kb - 1) && (kb & 1) == 0;    // odd K tail: this group is the stage's only consumer
    const uint32_t half = uint32_t(kb & 1);
    const int rs = rseq % RAW_N;
    const uint32_t rphase = uint32_t(rseq / RAW_N) & 1u;
    kb += GROUPS;
    const uint32_t raw = raw0 + uint32_t(rs) * C::RAW_STAGE_BYTES;
    mbar_wait(raw_full_addr + 8u * rs, rphase);
    const uint32_t b_dst = b_stage0 + stage * STAGE_BYTES;
#pragma unroll
    for (int s = 0; s < SUBS; ++s) {
      const uint32_t blk = raw + uint32_t(s) * C::RAW_SUB_BYTES + blk_off;
      uint32_t w[PASSES];
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) w[ps] = lds32(blk + (half * 64u + uint32_t(kr + ps * RPP)) * 8u);
      const uint32_t zw = lds32(blk + RP_ZW_OFF + half * 8u);
      uint32_t sp[4];
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(sp[0]), "=r"(sp[1]), "=r"(sp[2]), "=r"(sp[3])
                   : "r"(raw + uint32_t(s) * C::RAW_SUB_BYTES + uint32_t(wc >> 1) * RP_BLK_BYTES + RP_SC_OFF + half * 32u + uint32_t(wc & 1) * 16u));
      uint32_t zsub[4];
      if (BF16) {
#pragma unroll
        for (int q = 0; q < 4; ++q) zsub[q] = and_or(zw >> (4 * q), mask_lo, magic);   // 128 + z
      } else {
        const uint32_t zs = zw >> 8;
        zsub[0] = and_or(zw, mask_lo, magic);   // 1024 + z
        zsub[2] = and_or(zs, mask_lo, magic);
        const __half2 sixteenth = __float2half2_rn(0.0625f);   // high nibbles decode as 1024 + 16 z; /16 = 64 + z exactly
        const uint32_t z1 = and_or(zw, mask_hi, magic), z3 = and_or(zs, mask_hi, magic);
        __half2 h1 = __hmul2(*reinterpret_cast<const __half2*>(&z1), sixteenth);
        __half2 h3 = __hmul2(*reinterpret_cast<const __half2*>(&z3), sixteenth);
        zsub[1] = *reinterpret_cast<uint32_t*>(&h1);
        zsub[3] = *reinterpret_cast<uint32_t*>(&h3);
      }
      if (s == 0) mbar_wait(empty_addr + 8u * stage, phase ^ 1);
      if (in_tile) {   // columns past the tile part are never read by the MMA
#pragma unroll
        for (int ps = 0; ps < PASSES; ++ps) {
          const uint32_t wv = w[ps];
          uint32_t o[4];
          if (BF16) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t t = and_or(wv >> (4 * q), mask_lo, magic);   // {128 + q(col 2q), 128 + q(col 2q+1)}
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
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(b_dst + uint32_t(s) * C::B_SUB_BYTES + row_off(ps)), "r"(o[0]),
                       "r"(o[1]), "r"(o[2]), "r"(o[3])
                       : "memory");
        }
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      mbar_arrive_cluster(full_addr + 8u * stage);
      mbar_arrive_n(raw_empty_addr + 8u * rs, lone ? 2u : 1u);   // stands in for the absent second k-block's warps
    }
    stage += GROUPS;
    while (stage >= STAGES) { stage -= STAGES; phase ^= 1; }
  }
}

// ---------------------------------------------------------------- the kernel
// p.tile_n = sub_n (multiple of 32, <= 256): a pair's tile is 256 rows x SUBS * sub_n columns; CTA `rank` holds the
// columns [rank * sub_n / 2, (rank + 1) * sub_n / 2) of each sub-tile.
template <int SUBS, bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CfgRP<SUBS>::THREADS, 1)
qdm_w4rp_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_y,
                const __grid_constant__ CUtensorMap map_y16, const GemmParams p) {
  using C = CfgRP<SUBS>;
  constexpr int STAGES = C::STAGES;
  constexpr int ACCS = C::ACCS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // layout: [stages x (A | B sub-tiles)] [epilogue: 4 store stagings, 4 fp32 vectors] [raw ring] [barriers] [tmem ptr]
  const uint32_t epi_base = smem_base + STAGES * C::STAGE_BYTES;
  const uint32_t raw_base = epi_base + C::EPI_BYTES;
  const uint32_t bar_base = raw_base + C::RAW_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  auto raw_full_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 4 + s); };
  auto raw_empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 4 + RAW_STAGES + s); };
  static_assert(C::RAW_N <= RAW_STAGES && 8 * (2 * 8 + 4 + 2 * RAW_STAGES) + 8 <= 256, "barrier area");
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(
      smem_gen + (bar_base - smem_base) + 8 * (2 * STAGES + 4 + 2 * RAW_STAGES));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = int(blockIdx.x) >> 1, num_pairs = int(gridDim.x) >> 1;
  const int sub_n = p.tile_n, nloc = sub_n >> 1, tile_w = SUBS * sub_n;
  const int m_tiles = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int n_tiles = (p.N + tile_w - 1) / tile_w;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = p.K / 64;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&map_a);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), C::FULL_COUNT); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), 8); }   // 4 epilogue warps x 2 CTAs
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
  cluster_sync_all();   // barrier inits and TMEM allocation of BOTH CTAs are visible before anything remote happens
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t leader_full0 = mapa_shared(full_bar(0), 0);
  const uint32_t leader_tmem_empty0 = mapa_shared(tmem_empty_bar(0), 0);

  // accumulator columns of (accumulator set a, sub-tile s)
  auto acc_col = [&](int a, int s) { return uint32_t(SUBS == 1 ? a * 256 : s * 256); };
  // drain this warp's share (chunks c_first, c_first + c_step, ...) of every sub-tile of one tile
  auto drain_tile = [&](int tile, int a, int ew, uint32_t stg, float* vec_sm, int c_first, int c_step) {
    const int m0 = (tile / n_tiles) * (2 * BLOCK_M) + int(rank) * BLOCK_M, n0 = (tile % n_tiles) * tile_w;
#pragma unroll
    for (int s = 0; s < SUBS; ++s) {
      if (n0 + s * sub_n < p.N)
        epilogue_drain<256, G_W4, BF16>(p, &map_y, &map_y16, stg, vec_sm, tmem_base + (uint32_t(ew * 32) << 16) + acc_col(a, s),
                                        m0 + ew * 32, n0 + s * sub_n, lane, c_first, c_step, nullptr, 0, 0, sub_n);
    }
  };

  if (warp == 0) {
    // ===================================================== TMA producer: A (both CTAs, each its 128 rows)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m0 = (tile / n_tiles) * (2 * BLOCK_M) + int(rank) * BLOCK_M;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          // the peer's bytes land on the leader's barrier too; the peer itself does not arrive
          if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * A_STAGE_BYTES);
          tma_load_2d_pair(smem_base + stage * C::STAGE_BYTES, &map_a, leader_full0 + 8u * stage, kb * 64, m0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = make_idesc(1, BF16 ? 1 : 0, 1, 2 * BLOCK_M, sub_n);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * C::STAGE_BYTES;
#pragma unroll
          for (int s = 0; s < SUBS; ++s) {
            const uint32_t b_addr = a_addr + A_STAGE_BYTES + uint32_t(s) * C::B_SUB_BYTES;
            const uint32_t tmem_c = tmem_base + acc_col(acc, s);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = make_smem_desc(a_addr + k * 32, 16, 1024);
              const uint64_t db = make_smem_desc(b_addr + k * 2048, 64 * ROW_BYTES, 1024);
              umma_pair<G_W4>(tmem_c, da, db, idesc, (kb | k) != 0);
            }
          }
          umma_commit_pair(empty_bar(stage), 3);
          if (kb == num_kb - 1) umma_commit_pair(tmem_full_bar(acc), 3);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == ACCS) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ===================================================== raw producer: one bulk copy per sub-tile part per 128 k rows
    if (elect_one()) {
      int rs = 0;
      uint32_t rphase = 0;
      const int nblk_part = nloc / RP_BLK_COLS;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int n0 = (tile % n_tiles) * tile_w + int(rank) * nloc;
        for (int j = 0; 2 * j < num_kb; ++j) {
          mbar_wait(raw_empty_bar(rs), rphase ^ 1);
          const uint32_t raw = raw_base + uint32_t(rs) * C::RAW_STAGE_BYTES;
          int nblk[SUBS], nb0[SUBS];
          uint32_t bytes = 0;
#pragma unroll
          for (int s = 0; s < SUBS; ++s) {
            nb0[s] = (n0 + s * sub_n) / RP_BLK_COLS;
            nblk[s] = min(nblk_part, p.rp_nb - nb0[s]);
            if (nblk[s] < 0) nblk[s] = 0;
            bytes += uint32_t(nblk[s]) * RP_BLK_BYTES;
          }
          mbar_expect_tx(raw_full_bar(rs), bytes);   // with 0 bytes (tile part past N) this is a plain arrive
#pragma unroll
          for (int s = 0; s < SUBS; ++s)
            if (nblk[s] > 0)
              bulk_load(raw + uint32_t(s) * C::RAW_SUB_BYTES, p.rp_blob + (size_t(j) * p.rp_nb + nb0[s]) * RP_BLK_BYTES,
                        uint32_t(nblk[s]) * RP_BLK_BYTES, raw_full_bar(rs));
          if (++rs == C::RAW_N) { rs = 0; rphase ^= 1; }
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================================================== epilogue (each CTA drains its own 128 rows)
    const int ew = warp - 4;
    const uint32_t stg = epi_base + ew * EPI_STG_BYTES;
    float* vec_sm = reinterpret_cast<float*>(smem_gen + STAGES * C::STAGE_BYTES + 4 * EPI_STG_BYTES) + ew * (EPI_VEC_BYTES / 4);
    if (lane == 0) { tma_prefetch_desc(&map_y); tma_prefetch_desc(&map_y16); }
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      mbar_wait(tmem_full_bar(acc), acc_phase);
      tc_fence_after();
      const bool last = tile + num_pairs >= num_tiles;   // shared with the (by then idle) dequant warps, see below
      drain_tile(tile, acc, ew, stg, vec_sm, 0, last ? 3 : 1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tmem_empty0 + 8u * acc);
      if (++acc == ACCS) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();
  } else if (warp >= 8) {
    // ===================================================== int4 dequant (each CTA: its NLOC columns of every sub-tile)
    w4rp_dequant_loop<SUBS, BF16, STAGES, C::STAGE_BYTES, C::RAW_N>(
        threadIdx.x - 256, lane, pair, num_pairs, num_tiles, num_kb, nloc, smem_base + A_STAGE_BYTES, raw_base, empty_bar(0),
        leader_full0, raw_full_bar(0), raw_empty_bar(0));
    // ---- help drain the pair's last tile: two more warp sets (TMEM lane quarter = warp % 4) take chunks 1, 4, ... and
    // 2, 5, ...; staging lives in the B parts of pipeline stages 0..2, idle once the last accumulator is complete
    const int my_tiles = pair < num_tiles ? (num_tiles - pair + num_pairs - 1) / num_pairs : 0;
    if (my_tiles > 0) {
      const int lt = my_tiles - 1, tile = pair + lt * num_pairs;
      const int acc = lt % ACCS, dw = warp - 8, ew = dw & 3, set = 1 + (dw >> 2);
      mbar_wait(tmem_full_bar(acc), uint32_t(lt / ACCS) & 1u);
      tc_fence_after();
      const uint32_t stg = smem_base + uint32_t(dw >> 2) * C::STAGE_BYTES + A_STAGE_BYTES + uint32_t(dw & 3) * EPI_STG_BYTES;
      float* vec_sm = reinterpret_cast<float*>(smem_gen + 2 * C::STAGE_BYTES + A_STAGE_BYTES) + dw * 256;
      drain_tile(tile, acc, ew, stg, vec_sm, set, 3);
      if (lane == 0) tma_store_wait_all();
    }
  }

  tc_fence_before();
  // execution-only rendezvous (the peer may still be reading this CTA's operands / signalling its barriers)
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}

template <int SUBS, bool BF16>
int launch_rp(const CUtensorMap& a, const CUtensorMap& y, const CUtensorMap& y16, const GemmParams& p, cudaStream_t st) {
  using C = CfgRP<SUBS>;
  auto kern = qdm_w4rp_kernel<SUBS, BF16>;
  static bool attr_set = false;   // per instantiation; benign race (idempotent)
  if (!attr_set) {
    QDM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int tile_w = SUBS * p.tile_n;
  const int64_t tiles = int64_t((p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * ((p.N + tile_w - 1) / tile_w);
  const int pairs = int(tiles < QDM_NUM_SMS / 2 ? tiles : QDM_NUM_SMS / 2);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(2 * pairs));
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  QDM_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, a, y, y16, p));
  QDM_LAUNCH_CHECK();
  return QDM_OK;
}

}  // namespace

// Called by the W4A16 dispatcher (qdm_gemm.cu) once it has picked the RP path: `subs` sub-tiles of `sub_n` columns.
int qdm_w4rp_gemm(const void* x, const void* blob, const void* bias, void* y, int is_bf16, int64_t M, int64_t N, int64_t K,
                  int subs, int sub_n, cudaStream_t st) {
  QDM_REQUIRE(blob && qdm_aligned16(blob), "qdm_gemm_w4a16_rp: the repacked weight must be 16-byte aligned");
  QDM_REQUIRE((subs == 1 || subs == 2) && sub_n >= 32 && sub_n <= 256 && sub_n % 32 == 0, "qdm_gemm_w4a16_rp: bad tile %d x %d", subs, sub_n);
  QDM_REQUIRE(K % 64 == 0 && N % 8 == 0 && M > BLOCK_M, "qdm_gemm_w4a16_rp: shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  int rc = get_encode_fn();
  if (rc) return rc;
  CUtensorMap ma, my, my16;
  if ((rc = make_map(&ma, x, 2, M, K, 64, BLOCK_M))) return rc;
  if ((rc = make_map(&my, y, 2, M, N, EPI_COLS, 32))) return rc;
  if ((rc = make_map(&my16, y, 2, M, N, 16, 32, false))) return rc;
  GemmParams p{};
  p.M = int(M); p.N = int(N); p.K = int(K); p.tile_n = sub_n; p.bias = bias; p.y = y; p.is_bf16 = is_bf16;
  p.rp_blob = static_cast<const uint8_t*>(blob);
  p.rp_nb = int((N + RP_BLK_COLS - 1) / RP_BLK_COLS);
  if (subs == 2) return is_bf16 ? launch_rp<2, true>(ma, my, my16, p, st) : launch_rp<2, false>(ma, my, my16, p, st);
  return is_bf16 ? launch_rp<1, true>(ma, my, my16, p, st) : launch_rp<1, false>(ma, my, my16, p, st);
}

extern "C" size_t qdm_w4a16_repack_bytes(int64_t N, int64_t K) {
  if (N <= 0 || K <= 0) return 0;
  return size_t((K + 127) / 128) * size_t((N + RP_BLK_COLS - 1) / RP_BLK_COLS) * RP_BLK_BYTES;
}

extern "C" int qdm_w4a16_repack(const int32_t* qweight, const int32_t* qzeros, const void* scales, int64_t N, int64_t K,
                                int group, void* blob, size_t blob_bytes, void* stream) {
  QDM_REQUIRE(qweight && qzeros && scales && blob, "qdm_w4a16_repack: null pointer");
  QDM_REQUIRE(N > 0 && K > 0 && N % 8 == 0 && K % 64 == 0 && N < (1LL << 31) && K < (1LL << 31),
              "qdm_w4a16_repack: N=%lld must be a multiple of 8 and K=%lld of 64", (long long)N, (long long)K);
  QDM_REQUIRE(group > 0 && group % 64 == 0 && K % group == 0, "qdm_w4a16_repack: group=%d must be a multiple of 64 dividing K", group);
  QDM_REQUIRE(blob_bytes >= qdm_w4a16_repack_bytes(N, K) && qdm_aligned16(blob), "qdm_w4a16_repack: blob needs %zu bytes, 16-byte aligned",
              qdm_w4a16_repack_bytes(N, K));
  QDM_DEVICE_GATE();
  const int nb = int((N + RP_BLK_COLS - 1) / RP_BLK_COLS), kg = int((K + 127) / 128);
  const int64_t grid = int64_t(kg) * ((nb + RP_NB_PER_CTA - 1) / RP_NB_PER_CTA);
  QDM_REQUIRE(grid < (1LL << 31), "qdm_w4a16_repack: weight too large");
  w4rp_repack_kernel<<<unsigned(grid), 128, 0, (cudaStream_t)stream>>>(qweight, qzeros, static_cast<const uint16_t*>(scales), int(N), int(K),
                                                                    group, nb, static_cast<uint8_t*>(blob));
  QDM_LAUNCH_CHECK();
  return QDM_OK;
}
