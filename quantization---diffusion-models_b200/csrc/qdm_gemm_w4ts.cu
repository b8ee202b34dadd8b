// (c) W4A16 GEMM with the WEIGHTS as the tensor-core A operand in TENSOR MEMORY (the "TS" path):
//     y^T[N, M] = dequant(qweight, qzeros, scales)^T[N, K] . x^T[K, M]        (N output channels, M tokens)
//
// Why.  In the other W4 kernels the dequantised B tile travels through shared memory: the dequant warps write 16 KB per
// k-block with st.shared, the tensor core reads it back, next to the A tile written by TMA and read by the tensor core
// and the packed operand staged in between -- ~73 KB of shared-memory traffic per k-block per SM against 64 KB for a
// plain fp16 GEMM, and measured k-block periods of 850-960 cycles against 512 cycles of tensor-pipe time
// (profiles/README.md, round 2: both kernels sit at ~60 % of their shared-memory-port minimum).  tcgen05.mma can take
// its A operand from TMEM.  With the roles swapped -- output channels on the 128 TMEM lanes, tokens along the MMA N
// dimension -- the dequantised weights go registers -> TMEM (tcgen05.st) and never touch shared memory:
//   * shared memory only carries the activations (16 KB written by TMA + read by the tensor core per k-block per SM:
//     half of what an fp16 GEMM needs);
//   * a dequant thread owns ONE output channel (its TMEM lane) and the 64 k of a k-block: one scale and one zero point
//     per thread per k-block, 8 packed words -> 32 half2 registers -> one tcgen05.st.32x32b.x32;
//   * the packed operand is read straight from global memory into registers (a kernel-native copy of the AWQ tensors
//     built once by qdm_w4a16_repack_ts: per (k-block, channel) 8 words whose nibble order along k is the AWQ order
//     [0,2,4,6,1,3,5,7], so that one lop3 yields the half2 (k, k+1) = one 32-bit TMEM cell; a warp reads 1 KB
//     contiguous), prefetched three k-blocks ahead in registers.
// Dequant arithmetic is unchanged: (q - z) exact in the 16-bit type (magic-number trick), * s with one rounding --
// bit-identical to dequantize_gemm (utils/packing_utils.py:87-102).
//
// Tile: a CTA pair owns 256 output channels (128 TMEM lanes per CTA) x T <= 192 tokens (cta_group::2, M = 256, N = T).
// TMEM per CTA (512 columns): TWO fp32 accumulators of DW columns (the epilogue of tile t overlaps the main loop of tile
// t + 1 -- measured on the single-buffered first version: 5700 of 18400 cycles per tile were the exposed epilogue) and the
// weight ring in the rest: DW = 128 -> 4 stages of 64 columns (128 k as 64 half2), DW = 192 -> 2 stages.
// A pipeline stage is K = 128 (two k-blocks): the single MMA-issuing thread pays ~250 cycles of serial latency per
// barrier wait (measured), so it gets ONE full barrier per stage and 8 tcgen05.mma per wait.
// Roles (16 warps): warp 0 TMA producer of x (each CTA T/2 token rows), warp 1 MMA issuer (leader CTA), warp 2 TMEM
// allocator, warps 4-7 epilogue (TMEM -> registers -> +bias -> 16 bit -> shared [32 tokens][128 channels] -> one TMA store
// of 256-byte rows per 32 tokens), warps 8-11 / 12-15 two dequant sets taking even / odd stages.
#include "qdm_gemm_dev.cuh"
#include <stdlib.h>

using namespace qdmg;

namespace {

#ifndef TS_ORDER
#define TS_ORDER 0   // 0: token blocks fastest (pairs share a weight slab), 1: channel blocks fastest (pairs share a token tile)
#endif
// Tile order (TS_ORDER 0): PANELS of p.group_m token blocks; inside a panel the token blocks run fastest and the channel
// blocks slowest, so concurrently running pairs share a weight slab and the panel's activations (group_m * T * K * 2 bytes,
// sized by the host to stay in L2) are fetched from DRAM once and re-read from L2 by the panel's other channel blocks.
// group_m = m_blks is the plain "token blocks fastest" order (every shape whose activations fit L2 as a whole); without
// panels a problem whose activations exceed L2 streams them from DRAM once per channel block
// (65536 x 3072 x 3072: 12 x 403 MB = 750 us of DRAM time in a 1000 us kernel).
__device__ __forceinline__ int ts_nblk(int tile, int m_blks, int n_blks, int gm) {
  if (TS_ORDER) return tile % n_blks;
  const int per = gm * n_blks, pnl = tile / per, r = tile - pnl * per;
  return r / min(gm, m_blks - pnl * gm);
}
__device__ __forceinline__ int ts_mblk(int tile, int m_blks, int n_blks, int gm) {
  if (TS_ORDER) return tile / n_blks;
  const int per = gm * n_blks, pnl = tile / per, r = tile - pnl * per, gsz = min(gm, m_blks - pnl * gm);
  return pnl * gm + r % gsz;
}
#define TS_NBLK(tile) ts_nblk(tile, m_blks, n_blks, p.group_m)
#define TS_MBLK(tile) ts_mblk(tile, m_blks, n_blks, p.group_m)
constexpr int TS_KB = 2;                       // k-blocks (64 k) per pipeline stage: K = 128 per stage
constexpr int TS_DIST = 1;                     // packed words are prefetched this many of a set's stages ahead (one set stage = two
                                               // pipeline stages of lead time; the registers go to the unpacked values instead)
constexpr int TS_BAR_BYTES = 512;
constexpr int TS_THREADS = 512;

// DW = TMEM columns of ONE accumulator buffer = most tokens per tile.  Two buffers (the epilogue of tile t overlaps the
// main loop of tile t + 1) + the weight ring share the 512 columns: DW = 128 -> 4 weight stages, DW = 192 -> 2.
// WIDE: the two accumulator buffers are the two halves ("sub-tiles") of ONE tile of up to 2 DW tokens instead of two
// alternating tiles: every weight stage feeds 16 MMAs instead of 8 (the issuing thread's fixed cost per stage -- barrier
// test, commit -- and the dequant work are amortised over twice the tokens), at the price of an epilogue that no longer
// overlaps the next tile's main loop.  Taken where it removes a wave of tiles or the main loop is long (qdm_gemm.cu).
template <int DW, bool WIDE = false>
struct CfgTS {
  static constexpr int NS = (512 - 2 * DW) / 64;             // weight stages in TMEM (64 columns = 128 k each)
  static constexpr int A_COL0 = 2 * DW;
  static constexpr int KB_BYTES = DW * 64;                   // one k-block of activations: DW / 2 token rows x 128 B
  static constexpr int SUBS = WIDE ? 2 : 1;
  static constexpr int X_STAGE_BYTES = TS_KB * SUBS * KB_BYTES;   // [k-block][sub-tile][DW / 2 rows x 128 B]
  // activation stages in shared memory: deeper than the weight ring -- the L2 -> SM latency under load is 3000-5000
  // cycles (measured), several stages of tensor-pipe time
#ifndef TS_NXS192
#define TS_NXS192 6
#endif
#ifndef TS_NXS128
#define TS_NXS128 8
#endif
  static constexpr int NXS = WIDE ? (DW == 128 ? 5 : DW == 160 ? 4 : 3) : (DW <= 160 ? TS_NXS128 : TS_NXS192);
#ifndef TS_EPI_TILES
#define TS_EPI_TILES 2
#endif
  static constexpr int EPI_TILES = TS_EPI_TILES;
  static constexpr int EPI_BYTES = EPI_TILES * 32 * 256;     // staging tiles [32 tokens][128 channels] x 2 B
  static constexpr int SMEM_BYTES = NXS * X_STAGE_BYTES + EPI_BYTES + 1024 + TS_BAR_BYTES;
  static_assert(NS >= 2 && NXS > NS, "the dequant sets wait on the release barrier of the stage NS back: it must still be the x ring's current phase");
  static_assert(8 * (2 * NXS + NS + 4) + 8 <= TS_BAR_BYTES, "barrier area");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
};

// ---------------------------------------------------------------- one-time repack
// words[kb][n][8]: nibble i of word j = code of k = 64 kb + 8 j + {0,2,4,6,1,3,5,7}[i] of output channel n
// sz[kb][n]      : low 16 bits = scale (dtype bits) of the k-block's group, high 16 bits = zero point as a dtype value
template <bool BF16>
__global__ void __launch_bounds__(256) w4ts_repack_kernel(const int32_t* __restrict__ qweight, const int32_t* __restrict__ qzeros,
                                                          const uint16_t* __restrict__ scales, int N, int K, int group,
                                                          uint32_t* __restrict__ words, uint32_t* __restrict__ sz) {
  const int64_t idx = int64_t(blockIdx.x) * 256 + threadIdx.x;   // (kb, n)
  const int kbs = K / 64;
  if (idx >= int64_t(kbs) * N) return;
  const int kb = int(idx / N), n = int(idx % N);
  const int words_per_row = N / 8;
  const int pos = ((n & 7) >> 1) + ((n & 1) << 2);   // nibble of column n inside its AWQ word: order [0,2,4,6,1,3,5,7]
  const int shift = 4 * pos;
  uint32_t out[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t w = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int k = kb * 64 + 8 * j + t;
      const uint32_t code = (uint32_t(qweight[int64_t(k) * words_per_row + (n >> 3)]) >> shift) & 0xFu;
      const int nib = (t >> 1) + ((t & 1) << 2);     // k even -> nibbles 0..3, k odd -> nibbles 4..7
      w |= code << (4 * nib);
    }
    out[j] = w;
  }
  uint4* dst = reinterpret_cast<uint4*>(words + idx * 8);
  dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
  dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
  const int64_t g = (int64_t(kb) * 64) / group;
  const uint32_t z = (uint32_t(qzeros[g * words_per_row + (n >> 3)]) >> shift) & 0xFu;
  uint32_t zbits;
  if (BF16) { __nv_bfloat16 h = __float2bfloat16_rn(float(z)); zbits = *reinterpret_cast<uint16_t*>(&h); }
  else { __half h = __float2half_rn(float(z)); zbits = *reinterpret_cast<uint16_t*>(&h); }
  sz[idx] = uint32_t(scales[g * N + n]) | (zbits << 16);
}

// ---------------------------------------------------------------- PTX: TMEM store, A-from-TMEM MMA
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
      "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
      "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] . B[smem]; both CTAs of the pair hold their 128 rows of A at the same TMEM address
// One pipeline stage for the issuing thread: NKB x 4 tcgen05.mma (K = 16 each) with the NEXT stage's barrier tested at the
// top and its predicate read (selp) only after the last MMA.  Measured with the QDM_TRACE timeline (profiles/README.md,
// round 2): a barrier test costs the thread ~290 cycles when its result is consumed at once, each MMA ~80 cycles when its
// shared-memory descriptor is rebuilt from the address (five dependent uniform-datapath instructions), a commit ~170 --
// 1270 cycles of serial issue per K = 128 stage against 768 cycles of tensor-pipe work at 192 tokens.  Here the test's
// latency overlaps the MMAs, and every descriptor is the previous one plus 2 (32 bytes >> 4) in the low word.
template <int NKB>
__device__ __forceinline__ uint32_t ts_issue_stage(uint32_t tmem_c, uint32_t tmem_a, uint32_t desc_lo, uint32_t desc_lo_step,
                                                   uint32_t desc_hi, uint32_t idesc, uint32_t accum, uint32_t next_bar,
                                                   uint32_t next_parity) {
  uint32_t done;
  if (NKB == 2) {
    asm volatile(
        "{\n\t.reg .pred pa, pt, pd, pe;\n\t.reg .b64 d;\n\t.reg .b32 lo, ta;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 pd, [%8], %9;\n\t"
        "setp.ne.b32 pa, %7, 0;\n\t"
        "setp.eq.b32 pt, %7, %7;\n\t"
        "mov.b64 d, {%3, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], [%2], d, %6, pa;\n\t"
        "add.u32 lo, %3, 2;\n\tadd.u32 ta, %2, 8;\n\tmov.b64 d, {lo, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], [ta], d, %6, pt;\n\t"
        "add.u32 lo, %3, 4;\n\tadd.u32 ta, %2, 16;\n\tmov.b64 d, {lo, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], [ta], d, %6, pt;\n\t"
        "add.u32 lo, %3, 6;\n\tadd.u32 ta, %2, 24;\n\tmov.b64 d, {lo, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], [ta], d, %6, pt;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 pe, [%8], %9;\n\t"   // second look, ~300 cycles later
        "add.u32 lo, %3, %4;\n\tadd.u32 ta, %2, 32;\n\tmov.b64 d, {lo, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], [ta], d, %6, pt;\n\t"
        "add.u32 lo, lo, 2;\n\tadd.u32 ta, %2, 40;\n\tmov.b64 d, {lo, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], [ta], d, %6, pt;\n\t"
        "add.u32 lo, lo, 2;\n\tadd.u32 ta, %2, 48;\n\tmov.b64 d, {lo, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], [ta], d, %6, pt;\n\t"
        "add.u32 lo, lo, 2;\n\tadd.u32 ta, %2, 56;\n\tmov.b64 d, {lo, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], [ta], d, %6, pt;\n\t"
        "or.pred pd, pd, pe;\n\t"
        "selp.u32 %0, 1, 0, pd;\n\t}"
        : "=r"(done)
        : "r"(tmem_c), "r"(tmem_a), "r"(desc_lo), "r"(desc_lo_step), "r"(desc_hi), "r"(idesc), "r"(accum), "r"(next_bar),
          "r"(next_parity)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred pa, pt, pd;\n\t.reg .b64 d;\n\t.reg .b32 lo, ta;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 pd, [%8], %9;\n\t"
        "setp.ne.b32 pa, %7, 0;\n\t"
        "setp.eq.b32 pt, %7, %7;\n\t"
        "mov.b64 d, {%3, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], [%2], d, %6, pa;\n\t"
        "add.u32 lo, %3, 2;\n\tadd.u32 ta, %2, 8;\n\tmov.b64 d, {lo, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], [ta], d, %6, pt;\n\t"
        "add.u32 lo, %3, 4;\n\tadd.u32 ta, %2, 16;\n\tmov.b64 d, {lo, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], [ta], d, %6, pt;\n\t"
        "add.u32 lo, %3, 6;\n\tadd.u32 ta, %2, 24;\n\tmov.b64 d, {lo, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], [ta], d, %6, pt;\n\t"
        "selp.u32 %0, 1, 0, pd;\n\t}"
        : "=r"(done)
        : "r"(tmem_c), "r"(tmem_a), "r"(desc_lo), "r"(desc_lo_step), "r"(desc_hi), "r"(idesc), "r"(accum), "r"(next_bar),
          "r"(next_parity)
        : "memory");
  }
  return done;
}

// WIDE tiles: one k-block (4 x K 16) for BOTH accumulators = 8 MMAs; the next stage's barrier is tested first and its
// predicate read after the last MMA of the block (the caller ORs the two k-blocks' results).
__device__ __forceinline__ uint32_t ts_issue_kblock_wide(uint32_t tmem_c0, uint32_t tmem_c1, uint32_t tmem_a, uint32_t desc_lo0,
                                                        uint32_t desc_lo1, uint32_t desc_hi, uint32_t idesc, uint32_t accum,
                                                        uint32_t next_bar, uint32_t next_parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred pa, pt, pd;\n\t.reg .b64 d0, d1;\n\t.reg .b32 l0, l1, ta;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 pd, [%9], %10;\n\t"
      "setp.ne.b32 pa, %8, 0;\n\t"
      "setp.eq.b32 pt, %8, %8;\n\t"
      "mov.b64 d0, {%4, %6};\n\tmov.b64 d1, {%5, %6};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%1], [%3], d0, %7, pa;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%2], [%3], d1, %7, pa;\n\t"
      "add.u32 l0, %4, 2;\n\tadd.u32 l1, %5, 2;\n\tadd.u32 ta, %3, 8;\n\tmov.b64 d0, {l0, %6};\n\tmov.b64 d1, {l1, %6};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%1], [ta], d0, %7, pt;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%2], [ta], d1, %7, pt;\n\t"
      "add.u32 l0, %4, 4;\n\tadd.u32 l1, %5, 4;\n\tadd.u32 ta, %3, 16;\n\tmov.b64 d0, {l0, %6};\n\tmov.b64 d1, {l1, %6};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%1], [ta], d0, %7, pt;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%2], [ta], d1, %7, pt;\n\t"
      "add.u32 l0, %4, 6;\n\tadd.u32 l1, %5, 6;\n\tadd.u32 ta, %3, 24;\n\tmov.b64 d0, {l0, %6};\n\tmov.b64 d1, {l1, %6};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%1], [ta], d0, %7, pt;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%2], [ta], d1, %7, pt;\n\t"
      "selp.u32 %0, 1, 0, pd;\n\t}"
      : "=r"(done)
      : "r"(tmem_c0), "r"(tmem_c1), "r"(tmem_a), "r"(desc_lo0), "r"(desc_lo1), "r"(desc_hi), "r"(idesc), "r"(accum), "r"(next_bar),
        "r"(next_parity)
      : "memory");
  return done;
}

struct TsParams {
  int M, N, K;                 // tokens, output channels, reduction
  int tile_t;                  // tokens per tile (multiple of 16, <= 256)
  int group_m;                 // token blocks per panel of the tile order (see ts_nblk); >= 1, m_blks = no panels
  const uint32_t* words;       // [K/64][N][8]
  const uint32_t* sz;          // [K/64][N]
  const void* bias;            // [N] dtype or null
  void* y;                     // [M, N] dtype
  long long* trace;            // QDM_TRACE builds only
};

// ---------------------------------------------------------------- the kernel
template <int DW, bool BF16, bool WIDE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TS_THREADS, 1)
qdm_w4ts_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y, const TsParams p) {
  using C = CfgTS<DW, WIDE>;
  constexpr int NS = C::NS, NXS = C::NXS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + NXS * C::X_STAGE_BYTES;
  const uint32_t bar_base = epi_base + C::EPI_BYTES;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // ONE full barrier per stage: the MMA issuer is a single thread whose barrier waits are serial latency (measured: ~250
  // cycles per wait even on a complete barrier), so a stage of K = 128 costs it one wait.  full[g % NXS] collects, for
  // global stage g, the leader's expect_tx (TMA bytes of both CTAs) and the 4 dequant warps of the stage's set in both
  // CTAs.  The stage is released by ONE multicast commit (a commit costs the issuer ~170 cycles): x_empty[g % NXS], awaited
  // by the TMA producers (x slot g % NXS) and by the dequant set that writes TMEM slot g % NS next, i.e. for stage g + NS.
  // NXS and NS are even, so a slot of either ring always belongs to the same dequant set.
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto x_empty_bar = [&](int s) { return bar_base + 8u * (NXS + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * NXS + NS + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * NXS + NS + 2 + a); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + (bar_base - smem_base) + 8 * (2 * NXS + NS + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = int(blockIdx.x) >> 1, num_pairs = int(gridDim.x) >> 1;
  const int T = p.tile_t;                                          // tokens per tile
  const int Ts = WIDE ? T >> 1 : T;                                // ... per accumulator (sub-tile): the MMA's N
  const int th = Ts >> 1;                                          // token rows of one TMA box = per CTA and sub-tile
  const int n_blks = (p.N + 255) / 256, m_blks = (p.M + T - 1) / T;
  const int num_tiles = n_blks * m_blks;                           // tile = n_blk * m_blks + m_blk: token blocks fastest, so
  const int num_kb = p.K / 64;                                     // concurrently running pairs share a weight slab in L2
  const int num_st = (num_kb + TS_KB - 1) / TS_KB;                 // pipeline stages per tile (the last may hold one k-block)
  const int my_tiles = pair < num_tiles ? (num_tiles - pair + num_pairs - 1) / num_pairs : 0;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_y); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NXS; ++s) { mbar_init(full_bar(s), 9); mbar_init(x_empty_bar(s), 1); }   // 1 + 4 dequant warps x 2 CTAs
    for (int a = 0; a < 2; ++a) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), 8); }   // 4 epilogue warps x 2 CTAs
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(const_cast<uint32_t*>(tmem_ptr_smem))),
                 "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // Programmatic dependent launch: the next kernel may start its set-up now.  Only the roles that touch memory an EARLIER
  // kernel produces or still reads wait for it -- the TMA producer (activations) and the epilogue (the output buffer may be
  // an earlier kernel's input).  The dequant warps read constants (packed weights, scales, zeros), so their first loads --
  // a cold HBM round trip -- overlap the previous kernel's tail instead of following it.
  pdl_launch_dependents();
  const uint32_t leader_full0 = mapa_shared(full_bar(0), 0);
  const uint32_t leader_tmem_empty0 = mapa_shared(tmem_empty_bar(0), 0);

  if (warp == 0) {
    // ===================================================== TMA producer: x, this CTA's half of the tile's tokens
    if (elect_one()) {
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      TRC_DECL;
      int trc_it = 0;
      const uint32_t kb_bytes = uint32_t(th) * ROW_BYTES;               // the tensor map's box is th token rows x 64 k
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m0 = TS_MBLK(tile) * T + int(rank) * th;   // sub-tile s adds s * Ts: an accumulator's columns are consecutive tokens
        for (int st = 0; st < num_st; ++st) {
          mbar_wait(x_empty_bar(stage), phase ^ 1);
          TRC(p.trace, 0, 2000000 + trc_it);
          ++trc_it;
          const int nk = min(TS_KB, num_kb - st * TS_KB);
#ifdef TS_NO_X   // timing experiment only (wrong results): no activation loads at all
          if (rank == 0) mbar_arrive(full_bar(stage));
          for (int j = 0; j < 0; ++j)
#else
          if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * uint32_t(nk * C::SUBS) * kb_bytes);   // rows past M are zero-filled and counted
          for (int j = 0; j < nk * C::SUBS; ++j)
#endif
            tma_load_2d_pair(smem_base + stage * C::X_STAGE_BYTES + j * C::KB_BYTES, &map_x, leader_full0 + 8u * stage,
                             (st * TS_KB + j / C::SUBS) * 64, m0 + (j % C::SUBS) * Ts);
          if (++stage == NXS) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    if (rank == 0 && elect_one()) {   // elect.sync, not `lane == 0`: the compiler then knows ONE lane runs this (qdm_gemm_dev.cuh)
      const uint32_t idesc = make_idesc(1, BF16 ? 1 : 0, 0, 2 * BLOCK_M, Ts);
      int stage = 0, as = 0;
      uint32_t phase = 0;
      TRC_DECL;
      int trc_it = 0;
      uint32_t ready = 0;                                                // has the peek already seen this stage's barrier complete?
      for (int tl = 0; tl < my_tiles; ++tl) {
        const int acc = WIDE ? 0 : (tl & 1);      // WIDE: both buffers belong to this tile; one full / empty barrier pair
        mbar_wait(tmem_empty_bar(acc), (uint32_t(WIDE ? tl : (tl >> 1)) & 1u) ^ 1u);
        TRC(p.trace, 1, 1000000 + trc_it);
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + uint32_t(acc * DW);
        for (int st = 0; st < num_st; ++st) {
          if (!ready) mbar_wait(full_bar(stage), phase);
          TRC(p.trace, 1, (ready ? 2000000 : 9000000) + trc_it);    // 9 = the stage was not yet complete at the peek: waited
          if ((trc_it & 15) == 0) TRCG(p.trace, 1, 10000000 + trc_it);
          ++trc_it;
          tc_fence_after();
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == NXS) { nstage = 0; nphase ^= 1; }
          const int nk = min(TS_KB, num_kb - st * TS_KB);
          const uint64_t db = make_smem_desc(smem_base + stage * C::X_STAGE_BYTES, 16, 1024);
          const uint32_t ta = tmem_base + uint32_t(C::A_COL0 + as * 64);
          // the next stage's barrier is tested inside the block (top) and its result read after the stage's last MMA
          uint32_t ready_next;
          if (WIDE) {
            constexpr uint32_t KBD = uint32_t(C::KB_BYTES >> 4);   // descriptor units of one [sub-tile] box
            ready_next = ts_issue_kblock_wide(tmem_c, tmem_c + uint32_t(DW), ta, uint32_t(db), uint32_t(db) + KBD, uint32_t(db >> 32), idesc,
                                              uint32_t(st != 0), full_bar(nstage), nphase);
            if (nk == 2)
              ready_next |= ts_issue_kblock_wide(tmem_c, tmem_c + uint32_t(DW), ta + 32u, uint32_t(db) + 2u * KBD, uint32_t(db) + 3u * KBD,
                                                 uint32_t(db >> 32), idesc, 1u, full_bar(nstage), nphase);
          } else {
            ready_next =
                nk == 2 ? ts_issue_stage<2>(tmem_c, ta, uint32_t(db), uint32_t(C::KB_BYTES >> 4), uint32_t(db >> 32), idesc, uint32_t(st != 0),
                                            full_bar(nstage), nphase)
                        : ts_issue_stage<1>(tmem_c, ta, uint32_t(db), 0u, uint32_t(db >> 32), idesc, uint32_t(st != 0), full_bar(nstage), nphase);
          }
          TRC(p.trace, 1, 7000000 + trc_it - 1);
          umma_commit_pair(x_empty_bar(stage), 3);   // the stage's x slot AND its TMEM weight slot: one commit, two kinds of waiters
          TRC(p.trace, 1, 3000000 + trc_it - 1);
          if (st == num_st - 1) umma_commit_pair(tmem_full_bar(acc), 3);
          stage = nstage; phase = nphase;
          ready = ready_next;
          if (++as == NS) as = 0;
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    pdl_wait();
    // ===================================================== epilogue: lanes = output channels, columns = tokens
    // The four warps (one per TMEM lane quarter = 32 channels each) fill ONE staging tile [32 tokens][128 channels]
    // (256-byte rows: a quarter of the TMA rows that per-warp 64-byte boxes need) and warp 4 stores it with one TMA
    // store per 32 tokens; two staging tiles alternate.  The whole epilogue overlaps the next tile's main loop.
    const int ew = warp - 4;
    // the thread that issues (and later waits for) the TMA stores: elected once, so that both happen on the same lane and the
    // compiler knows the branch is taken by one lane (no per-instruction active-lane loop around the store)
    const bool el = elect_one();
    const bool epi_leader = el && warp == 4;
    TRC_DECL;
    int trc_it = 0;
    uint32_t chunk = 0;
    for (int tl = 0; tl < my_tiles; ++tl) {
      const int tile = pair + tl * num_pairs, acc = WIDE ? 0 : (tl & 1);
      const int n0 = TS_NBLK(tile) * 256 + int(rank) * 128;               // this CTA's 128 channels
      const int m_tile = TS_MBLK(tile) * T;
      const int n = n0 + ew * 32 + lane;
      float bias = 0.f;
      if (p.bias && n < p.N) {
        if (BF16) bias = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[n]);
        else bias = __half2float(reinterpret_cast<const __half*>(p.bias)[n]);
      }
      mbar_wait(tmem_full_bar(acc), uint32_t(WIDE ? tl : (tl >> 1)) & 1u);
      if (threadIdx.x == 128) TRC(p.trace, 2, 1000000 + trc_it);
      tc_fence_after();
      for (int sub = 0; sub < C::SUBS; ++sub) {
      const int m0 = m_tile + sub * Ts;                                   // first token of this accumulator
      const int t_end = min(Ts, p.M - m0);
      const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + uint32_t((WIDE ? sub : acc) * DW);
      for (int c = 0; c < t_end; c += 32, ++chunk) {
        uint32_t v[32];
        tmem_ld32(taddr + uint32_t(c), v);
        const uint32_t stg = epi_base + (chunk % uint32_t(C::EPI_TILES)) * 8192u;
        if (epi_leader) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(C::EPI_TILES - 1) : "memory");   // the store of EPI_TILES chunks ago has read this tile
        asm volatile("bar.sync 1, 128;" ::: "memory");
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float f = __uint_as_float(v[j]) + bias;
          uint16_t h;
          if (BF16) { __nv_bfloat16 bb = __float2bfloat16_rn(f); h = *reinterpret_cast<uint16_t*>(&bb); }
          else { __half bb = __float2half_rn(f); h = *reinterpret_cast<uint16_t*>(&bb); }
          asm volatile("st.shared.b16 [%0], %1;" ::"r"(stg + uint32_t(j) * 256u + uint32_t(ew * 32 + lane) * 2u), "h"(h) : "memory");
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (epi_leader && n0 < p.N) tma_store_2d(&map_y, stg, n0, m0 + c);   // 128 channels x 32 tokens, clipped at N and M
      }
      }
      tc_fence_before();
      __syncwarp();
      if (threadIdx.x == 128) TRC(p.trace, 2, 2000000 + trc_it);
      ++trc_it;
      if (lane == 0) mbar_arrive_cluster(leader_tmem_empty0 + 8u * acc);
    }
    if (epi_leader) tma_store_wait_all();
  } else if (warp >= 8) {
    // ===================================================== dequant: registers -> TMEM
    const int set = (warp - 8) >> 2, q = warp & 3;                       // TMEM lane quarter = warp % 4
    uint32_t mask_lo = 0x000F000Fu, mask_hi = 0x00F000F0u;
    uint32_t magic = BF16 ? 0x43004300u : 0x64006400u;
    asm volatile("" : "+r"(mask_lo), "+r"(mask_hi), "+r"(magic));
    auto and_or = [](uint32_t a, uint32_t b, uint32_t c) {
      uint32_t d;
      asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));   // (a & b) | c
      return d;
    };
    struct Pf { uint4 w0, w1; uint32_t sz; };   // one k-block of this thread's channel: 8 packed words, (scale, zero)
    Pf ring[TS_DIST][TS_KB];
    const int total = my_tiles * num_st;                                 // stages of this pair, global sequence
    const int mine = (total - set + 1) / 2;                              // ... of which this set takes every other one
    // prefetch cursor: the set's next stage to load = global stage pf_g = (pf_tl, pf_st)
    int pf_g = set, pf_tl = 0, pf_st = set;
    while (pf_st >= num_st) { pf_st -= num_st; ++pf_tl; }
    auto prefetch = [&](Pf (&f)[TS_KB]) {
#pragma unroll
      for (int j = 0; j < TS_KB; ++j) { f[j].w0 = make_uint4(0, 0, 0, 0); f[j].w1 = f[j].w0; f[j].sz = 0; }
#ifdef TS_NO_W    // timing experiment only (wrong results): no packed-weight loads
      if (false) {
#else
      if (pf_g < total) {
#endif
        const int tile = pair + pf_tl * num_pairs;
        const int n = TS_NBLK(tile) * 256 + int(rank) * 128 + q * 32 + lane;
        if (n < p.N) {
#pragma unroll
          for (int j = 0; j < TS_KB; ++j) {
            const int kb = pf_st * TS_KB + j;
            if (kb < num_kb) {
              const int64_t idx = int64_t(kb) * p.N + n;
              const uint4* src = reinterpret_cast<const uint4*>(p.words + idx * 8);
              f[j].w0 = __ldg(src);
              f[j].w1 = __ldg(src + 1);
              f[j].sz = __ldg(p.sz + idx);
            }
          }
        }
      }
      pf_g += 2; pf_st += 2;
      while (pf_st >= num_st) { pf_st -= num_st; ++pf_tl; }
    };
#pragma unroll
    for (int d = 0; d < TS_DIST; ++d) prefetch(ring[d]);
    auto unpack = [&](const Pf& f, uint32_t (&o)[32]) {
      const uint32_t s2 = (f.sz & 0xFFFFu) * 0x10001u;                   // (s, s)
      const uint32_t zz = (f.sz >> 16) * 0x10001u;                       // (z, z) as dtype values
      uint32_t zlo, zhi = 0;
      if (BF16) {
        const __nv_bfloat162 m = __float2bfloat162_rn(128.f);
        __nv_bfloat162 t = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&zz), m);   // 128 + z, exact
        zlo = *reinterpret_cast<uint32_t*>(&t);
      } else {
        __half2 t = __hadd2(*reinterpret_cast<const __half2*>(&zz), __float2half2_rn(1024.f));   // 1024 + z
        __half2 u = __hadd2(*reinterpret_cast<const __half2*>(&zz), __float2half2_rn(64.f));     // 64 + z
        zlo = *reinterpret_cast<uint32_t*>(&t);
        zhi = *reinterpret_cast<uint32_t*>(&u);
      }
      const uint32_t wsrc[8] = {f.w0.x, f.w0.y, f.w0.z, f.w0.w, f.w1.x, f.w1.y, f.w1.z, f.w1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t wv = wsrc[j];
        if (BF16) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t t = and_or(wv >> (4 * c), mask_lo, magic);    // {128 + q(k = 8j + 2c), 128 + q(k + 1)}
            __nv_bfloat162 d = __hsub2(*reinterpret_cast<const __nv_bfloat162*>(&t), *reinterpret_cast<const __nv_bfloat162*>(&zlo));
            d = __hmul2(d, *reinterpret_cast<const __nv_bfloat162*>(&s2));
            o[4 * j + c] = *reinterpret_cast<uint32_t*>(&d);
          }
        } else {
          const uint32_t ws = wv >> 8;
          const __half2 sixteenth = __float2half2_rn(0.0625f);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t src = (c < 2) ? wv : ws;
            __half2 d;
            if ((c & 1) == 0) {   // nibbles c, c + 4 -> 1024 + q, exact subtract
              const uint32_t t = and_or(src, mask_lo, magic);
              d = __hsub2(*reinterpret_cast<const __half2*>(&t), *reinterpret_cast<const __half2*>(&zlo));
            } else {              // 1024 + 16 q; fma(., 1/16, -(64 + z)) = q - z exactly
              const uint32_t t = and_or(src, mask_hi, magic);
              d = __hfma2(*reinterpret_cast<const __half2*>(&t), sixteenth, __hneg2(*reinterpret_cast<const __half2*>(&zhi)));
            }
            d = __hmul2(d, *reinterpret_cast<const __half2*>(&s2));      // (q - z) * s, one rounding
            o[4 * j + c] = *reinterpret_cast<uint32_t*>(&d);
          }
        }
      }
    };
    int as = set % NS, fs = set % NXS;                                   // TMEM slot, full-barrier slot of the set's next stage
    int wg = set - NS;                                                   // the global stage whose MMAs last read that TMEM slot
    TRC_DECL;
    int trc_it = 0;
    auto process = [&](Pf (&f)[TS_KB]) {
      // BOTH k-blocks are unpacked before the TMEM slot is waited for: with a two-slot weight ring (DW = 192) the slot
      // frees only when the MMAs two stages back have finished, and everything after the wait is exposed latency
      // (measured: 780 cycles from slot-free to arrive when the second k-block was unpacked after the wait)
      uint32_t o[TS_KB][32];
      if (q == 0 && lane == 0) TRC(p.trace, 3 + set, 1000000 + trc_it);
#pragma unroll
      for (int j = 0; j < TS_KB; ++j) unpack(f[j], o[j]);
      // the packed words are dead once unpacked: load the set's NEXT stage into the same registers now, so the loads fly
      // while this stage waits for its TMEM slot, stores and arrives
      prefetch(f);
      // slot free = the MMAs of stage wg = g - NS are complete = phase wg / NXS of that stage's release barrier.  The phase
      // after it needs this very stage to be consumed first (NXS > NS), so the waiter never sees the barrier two phases on.
      if (wg >= 0) mbar_wait(x_empty_bar(wg % NXS), uint32_t(wg / NXS) & 1u);
      if (q == 0 && lane == 0) TRC(p.trace, 3 + set, 2000000 + trc_it);
      tc_fence_after();
      const uint32_t a_addr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(C::A_COL0 + as * 64);
#pragma unroll
      for (int j = 0; j < TS_KB; ++j) tmem_st32(a_addr + uint32_t(j) * 32u, o[j]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_full0 + 8u * fs);
      if (q == 0 && lane == 0) TRC(p.trace, 3 + set, 3000000 + trc_it);
      ++trc_it;
      as += 2;
      if (as >= NS) as -= NS;
      wg += 2;
      fs += 2;
      if (fs >= NXS) fs -= NXS;
    };
    for (int it = 0; it < mine; it += TS_DIST) {
#pragma unroll
      for (int u = 0; u < TS_DIST; ++u) {
        if (it + u < mine) {
          process(ring[u]);
        }
      }
    }
  }

  tc_fence_before();
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

template <int DW, bool BF16, bool WIDE = false>
int launch_ts(const CUtensorMap& mx, const CUtensorMap& my, const TsParams& p, cudaStream_t st) {
  using C = CfgTS<DW, WIDE>;
  auto kern = qdm_w4ts_kernel<DW, BF16, WIDE>;
  static bool attr_set = false;
  if (!attr_set) {
    QDM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int64_t tiles = int64_t((p.N + 255) / 256) * ((p.M + p.tile_t - 1) / p.tile_t);
  const int pairs = int(tiles < QDM_NUM_SMS / 2 ? tiles : QDM_NUM_SMS / 2);
#ifdef QDM_TRACE
  if (getenv("QDM_TRACE")) {   // debug build only: per-role clock64 timeline of the first CTA pair (tools/trace_view.py)
    static long long* tbuf = nullptr;
    if (!tbuf) cudaMalloc(&tbuf, 16 * 2048 * sizeof(long long));
    cudaMemset(tbuf, 0, 16 * 2048 * sizeof(long long));
    TsParams pt = p;
    pt.trace = tbuf;
    kern<<<2 * pairs, TS_THREADS, C::SMEM_BYTES, st>>>(mx, my, pt);
    cudaDeviceSynchronize();
    static long long host[16 * 2048];
    cudaMemcpy(host, tbuf, sizeof(host), cudaMemcpyDeviceToHost);
    fprintf(stderr, "QDMTRACE begin M=%d N=%d K=%d tile_t=%d (TS)\n", p.M, p.N, p.K, p.tile_t);
    for (int r = 0; r < 16; ++r)
      for (int i = 0; i < 1000 && host[r * 2048 + 2 * i]; ++i)
        fprintf(stderr, "QDMTRACE %d %lld %lld\n", r, host[r * 2048 + 2 * i], host[r * 2048 + 2 * i + 1]);
    QDM_LAUNCH_CHECK();
    return QDM_OK;
  }
#endif
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(2 * pairs));
  cfg.blockDim = dim3(TS_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  static const bool no_pdl = getenv("QDM_NO_PDL") != nullptr;   // A/B switch, read once
  cfg.attrs = &attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  QDM_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, mx, my, p));
  QDM_LAUNCH_CHECK();
  return QDM_OK;
}

}  // namespace

// layout of the TS blob: [words: K/64 x N x 32 bytes][sz: K/64 x N x 4 bytes]
extern "C" size_t qdm_w4a16_repack_ts_bytes(int64_t N, int64_t K) {
  if (N <= 0 || K <= 0 || K % 64) return 0;
  return size_t(K / 64) * size_t(N) * 36;
}

extern "C" int qdm_w4a16_repack_ts(const int32_t* qweight, const int32_t* qzeros, const void* scales, int dtype, int64_t N, int64_t K,
                                   int group, void* blob, size_t blob_bytes, void* stream) {
  QDM_REQUIRE(qweight && qzeros && scales && blob, "qdm_w4a16_repack_ts: null pointer");
  QDM_REQUIRE(dtype == QDM_F16 || dtype == QDM_BF16, "qdm_w4a16_repack_ts: dtype must be f16 or bf16");
  QDM_REQUIRE(N > 0 && K > 0 && N % 8 == 0 && K % 64 == 0 && N < (1LL << 31) && K < (1LL << 31),
              "qdm_w4a16_repack_ts: N=%lld must be a multiple of 8 and K=%lld of 64", (long long)N, (long long)K);
  QDM_REQUIRE(group > 0 && group % 64 == 0 && K % group == 0, "qdm_w4a16_repack_ts: group=%d must be a multiple of 64 dividing K", group);
  QDM_REQUIRE(blob_bytes >= qdm_w4a16_repack_ts_bytes(N, K) && qdm_aligned16(blob), "qdm_w4a16_repack_ts: blob needs %zu bytes, 16-byte aligned",
              qdm_w4a16_repack_ts_bytes(N, K));
  QDM_DEVICE_GATE();
  const int64_t items = (K / 64) * N;
  uint32_t* words = static_cast<uint32_t*>(blob);
  uint32_t* sz = words + items * 8;
  const unsigned grid = unsigned((items + 255) / 256);
  if (dtype == QDM_BF16)
    w4ts_repack_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(qweight, qzeros, static_cast<const uint16_t*>(scales), int(N), int(K), group, words, sz);
  else
    w4ts_repack_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(qweight, qzeros, static_cast<const uint16_t*>(scales), int(N), int(K), group, words, sz);
  QDM_LAUNCH_CHECK();
  return QDM_OK;
}

// Called by the W4A16 dispatcher (qdm_gemm.cu) once it has picked the TS path with `tile_t` tokens per tile.
int qdm_w4ts_gemm(const void* x, const void* blob, const void* bias, void* y, int is_bf16, int64_t M, int64_t N, int64_t K, int tile_t,
                  cudaStream_t st) {
  QDM_REQUIRE(blob && qdm_aligned16(blob), "qdm_gemm_w4a16_ts: the repacked weight must be 16-byte aligned");
  // tile_t <= 192: two alternating tiles of tile_t tokens; 256 / 320 / 384: ONE tile of two sub-tiles of tile_t / 2 tokens (WIDE)
  const bool wide = tile_t > 192;
  QDM_REQUIRE(tile_t >= 32 && tile_t <= 384 && tile_t % (wide ? 64 : 32) == 0, "qdm_gemm_w4a16_ts: bad token tile %d", tile_t);
  QDM_REQUIRE(K % 64 == 0 && N % 8 == 0, "qdm_gemm_w4a16_ts: shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  int rc = get_encode_fn();
  if (rc) return rc;
  CUtensorMap mx, my;
  if ((rc = make_map(&mx, x, 2, M, K, 64, wide ? tile_t / 4 : tile_t / 2))) return rc;   // 64 k x token rows per CTA (and sub-tile), SWIZZLE_128B
  if ((rc = make_map(&my, y, 2, M, N, 128, 32, false))) return rc;        // 128 channels x 32 tokens, dense 256-byte rows
  TsParams p{};
  p.M = int(M); p.N = int(N); p.K = int(K); p.tile_t = tile_t; p.bias = bias; p.y = y;
  p.words = static_cast<const uint32_t*>(blob);
  p.sz = p.words + (K / 64) * N * 8;
  {
    // panel height of the tile order: the panel's activations must survive in L2 while the panel's channel blocks pass
    // over them (32 MB of the 126 MB: the weights, the output stream and the other die's copies share it); at least 8
    // token blocks so that the (L2-resident) weights are not re-read more often than necessary.
    // QDM_W4_TS_GM: A/B knob, read once (-1: no panels, n > 0: that many token blocks).
    static const int gm_env = getenv("QDM_W4_TS_GM") ? atoi(getenv("QDM_W4_TS_GM")) : 0;
    const int64_t m_blks = (M + tile_t - 1) / tile_t, per_blk = int64_t(tile_t) * K * 2;
    int64_t gm = (int64_t(32) << 20) / per_blk;
    if (gm < 8) gm = 8;
    if (gm_env > 0) gm = gm_env;
    if (gm > m_blks || gm_env < 0) gm = m_blks;
    p.group_m = int(gm);
  }
  if (wide) {
    if (tile_t <= 256) return is_bf16 ? launch_ts<128, true, true>(mx, my, p, st) : launch_ts<128, false, true>(mx, my, p, st);
    if (tile_t <= 320) return is_bf16 ? launch_ts<160, true, true>(mx, my, p, st) : launch_ts<160, false, true>(mx, my, p, st);
    return is_bf16 ? launch_ts<192, true, true>(mx, my, p, st) : launch_ts<192, false, true>(mx, my, p, st);
  }
  if (tile_t <= 128) return is_bf16 ? launch_ts<128, true>(mx, my, p, st) : launch_ts<128, false>(mx, my, p, st);
  if (tile_t <= 160) return is_bf16 ? launch_ts<160, true>(mx, my, p, st) : launch_ts<160, false>(mx, my, p, st);
  return is_bf16 ? launch_ts<192, true>(mx, my, p, st) : launch_ts<192, false>(mx, my, p, st);
}
