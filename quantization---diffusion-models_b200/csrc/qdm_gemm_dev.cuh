// Device-side building blocks shared by the tcgen05 GEMM kernels of libqdm (qdm_gemm.cu, qdm_gemm_w4rp.cu):
// PTX wrappers (mbarrier, TMA, tcgen05, cluster), descriptors, GemmParams and the TMEM -> TMA-store epilogue.
#pragma once
#include "qdm_common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>

namespace qdmg {

// Debug timeline (compile with -DQDM_TRACE, run with QDM_TRACE=1): every role of block 0 logs (tag, clock64) pairs.
#ifdef QDM_TRACE
#define TRC_DECL int trc_n = 0
#define TRC(ptr, region, tag)                                            \
  do {                                                                   \
    if ((ptr) && blockIdx.x < 2 && trc_n < 1000) {                       \
      long long* t_ = (ptr) + ((region) + 8 * blockIdx.x) * 2048 + 2 * trc_n; \
      t_[0] = (tag);                                                     \
      t_[1] = clock64();                                                 \
      ++trc_n;                                                           \
    }                                                                    \
  } while (0)
// wall-clock companion (ns): two TRCG events around a region give the SM clock the region actually ran at
#define TRCG(ptr, region, tag)                                           \
  do {                                                                   \
    if ((ptr) && blockIdx.x < 2 && trc_n < 1000) {                       \
      long long* t_ = (ptr) + ((region) + 8 * blockIdx.x) * 2048 + 2 * trc_n; \
      unsigned long long g_;                                             \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_));             \
      t_[0] = (tag);                                                     \
      t_[1] = (long long)g_;                                             \
      ++trc_n;                                                           \
    }                                                                    \
  } while (0)
#else
#define TRC_DECL
#define TRC(ptr, region, tag)
#define TRCG(ptr, region, tag)
#endif

enum GemmKind { G_F16 = 0, G_F16_KN = 1, G_W4 = 2, G_I8 = 3 };

constexpr int BLOCK_M = 128;
constexpr int ROW_BYTES = 128;                 // one swizzle-128B row = one k-block of an operand row
constexpr int A_STAGE_BYTES = BLOCK_M * ROW_BYTES;
constexpr int EPI_COLS = 64;                   // accumulator columns per epilogue chunk = one 128-byte output row segment
constexpr int EPI_STG_BYTES = 32 * 128;        // per-warp TMA-store staging: 32 rows x 128 B, SWIZZLE_128B
constexpr int EPI_VEC_BYTES = 256 * 4;         // per-warp fp32 copy of the tile's bias (and of the W8A8 column scales)
constexpr int NUM_DQ_WARPS = 8;

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint32_t bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
// Bounded wait: a pipeline bug traps (kernel error) instead of hanging the GPU.
#define mbar_wait(bar, parity) mbar_wait_(bar, parity, __LINE__)
__device__ __forceinline__ void mbar_wait_(uint32_t bar, uint32_t parity, int line) {
  uint32_t done = 0;
  uint32_t polls = 0;
  long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++polls == 4096) t0 = clock64();
    if (polls > 4096 && (polls & 1023) == 0 && clock64() - t0 > 2000000000LL) {
      printf("qdm mbar_wait timeout: line %d block %d thread %d bar 0x%x parity %u\n", line, blockIdx.x, threadIdx.x, bar, parity);
#ifdef QDM_DEBUG_WAIT
      return;
#else
      __trap();
#endif
    }
  }
}
// Non-blocking phase test: 1 if the phase with this parity has completed.  The MMA issuer PEEKS at the next stage's barrier
// before issuing the current stage's tcgen05.mma: the result is consumed only after those are queued, so the ~250 cycles
// of barrier latency overlap tensor-pipe work instead of idling it once per stage.
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}
// One thread of a converged warp, the way the compiler understands it: a branch on elect.sync's predicate is known to be
// taken by exactly one lane, so instructions with uniform-register operands inside it (tcgen05.mma / commit, TMA) need no
// per-instruction "for each active lane" loop (ELECT + BRA.U.ANY around every UTCHMMA when the condition is `lane == 0`).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// 4-D tiled load (channels, w, h, image) for the direct implicit-GEMM convolution: coordinates outside the tensor are
// zero-filled, which IS the convolution's zero padding
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  if (KIND == G_I8) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
  }
}
// arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- programmatic dependent launch: the next kernel of the stream may be scheduled while this one drains (its CTAs
// start as SMs free up, its prologue overlaps our tail); it must not touch global memory before pdl_wait().
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- cluster / CTA-pair helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Remote arrive with the default (.release.cta) semantics, as CUTLASS' ClusterBarrier::arrive(cta_id) does.
// A .release.cluster arrive compiles to MEMBAR.ALL.GPU and an .acquire.cluster wait to CCTL.IVALL; measured: they
// halve the throughput of the CTA-pair kernel.  What is published here is shared memory of the ARRIVING CTA, already
// made visible to its own async proxy by fence.proxy.async (or TMEM reads completed by tcgen05.wait::ld), and it is
// consumed by that same SM's tensor core / by the leader's next tcgen05.mma, so CTA scope is sufficient.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// CTA-pair TMA load: data lands in THIS CTA's shared memory, completion bytes are counted on `bar`, a
// shared::cluster address (the leader CTA's full barrier)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// CTA-pair TMA load multicast to the CTAs of `mask` (same CTA-relative destination offset in each); the completion
// bytes of every destination CTA are counted on the barrier at `bar`'s offset in the leader (even rank) of THAT CTA's pair
__device__ __forceinline__ void tma_load_2d_pair_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
      ::"r"(dst), "l"(map), "r"(bar), "h"(mask), "r"(c0), "r"(c1)
      : "memory");
}
template <int KIND>
__device__ __forceinline__ void umma_pair(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  if (KIND == G_I8) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
  }
}
// One k-block (4 MMAs over a 128-byte operand row) for the issuing thread, shared-memory operands (SS form).  Why one asm
// block: measured on the issuing thread (QDM_TRACE timeline, profiles/README.md round 2) a barrier test costs ~290 cycles
// when its predicate is read at once, an MMA ~80 cycles when its two descriptors are rebuilt from addresses, a commit ~170
// -- ~870 cycles of serial issue per k-block against 512 cycles of tensor-pipe work at N = 256, so the thread, not the
// tensor pipe, set the pace of every kernel.  Here the NEXT stage's barrier is tested first and its predicate read after the
// fourth MMA (the latency overlaps the MMAs), and each descriptor is the previous one plus a constant in the low word
// (a: 32 bytes >> 4; b: 32 bytes >> 4 K-major, 2048 bytes >> 4 MN-major).  Returns 1 if the next stage was already full.
template <int KIND, bool PAIR>
__device__ __forceinline__ uint32_t issue_kblock_ss(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t b_step, uint32_t idesc,
                                                   uint32_t accum, uint32_t next_bar, uint32_t next_parity) {
  uint32_t done;
#define QDM_KBLOCK_ASM(MMA)                                                                                          \
  asm volatile(                                                                                                      \
      "{\n\t.reg .pred pa, pt, pd;\n\t.reg .b64 xa, xb;\n\t.reg .b32 la, lb;\n\t"                                  \
      "mbarrier.test_wait.parity.shared::cta.b64 pd, [%9], %10;\n\t"                                                 \
      "setp.ne.b32 pa, %8, 0;\n\t"                                                                                   \
      "setp.eq.b32 pt, %8, %8;\n\t"                                                                                  \
      "mov.b64 xa, {%2, %3};\n\tmov.b64 xb, {%4, %5};\n\t"                                                          \
      MMA " [%1], xa, xb, %7, pa;\n\t"                                                                               \
      "add.u32 la, %2, 2;\n\tadd.u32 lb, %4, %6;\n\tmov.b64 xa, {la, %3};\n\tmov.b64 xb, {lb, %5};\n\t"            \
      MMA " [%1], xa, xb, %7, pt;\n\t"                                                                               \
      "add.u32 la, la, 2;\n\tadd.u32 lb, lb, %6;\n\tmov.b64 xa, {la, %3};\n\tmov.b64 xb, {lb, %5};\n\t"            \
      MMA " [%1], xa, xb, %7, pt;\n\t"                                                                               \
      "add.u32 la, la, 2;\n\tadd.u32 lb, lb, %6;\n\tmov.b64 xa, {la, %3};\n\tmov.b64 xb, {lb, %5};\n\t"            \
      MMA " [%1], xa, xb, %7, pt;\n\t"                                                                               \
      "selp.u32 %0, 1, 0, pd;\n\t}"                                                                                  \
      : "=r"(done)                                                                                                   \
      : "r"(tmem_c), "r"(uint32_t(da)), "r"(uint32_t(da >> 32)), "r"(uint32_t(db)), "r"(uint32_t(db >> 32)), "r"(b_step),  \
        "r"(idesc), "r"(accum), "r"(next_bar), "r"(next_parity)                                                      \
      : "memory")
  if (PAIR) {
    if (KIND == G_I8) QDM_KBLOCK_ASM("tcgen05.mma.cta_group::2.kind::i8");
    else QDM_KBLOCK_ASM("tcgen05.mma.cta_group::2.kind::f16");
  } else {
    if (KIND == G_I8) QDM_KBLOCK_ASM("tcgen05.mma.cta_group::1.kind::i8");
    else QDM_KBLOCK_ASM("tcgen05.mma.cta_group::1.kind::f16");
  }
#undef QDM_KBLOCK_ASM
  return done;
}
// commit to the same barrier offset in the CTAs of `mask` (both CTAs of the pair; all four CTAs of a quad cluster)
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (sm_100): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((addr >> 4) & 0x3FFF);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}
// Instruction descriptor: c_format [4,6) | a_format [7,10) | b_format [10,13) | a_major 15 | b_major 16 |
// N>>3 [17,23) | M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int c_fmt, int ab_fmt, int b_mn_major, int M, int N) {
  return (uint32_t(c_fmt) << 4) | (uint32_t(ab_fmt) << 7) | (uint32_t(ab_fmt) << 10) | (uint32_t(b_mn_major) << 16) |
         (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

struct GemmParams {
  int M, N, K;
  int group;                 // G_W4
  int tile_n;                // columns per output tile (multiple of 16, <= the BLOCK_N the kernel was built for)
  const int32_t* qweight;    // G_W4  [K, N/8]
  const int32_t* qzeros;     // G_W4  [K/group, N/8]
  const void* scales;        // G_W4  [K/group, N] dtype
  const void* bias;          // [N] out dtype or null
  const float* sx;           // G_I8  [M]
  const float* sw;           // G_I8  [N]
  void* y;                   // [M, N] out dtype
  int is_bf16;               // element / output type: 0 fp16, 1 bf16
  long long* trace;          // QDM_TRACE builds only: per-role (tag, clock64) event log of block 0
  // 3x3 convolution as an implicit GEMM over a zero-padded NHWC grid (qdm_conv3x3_*): K = 9 * conv_cin, the A rows of
  // k-block kb come from the activation rows shifted by conv_off[tap], tap = kb * 64 / conv_cin.  0 = plain GEMM.
  int conv_cin;
  int conv_off[9];
  // direct form (conv_w != 0): A is the UNPADDED NHWC tensor behind a 4-D tensor map (c, w, h, image); a 128-row tile is
  // a box of whole image rows (W divides 128) and tap (dy, dx) shifts its (w, h) start by (dx-1, dy-1) -- the hardware's
  // out-of-bounds zero fill is the padding, so there is neither a padded copy nor a wasted border row.
  int conv_w, conv_h;
  // stride of the direct form (1 or 2): conv_w / conv_h are then the OUTPUT grid, the box start moves by conv_stride
  // input rows per output row and the tensor map's element strides skip every other pixel (qdm_conv3x3s2_*)
  int conv_stride;
  const uint8_t* rp_blob;    // W4 repacked weights (qdm_w4a16_repack): [K/128][rp_nb] blocks of RP_BLK_BYTES
  int rp_nb;                 // 16-column blocks per k-group row of the blob = ceil(N / 16)
  float* sk_data;            // stream-K: partial accumulators, [pair][rank][128 rows][256] fp32
  uint32_t* sk_flags;        // stream-K: [pair][rank][4 warps], 0 = empty, 1 = partial written (reset by its reader)
};

// Packed-int4 staging ring of the W4 kernels: per k-block the TMA producer drops the tile part's packed words
// (64 k rows x NLOC/8 words), its NLOC scales and NLOC/8 zero-point words here; the dequant warps read them from
// shared memory, so they never have global loads outstanding when they reach fence.proxy.async.
constexpr int RAW_STAGES = 4;
template <int NLOC>
struct RawCfg {
  // One raw stage = TWO k-blocks (128 k rows): a thread issues a TMA every ~250 cycles, so the packed operands must
  // cost fewer than one TMA per k-block each (measured: 4 TMAs per k-block made the producer the bottleneck).
  // TMA needs a 16-byte aligned box start: the word box starts at the tile part's first word rounded DOWN to a
  // multiple of 4 words (32 columns) and is 4 words wider; the dequant threads add the remainder (0..3 words).
  static constexpr int WPRX = NLOC / 8 + 4;            // words per staged k row
  static constexpr int QW_BYTES = 128 * WPRX * 4;      // [128 k rows][WPRX]
  static constexpr int SC_ROW_BYTES = NLOC * 2;        // up to 2 quantisation groups per stage (group 64)
  static constexpr int SC_BYTES = 2 * SC_ROW_BYTES;
  static constexpr int ZW_ROW_BYTES = WPRX * 4;
  static constexpr int ZW_BYTES = 2 * ZW_ROW_BYTES;
  static constexpr int BYTES = (QW_BYTES + SC_BYTES + ZW_BYTES + 127) / 128 * 128;
  static_assert(QW_BYTES % 128 == 0 && SC_BYTES % 128 == 0, "TMA destinations must stay 128-byte aligned");
  __host__ __device__ static constexpr int tx_bytes(int srows) { return QW_BYTES + srows * (SC_ROW_BYTES + ZW_ROW_BYTES); }
};

template <int BLOCK_N, int KIND>
struct Cfg {
  static constexpr int B_STAGE_BYTES = BLOCK_N * ROW_BYTES;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int EPI_BYTES = 4 * (EPI_STG_BYTES + EPI_VEC_BYTES * (KIND == G_I8 ? 2 : 1));
  static constexpr int RAW_N = 2;            // raw stages in use (of the RAW_STAGES barrier slots)
  static constexpr int RAW_BYTES = (KIND == G_W4) ? RAW_N * RawCfg<BLOCK_N>::BYTES : 0;
  static constexpr int STAGES = (227 * 1024 - 2048 - EPI_BYTES - RAW_BYTES) / STAGE_BYTES > 8 ? 8 : (227 * 1024 - 2048 - EPI_BYTES - RAW_BYTES) / STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + RAW_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int THREADS = (KIND == G_W4) ? 512 : 256;
  static constexpr int TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256) ? 256 : 512;
  static constexpr int K_PER_BLOCK = (KIND == G_I8) ? 128 : 64;   // elements per k-block (128 bytes)
  static constexpr int FULL_COUNT = (KIND == G_W4) ? 1 + NUM_DQ_WARPS / 2 : 1;   // TMA + one producer group (global-load path)
  static constexpr int RAW_GROUPS = 2;                                            // dequant groups of the raw-ring path
  static constexpr int FULL_COUNT_RAW = 1 + NUM_DQ_WARPS / RAW_GROUPS;
  static constexpr int RAW_EMPTY_COUNT = 2 * NUM_DQ_WARPS / RAW_GROUPS;           // the warps of the stage's two k-blocks
};

template <int KIND, bool BF16>
__device__ __forceinline__ uint32_t pack_out2(float a, float b) {
  if (BF16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}

template <bool BF16>
__device__ __forceinline__ void load8_as_float(const void* base, int64_t idx, bool ok, float* out) {
  if (!base || !ok) {
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = 0.f;
    return;
  }
  uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(base) + idx);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (BF16) {
      __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
      out[2 * i] = __low2float(h);
      out[2 * i + 1] = __high2float(h);
    } else {
      __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      out[2 * i] = __low2float(h);
      out[2 * i + 1] = __high2float(h);
    }
  }
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Drain (a share of) one accumulator tile: this warp owns 32 TMEM lanes (= 32 output rows starting at `row0`) and
// takes the 64-column chunks c_first, c_first + c_step, ...  Per chunk: tcgen05.ld (32-column halves, double
// buffered in registers so that the load of the next half is in flight while this one is converted) -> scale /
// bias in fp32 from a per-warp shared copy staged once per tile -> pack -> shared staging -> TMA store:
//   * whole chunk: ONE store of the 32 x 64 box from a SWIZZLE_128B staging tile (map_y);
//   * trailing chunk narrower than 64 columns: one store per 16-column slice (map_y16, dense 32 x 32 B slices).
// Rows past M and columns past N are clipped by the tensor maps.  The stores are asynchronous; the staging buffer
// is reclaimed by wait_group.read when the next chunk needs it.
// Measured before this (padded transpose + 8 x (ld.shared, predicated st.global) per chunk, bias from global per
// 8 columns, no load/convert overlap): 5500 cycles per 128 x 160 tile part -- the critical path of every K <= 640 shape.
template <int BLOCK_N, int KIND, bool BF16>
__device__ __forceinline__ void epilogue_drain(const GemmParams& p, const CUtensorMap* map_y, const CUtensorMap* map_y16,
                                               uint32_t stg, float* vec_sm, uint32_t taddr0, int row0, int n0, int lane,
                                               int c_first, int c_step, const float* part_row = nullptr, int n_parts = 0,
                                               int64_t part_stride = 0, int tile_w = 0) {
  // stream-K: part_row + q * part_stride is this lane's entry in the q-th earlier partial accumulator of the tile (fp32,
  // layout [half][16-byte chunk][128 rows], see qdm_gemm2_sk_kernel); the partials are added in slot order, so the
  // result does not depend on timing
  // the lane that issues and later waits for this warp's TMA stores: elect.sync tells the compiler that one lane runs those
  // branches (no active-lane loop around the store); `&& lane == 0` pins the choice so that successive calls agree on it
  const bool store_leader = elect_one() && lane == 0;
  const int n_end = min(n0 + (tile_w ? tile_w : p.tile_n), p.N);   // columns of this (sub-)tile that exist
  const int row = row0 + lane;
  float sxr = 1.f;
  if (KIND == G_I8) sxr = (row < p.M) ? p.sx[row] : 0.f;
  {  // lane l stages columns n0 + 8 l .. + 8 (whole groups of 8 are inside or outside: N % 8 == 0, tile_n % 16 == 0)
    const int n = n0 + lane * 8;
    const bool ok = n < n_end;
    float b8[8];
    load8_as_float<BF16>(p.bias, n, ok, b8);
    __syncwarp();   // every lane is done reading the previous tile's copy
    *reinterpret_cast<float4*>(vec_sm + lane * 8) = make_float4(b8[0], b8[1], b8[2], b8[3]);
    *reinterpret_cast<float4*>(vec_sm + lane * 8 + 4) = make_float4(b8[4], b8[5], b8[6], b8[7]);
    if (KIND == G_I8) {
      float4 s0 = make_float4(0, 0, 0, 0), s1 = s0;
      if (ok) {
        s0 = *reinterpret_cast<const float4*>(p.sw + n);
        s1 = *reinterpret_cast<const float4*>(p.sw + n + 4);
      }
      *reinterpret_cast<float4*>(vec_sm + 256 + lane * 8) = s0;
      *reinterpret_cast<float4*>(vec_sm + 256 + lane * 8 + 4) = s1;
    }
    __syncwarp();
  }
  auto halves_in = [&](int c) {   // 32-column halves of chunk c that hold columns of the tile
    const int w = n_end - (n0 + c * EPI_COLS);
    return w <= 0 ? 0 : (w > 32 ? 2 : 1);
  };
  auto process_half = [&](const uint32_t (&v)[32], int h, auto with_ex, const uint32_t (&ex)[32]) {
    constexpr bool kEx = decltype(with_ex)::value;           // ex = fp32 bits to add (stream-K partial sums)
    const int c = h >> 1, hh = h & 1;
    const int nc = n0 + c * EPI_COLS;                       // first column of the 64-column chunk
    const int width = min(EPI_COLS, n_end - nc);            // columns of the chunk inside the tile
    const bool whole = width == EPI_COLS;
    if (hh == 0) {   // the previous stores of this warp must have read the staging buffer before it is overwritten
      if (store_leader) tma_store_wait_read();
      __syncwarp();
    }
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
      const int col = h * 32 + j8 * 8;                       // column inside the tile
      const float4 b0 = *reinterpret_cast<const float4*>(vec_sm + col), b1 = *reinterpret_cast<const float4*>(vec_sm + col + 4);
      const float bias8[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float f[8];
      if (KIND == G_I8) {
        const float4 s0 = *reinterpret_cast<const float4*>(vec_sm + 256 + col), s1 = *reinterpret_cast<const float4*>(vec_sm + 256 + col + 4);
        const float sw8[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
          f[i] = __fmaf_rn(__fmul_rn(float(int(v[j8 * 8 + i])), __fmul_rn(sxr, sw8[i])), 1.f, bias8[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          f[i] = (kEx ? __uint_as_float(v[j8 * 8 + i]) + __uint_as_float(ex[j8 * 8 + i]) : __uint_as_float(v[j8 * 8 + i])) + bias8[i];
      }
      uint4 o;
      o.x = pack_out2<KIND, BF16>(f[0], f[1]);
      o.y = pack_out2<KIND, BF16>(f[2], f[3]);
      o.z = pack_out2<KIND, BF16>(f[4], f[5]);
      o.w = pack_out2<KIND, BF16>(f[6], f[7]);
      const uint32_t g16 = uint32_t(hh * 4 + j8);            // 16-byte column group inside the chunk
      const uint32_t dst = whole ? stg + uint32_t(lane) * 128u + ((g16 ^ uint32_t(lane & 7)) << 4)
                                 : stg + (g16 >> 1) * 1024u + uint32_t(lane) * 32u + (g16 & 1u) * 16u;
      if (whole || int(g16) * 8 < width)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
    }
    if (hh == halves_in(c) - 1) {   // chunk complete
      fence_proxy_async();
      __syncwarp();
#ifdef QDM_EXP_NOSTORE   // timing experiment only (no output): what the epilogue's TMA stores cost
      if (false) {
#else
      if (store_leader) {
#endif
        if (whole) {
          tma_store_2d(map_y, stg, nc, row0);
        } else {
          for (int sl = 0; sl * 16 < width; ++sl) tma_store_2d(map_y16, stg + uint32_t(sl) * 1024u, nc + sl * 16, row0);
        }
      }
    }
  };
  // walk the halves of this warp's chunks with two register buffers
  int c = c_first;
  if (halves_in(c) == 0) return;
  int h = 2 * c;
  auto next_half = [&](int& hn, int& cn) {   // successor of half h of chunk c, or hn = -1
    if ((h & 1) == 0 && halves_in(c) == 2) { hn = h + 1; cn = c; return; }
    cn = c + c_step;
    hn = (cn * EPI_COLS < BLOCK_N && halves_in(cn) > 0) ? 2 * cn : -1;
  };
  uint32_t va[32], vb[32];
  if (n_parts > 0) {
    // stream-K final part: the second register buffer holds the sum of the earlier partials of the half (8 x 16-byte
    // loads per part in flight together with the tcgen05.ld), added in slot order
#pragma unroll 1
    for (;;) {
      tmem_ld32(taddr0 + h * 32, va);
#pragma unroll
      for (int i = 0; i < 32; ++i) vb[i] = 0u;
      for (int q = 0; q < n_parts; ++q) {
        float4 t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = __ldcg(reinterpret_cast<const float4*>(part_row + q * part_stride) + (h * 8 + i) * 128);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          vb[4 * i] = __float_as_uint(__uint_as_float(vb[4 * i]) + t[i].x);
          vb[4 * i + 1] = __float_as_uint(__uint_as_float(vb[4 * i + 1]) + t[i].y);
          vb[4 * i + 2] = __float_as_uint(__uint_as_float(vb[4 * i + 2]) + t[i].z);
          vb[4 * i + 3] = __float_as_uint(__uint_as_float(vb[4 * i + 3]) + t[i].w);
        }
      }
      tmem_ld_wait();
      process_half(va, h, std::true_type{}, vb);
      int hn, cn;
      next_half(hn, cn);
      if (hn < 0) break;
      h = hn; c = cn;
    }
    return;
  }
  tmem_ld32(taddr0 + h * 32, va);
#pragma unroll 1
  for (;;) {
    int hn, cn;
    next_half(hn, cn);
    tmem_ld_wait();
    if (hn >= 0) tmem_ld32(taddr0 + hn * 32, vb);
    process_half(va, h, std::false_type{}, va);
    if (hn < 0) break;
    h = hn; c = cn;
    next_half(hn, cn);
    tmem_ld_wait();
    if (hn >= 0) tmem_ld32(taddr0 + hn * 32, va);
    process_half(vb, h, std::false_type{}, vb);
    if (hn < 0) break;
    h = hn; c = cn;
  }
}

// host helpers defined in qdm_gemm.cu
int get_encode_fn();
extern PFN_cuTensorMapEncodeTiled g_encode;
// 2-D row-major tensor [rows, cols] of `elem_bytes` elements; box = {box_cols, box_rows}, 128-byte swizzle by default
int make_map(CUtensorMap* map, const void* ptr, int elem_bytes, int64_t rows, int64_t cols, int box_cols, int box_rows,
             bool swizzle128 = true);

}  // namespace qdmg
