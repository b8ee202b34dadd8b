/*
 * qdm.h -- C ABI of libqdm.so: the B200 (sm_100a) quantized-linear hot path.
 *
 * This is the drop-in boundary for the quantized-linear path of
 * maani3/Quantization---Diffusion-Models.  The reference has no FFI of its own
 * (it is pure Python on torch); every entry point below replaces a torch-op
 * sequence inside one reference function, cited as `file:line` relative to the
 * reference root.  INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - the caller allocates every output; the library never retains or frees
 *     user memory and never allocates in a hot call (workspaces are passed in,
 *     sized by the matching qdm_*_workspace_bytes());
 *   - all calls are asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return value: QDM_OK or a negative code; qdm_last_error() gives the
 *     thread-local message.  The Python mirror raises ValueError for
 *     QDM_ERR_INVALID and RuntimeError for the rest (the reference's own
 *     exception types: quantize/fake_quant.py:198,253; quantize/scale.py:71);
 *   - dtype enum: QDM_F16 / QDM_BF16 / QDM_F32 (tensor element type).  All
 *     floating-point arithmetic replays torch's op-by-op rounding: each
 *     reference torch op is evaluated in fp32 and rounded to `dtype` before the
 *     next op (true IEEE division, round-half-even), so integer codes, scales
 *     and zeros are bit-exact with the reference.
 *   - no CPU fallback exists: on a non-sm_100 device every call fails.
 */
#ifndef QDM_H_
#define QDM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QDM_OK               0
#define QDM_ERR_INVALID     (-1) /* bad argument            -> ValueError   */
#define QDM_ERR_CUDA        (-2) /* CUDA runtime failure    -> RuntimeError */
#define QDM_ERR_UNSUPPORTED (-3) /* shape/dtype not built   -> RuntimeError */
#define QDM_ERR_DEVICE      (-4) /* not a cc 10.0 device    -> RuntimeError */

#define QDM_F16  0
#define QDM_BF16 1
#define QDM_F32  2

/* flags for qdm_quant_group */
#define QDM_Q_ZERO_POINT 1u /* asymmetric min/max + zero point (quantizer.py:172-182)      */
#define QDM_Q_NO_CLAMP   2u /* symmetric, codes not clamped (fake_quant.py:44-46,72)        */

int         qdm_version(void);
const char* qdm_last_error(void);
/* fails with QDM_ERR_DEVICE unless `device` has compute capability 10.0 */
int         qdm_device_check(int device);

/* ------------------------------------------------------------------ (a) reductions */

/* out[c] = max_r |x[r,c]|, x row-major [rows, cols] with row stride `ld` elements.
 * Replaces `x.reshape(-1, C).abs().amax(dim=0)`  utils/calib_data.py:117-118 and
 * `fc.weight.abs().max(dim=0)`                   quantize/quantizer_SQ.py:417-418.
 * mode 0: out = colmax; mode 1: out = max(out, colmax) (running max,
 * quantize/quantizer_SQ.py:1080-1084).  out has `dtype`.  Exact (max is order-free). */
size_t qdm_colreduce_workspace_bytes(int64_t rows, int64_t cols);
int qdm_colabsmax(const void* x, int dtype, int64_t rows, int64_t cols, int64_t ld,
                  void* out, int mode, void* workspace, size_t workspace_bytes, void* stream);

/* out_sum[c] = sum_r |x[r,c]| accumulated in fp32 with a fixed reduction tree
 * (deterministic run to run).  out_sum is fp32[cols]; the caller finishes
 * `(sum / rows).to(dtype)` as quantize/quantizer.py:652-659 does. */
int qdm_colabssum(const void* x, int dtype, int64_t rows, int64_t cols, int64_t ld,
                  float* out_sum, void* workspace, size_t workspace_bytes, void* stream);

/* One-pass hook statistic: ONE read of x[rows, cols] yields the per-call column
 * |x| max and |x| sum, and folds them in place into the caller's running accumulators, replacing the per-call
 * `abs().amax(0)` + retained `max_scales[step]` tensors of utils/calib_data.py:112-121 and the later
 * `torch.mean(torch.stack(...))` of models/StableDiffusion1_x.py:104-112, and the chunked `abs().sum(0)` of
 * quantize/quantizer.py:642-659.  Any of the three outputs may be NULL (not all):
 *   out_max[c]    (dtype)  = colmax (max_mode 0) or max(out_max[c], colmax) (max_mode 1)
 *   acc_maxsum[c] (double) += colmax            -> mean over calls of the per-call max = acc / n_calls
 *   acc_abssum[c] (double) += sum_r |x[r,c]|     (per-call sum in fp32 by a fixed tree) -> x_mean = acc / n_rows
 * Accumulators are updated by one thread per column in stream order: deterministic; the fp64 sum of fp16 maxima
 * is exact (order-free), so it may also be all-reduced across data-parallel ranks without changing a bit. */
size_t qdm_colstats_workspace_bytes(int64_t rows, int64_t cols);
int qdm_colstats(const void* x, int dtype, int64_t rows, int64_t cols, int64_t ld, void* out_max, int max_mode,
                 double* acc_maxsum, double* acc_abssum, void* workspace, size_t workspace_bytes, void* stream);

/* out[r] = max_c |x[r,c]| (dtype), rows of `cols` contiguous elements.
 * `w.abs().max(dim=-1)` quantize/fake_quant.py:89,114; cols may be tiny (conv kw). */
int qdm_rowabsmax(const void* x, int dtype, int64_t rows, int64_t cols, void* out, void* stream);

/* out[0] = max |x| over numel elements (dtype).  quantize/fake_quant.py:101,163. */
size_t qdm_absmax_workspace_bytes(int64_t numel);
int qdm_absmax(const void* x, int dtype, int64_t numel, void* out,
               void* workspace, size_t workspace_bytes, void* stream);

/* AWQ weight statistic, quantize/quantizer.py:627-637:
 *   w_scale = |W| / (groupmax(|W|) + 1e-6)   (per group of `group` along K, ops rounded in dtype)
 *   out_sum[k] = sum_n w_scale[n,k]  in fp32 (fixed tree); caller divides by N and casts.
 * W is row-major [n_rows, k_cols]; group must be 8*2^j <= 256 and divide k_cols. */
int qdm_awq_wsum(const void* w, int dtype, int64_t n_rows, int64_t k_cols, int group,
                 float* out_sum, void* workspace, size_t workspace_bytes, void* stream);

/* out[0] (double) = sum_i ( float( dtype(a[i] - b[i]) ) )^2, fixed tree.
 * `(a - b).float().pow(2).sum()` quantize/quantizer.py:777. */
size_t qdm_sqdiff_workspace_bytes(int64_t numel);
int qdm_sqdiff_sum(const void* a, const void* b, int dtype, int64_t numel, double* out,
                   void* workspace, size_t workspace_bytes, void* stream);

/* AWQ clip search, AwqQuantizer._compute_best_clip (quantize/quantizer.py:805-863), one call per Linear:
 *   for every out-row r and group g:  org = max |w[r, g]|;  for i in 0 .. int(max_shrink * n_grid) - 1:
 *     max_i = org * (1 - i / n_grid);  q_i = pseudo_quantize(clamp(w[r, g], -max_i, max_i))   (bit-exact RTN chain)
 *     err_i = mean_t ( sum_k x[t, g, k] * (q_i[k] - w[r, g, k]) )^2
 *   best_max[r, g] = max_i of the FIRST minimal err_i (strict <, :851).
 * The reference forms co_b x n_tok x K products eleven times per batch of rows; here err_i = d^T C_g d with the group's
 * g x g Gram matrix C_g = X_g^T X_g / n_tok (built once per call in `workspace`, fp32) -- no temporaries, g^2 instead of
 * n_tok * g multiply-adds per (row, group, level).  x: [n_tok, ci] rows of `ld_x` elements (the caller passes the token
 * subsample of :822-823 as a strided view), w: [co, ci] row-major, best_max: [co, ci / group] dtype.
 * group in {64, 128}; flags: QDM_Q_ZERO_POINT or 0; n_bits 2..8.  The quadratic form is fp32 whereas the reference
 * rounds products and sums to fp16, so near-ties may resolve to the neighbouring level (tests bound agreement and loss). */
size_t qdm_awq_clip_workspace_bytes(int64_t ci, int group);
int qdm_awq_clip_search(const void* w, int dtype, int64_t co, int64_t ci, int group, int n_bits, unsigned flags,
                        const void* x, int64_t n_tok, int64_t ld_x, int n_grid, float max_shrink,
                        void* best_max, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ (b) quantize / pack */

/* Per-group RTN.  w is [n_rows, k_cols] row-major, groups of `group` consecutive
 * elements along K (group divides k_cols, multiple of 16 bytes).
 *   flags & QDM_Q_ZERO_POINT : AwqQuantizer.pseudo_quantize_tensor, zero_point branch
 *                              quantize/quantizer.py:172-182
 *   flags == 0               : same function, symmetric branch  quantizer.py:183-190
 *   flags & QDM_Q_NO_CLAMP   : quantize_weight_absmax            fake_quant.py:44-46,72
 * Optional inputs (NULL to skip), each applied as its own rounded torch op:
 *   pre_mul[k_cols]  : w <- w * pre_mul[k]         (fc.weight.mul_(scales_view), quantizer.py:727)
 *   clip_max[n_rows*k_cols/group] : w <- clamp(w, -c, c) (scale.py:31, quantizer.py:845)
 *   post_div[k_cols] : dq <- dq / post_div[k]      (quantizer.py:728-730)
 * Optional outputs (NULL to skip):
 *   dq     [n_rows,k_cols] dtype : fake-quantised weight (may alias w)
 *   codes  [n_rows,k_cols] 1 byte: integer codes; unsigned 0..2^b-1 with zero point, two's
 *                                   complement otherwise (saturated to [-128,127] when QDM_Q_NO_CLAMP)
 *   scales [n_rows,k_cols/group] dtype, zeros [same] dtype (float-typed integers; zero-point only) */
int qdm_quant_group(const void* w, int dtype, int64_t n_rows, int64_t k_cols, int group,
                    int n_bits, unsigned flags,
                    const void* pre_mul, const void* clip_max, const void* post_div,
                    void* dq, int8_t* codes, void* scales, void* zeros, void* stream);

/* Row-wise RTN: one scale (and zero point) per row of `cols` contiguous elements; same `flags`
 * as qdm_quant_group.  With QDM_Q_NO_CLAMP it is
 *   s = absmax_row.clamp(1e-5) / (2^(b-1)-1);  q = round(x / s);  dq = q * s
 * = quantize_weight_per_channel_absmax fake_quant.py:86-93 (rows = numel / last_dim) and
 *   quantize_activation_per_token_absmax fake_quant.py:109-118;
 * with QDM_Q_ZERO_POINT / 0 it is pseudo_quantize_tensor with group_size <= 0 (quantizer.py:165-167).
 * Outputs optional: dq (dtype, may alias x), codes int8, scales[rows] dtype, zeros[rows] dtype. */
int qdm_quant_rowwise(const void* x, int dtype, int64_t rows, int64_t cols, int n_bits, unsigned flags,
                      void* dq, int8_t* codes, void* scales, void* zeros, void* stream);

/* Per-token int8 activation codes, the A8 of W8A8: same arithmetic as
 * quantize_activation_per_token_absmax (fake_quant.py:109-118) with n_bits = 8, but emits
 * xq[rows, cols] int8 and sx[rows] fp32 (= the dtype-rounded scale) instead of q * s.
 * smooth[cols] (dtype, optional): x is divided by it first (SmoothQuant activation side,
 * quantize/quantizer_SQ.py:425-431 when the divide cannot be folded into a previous op). */
int qdm_actquant_token_i8(const void* x, int dtype, int64_t rows, int64_t cols, const void* smooth,
                          int8_t* xq, float* sx, void* stream);

/* Whole-tensor symmetric absmax RTN, fake_quant.py:97-105,158-167.
 * scale_out[0] (dtype) optional. */
size_t qdm_quant_tensor_workspace_bytes(int64_t numel);
int qdm_quant_tensor(const void* x, int dtype, int64_t numel, int n_bits,
                     void* dq, int8_t* codes, void* scale_out,
                     void* workspace, size_t workspace_bytes, void* stream);

/* AWQ GEMM int4 layout (utils/packing_utils.py:4-27,87-102; utils/quant_utils.py:10-39):
 *   qweight[k, c] int32, nibble i = code of column 8c + {0,2,4,6,1,3,5,7}[i].
 * codes_nk is int8 [n_rows, k_cols] (low 4 bits used) in nn.Linear orientation; the
 * kernel transposes: qweight is [k_cols, n_rows/8].  n_rows % 8 == 0. */
int qdm_pack_awq(const int8_t* codes_nk, int64_t n_rows, int64_t k_cols, int32_t* qweight, void* stream);
/* inverse: qweight [k_rows, n_cols/8] -> codes_kn int8 [k_rows, n_cols] in natural column order */
int qdm_unpack_awq(const int32_t* qweight, int64_t k_rows, int64_t n_cols, int8_t* codes_kn, void* stream);

/* Fused zero-point W4 RTN + AWQ pack straight from the nn.Linear weight [n_rows,k_cols]:
 *   qweight [k_cols, n_rows/8] int32, qzeros [k_cols/group, n_rows/8] int32,
 *   scales_t [k_cols/group, n_rows] dtype   (quantizer.py:540-547 + WQLinear_GEMM.from_linear)
 * group in {32,64,128}; n_rows % 64 == 0; k_cols % group == 0.  Optional dq as above. */
int qdm_quant_pack_awq(const void* w, int dtype, int64_t n_rows, int64_t k_cols, int group,
                       int32_t* qweight, int32_t* qzeros, void* scales_t, void* dq, void* stream);

/* W_kn[k, n] = (q - z) * s in dtype (utils/packing_utils.py:87-102), out [k_rows, n_cols] */
int qdm_dequant_awq(const int32_t* qweight, const int32_t* qzeros, const void* scales_t, int dtype,
                    int64_t k_rows, int64_t n_cols, int group, void* out_kn, void* stream);

/* ------------------------------------------------------------------ (c)(d) GEMMs (tcgen05) */

/* y[M,N] = x[M,K] @ w[N,K]^T + bias, fp32 accumulate in TMEM, x/w/y/bias in dtype (f16|bf16).
 * This is WxAxLinear.forward on fake-quant weights, quantize/fake_quant.py:223.
 * K % 8 == 0, N % 8 == 0 (16-byte TMA strides).  bias may be NULL. */
int qdm_gemm_f16(const void* x, const void* w, const void* bias, void* y, int dtype,
                 int64_t M, int64_t N, int64_t K, void* stream);

/* y[M,N] = x[M,K] @ w_kn[K,N] + bias: same as qdm_gemm_f16 with the weight stored [K, N]
 * (the orientation dequantize_gemm returns, utils/packing_utils.py:87-102). */
int qdm_gemm_f16_kn(const void* x, const void* w_kn, const void* bias, void* y, int dtype,
                    int64_t M, int64_t N, int64_t K, void* stream);

/* y[M,N] = x[M,K] @ dequant(qweight,qzeros,scales)[K,N] + bias; dequant inside the main loop.
 * Same storage as WQLinear_GEMM (call sites quantize/quantizer.py:544-569):
 *   qweight [K, N/8] int32, qzeros [K/group, N/8] int32, scales [K/group, N] dtype.
 * N % 8 == 0, K % 64 == 0, group = 64 * 2^j dividing K.  M <= 128 runs single-CTA tiles, larger M CTA-pair tiles. */
int qdm_gemm_w4a16(const void* x, const int32_t* qweight, const int32_t* qzeros, const void* scales,
                   const void* bias, void* y, int dtype, int64_t M, int64_t N, int64_t K, int group,
                   void* stream);

/* Kernel-native copy of an AWQ weight for qdm_gemm_w4a16_rp ("repacked" weights; built once per weight, e.g. at module
 * load -- the checkpoint / module buffers keep the utils/packing_utils.py layout).  Words, scales and zero points of
 * 128 k rows x 16 output columns become one 1104-byte block, blocks of one k group are contiguous along N, so a CTA
 * fetches everything it needs for 128 k rows of its tile part with ONE bulk copy of whole 128-byte lines instead of
 * 64-128 short rows per k-block through three tensor maps.  blob: qdm_w4a16_repack_bytes(N, K) bytes, 16-byte aligned,
 * owned by the caller.  N % 8 == 0, K % 64 == 0, group a multiple of 64 dividing K. */
size_t qdm_w4a16_repack_bytes(int64_t N, int64_t K);
int qdm_w4a16_repack(const int32_t* qweight, const int32_t* qzeros, const void* scales, int64_t N, int64_t K, int group,
                     void* blob, size_t blob_bytes, void* stream);

/* qdm_gemm_w4a16 with the repacked copy of the same weight in `blob` (may be NULL = plain qdm_gemm_w4a16).  Problems
 * with M > 128 run the repacked-weight CTA-pair kernel: a pair owns 256 rows x one or two sub-tiles of <= 256 columns
 * (two: all 512 TMEM columns as one accumulator set, the A operand amortised over twice the columns, mid-sized layers in
 * one wave instead of two).  Small-M problems and the small-K B-stationary case read the AWQ tensors as before, so all
 * four tensors must describe the same weight.  Results are identical to qdm_gemm_w4a16 up to fp32 summation order. */
int qdm_gemm_w4a16_rp(const void* x, const int32_t* qweight, const int32_t* qzeros, const void* scales, const void* blob,
                      const void* bias, void* y, int dtype, int64_t M, int64_t N, int64_t K, int group, void* stream);

/* Second kernel-native form, for the kernel that feeds the dequantised WEIGHTS to the tensor core as its A operand from
 * tensor memory (output channels on the 128 TMEM lanes, tokens along the MMA N dimension; the weights never touch shared
 * memory): per (k-block, output channel) 8 words whose nibble order along k is {0,2,4,6,1,3,5,7} -- one lop3 gives the
 * half2 (k, k+1) of one TMEM cell -- followed by one (scale, zero-point-as-dtype-value) pair per (k-block, channel).
 * blob: qdm_w4a16_repack_ts_bytes(N, K) bytes (0.56 B / weight).  qdm_gemm_w4a16_plan takes both forms (either may be
 * NULL) next to the AWQ tensors and picks the kernel per problem. */
size_t qdm_w4a16_repack_ts_bytes(int64_t N, int64_t K);
int qdm_w4a16_repack_ts(const int32_t* qweight, const int32_t* qzeros, const void* scales, int dtype, int64_t N, int64_t K,
                        int group, void* blob, size_t blob_bytes, void* stream);
int qdm_gemm_w4a16_plan(const void* x, const int32_t* qweight, const int32_t* qzeros, const void* scales, const void* blob_rp,
                        const void* blob_ts, const void* bias, void* y, int dtype, int64_t M, int64_t N, int64_t K, int group,
                        void* stream);

/* Which kernel the calling thread's last W4A16 GEMM call launched, and its tile width in columns (tests assert the
 * dispatch with this; not part of the reference interface). */
#define QDM_GEMM_SINGLE  1 /* single-CTA tcgen05 tiles (M <= 128)              */
#define QDM_GEMM_PAIR    2 /* CTA-pair tiles, AWQ tensors through tensor maps  */
#define QDM_GEMM_BSTAT   3 /* B-stationary CTA-pair kernel (K <= 384)          */
#define QDM_GEMM_STREAMK 4 /* stream-K CTA-pair kernel                         */
#define QDM_GEMM_SKINNY  5 /* M <= 32, sector-wide cluster split-K (mma.sync)  */
#define QDM_GEMM_SMALLM  6 /* M <= 32, small weights (mma.sync)                */
#define QDM_GEMM_RP1     7 /* repacked weights, one sub-tile per pair          */
#define QDM_GEMM_RP2     8 /* repacked weights, two sub-tiles per pair         */
#define QDM_GEMM_QUAD    9 /* quad clusters (forced only)                      */
#define QDM_GEMM_TS     10 /* weights as the A operand from tensor memory      */
int qdm_gemm_last_variant(int* tile_n);

/* y[m,n] = (sum_k xq[m,k]*wq[n,k]) * sx[m] * sw[n] + bias[n]; int8 x int8 -> int32 in TMEM.
 * xq [M,K] int8 (per-token codes, fake_quant.py:109-118), wq [N,K] int8 (per-channel codes,
 * fake_quant.py:86-93), sx [M] / sw [N] fp32, bias/y in out_dtype.  K % 16 == 0, N % 8 == 0. */
int qdm_gemm_w8a8(const int8_t* xq, const float* sx, const int8_t* wq, const float* sw,
                  const void* bias, void* y, int out_dtype, int64_t M, int64_t N, int64_t K,
                  void* stream);

/* 3x3 / stride 1 / padding 1 convolution as an implicit GEMM on the same tcgen05 kernels (SURVEY.md section 8(f)
 * row 3; replaces the cuDNN call of WxAxConv2d.forward, quantize/fake_quant.py:337-341, for these geometries).
 *   x_pad [B, H+2, W+2, C]  NHWC activations with a one-pixel zero border (dtype f16/bf16), C % 64 == 0
 *   w_tap [N, 9*C]          weight [N, C, 3, 3] permuted to (n, dy, dx, c), fake-quantised by the caller
 *   y_pad [B, H+2, W+2, N]  NHWC output on the same padded grid: y_pad[b, h+1, w+1, :] is output pixel (h, w);
 *                           the border positions are scratch (written, meaningless).
 * One GEMM with M = B*(H+2)*(W+2), K = 9*C: the TMA producer fetches the A rows of tap (dy, dx) from rows shifted by
 * (dy-1)*(W+2) + (dx-1); nothing is materialised (no im2col buffer).  N % 8 == 0. */
int qdm_conv3x3_f16(const void* x_pad, const void* w_tap, const void* bias, void* y_pad, int dtype,
                    int64_t B, int64_t H, int64_t W, int64_t C, int64_t N, void* stream);

/* Direct form of the same convolution for geometries whose 128-row tiles are whole image rows
 * (qdm_conv3x3_direct_ok(H, W) == 1: W divides 128 and 128/W divides H or is a multiple of H -- every power-of-two
 * latent size up to 128 x 128):
 *   x [B, H, W, C] NHWC, UNPADDED;  y [B, H, W, N] NHWC, every element valid.
 * The activation sits behind a 4-D tensor map (c, w, h, image); tap (dy, dx) moves the box start by (dx-1, dy-1) and
 * the TMA unit's out-of-bounds zero fill is the convolution's padding: no padded copy, no border rows, no im2col.
 * Other geometries return QDM_ERR_UNSUPPORTED (use the padded-grid entry above). */
int qdm_conv3x3_direct_ok(int64_t H, int64_t W);
int qdm_conv3x3_nhwc_f16(const void* x, const void* w_tap, const void* bias, void* y, int dtype,
                         int64_t B, int64_t H, int64_t W, int64_t C, int64_t N, void* stream);

/* Same convolution from packed int4 weights: qweight [9*C, N/8], qzeros [9*C/group, N/8], scales [9*C/group, N]
 * are the AWQ GEMM layout (utils/packing_utils.py) of w_tap; group = 64 * 2^j dividing 9*C. */
int qdm_conv3x3_w4a16(const void* x_pad, const int32_t* qweight, const int32_t* qzeros, const void* scales,
                      const void* bias, void* y_pad, int dtype, int64_t B, int64_t H, int64_t W, int64_t C,
                      int64_t N, int group, void* stream);
int qdm_conv3x3_nhwc_w4a16(const void* x, const int32_t* qweight, const int32_t* qzeros, const void* scales,
                           const void* bias, void* y, int dtype, int64_t B, int64_t H, int64_t W, int64_t C,
                           int64_t N, int group, void* stream);

/* 3x3 / STRIDE 2 / padding 1 convolution (the UNet down-samplers, Downsample2D: quantize/fake_quant.py:337-341 runs them
 * through cuDNN with the module's stride) in the direct form: x [B, H, W, C] NHWC unpadded, H and W even,
 * y [B, H/2, W/2, N] NHWC; output pixel (ho, wo) = sum over taps of x[b, 2*ho + dy - 1, 2*wo + dx - 1, :] @ W[:, dy, dx, :]^T.
 * Same kernels and the same 4-D tensor map as the stride-1 direct form; the map's element (traversal) strides of 2 along
 * w and h make the TMA unit fetch every other pixel of a box twice as large, so a tile in shared memory is the same
 * 128 output pixels x 64 channels and nothing is gathered or materialised.  Needs qdm_conv3x3_direct_ok(H/2, W/2);
 * other geometries return QDM_ERR_UNSUPPORTED (the caller keeps cuDNN for those). */
int qdm_conv3x3s2_nhwc_f16(const void* x, const void* w_tap, const void* bias, void* y, int dtype,
                           int64_t B, int64_t H, int64_t W, int64_t C, int64_t N, void* stream);
int qdm_conv3x3s2_nhwc_w4a16(const void* x, const int32_t* qweight, const int32_t* qzeros, const void* scales,
                             const void* bias, void* y, int dtype, int64_t B, int64_t H, int64_t W, int64_t C,
                             int64_t N, int group, void* stream);

/* Optional workspace of the W4A16 GEMM (device memory owned by the caller, qdm_gemm_workspace_bytes() bytes, kept until
 * replaced or cleared with NULL; one per device).  With it, problems whose whole-tile waves would leave CTA pairs idle
 * (few tiles with a long K, a nearly empty last wave) run the stream-K kernel, which parks partial fp32 accumulators
 * there; without it every problem runs the whole-tile kernels.  The call zeroes the flag area on `stream`.  GEMM calls
 * that share a workspace must be ordered on one stream (the Python layer keeps one workspace per device). */
size_t qdm_gemm_workspace_bytes(void);
int qdm_gemm_set_workspace(void* workspace, size_t bytes, void* stream);

/* GEGLU of a diffusers FeedForward between its two quantized Linears (`ff.net.0.proj` -> `ff.net.2`, the pair the reference
 * quantizes in models/StableDiffusion1_x.py:121-137):  y[m, f] = x[m, f] * gelu_erf(x[m, F + f]),  x [M, 2F], y [M, F].
 * fp32 math with the dtype roundings of the two torch ops (F.gelu, mul).  f16 / bf16; F % 8 == 0; 16-byte aligned. */
int qdm_geglu(const void* x, int dtype, int64_t M, int64_t F, void* y, void* stream);

/* Tile-shape override for bring-up and A/B timing: 0 = heuristic, 1 = single-CTA tiles (128 x N),
 * 2 = CTA-pair tiles (cta_group::2, 256 x N), 4 = quad clusters, 8 = stream-K, 16 / 32 = repacked-weight kernel with one /
 * two sub-tiles (+ (sub-tile width << 8) to pin the width), 64 = ignore the repacked copy.  Process-wide, test-only;
 * not part of the reference interface. */
int qdm_set_gemm_mode(int ctas);

/* Kernel-family switches of the W4A16 dispatcher for tests and A/B timing, a process-wide OR-mask on top of the
 * QDM_W4_NO_* environment switches (which are read once, at the first GEMM call): a set bit takes that family out of the
 * dispatch -- e.g. QDM_W4_NO_SMALLM sends every M <= 32 problem to the cluster-split-K skinny kernel, QDM_W4_NO_SMALLM |
 * QDM_W4_NO_SKINNY to the tcgen05 kernels.  0 restores the heuristic.  Test-only; not part of the reference interface. */
#define QDM_W4_NO_SMALLM 1
#define QDM_W4_NO_SKINNY 2
#define QDM_W4_NO_TMA    4
#define QDM_W4_NO_BSTAT  8
#define QDM_W4_NO_SK     16
#define QDM_W4_NO_RP     32
int qdm_set_w4_disable(int mask);

/* Test hook (synchronous, allocates 16 bytes): sweeps EVERY (dividend, divisor) pair of 16-bit values of
 * `dtype` (QDM_F16 | QDM_BF16) inside the window in which the quantise kernels replace IEEE division by a
 * reciprocal + two FMAs, and compares against __fdiv_rn after the dtype rounding.
 * out_host[0] = pairs tested, out_host[1] = pairs that differ (must be 0).  See csrc/qdm_common.cuh. */
int qdm_selftest_fastdiv(int dtype, uint64_t* out_host);

/* number of kernels launched by this library in the calling thread since the last reset */
int64_t qdm_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* QDM_H_ */
