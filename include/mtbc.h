/*
 * mtbc.h -- C ABI of the B200-native (sm_100a) kernels behind the multi-task encoder-decoder hot path of
 * caumente/multi_task_breast_cancer.
 *
 * The reference has no FFI layer of its own: its operator API is torch.nn modules.  Each entry
 * point below names the reference call site (file:line under /root/reference) whose ATen/cuDNN work it replaces.
 * Everything is plain C: raw device pointers, sizes, a cudaStream_t passed as void*.  All launches are asynchronous
 * and stream ordered; functions return 0 on success or a negative mtbc_status (never abort, never synchronise).
 *
 * Activation layout everywhere: bf16, NHWC, channel count padded to a multiple of 32 ("Cp"); pad lanes are kept 0.
 */
#ifndef MTBC_H_
#define MTBC_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum mtbc_status {
  MTBC_OK = 0,
  MTBC_ERR_INVALID = -1, /* bad argument / unsupported shape        */
  MTBC_ERR_CUDA = -2,    /* CUDA runtime or driver call failed       */
  MTBC_ERR_NO_DEVICE = -3 /* no sm_100 device / driver entry missing */
};

/* Human readable text for the last error raised on the calling thread. */
const char* mtbc_last_error(void);
/* Library ABI version (bumped on any struct change). */
int mtbc_abi_version(void);
/* SHA-256 (hex) of the sources and compiler flags this binary was built from; the host binding refuses to run a
 * library whose digest differs from the sources next to it (a stale .so after a checkout). */
const char* mtbc_build_digest(void);
/* 0 if the current device is sm_100 and the tensor-map driver entry point resolves. */
int mtbc_device_check(void);

/* Execution mode of the calling thread for the element-type generic entry points below (every function that takes an
 * NHWC activation as `const void*` / `void*`).  Default 0 = the product path: bf16 activations.
 *   MTBC_MODE_ACT_FP32      the activation pointers hold fp32 (TF32 parity mode, north_star "1e-3 (TF32 mode)": fp32
 *                           storage, tcgen05.mma kind::tf32 through mtbc_conv_gemm_desc.dtype = 1 or 3)
 *   MTBC_MODE_DETERMINISTIC reductions whose order would depend on block scheduling take an order-independent path
 * A launch list is built under one mode and must be launched under the same mode (plan.py sets it around each call;
 * a CUDA graph bakes the choice in at capture). */
#define MTBC_MODE_ACT_FP32 1
#define MTBC_MODE_DETERMINISTIC 2
int mtbc_set_mode(int32_t flags);
int mtbc_get_mode(void);
/* Scratch an entry point needs beyond the buffers in its signature, in bytes (0 for every entry point but the ones
 * named here; -1 for bad arguments).  op = "in_stats_det": N samples, HW pixels per plane, channel pitch Cp. */
int64_t mtbc_query_workspace_bytes(const char* op, int32_t N, int64_t HW, int32_t Cp);

/* ------------------------------------------------------------------------------------------------------------- */
/* Strided NHWC view of a bf16 activation (channel stride 1).  Strides in elements.                               */
typedef struct {
  const void* ptr;
  int32_t C, W, H, N;
  int64_t sW, sH, sN;
} mtbc_act_view;

/* One K-segment of the implicit GEMM: a (tap, source) pair.  The A box of a 128-pixel tile is read from
 * views[view] shifted by (dh, dw) (out-of-range pixels read as 0 = the conv's zero padding); the matching B rows
 * are wpack[wtap][n][wk0 + c]. */
typedef struct {
  int32_t view, dh, dw, wk0, wtap;
} mtbc_gemm_seg;

#define MTBC_MAX_VIEWS 8
#define MTBC_MAX_SEGS 48

/* Implicit-GEMM forward/data-gradient op on tcgen05 tensor cores (TMA-fed, TMEM accumulators):
 *   D[pixel (n,h,w)][col] = sum_seg sum_c  A_seg[pixel + (dh,dw)][c] * wpack[wtap][col][wk0 + c]
 * Replaces: nn.Conv2d 3x3 forward  (MTnnUNet.py:12-16,34; MONAI Convolution via MTUNetPlusPlus.py:47-71) with the
 * skip concatenation (MTUNetPlusPlus.py:107-118, MTnnUNet.py:161-169) folded in as extra segments; its data gradient
 * (aten::convolution_backward); nn.ConvTranspose2d k=s forward and data gradient (MTnnUNet.py:96-100, MONAI UpCat).
 * epi_mode 0: out[n,h,w,col] (bf16, channel stride out_C);  1: pixel shuffle, col = q*up_cp + co, q = i*up_k + j,
 * out[n, up_k*h+i, up_k*w+j, co].  bias (fp32[ncols or up_cp]) optional.  stat_sum/stat_sq (fp32 [N][stat_C]) optional:
 * per-(n,channel) sum and sum of squares of the fp32 result = InstanceNorm statistics (MTnnUNet.py:35).
 * accumulate != 0: out += D (gradient accumulation for tensors with several consumers). */
/* Routing of a range of GEMM output columns to its own tensor (fused data gradient of a conv over a folded concat:
 * one launch writes every source's gradient).  Column ranges are multiples of 32 and must tile [0, ncols). */
typedef struct {
  void* ptr;       /* bf16 NHWC, channel stride out_C                 */
  int32_t out_C;
  int32_t col0;    /* first GEMM column of this slice                 */
  int32_t ncols;   /* columns of this slice                           */
  int32_t accumulate;
} mtbc_out_slice;

typedef struct {
  int32_t nviews;
  mtbc_act_view views[MTBC_MAX_VIEWS];
  int32_t nseg;
  mtbc_gemm_seg seg[MTBC_MAX_SEGS];
  const void* wpack; /* bf16 [w_ntaps][ncols][w_ktot] */
  int32_t w_ntaps, w_ktot;
  int32_t ncols;   /* GEMM N (padded)           */
  int32_t W, H, N; /* GEMM row space = pixels   */
  int32_t epi_mode;
  void* out;
  int32_t out_C;
  int32_t up_k, up_cp;
  const float* bias;
  float* stat_sum;
  float* stat_sq;
  int32_t stat_C;
  int32_t accumulate;
  /* nouts == 0: single output (out / out_C / accumulate above).  nouts > 0 (epi_mode 0, halo-eligible 3x3 only, no
   * bias/statistics): columns are routed to outs[]; creation fails with MTBC_ERR_INVALID if the shape is not eligible. */
  int32_t nouts;
  mtbc_out_slice outs[MTBC_MAX_VIEWS];
  /* Element type of views / wpack / out.  0: bf16 operands, tcgen05.mma kind::f16 (the product path).
   * 1: fp32 storage, tcgen05.mma kind::tf32 (the TF32 parity mode of north_star; wpack holds TF32-rounded values).
   * 3: fp32 storage, 3xTF32: wpack_lo holds the TF32 rounding of (w - tf32(w)), the kernel splits the activations
   *    the same way in shared memory and issues a_hi*w_hi + a_lo*w_hi + a_hi*w_lo (fp32-grade products). */
  int32_t dtype;
  const void* wpack_lo;
  /* Fused InstanceNorm + LeakyReLU backward statistics (dtype 0, single-output non-accumulating 3x3 data gradient on a
   * halo-eligible shape with at most 64 GEMM columns per N tile; creation fails with MTBC_ERR_INVALID otherwise and the
   * caller keeps mtbc_in_bwd_reduce).  With bwd_y != NULL the GEMM result D is the gradient of a = LeakyReLU(IN(y))
   * (training_multitask.py:100 `loss.backward()` through MTnnUNet.py:35-36 / MONAI ADN): the epilogue reads y (bf16 NHWC,
   * the geometry of `out`), forms gg = z > 0 ? D : bwd_slope * D with z = gamma * (y - mean) * rstd + beta from the fp32
   * accumulator, stores gg in `out` instead of D, and adds sum(gg) to stat_sum[n][c] and sum(gg * (y - mean) * rstd) to
   * stat_sq[n][c] -- exactly what mtbc_in_bwd_reduce leaves in s1 / s2, so only mtbc_in_bwd_apply (called with
   * slope = 1, its input already carries the LeakyReLU factor) remains: 6 instead of 10 bytes per element for the
   * normalisation backward of a single-consumer tensor.  bwd_mean / bwd_rstd: fp32 [N][stat_C]; bwd_gamma / bwd_beta:
   * fp32 [>= stat_C] or NULL. */
  const void* bwd_y;
  const float* bwd_mean;
  const float* bwd_rstd;
  const float* bwd_gamma;
  const float* bwd_beta;
  float bwd_slope;
  /* Pixel-pair view (stat_fold > 0; halo-eligible 3x3 convolutions / data gradients, single output): the caller describes a conv
   * over the tensors viewed as (N, H, W/2, 2C) -- one TMA box row = two neighbouring pixels of a dense 24-channel tensor
   * (96 bytes) instead of one (48) -- with the paired weights of MTBC_JOB_PACK_CONV_PAIR, so GEMM column c = op * C + co
   * is output pixel 2q + op, channel co.  Statistics and bias are per CHANNEL: column c (< stat_C) adds into
   * stat_sum[n][c % stat_fold] / stat_sq[n][c % stat_fold] (rows of stat_fold floats) and takes bias[c % stat_fold];
   * with bwd_y (the pair view of y) bwd_mean / bwd_rstd are [N][stat_fold] and bwd_gamma / bwd_beta indexed c % stat_fold.
   * Creation fails with MTBC_ERR_INVALID if the shape is not served by the halo kernel. */
  int32_t stat_fold;
} mtbc_conv_gemm_desc;

/* One tap of a weight-gradient GEMM: dW[tap][co][k0 + ci] += sum_pixels A[a_view][pixel + (a_dh,a_dw)][ci] *
 * B[b_view][pixel][co]. */
typedef struct {
  int32_t a_view, a_dh, a_dw, b_view;
} mtbc_wgrad_tap;

/* Weight gradient on tcgen05 (both operands MN-major, reduction over pixels, split over CTAs, fp32 atomics into a
 * zeroed accumulator laid out like wpack: dw_acc[tap][row][ld_k]).
 * Replaces: aten::convolution_backward weight path for Conv2d 3x3 and ConvTranspose2d k=s. */
typedef struct {
  int32_t a_nviews;
  mtbc_act_view a_views[4];
  int32_t b_nviews;
  mtbc_act_view b_views[4];
  int32_t ntaps;
  mtbc_wgrad_tap taps[9];
  int32_t W, H, N; /* pixel grid reduced over */
  float* dw_acc;
  int32_t n_rows; /* rows per tap plane of dw_acc (>= b C) */
  int32_t ld_k;   /* row length of dw_acc                   */
  int32_t k0;     /* column offset of this source           */
  int32_t splits; /* 0 = auto */
  /* 0: bf16 views, tcgen05 kind::f16.  1: fp32 views (TF32 parity modes): the reduction runs in plain fp32 on the CUDA
   * cores (exact products, fp32 accumulation) -- a checker-grade path, not a fast one. */
  int32_t dtype;
} mtbc_wgrad_desc;

/* Weight gradient of a 3x3 Conv2d over a folded concatenation, ALL sources in one launch: dy is read once per pixel
 * tile instead of once per source (MTUNetPlusPlus.py:107-118 concatenates up to five tensors in front of a conv; the
 * reference's aten::convolution_backward sees the materialised concat).  Same accumulator layout as mtbc_wgrad_desc;
 * source i lands at columns k0[i]...  Only planes with H % 16 == 0 and W % 8 == 0 (the caller falls back to one
 * mtbc_wgrad_create per source otherwise: creation fails with MTBC_ERR_INVALID). */
typedef struct {
  int32_t nsrc;
  mtbc_act_view x[MTBC_MAX_VIEWS];
  int32_t k0[MTBC_MAX_VIEWS];
  mtbc_act_view dy;
  int32_t W, H, N;
  float* dw_acc;
  int32_t n_rows, ld_k;
  int32_t splits; /* 0 = auto */
} mtbc_wgrad_multi_desc;

/* Fused backward of nn.ConvTranspose2d(kernel = stride = 2) (MTnnUNet.py:96-100; MONAI UpCat via
 * MTUNetPlusPlus.py:107-118): ONE pass over the output gradient produces the data gradient dx (+)= sum_q dy_q . wd[q],
 * the weight gradient dw_acc[q][co][ci] += sum_pixels x[.][ci] * dy_q[.][co] and the bias gradient dbias[co] += sum dy --
 * what aten::convolution_backward returns for the transposed conv, and what mtbc_conv_gemm (data gradient) +
 * mtbc_wgrad + mtbc_channel_sum compute in three passes.  dy[q], q = 2 i + j, is the sub-lattice dy[n, 2h+i, 2w+j, :] as
 * a strided view on the input's pixel grid.  Shapes served: Cin % 8 == 0 and Cin < 64, Cout % 8 == 0 and Cout <= 64,
 * W % 16 == 0, H % 8 == 0, bf16; anything else fails with MTBC_ERR_INVALID (the caller keeps the three launches). */
typedef struct {
  mtbc_act_view x;       /* input of the transposed conv: bf16 NHWC (N, H, W, Cin)                              */
  mtbc_act_view dy[4];   /* sub-lattices of the output gradient (N, H, W, Cout), strided                         */
  const void* wd;        /* bf16 [4][wd_rows][wd_ld]: data-gradient pack of mtbc_pack_convT_weight               */
  int32_t wd_rows, wd_ld;
  float* dw_acc;         /* fp32 [4][n_rows][ld_k], zeroed by the caller (layout of mtbc_wgrad_desc)             */
  int32_t n_rows, ld_k;
  float* dbias;          /* fp32 [Cout], accumulated into; NULL: no bias                                         */
  void* dx;              /* bf16 NHWC (N, H, W, dx_C)                                                            */
  int32_t dx_C, accumulate;
  int32_t Cout;
} mtbc_convT_bwd_desc;

typedef struct mtbc_op mtbc_op; /* opaque: encoded tensor maps + launch geometry */

int mtbc_conv_gemm_create(const mtbc_conv_gemm_desc* d, mtbc_op** out);
int mtbc_wgrad_create(const mtbc_wgrad_desc* d, mtbc_op** out);
int mtbc_wgrad_multi_create(const mtbc_wgrad_multi_desc* d, mtbc_op** out);
int mtbc_convT_bwd_create(const mtbc_convT_bwd_desc* d, mtbc_op** out);
int mtbc_op_launch(mtbc_op* op, void* stream);
int mtbc_ops_launch(mtbc_op* const* ops, int32_t n, void* stream);
void mtbc_op_destroy(mtbc_op* op);
/* Algorithmic FLOPs (2*M*N*K over padded dims) of one launch of the op, for roofline accounting. */
double mtbc_op_flops(const mtbc_op* op);

/* ------------------------------------------------------------------------------------------------------------- */
/* Weight packing.  fp32 parameters in PyTorch layout -> bf16 GEMM operands (run once per optimizer step).        */

/* Conv2d weight [Cout][Cin][kh][kw] (fp32) -> forward pack  wf[tap][co (rows=ncols)][ld_k] at column k0.. (tap =
 * r*kw+s) and, if wd != NULL, data-gradient pack wd[tap'][ci (rows=wd_rows)][wd_ld] with tap' = flipped tap,
 * for the channel slice [c_begin, c_begin+c_count) of Cin (one concat source). */
int mtbc_pack_conv_weight(const float* w, int32_t Cout, int32_t Cin, int32_t ksz, int32_t c_begin, int32_t c_count,
                          void* wf, int32_t wf_rows, int32_t wf_ld, int32_t wf_k0, void* wd, int32_t wd_rows,
                          int32_t wd_ld, void* stream);
/* ConvTranspose2d weight [Cin][Cout][k][k] (fp32) -> forward pack wf[0][q*cp + co][ci] (q = i*k+j, rows = k*k*cp,
 * ld = wf_ld) and data-gradient pack wd[q][ci][co] (rows = wd_rows, ld = wd_ld). */
int mtbc_pack_convT_weight(const float* w, int32_t Cin, int32_t Cout, int32_t k, int32_t cp, void* wf, int32_t wf_ld,
                           void* wd, int32_t wd_rows, int32_t wd_ld, void* stream);
/* dw_acc[tap][row][ld] (fp32) -> Conv2d grad [Cout][Cin][kh][kw] slice; add != 0 accumulates into grad. */
int mtbc_unpack_conv_wgrad(const float* acc, int32_t rows, int32_t ld, int32_t k0, float* grad, int32_t Cout,
                           int32_t Cin, int32_t ksz, int32_t c_begin, int32_t c_count, int32_t add, void* stream);
/* dw_acc[q][co][ld] -> ConvTranspose2d grad [Cin][Cout][k][k]. */
int mtbc_unpack_convT_wgrad(const float* acc, int32_t rows, int32_t ld, float* grad, int32_t Cin, int32_t Cout,
                            int32_t k, int32_t add, void* stream);

/* Batched form of the five parameter-side jobs above plus mtbc_copy_f32: one launch runs a whole table (a U-Net++ step
 * has ~250 of them; nn.Module parameters stay in PyTorch layout, MTUNetPlusPlus.py:47-87).  Integer arguments, in
 * the order of the single-job entry points:
 *   COPY_F32      i = {n}                                                            src -> dst0
 *   PACK_CONV     i = {Cout, Cin, ksz, c_begin, c_count, wf_rows, wf_ld, wf_k0, wd_rows, wd_ld}   src=w dst0=wf|NULL dst1=wd|NULL
 *   PACK_CONVT    i = {Cin, Cout, k, cp, wf_ld, wd_rows, wd_ld}                      src=w dst0=wf dst1=wd|NULL
 *   UNPACK_CONV   i = {rows, ld, k0, Cout, Cin, ksz, c_begin, c_count, add}          src=acc dst0=grad
 *   UNPACK_CONVT  i = {rows, ld, Cin, Cout, k, add}                                  src=acc dst0=grad
 *   PACK_CONV_PAIR i = {Cout, Cin, c_begin, c_count, rows, ld, k0, n0, Ks, Np, dgrad}  src=w dst0=wp (bf16, pre-zeroed)
 *     3x3 Conv2d weight -> operand of the SAME convolution over pixel pairs (mtbc_conv_gemm_desc.stat_fold): with
 *     E[dh][dw][k][n] the per-pixel operand (forward: k = ci - c_begin, n = co, E = w[co][ci][dh+1][dw+1]; dgrad != 0:
 *     k = co, n = ci - c_begin, E = w[co][ci][1-dh][1-dw]) the pair operand is
 *       wp[(dh+1)*3 + (dq+1)][n0 + op*Np + n][k0 + par*Ks + k] = E[dh][2*dq + par - op][k][n]   (|2*dq + par - op| <= 1)
 *     for input pixel parity par and output pixel parity op (Ks / Np = channel pitch of the source / output tensor);
 *     every other element of wp stays zero.
 * The table is copied to the device at creation; launch with mtbc_op_launch. */
enum mtbc_job_kind {
  MTBC_JOB_COPY_F32 = 0, MTBC_JOB_PACK_CONV = 1, MTBC_JOB_PACK_CONVT = 2, MTBC_JOB_UNPACK_CONV = 3,
  MTBC_JOB_UNPACK_CONVT = 4, MTBC_JOB_PACK_CONV_PAIR = 5
};
typedef struct {
  int32_t kind;
  int32_t i[11];
  const void* src;
  void* dst0;
  void* dst1;
} mtbc_param_job;
int mtbc_param_jobs_create(const mtbc_param_job* jobs, int32_t n, mtbc_op** out);

/* ------------------------------------------------------------------------------------------------------------- */
/* Sibling backbone ResidualUNet (src/models/segmentation/ResidualUNet.py): BatchNorm2d, residual add, F.dropout,    */
/* stride-2 3x3 convolutions.  Element-type generic (mtbc_set_mode), memory bound.                                  */

/* out = a + b over n elements (n % 8 == 0): `path + residual` (ResidualUNet.py:69,155). */
int mtbc_add(const void* a, const void* b, void* out, int64_t n, void* stream);
/* dst = src (accumulate == 0) or dst += src: hands a gradient to a tensor that may already hold one. */
int mtbc_accumulate(const void* src, void* dst, int64_t n, int32_t accumulate, void* stream);
/* F.dropout(x, p) (ResidualUNet.py:61,139,145; called with the default training=True, i.e. active in eval mode too):
 * out = x * keep / (1 - p).  keep (one byte per element) is drawn from a counter-based generator keyed on (seed,
 * counter[0], layer, element) and written to `mask`; external_mask != 0 reads `mask` instead (parity tests). */
int mtbc_dropout_fwd(const void* x, void* out, uint8_t* mask, int64_t n, float p, uint64_t seed, const int32_t* counter,
                     int32_t layer, int32_t external_mask, void* stream);
int mtbc_dropout_bwd(const void* g, const uint8_t* mask, void* out, int64_t n, float p, int32_t accumulate, void* stream);
/* out (N, 2H, 2W, Cp) = dy on the even lattice, zero elsewhere: the data gradient of a stride-2 3x3 convolution
 * (ResidualUNet.py:115-131) is the stride-1 data gradient of this tensor. */
int mtbc_zero_stuff2(const void* dy, int32_t N, int32_t H, int32_t W, int32_t Cp, void* out, void* stream);
/* nn.BatchNorm2d (eps 1e-5, momentum 0.1) on top of the InstanceNorm passes: pools the per-(sample, channel) sums over
 * the batch in place (training: batch average + running-statistics update with the unbiased variance +
 * num_batches_tracked; eval: rows synthesised from the running statistics), after which mtbc_in_apply normalises with
 * the batch / running statistics.  Backward: mtbc_in_bwd_reduce, mtbc_bn_pool_bwd (adds dgamma / dbeta from the
 * un-pooled totals, then pools; eval: zeroes the sums), mtbc_in_bwd_apply with dgamma = dbeta = NULL. */
int mtbc_bn_pool_fwd(float* stat_sum, float* stat_sq, int32_t N, int32_t Cp, int32_t C, int64_t HW, int32_t training,
                     float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked, void* stream);
int mtbc_bn_pool_bwd(float* s1, float* s2, int32_t N, int32_t Cp, int32_t C, int32_t training, float* dgamma,
                     float* dbeta, void* stream);

/* ------------------------------------------------------------------------------------------------------------- */
/* First layer: Conv2d 3x3 with Cin <= 4 on the fp32 NCHW input image (K = 9*Cin is too small for a tensor tile).  */
/* Replaces conv_0_0.conv_0 / encoder1.ConvInNormLRelu1.Conv forward + weight gradient (no data gradient: the image
 * needs none, training_multitask.py:82-90).
 * center_scratch (fp32 [N][Cin*9], optional): when given, y is stored with its exact per-(n,channel) mean removed
 * (computed from nine shifted plane sums of x, which land in the scratch).  Only valid when InstanceNorm follows
 * (shift invariant): the raw 0..255 image puts a large DC level on this layer, and centred storage keeps bf16's
 * mantissa for the signal and sum(y^2) free of cancellation.  bias is ignored in that mode (it cancels). */
int mtbc_conv_first_fwd(const float* x, int32_t N, int32_t Cin, int32_t H, int32_t W, const float* w, const float* bias,
                        int32_t Cout, void* y, int32_t Cp, float* stat_sum, float* stat_sq, float* center_scratch,
                        void* stream);
int mtbc_conv_first_wgrad(const float* x, int32_t N, int32_t Cin, int32_t H, int32_t W, const void* dy, int32_t Cp,
                          int32_t Cout, float* dw, void* stream);

/* ------------------------------------------------------------------------------------------------------------- */
/* InstanceNorm2d (+affine) + LeakyReLU (+ optional 2x2 max-pool output), memory bound, 128-bit accesses.          */
/* Replaces nn.InstanceNorm2d / nn.LeakyReLU / nn.MaxPool2d (MTnnUNet.py:35-36,103; MONAI ADN + Down).             */

/* Per-(n,c) sum / sum of squares of a bf16 NHWC tensor (used when the producing GEMM could not fuse them). */
int mtbc_in_stats(const void* y, int32_t N, int32_t HW, int32_t Cp, float* stat_sum, float* stat_sq, void* stream);
/* Same result, bit-identical from run to run: no floating-point atomics (block partials in `workspace`, added in block
 * order by the last block to arrive).  stat_sum / stat_sq are STORED, not accumulated.  workspace:
 * mtbc_query_workspace_bytes("in_stats_det", N, HW, Cp) bytes, zeroed once by the caller (the kernel re-arms it). */
int mtbc_in_stats_det(const void* y, int32_t N, int32_t HW, int32_t Cp, float* stat_sum, float* stat_sq,
                      void* workspace, void* stream);
/* a = lrelu(((y-mean)*rstd)*gamma+beta); mean/rstd (fp32 [N][Cp]) are written for the backward pass.
 * gamma/beta may be NULL (affine=False).  pooled (bf16 [N][H/2][W/2][Cp]) optional. */
int mtbc_in_apply(const void* y, int32_t N, int32_t H, int32_t W, int32_t Cp, const float* stat_sum,
                  const float* stat_sq, const float* gamma, const float* beta, int32_t C_true, float eps, float slope,
                  void* a, void* pooled, float* mean, float* rstd, void* stream);
/* Backward pass 1: s1[n][c] = sum g, s2[n][c] = sum g*xhat with g = dA * lrelu'(.)  (s1,s2 zeroed by the caller). */
int mtbc_in_bwd_reduce(const void* dA, const void* y, int32_t N, int32_t HW, int32_t Cp, const float* mean,
                       const float* rstd, const float* gamma, const float* beta, float slope, float* s1, float* s2,
                       void* stream);
/* Backward pass 2: dy = rstd*gamma*(g - s1/HW - xhat*s2/HW)  (bf16).  dgamma[c] += sum_n s2, dbeta[c] += sum_n s1. */
int mtbc_in_bwd_apply(const void* dA, const void* y, int32_t N, int32_t HW, int32_t Cp, const float* mean,
                      const float* rstd, const float* gamma, const float* beta, float slope, const float* s1,
                      const float* s2, void* dy, float* dgamma, float* dbeta, int32_t C_true, void* stream);
/* Both backward passes in one call (replaces the autograd of nn.InstanceNorm2d + LeakyReLU, MTnnUNet.py:35-36 / MONAI
 * ADN "NDA").  When dA and y together exceed L2 this is ONE cooperative launch whose second pass re-reads the planes out
 * of L2 (6 instead of 10 bytes of DRAM traffic per element); otherwise it is mtbc_in_bwd_reduce + mtbc_in_bwd_apply.
 * s1, s2: fp32 [N][Cp] zeroed by the caller; counters: int32 [N] zeroed by the caller (group-barrier arrivals). */
int mtbc_in_bwd(const void* dA, const void* y, int32_t N, int32_t HW, int32_t Cp, const float* mean, const float* rstd,
                const float* gamma, const float* beta, float slope, float* s1, float* s2, void* dy, float* dgamma,
                float* dbeta, int32_t C_true, int32_t* counters, void* stream);
/* Stand-alone 2x2/2 max-pool backward: dA (+)= route(dP) to the first maximum of each window of a. */
int mtbc_maxpool2_bwd(const void* a, const void* dP, int32_t N, int32_t H, int32_t W, int32_t Cp, void* dA,
                      int32_t accumulate, void* stream);
/* Nearest x2 upsample forward / backward (Multi_BTS_UNet.py:100,154-158). */
int mtbc_upsample2_fwd(const void* x, int32_t N, int32_t H, int32_t W, int32_t Cp, void* y, void* stream);
int mtbc_upsample2_bwd(const void* dy, int32_t N, int32_t H, int32_t W, int32_t Cp, void* dx, int32_t accumulate,
                       void* stream);
/* Per-channel sum over all pixels of a bf16 NHWC tensor -> fp32 [C_true] (ConvTranspose2d bias gradient). */
int mtbc_channel_sum(const void* t, int64_t npix, int32_t Cp, int32_t C_true, float* out, int32_t add, void* stream);

/* ------------------------------------------------------------------------------------------------------------- */
/* Mask heads: Conv2d 1x1 C -> 1 (MTnnUNet.py:6-9,118; MTUNetPlusPlus.py:73-76,120-123), optionally preceded by a
 * ConvTranspose2d k=s (deep-supervision heads MTnnUNet.py:106-117, Multi_BTS_UNet.py:118-127) which is composed
 * algebraically with the 1x1 projection so the k*k*C intermediate is never materialised. */
int mtbc_head1x1_fwd(const void* a, int64_t npix, int32_t Cp, int32_t C, const float* w, const float* b,
                     float* logits, void* stream);
int mtbc_head1x1_bwd(const void* a, const float* dlogits, int64_t npix, int32_t Cp, int32_t C, const float* w,
                     void* dA, int32_t accumulate, float* dw, float* db, void* stream);
/* Composed head: Wc[ci][q] = sum_co Wt[ci][co][q]*w1[co], bc = sum_co bt[co]*w1[co] + b1. */
int mtbc_dshead_compose(const float* wt, const float* bt, const float* w1, const float* b1, int32_t C, int32_t k,
                        float* wc, float* bc, void* stream);
int mtbc_dshead_fwd(const void* a, int32_t N, int32_t H, int32_t W, int32_t Cp, int32_t C, int32_t k, const float* wc,
                    const float* bc, float* logits, void* stream);
/* dA (+)= dlogits . Wc^T.  The weight-gradient partials of block r go to row r of dwc_part ([nparts + 1][C*k*k]) and
 * dbc_part ([nparts + 1]) -- no atomics, no zeroing by the caller; nparts = mtbc_dshead_bwd_parts(N, H, W). */
int mtbc_dshead_bwd_parts(int32_t N, int32_t H, int32_t W);
int mtbc_dshead_bwd(const void* a, const float* dlogits, int32_t N, int32_t H, int32_t W, int32_t Cp, int32_t C,
                    int32_t k, const float* wc, void* dA, int32_t accumulate, float* dwc_part, float* dbc_part,
                    int32_t nparts, void* stream);
/* Adds the partial rows (into row nparts of both buffers), then the chain rule back to the two original parameter sets
 * (state_dict layout stays the reference's); dwt/dbt/dw1/db1 accumulate. */
int mtbc_dshead_decompose(float* dwc_part, float* dbc_part, int32_t nparts, const float* wt, const float* bt,
                          const float* w1, int32_t C, int32_t k, float* dwt, float* dbt, float* dw1, float* db1,
                          void* stream);

/* ------------------------------------------------------------------------------------------------------------- */
/* Classification head: AdaptiveAvgPool2d(1) -> Flatten -> Linear(F,Hd) -> ReLU -> Linear(Hd,K)
 * (MTnnUNet.py:125-132, MTUNetPlusPlus.py:80-87).  hidden (fp32 [N][Hd]) and gap (fp32 [N][F]) are saved by the
 * forward; the backward overwrites hidden with d(hidden).  dw1/db1/dw2/db2 accumulate (zeroed by the caller);
 * dgap is fp32 [N][F] scratch (the pooled gradient, broadcast over the plane by a grid-wide pass). */
int mtbc_gap_fc_fwd(const void* a, int32_t N, int32_t HW, int32_t Cp, int32_t F, const float* w1, const float* b1,
                    int32_t Hd, const float* w2, const float* b2, int32_t K, float* gap, float* hidden, float* logits,
                    void* stream);
int mtbc_gap_fc_bwd(const float* dlogits, int32_t N, int32_t HW, int32_t Cp, int32_t F, const float* w1, int32_t Hd,
                    const float* w2, int32_t K, const float* gap, float* hidden, void* dA, int32_t accumulate,
                    float* dw1, float* db1, float* dw2, float* db2, float* dgap, void* stream);
/* Flatten -> Linear(C*HW, Hd) -> ReLU -> Linear(Hd,K) head of Multi_BTS_UNet (Multi_BTS_UNet.py:107-115); the
 * weight columns are indexed in NCHW flatten order (c*HW + hw) as in the reference. */
int mtbc_flat_fc_fwd(const void* a, int32_t N, int32_t HW, int32_t Cp, int32_t C, const float* w1, const float* b1,
                     int32_t Hd, const float* w2, const float* b2, int32_t K, float* hidden, float* logits,
                     void* stream);
int mtbc_flat_fc_bwd(const void* a, const float* dlogits, int32_t N, int32_t HW, int32_t Cp, int32_t C,
                     const float* w1, int32_t Hd, const float* w2, int32_t K, const float* hidden, void* dA,
                     int32_t accumulate, float* dw1, float* db1, float* dw2, float* db2, float* scratch, void* stream);

/* nn.Softmax(dim=1) over the (N, K) class logits (nnUNet_classifier.py:110,165-166; K <= 32) and its backward
 * dlogits = probs * (dprobs - sum_k dprobs*probs).  fp32, overwriting. */
int mtbc_softmax_rows_fwd(const float* logits, int32_t N, int32_t K, float* probs, void* stream);
int mtbc_softmax_rows_bwd(const float* probs, const float* dprobs, int32_t N, int32_t K, float* dlogits, void* stream);

/* ------------------------------------------------------------------------------------------------------------- */
/* Losses.  Dice = monai.losses.DiceLoss(sigmoid=True, squared_pred=True, smooth_nr=1, smooth_dr=1)
 * (experiment_init.py:209-211); focal = FocalLoss(alpha, gamma=2) on soft targets (criterions.py:6-24). */

/* sums[n][3] = (sum t*p, sum t*t, sum p*p) with p = sigmoid(logit); zeroed by the caller. */
int mtbc_dice_sums(const float* logits, const float* target, int32_t N, int64_t HW, float* sums, void* stream);
/* loss = mean_n 1 - (2I+1)/(G+P+1)  -> *loss (fp32 scalar, overwritten). */
int mtbc_dice_finalize(const float* sums, int32_t N, float* loss, void* stream);
/* dlogits = gscale[0] * dLoss/dlogit   (gscale: device fp32 scalar = upstream gradient * head weight). */
int mtbc_dice_bwd(const float* logits, const float* target, int32_t N, int64_t HW, const float* sums,
                  const float* gscale, float gmul, float* dlogits, void* stream);
int mtbc_focal_fwd(const float* logits, const float* target, int32_t N, int32_t K, float alpha, float gamma,
                   float* loss, void* stream);
int mtbc_focal_bwd(const float* logits, const float* target, int32_t N, int32_t K, float alpha, float gamma,
                   const float* gscale, float gmul, float* dlogits, void* stream);
/* Fused multi-task objective (criterions.py:52-76 + training_multitask.py:98): given per-head dice sums (head 0 =
 * full decoder, weight 1/(j+1) if inversely_weighted) and the focal loss, writes out[0..3] = total, seg, cls, nan
 * flag (1.0 if any is NaN) -- the NaN guard is a device flag, not a host sync. */
int mtbc_multitask_loss(const float* dice_losses, int32_t nheads, int32_t inversely_weighted, const float* focal,
                        float alpha_mix, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------- */
/* Prediction-refining module (utils/models.py:316-332,366-386) + hard Dice inputs (training_multitask.py:65-71):
 * mask = logit > 0; cnt[n] = #mask; cls[n] = argmax(class logits); refined mask = 0 if cls==normal_id (when
 * seg_by_class); refined cls = normal_id if cnt <= threshold... (when class_by_seg, cnt == 0).  Both refinements read
 * the initial predictions. */
int mtbc_refine_predictions(const float* mask_logits, const float* class_logits, int32_t N, int64_t HW, int32_t K,
                            int32_t normal_id, int32_t seg_by_class, int32_t class_by_seg, int32_t pixel_threshold,
                            uint8_t* mask_out, int32_t* class_out, int32_t* count_out, void* stream);
/* tp/fp/fn of (logit > 0) against a {0,1} mask over the whole batch -> int64 out[3] (zeroed by caller). */
int mtbc_hard_dice_counts(const float* logits, const float* target, int64_t n, long long* out, void* stream);
/* Per-sample confusion counts of a refined uint8 mask against a {0,1} mask: out[n][4] = tp, fp, fn, tn (int64, zeroed
 * by the caller) -- the inputs of calculate_metrics (utils/metrics.py, called at utils/models.py:335) without moving
 * the masks to the host. */
int mtbc_confusion_counts(const uint8_t* mask, const float* target, int32_t N, int64_t HW, long long* out, void* stream);
/* The reference's Hausdorff distance (utils/metrics.py:236-252: scipy directed_hausdorff on the (H, W) boolean images,
 * each image ROW a point of {0,1}^W) as integers: out[n][4] (int32, overwritten) = max_i min_j Hamming(mask row i,
 * target row j), the same with the roles swapped, #mask pixels, #target pixels.  The distance is the square root of
 * the larger of the first two; the caller applies the reference's empty-mask rules. */
int mtbc_row_hausdorff(const uint8_t* mask, const float* target, int32_t N, int32_t H, int32_t W, int32_t* out,
                       void* stream);
/* Epoch bookkeeping of train_one_epoch / validate_one_epoch (training_multitask.py:99,108-109,146-152) without host
 * syncs: acc[0..5] (double) += total, seg, cls, nan flag, hard Dice of the batch (metrics.py:255-267 from counts =
 * tp/fp/fn of mtbc_hard_dice_counts, which is zeroed afterwards), 1; confusion[gt*K+pred] (int64) += 1 per sample. */
int mtbc_metrics_accumulate(const float* loss4, long long* counts, const float* class_logits, const float* onehot,
                            int32_t B, int32_t K, double* acc, long long* confusion, void* stream);

/* ------------------------------------------------------------------------------------------------------------- */
/* GPU input pipeline (next-row f3; replaces the per-sample host transforms of src/dataset/BUSI_dataset.py:151-163 and
 * the pageable host->device copies of src/training_multitask.py:82): gathers B samples idx[b] from a uint8 dataset
 * resident on the device (images, masks: [n][H][W]; labels: [n] int32), applies RandomHorizontalFlip ->
 * RandomVerticalFlip -> RandomRotation (nearest, zero fill) with the per-sample draw in flips[b] (bit 0 h-flip,
 * bit 1 v-flip, bit 2 rotation present) and theta[b][6] (the inverse affine matrix of torchvision's rotate divided by
 * (W/2, H/2), row major), and writes fp32 image / mask (B,1,H,W) and the one-hot label (B,K; may be NULL). */
int mtbc_augment_batch(const uint8_t* images, const uint8_t* masks, const int32_t* labels, const int32_t* idx,
                       const uint8_t* flips, const float* theta, int32_t B, int32_t H, int32_t W, int32_t K,
                       float* out_img, float* out_mask, float* out_onehot, void* stream);

/* ------------------------------------------------------------------------------------------------------------- */
/* Optimizer (next-row f1): torch.optim.Adam(lr, betas, eps) semantics (experiment_init.py:186-187), fused with the
 * 1/world gradient scaling of the data-parallel all-reduce.  step is the 1-based step count. */
int mtbc_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                   float beta1, float beta2, float eps, float grad_scale, int32_t step, void* stream);

/* Same, with the step count and learning rate read from device memory (valid inside a captured CUDA graph). */
int mtbc_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                       const float* lr_dev, float beta1, float beta2, float eps, float grad_scale,
                       const int32_t* step_dev, void* stream);
int mtbc_increment_i32(int32_t* p, void* stream);

/* Utilities. */
int mtbc_fill_f32(float* p, int64_t n, float v, void* stream);
int mtbc_zero_bytes(void* p, int64_t nbytes, void* stream);
int mtbc_copy_f32(float* dst, const float* src, int64_t n, void* stream);
int mtbc_f32_to_bf16_nhwc(const float* x, int32_t N, int32_t C, int32_t H, int32_t W, void* y, int32_t Cp,
                          void* stream);
int mtbc_bf16_nhwc_to_f32(const void* x, int32_t N, int32_t C, int32_t H, int32_t W, int32_t Cp, float* y,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MTBC_H_ */
