// Halo-tile 3x3 convolution kernels for sm_100a (large planes: H % 16 == 0, W % 8 == 0).
//
// The generic kernels of conv_gemm.cu fetch one 128-pixel box per (tap, source chunk): every input pixel crosses
// L2 -> SM nine times and the 24/48-channel full-resolution layers of U-Net++ end up L2-bound.  Here one TMA load
// brings the (16+2) x (8+2) pixel halo of an 8x16 output tile into shared memory ONCE per source chunk and the nine taps
// are nine UMMA descriptors that start at different rows of that same tile (row pitch 10 pixels -> 8-row group stride
// SBO = 10 rows).  tools/exp_swizzle.cu verified on a B200 that tcgen05 applies the 64B/128B swizzle on absolute
// shared-memory address bits, so a descriptor whose start is shifted by whole rows (base_offset = 0) reads a
// TMA-written swizzled tile correctly, for K-major and MN-major operands alike.
//
//  * conv_halo_kernel  (forward / data gradient): persistent CTAs, weights of the layer RESIDENT in shared memory,
//    multi-stage halo ring, two TMEM accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
//  * wgrad_halo_kernel (weight gradient): per CTA one 32-channel chunk of one source; the three horizontal taps of a
//    kernel row are stacked along M (they are the same halo rows shifted by one pixel = leading-dimension offset of one
//    row), so 9 taps cost 3 MMA groups instead of 9 and no M lane is wasted on padding.
#include "ptx.cuh"
#include "internal.h"
#include <stdlib.h>
#include <string.h>

namespace mtbc {

constexpr int kHaloW = 10;   // 8 + 2
constexpr int kHaloH = 18;   // 16 + 2
constexpr int kHaloRows = kHaloW * kHaloH;  // 180

// ================================================================================================ forward / dgrad
struct HaloSrc {
  int16_t nchunk, kc;   // chunks of kc channels
  int32_t wk0;          // first column in the packed weight row
  int32_t b_off;        // byte offset of this source's first resident weight block inside one tap plane
  int32_t drop_last;    // the last 16 channels of the last chunk do not exist in the (dense) source: TMA zero-fills
                        // them, so that K step is not issued at all (48-channel tensors: 3 instead of 4 K steps)
};

struct HaloOut {
  __nv_bfloat16* ptr;
  int32_t out_C, col0, col_end, accumulate;   // out_C = channel pitch of the destination = number of columns that exist
};

struct ConvHaloParams {
  CUtensorMap amap[MTBC_MAX_VIEWS];
  CUtensorMap wmap[2];
  CUtensorMap ymap;     // fused InstanceNorm backward: the y tile of an output tile as 16 G image-row segments of 8 pixels
  HaloSrc src[MTBC_MAX_VIEWS];
  HaloOut outs[MTBC_MAX_VIEWS];  // column ranges of the GEMM output -> destination tensors (>= 1 entry)
  int32_t nsrc, nouts;
  int32_t W, H, N, tiles_w, tiles_h, n_mtiles;
  int32_t BN, tmem_cols, stages, a_stage_bytes, b_tap_bytes, b_total_bytes;
  int32_t lanes;          // MMA-issuing lanes (2: alternate tiles on two accumulators; MTBC_HALO_LANES=1: one)
  int32_t dbg;            // debug (MTBC_HALO_DBG=1): accumulate the MMA lane's cycle breakdown into g_halo_dbg
  int32_t late_release;   // debug (MTBC_HALO_LATE_RELEASE=1): hand the accumulator back after the stores, not before
  int32_t G, bn1;   // G = output rows stacked along N (1 or 2); bn1 = columns per output pixel; BN = G * bn1
  int32_t stat_C;
  int32_t stat_fold;      // pixel-pair view (mtbc_conv_gemm_desc.stat_fold): column c -> channel c % stat_fold
  int32_t pair24;         // ... of a 24 -> 24 layer: only the non-zero blocks of the pair operand are issued

  const float* bias;
  float* stat_sum;
  float* stat_sq;
  // Fused InstanceNorm + LeakyReLU backward statistics (mtbc_conv_gemm_desc.bwd_y): the accumulator is the gradient of
  // a = LeakyReLU(IN(y)); the epilogue reads the y tile (same geometry as the output), turns the result into
  // gg = z > 0 ? D : slope * D, stores gg and keeps sum(gg), sum(gg * (y - mean)) in its statistics registers.
  const __nv_bfloat16* bwd_y;
  const float *bwd_mean, *bwd_rstd, *bwd_gamma, *bwd_beta;
  float bwd_slope;
  int32_t y_stages, y_stage_bytes, y_rowb;   // y tile ring behind the halo ring (dense: pixel pitch y_rowb = 2 * out_C bytes)
};

constexpr int kHaloMaxStages = 8;

// Debug (MTBC_HALO_DBG=1): where CTA 0's MMA lane spends its cycles: [0] wait accumulator, [1] wait halo data,
// [2] issuing MMAs, [3] commits, [4] tiles, [5] chunks, [6] total.  Read with mtbc_debug_halo_times().
__device__ long long g_halo_dbg[8];

// The 9 taps x KC/16 K-steps of one halo chunk, fully unrolled: only the low descriptor words change, by compile-time
// (A) or per-kernel (B tap plane) offsets, so the single issuing lane stays at the tensor pipe's ~45-cycle
// per-instruction floor (tools/exp_mma_rate2.cu) instead of ~160 cycles when descriptors are rebuilt in 64-bit.
template <int KC, int KS = KC / 16>
__device__ __forceinline__ void halo_issue_chunk(uint32_t d_addr, uint32_t a_lo, uint32_t b_lo, uint32_t b_tap16,
                                                 uint32_t idesc, uint32_t accumulate) {
  constexpr uint32_t rowb = KC * 2u;
  constexpr uint32_t layout = KC == 64 ? 2u : 4u;
  const uint32_t a_hi = umma_desc_hi(kHaloW * rowb, layout);
  const uint32_t b_hi = umma_desc_hi(8u * rowb, layout);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const uint32_t a_t = a_lo + (((tap / 3) * kHaloW + (tap % 3)) * rowb >> 4);
    const uint32_t b_t = b_lo + tap * b_tap16;
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      umma_bf16_lohi(d_addr, a_t + 2 * k, a_hi, b_t + 2 * k, b_hi, idesc, (tap | k) ? 1u : accumulate);
    }
  }
}

// G = 2: two vertically adjacent output pixels share one accumulator row (columns [0, bn1) = row 2t, [bn1, 2*bn1) =
// row 2t+1).  An instruction costs the same ~45 cycles for any N <= ~96 (tools/exp_mma_rate2.cu), so the 24-channel
// layers (bn1 = 32) waste two thirds of every MMA; stacking the two rows makes the two middle input rows (dh' = 1, 2)
// single N = 2*bn1 instructions: 12 instructions per K step and 256 pixels instead of 18.  The resident weights are
// stored [dw][dh descending][bn1 x KC] so the N = 2*bn1 operand is just a window over two neighbouring tap blocks:
// input row dh' feeds tap dh' of the upper pixel and tap dh'-1 of the lower one.
template <int KC, int BN1, int KS = KC / 16>
__device__ __forceinline__ void halo_issue_chunk_g2(uint32_t d_addr, uint32_t a_lo, uint32_t b_lo, uint32_t b_dw16,
                                                    uint32_t idesc1, uint32_t idesc2, uint32_t accumulate) {
  constexpr uint32_t rowb = KC * 2u;
  constexpr uint32_t layout = KC == 64 ? 2u : 4u;
  constexpr uint32_t blk16 = BN1 * KC * 2u / 16u;   // one tap block of the resident weights, in 16-byte units
  const uint32_t a_hi = umma_desc_hi(2 * kHaloW * rowb, layout);   // 8-row group stride = two halo rows
  const uint32_t b_hi = umma_desc_hi(8u * rowb, layout);
  // The single issuing lane is also bound by its scalar instruction stream (every operand must reach a uniform
  // register): all offsets except the dw plane are compile-time constants, like in halo_issue_chunk.
  // All N = BN1 instructions first, then all N = 2*BN1 ones (they also initialise the two column halves).
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int dw = 0; dw < 3; ++dw) {
      const uint32_t b_w = b_lo + dw * b_dw16;   // blocks of this dw: [dh=2][dh=1][dh=0]
#pragma unroll
      for (int step = 0; step < 2; ++step) {
        const int dhp = pass == 0 ? (step == 0 ? 3 : 0) : (step + 1);   // pass 0: 3, 0;  pass 1: 1, 2
        const uint32_t a_t = a_lo + ((dhp * kHaloW + dw) * rowb >> 4);
        // dhp = 0: block dh=0 -> upper half; 3: block dh=2 -> lower half; 1: blocks [dh=1, dh=0]; 2: blocks [dh=2, dh=1]
        const uint32_t b_t = b_w + (dhp == 0 ? 2 * blk16 : (dhp == 1 ? blk16 : 0u));
        const uint32_t d_t = d_addr + (dhp == 3 ? BN1 : 0);
        const uint32_t idesc = pass == 0 ? idesc1 : idesc2;
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          const uint32_t acc = (pass == 0 && dw == 0 && k == 0) ? accumulate : 1u;
          umma_bf16_lohi(d_t, a_t + 2 * k, a_hi, b_t + 2 * k, b_hi, idesc, acc);
        }
      }
    }
  }
}

// Pixel-pair view of a dense 24-channel source and 24-channel output (mtbc_conv_gemm_desc.stat_fold = 24, one 64-wide
// chunk: K elements [0, 24) = pixel 2q, [24, 48) = pixel 2q+1; columns [0, 24) = output pixel 2q, [24, 48) = 2q+1).  Of the
// nine (dh, dq) tap blocks only the three dq = 0 ones are dense; a dq = -1 block holds a single 24 x 24 sub-block
// (input pixel 2q-1 -> output pixel 2q: K elements 24..47, columns 0..23) and a dq = +1 block the opposite corner
// (input pixel 2q+2 -> output pixel 2q+1: K elements 0..23, columns 24..47).  Issued: dq = 0 as three K steps of N = 48,
// dq = -1 as K steps 1, 2 with N = 32 (columns 0..31), dq = +1 as K steps 0, 1 with N = 32 on columns 16..47 (B rows and
// accumulator columns shifted by 16): 21 instructions instead of 27, and everything skipped is exactly zero.
__device__ __forceinline__ void halo_issue_chunk_pair24(uint32_t d_addr, uint32_t a_lo, uint32_t b_lo, uint32_t b_tap16,
                                                        uint32_t idesc48, uint32_t idesc32, uint32_t accumulate) {
  constexpr uint32_t rowb = 128u;
  const uint32_t a_hi = umma_desc_hi(kHaloW * rowb, 2u);
  const uint32_t b_hi = umma_desc_hi(8u * rowb, 2u);
#pragma unroll
  for (int dh = 0; dh < 3; ++dh) {   // dq = 0: initialises all 48 columns
    const uint32_t a_t = a_lo + ((dh * kHaloW + 1) * rowb >> 4);
    const uint32_t b_t = b_lo + (dh * 3 + 1) * b_tap16;
#pragma unroll
    for (int k = 0; k < 3; ++k)
      umma_bf16_lohi(d_addr, a_t + 2 * k, a_hi, b_t + 2 * k, b_hi, idesc48, (dh | k) ? 1u : accumulate);
  }
#pragma unroll
  for (int dh = 0; dh < 3; ++dh) {   // dq = -1
    const uint32_t a_t = a_lo + ((dh * kHaloW) * rowb >> 4);
    const uint32_t b_t = b_lo + (dh * 3) * b_tap16;
#pragma unroll
    for (int k = 1; k < 3; ++k) umma_bf16_lohi(d_addr, a_t + 2 * k, a_hi, b_t + 2 * k, b_hi, idesc32, 1u);
  }
#pragma unroll
  for (int dh = 0; dh < 3; ++dh) {   // dq = +1
    const uint32_t a_t = a_lo + ((dh * kHaloW + 2) * rowb >> 4);
    const uint32_t b_t = b_lo + (dh * 3 + 2) * b_tap16 + (16u * rowb >> 4);
#pragma unroll
    for (int k = 0; k < 2; ++k) umma_bf16_lohi(d_addr + 16u, a_t + 2 * k, a_hi, b_t + 2 * k, b_hi, idesc32, 1u);
  }
}

// Epilogue of the halo kernel: 8 warps, two per TMEM lane quarter (thread <-> output pixel), each pair splitting the
// tile's columns in halves.  The epilogue warps share their schedulers with nobody, so the loop is written for a short
// dependent chain: destinations per 16-column chunk are precomputed in shared memory (s_chunk), TMEM loads are issued
// in pairs before one wait, stores are 32-byte (one full sector per thread) and gradient accumulation is a
// fire-and-forget bf16 reduction in L2 instead of a read-modify-write.
//   RACC > 0 (= BN, 32 or 64): InstanceNorm partial sums (sum y, sum y^2 per column) live in registers across the
//     tiles of one sample and are combined (butterfly shuffles -> per-quarter partials in smem -> one global atomic per
//     column) only when the sample changes: per tile the statistics cost 2 FP32 ops per element and nothing else.
//   RACC == 0: any BN; statistics (if requested) are combined per tile.
struct HaloChunk {
  __nv_bfloat16* base;   // destination of column 0 of this chunk for pixel 0
  int32_t out_C;
  int16_t accumulate;
  int8_t nvalid;         // 0 / 8 / 16 columns of this chunk exist in the destination (dense pitch < GEMM columns)
  int8_t wide;           // 32-byte aligned chunk addresses: 256-bit stores allowed
};

template <int V> struct IntC { static constexpr int value = V; };

// P = column parts per TMEM lane quarter: 4 * P epilogue warps.  P = 2 everywhere except the 64-column G = 2 layers,
// whose epilogue is a per-warp dependent chain (TMEM load -> statistics -> pack -> store, ~2600 cycles per tile
// against ~1100 cycles of MMAs, profiles/r01f): with P = 4 every warp owns one 16-column chunk and the chain halves.
template <int RACC, int P, bool BWD = false>
__device__ __forceinline__ void halo_epilogue(const ConvHaloParams& p, uint32_t tmem_base, int t_begin, int t_end,
                                              int ntile, int warp, int lane, const float* s_bias,
                                              float (*s_part)[2][256], const HaloChunk* s_chunk,
                                              uint64_t* s_accfull, uint64_t* s_accempty, float (*s_bw)[64],
                                              const uint8_t* smem_y, uint64_t* s_yfull, uint64_t* s_yempty) {
  constexpr int kEpiWarps = 4 * P;
  constexpr int kEpiThreads = kEpiWarps * 32;
  const int q = warp & 3;                  // TMEM lane quarter this warp may read
  const int part = (warp - 2) >> 2;        // column part
  const int row = q * 32 + lane;
  const int tw = row & 7, th = row >> 3;
  const int BN = p.BN;
  const int nchunks = BN >> 4;
  const int k_begin = (nchunks * part + P - 1) / P;
  const int k_end = (nchunks * (part + 1) + P - 1) / P;
  const bool do_stats = (p.stat_sum != nullptr);
  const bool has_bias = (p.bias != nullptr);
  const int etid = (warp - 2) * 32 + lane;
  const int tiles_w = p.tiles_w, tiles_per_n = p.tiles_w * p.tiles_h;
  const int H = p.H, W = p.W, stat_C = p.stat_C;
  const int G = p.G, bn1 = p.bn1;
  const int fold = p.stat_fold;
  float* const stat_sum = p.stat_sum;
  float* const stat_sq = p.stat_sq;
  constexpr int NR = RACC > 0 ? RACC / P : 1;   // columns of this warp's part
  constexpr bool bwd = BWD && RACC > 0;   // fused InstanceNorm backward statistics (its own instantiation)
  const float bwd_slope = p.bwd_slope;
  const __nv_bfloat16* const out0 = p.outs[0].ptr;
  float rs[NR], rq[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) { rs[i] = 0.f; rq[i] = 0.f; }
  int cur_n = t_begin < t_end ? t_begin / tiles_per_n : 0;

  // combine the 4 quarters' column partials and add them to the per-(n, channel) statistics
  auto flush_cols = [&](int n) {
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
    for (int i = etid; i < BN; i += kEpiThreads) {
      const int gc = ntile * bn1 + (i % bn1);            // G = 2: both rows -> one channel
      if (gc >= stat_C) continue;                        // GEMM pad column of a dense tensor: no such channel
      // (pixel-pair view: the columns of both pixels of a pair belong to one channel)
      const int64_t o = fold > 0 ? static_cast<int64_t>(n) * fold + gc % fold : static_cast<int64_t>(n) * stat_C + gc;
      const float t0 = s_part[0][0][i] + s_part[1][0][i] + s_part[2][0][i] + s_part[3][0][i];
      float t1 = s_part[0][1][i] + s_part[1][1][i] + s_part[2][1][i] + s_part[3][1][i];
      if (bwd) t1 *= p.bwd_rstd[o];   // sum gg * (y - mean)  ->  sum gg * xhat
      atomicAdd(stat_sum + o, t0);
      atomicAdd(stat_sq + o, t1);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
  };
  auto flush_regs = [&](int n) {
    if constexpr (RACC > 0) {
#pragma unroll
      for (int c = 0; c < NR; c += 16) {
        float a[16], b[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { a[i] = rs[c + i]; b[i] = rq[c + i]; rs[c + i] = 0.f; rq[c + i] = 0.f; }
        const float cs = warp_colsum16(a, lane);
        const float cq = warp_colsum16(b, lane);
        if ((lane & 1) == 0) {
          const int cc = k_begin * 16 + c + col16_of_lane(lane);
          s_part[q][0][cc] = cs;
          s_part[q][1][cc] = cq;
        }
      }
      flush_cols(n);
    }
  };

  // fused InstanceNorm backward: per accumulator column [0] mean, [1] gamma * rstd, [2] beta of sample n
  auto load_consts = [&](int n) {
    for (int i = etid; i < BN; i += kEpiThreads) {
      const int gc = ntile * bn1 + (i % bn1);
      float mu = 0.f, zc = 0.f, zb = 1.f;   // GEMM pad column: accumulator is exactly 0, any branch will do
      if (gc < stat_C) {
        const int c = fold > 0 ? gc % fold : gc;   // pixel-pair view: both pixels of a pair share the channel constants
        const int64_t o = static_cast<int64_t>(n) * (fold > 0 ? fold : stat_C) + c;
        mu = p.bwd_mean[o];
        zc = (p.bwd_gamma ? p.bwd_gamma[c] : 1.f) * p.bwd_rstd[o];
        zb = p.bwd_beta ? p.bwd_beta[c] : 0.f;
      }
      s_bw[0][i] = mu; s_bw[1][i] = zc; s_bw[2][i] = zb;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
  };

  // one 16-column chunk: bias, statistics, bf16 store / accumulate
  auto do_chunk = [&](int k, auto jc, uint32_t (&r)[16], int64_t pix, const HaloChunk& hc, const uint4 (&yq)[2]) {
    constexpr int J = decltype(jc)::value;   // position of the chunk inside this warp's half (RACC > 0 only)
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
    if constexpr (bwd) {
      const uint32_t yw[8] = {yq[0].x, yq[0].y, yq[0].z, yq[0].w, yq[1].x, yq[1].y, yq[1].z, yq[1].w};
      const float* mu = s_bw[0] + k * 16;
      const float* zc = s_bw[1] + k * 16;
      const float* zb = s_bw[2] + k * 16;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float yv = __uint_as_float((i & 1) ? (yw[i >> 1] & 0xffff0000u) : (yw[i >> 1] << 16));
        const float yc = yv - mu[i];
        const float z = fmaf(yc, zc[i], zb[i]);
        const float gg = z > 0.f ? v[i] : v[i] * bwd_slope;
        v[i] = gg;
        rs[J * 16 + i] += gg;
        rq[J * 16 + i] = fmaf(gg, yc, rq[J * 16 + i]);
      }
      emit_bf16x16_n(hc.base + pix * hc.out_C, v, false, hc.nvalid, hc.wide != 0);
      return;
    }
    if (has_bias) {
      const float4* b4 = reinterpret_cast<const float4*>(s_bias + k * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 bb = b4[i];
        v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
      }
    }
    if constexpr (RACC > 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { rs[J * 16 + i] += v[i]; rq[J * 16 + i] = fmaf(v[i], v[i], rq[J * 16 + i]); }
    } else if (do_stats) {
      float sq[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) sq[i] = v[i] * v[i];
      const float cs = warp_colsum16(v, lane);
      const float cq = warp_colsum16(sq, lane);
      if ((lane & 1) == 0) {
        const int cc = k * 16 + col16_of_lane(lane);
        s_part[q][0][cc] = cs;
        s_part[q][1][cc] = cq;
      }
    }
    emit_bf16x16_n(hc.base + pix * hc.out_C, v, hc.accumulate != 0, hc.nvalid, hc.wide != 0);
  };

  // tile coordinates advance incrementally (the 8 epilogue warps are the critical path of the narrow layers: two
  // integer divisions per tile were ~25 % of their instructions)
  int n = t_begin < t_end ? t_begin / tiles_per_n : 0;
  int tr, tc;
  {
    const int rem = t_begin - n * tiles_per_n;
    tr = rem / tiles_w;
    tc = rem - tr * tiles_w;
  }
  // statistics path: this warp's one or two chunk descriptors live in registers
  HaloChunk hc0 = s_chunk[k_begin], hc1 = s_chunk[(NR > 16) ? k_begin + 1 : k_begin];
  // fused InstanceNorm backward: the producer lane streams the y tile of every output tile into a ring behind the halo
  // ring (TMA, swizzled rows of y_rowb bytes); this thread's two 16-byte pieces per chunk sit at tile-invariant offsets
  constexpr int NC = NR > 16 ? 2 : 1;
  int yoff[NC][2];
  int ys = 0;
  uint32_t yph = 0;
  if constexpr (bwd) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int col = (k_begin + c) * 16;
      const int prow = col / bn1, ch0 = col - prow * bn1;
      const int r = G == 2 ? (2 * th + prow) * 8 + tw : row;   // pixel inside the tile, image-row major
      yoff[c][0] = r * p.y_rowb + ch0 * 2;
      yoff[c][1] = yoff[c][0] + 16;
    }
    if (t_begin < t_end) load_consts(cur_n);
  }
  int it = 0;
  for (int t = t_begin; t < t_end; ++t, ++it) {
    const int buf = it & 1;
    const uint32_t acc_phase = (it >> 1) & 1;
    const int w = tc * 8 + tw;
    const int h = (tr * 16 + th) * G;   // G = 2: row of the upper pixel; the lower one is folded into s_chunk[k].base
    if (RACC > 0 && n != cur_n) { flush_regs(cur_n); cur_n = n; if (bwd) load_consts(n); }
    const int64_t pix = (static_cast<int64_t>(n) * H + h) * W + w;
    mbar_wait(&s_accfull[buf], acc_phase);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * BN);
    if (p.dbg & 8) {   // debug: epilogue releases the accumulator without reading it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_accempty[buf]);
      if (++tc == tiles_w) { tc = 0; if (++tr == p.tiles_h) { tr = 0; ++n; } }
      continue;
    }
    if constexpr (RACC > 0) {
      // this warp's half is NR / 16 = 1 or 2 chunks, known at compile time: statistics stay in registers
      uint32_t r[NR / 16][16];
      tmem_ld16_nowait(taddr + k_begin * 16, r[0]);
      if constexpr (NR > 16) tmem_ld16_nowait(taddr + k_begin * 16 + 16, r[1]);
      tmem_wait_ld();
      // the accumulator is in registers: hand it back to the MMA warp before the math and the stores
      tc_fence_before();
      __syncwarp();
      if (lane == 0 && !p.late_release) mbar_arrive(&s_accempty[buf]);
      uint4 yq[NC][2];
      if constexpr (bwd) {
        mbar_wait(&s_yfull[ys], yph);
        const uint8_t* yb = smem_y + ys * p.y_stage_bytes;
        // (channels the dense tensor lacks read as 0: their accumulator columns are exactly 0 anyway)
        yq[0][0] = hc0.nvalid >= 8 ? *reinterpret_cast<const uint4*>(yb + yoff[0][0]) : make_uint4(0, 0, 0, 0);
        yq[0][1] = hc0.nvalid >= 16 ? *reinterpret_cast<const uint4*>(yb + yoff[0][1]) : make_uint4(0, 0, 0, 0);
        if constexpr (NC > 1) {
          yq[NC - 1][0] = hc1.nvalid >= 8 ? *reinterpret_cast<const uint4*>(yb + yoff[NC - 1][0]) : make_uint4(0, 0, 0, 0);
          yq[NC - 1][1] = hc1.nvalid >= 16 ? *reinterpret_cast<const uint4*>(yb + yoff[NC - 1][1]) : make_uint4(0, 0, 0, 0);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_yempty[ys]);   // values are in registers: hand the slot back
        if (++ys == p.y_stages) { ys = 0; yph ^= 1u; }
      }
      do_chunk(k_begin, IntC<0>{}, r[0], pix, hc0, yq[0]);
      if constexpr (NR > 16) do_chunk(k_begin + 1, IntC<1>{}, r[1], pix, hc1, yq[NC - 1]);
      if (p.late_release) { __syncwarp(); if (lane == 0) mbar_arrive(&s_accempty[buf]); }
    } else {
      const uint4 ynone[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
#pragma unroll 1
      for (int k = k_begin; k < k_end; k += 2) {
        const bool two = (k + 1 < k_end);   // warp uniform
        uint32_t r[2][16];
        tmem_ld16_nowait(taddr + k * 16, r[0]);
        if (two) tmem_ld16_nowait(taddr + k * 16 + 16, r[1]);
        tmem_wait_ld();
        if (k + 2 >= k_end && !p.late_release) {   // last loads of this tile are in registers: release the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_accempty[buf]);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (u == 0 || two) do_chunk(k + u, IntC<0>{}, r[u], pix, s_chunk[k + u], ynone);
      }
      if (k_begin >= k_end || p.late_release) {   // (no chunk for this warp: still one arrival per warp and tile)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_accempty[buf]);
      }
      if (do_stats) flush_cols(n);
    }
    if (++tc == tiles_w) { tc = 0; if (++tr == p.tiles_h) { tr = 0; ++n; } }
  }
  if (RACC > 0 && t_begin < t_end) flush_regs(cur_n);
}

// MINB = resident CTAs per SM the register budget is compiled for: 2 for the narrow layers (small resident weights, two
// CTAs overlap each other's pipeline bubbles), 1 for wide N tiles (no spills, one CTA owns the SM).
// BWD = the fused InstanceNorm-backward-statistics epilogue (its own kernels, so the forward / plain data-gradient kernels'
// register allocation does not depend on it)
template <int MINB, int P, bool BWD = false>
__global__ void __launch_bounds__((MINB == 2 ? 64 : 96) + 128 * P, MINB) conv_halo_kernel(const __grid_constant__ ConvHaloParams p) {
  constexpr int kEpiWarps = 4 * P;
  constexpr int kEpiThreads = kEpiWarps * 32;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_afull[kHaloMaxStages], s_aempty[kHaloMaxStages];
  __shared__ uint64_t s_bfull, s_accfull[2], s_accempty[2];
  __shared__ uint64_t s_init[2];   // split-chunk MMA lanes: "the first chunk of this tile has initialised the accumulator"
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float s_bias[256];
  __shared__ float s_part[4][2][256];  // [lane quarter][sum|sumsq][col]: per-warp column partials at a flush
  __shared__ HaloChunk s_chunk[16];    // destination of every 16-column chunk of this CTA's N tile
  __shared__ uint32_t s_cb[2][64];     // MMA lanes' chunk schedule: weight-block descriptor (low word) | 64-wide flag
  __shared__ float s_bw[3][64];        // fused InstanceNorm backward: per accumulator column mean, gamma * rstd, beta
  __shared__ uint64_t s_yfull[4], s_yempty[4];   // ... and the y tile ring

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem;                       // resident weights: [9 taps][source][chunk][BN x kc]
  uint8_t* smem_a = smem + p.b_total_bytes;     // halo ring
  uint8_t* smem_y = smem_a + p.stages * p.a_stage_bytes;   // y tile ring (fused InstanceNorm backward only)
  const int ntile = blockIdx.y;
  const int BN = p.BN;
  // contiguous tile range per CTA: neighbouring tiles share halo rows in L2 and mostly belong to one sample, so the
  // epilogue can keep InstanceNorm partial sums in registers across tiles
  const int t_begin = static_cast<int>(static_cast<int64_t>(p.n_mtiles) * blockIdx.x / gridDim.x);
  const int t_end = static_cast<int>(static_cast<int64_t>(p.n_mtiles) * (blockIdx.x + 1) / gridDim.x);

  const int bn1 = p.bn1;
  if (tid < (BN >> 4)) {
    const int lc = tid * 16;                      // column inside the accumulator
    const int prow = lc / bn1;                    // G = 2: 0 = upper pixel, 1 = lower pixel
    const int gc = ntile * bn1 + (lc - prow * bn1);
    int o = 0;
    while (o + 1 < p.nouts && gc >= p.outs[o].col_end) ++o;   // slices are ordered by column
    s_chunk[tid].base = p.outs[o].ptr + (gc - p.outs[o].col0) + static_cast<int64_t>(prow) * p.W * p.outs[o].out_C;
    s_chunk[tid].out_C = p.outs[o].out_C;
    s_chunk[tid].accumulate = static_cast<int16_t>(p.outs[o].accumulate);
    const int left = p.outs[o].out_C - (gc - p.outs[o].col0);
    s_chunk[tid].nvalid = static_cast<int8_t>(left >= 16 ? 16 : (left >= 8 ? 8 : 0));
    s_chunk[tid].wide = static_cast<int8_t>(p.outs[o].out_C % 16 == 0);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&s_afull[s], 1); mbar_init(&s_aempty[s], 1); }
    mbar_init(&s_bfull, 1);
    for (int s = 0; s < p.y_stages; ++s) { mbar_init(&s_yfull[s], 1); mbar_init(&s_yempty[s], kEpiWarps); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_accfull[b], p.lanes == 3 ? 2 : 1);   // split-chunk mode: both lanes commit every tile
      mbar_init(&s_accempty[b], kEpiWarps);
      mbar_init(&s_init[b], 1);
    }
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&s_tmem, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  // everything above touched kernel parameters and shared memory only: the previous kernel may still be running
  pdl_trigger();
  pdl_wait();
  if (p.bias != nullptr) {   // (uniform) the bias vector comes out of the parameter pack of this step
    for (int i = tid; i < 256; i += blockDim.x) {
      const int gc = ntile * bn1 + (i % bn1);
      s_bias[i] = i >= BN ? 0.f : (p.stat_fold > 0 ? (gc < p.stat_C ? p.bias[gc % p.stat_fold] : 0.f) : p.bias[gc]);
    }
    __syncthreads();
  }

  if (warp == 0) {
    if (elect_one()) {
      // ------------------------------------------------------------ producer: weights once, then halo tiles
      mbar_arrive_expect_tx(&s_bfull, static_cast<uint32_t>(p.b_total_bytes));
      for (int tap = 0; tap < 9; ++tap)
        for (int s = 0; s < p.nsrc; ++s) {
          const HaloSrc sc = p.src[s];
          const CUtensorMap* wm = &p.wmap[sc.kc == 64 ? 1 : 0];
          for (int ch = 0; ch < sc.nchunk; ++ch) {
            const int blk = bn1 * sc.kc * 2;
            // G = 1: [tap][source][chunk];  G = 2: [dw][source][chunk][dh = 2, 1, 0]
            uint8_t* dst = p.G == 1 ? smem_b + tap * p.b_tap_bytes + sc.b_off + ch * blk
                                    : smem_b + (tap % 3) * (3 * p.b_tap_bytes) + 3 * (sc.b_off + ch * blk) + (2 - tap / 3) * blk;
            tma_load_3d(dst, wm, &s_bfull, sc.wk0 + ch * sc.kc, ntile * bn1, tap);
          }
        }
      int stage = 0;
      uint32_t phase = 0;
      int ystage = 0;
      uint32_t yphase = 0;
      const uint32_t ybytes = static_cast<uint32_t>(p.y_stage_bytes);
      for (int t = t_begin; t < t_end; ++t) {
        const int w0 = (t % p.tiles_w) * 8;
        const int h0 = ((t / p.tiles_w) % p.tiles_h) * 16 * p.G;
        const int n = t / (p.tiles_w * p.tiles_h);
        if (p.y_stages > 0) {
          mbar_wait(&s_yempty[ystage], yphase ^ 1u);
          mbar_arrive_expect_tx(&s_yfull[ystage], ybytes);
          tma_load_3d(smem_y + ystage * p.y_stage_bytes, &p.ymap, &s_yfull[ystage], w0 * (p.y_rowb >> 2), h0, n);
          if (++ystage == p.y_stages) { ystage = 0; yphase ^= 1u; }
        }
        for (int s = 0; s < p.nsrc; ++s) {
          const HaloSrc sc = p.src[s];
          const uint32_t bytes = static_cast<uint32_t>(kHaloW * (16 * p.G + 2) * sc.kc * 2);
          for (int ch = 0; ch < sc.nchunk; ++ch) {
            mbar_wait(&s_aempty[stage], phase ^ 1u);
            if ((p.dbg & 4) && t > t_begin) { mbar_arrive(&s_afull[stage]); }   // debug: no TMA traffic after tile 0
            else {
              mbar_arrive_expect_tx(&s_afull[stage], bytes);
              tma_load_4d(smem_a + stage * p.a_stage_bytes, &p.amap[s], &s_afull[stage], ch * sc.kc, w0 - 1, h0 - 1, n);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 2 + kEpiWarps) {
    // ------------------------------------------------------------ MMA issuers
    // TWO issuing lanes (warp 1 and the last warp) on alternating tiles, each with its own TMEM accumulator.  A single
    // lane spends ~45 % of a narrow-layer tile in tcgen05.commit, mbarrier round trips and loop overhead while the
    // tensor pipe -- which only queues a few instructions -- idles (tools/diag_halo_dbg.py); with two lanes the pipe is
    // fed by one while the other sits in its overhead.  The chunk ring is consumed in tile order, so lane L takes the
    // chunks of tiles it = L, L + 2, ... and skips the other lane's.  The per-tile chunk schedule is a flat table in
    // shared memory built once (weight-block descriptor + chunk width per chunk).
    const int L = warp == 1 ? 0 : 1;
    const int nlanes = p.lanes;
    if (nlanes == 3) {
      // SPLIT-CHUNK mode (tiles of >= 2 chunks, one CTA per SM): the two lanes share every tile and the same
      // accumulator; lane L takes the chunks whose GLOBAL index (over the CTA's whole tile range) has parity L.  With an
      // even ring depth every ring slot then has one fixed consumer lane, so the one-bit mbarrier parities stay exact.
      // Measured on the one-lane version (tools/diag_halo_dbg.py, [24,24,24,24,48] -> 24): per 24-MMA chunk 925 cycles
      // of issue (the pipe's 42 cycles per instruction) + 188 in tcgen05.commit + 225 waiting for the next chunk's
      // barrier although its data had landed + ~100 loop: the pipe idles a third of the time behind ONE lane's
      // bookkeeping; a second lane issues its chunk meanwhile.  The lane that owns chunk 0 of a tile zero-initialises
      // the accumulator (accumulate = 0) and commits s_init; the other lane waits for that commit before its first
      // MMA of the tile (MMAs of different threads are not ordered otherwise).  Both lanes commit s_accfull (count 2).
      if (elect_one()) {
        const uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
        const uint32_t idesc1 = umma_idesc_bf16(128, bn1, 0, 0);
        const uint32_t b_tap16 = static_cast<uint32_t>(p.b_tap_bytes) >> 4;
        const bool g2 = p.G == 2;
        int nch = 0;
        for (int s = 0; s < p.nsrc; ++s) {
          const HaloSrc sc = p.src[s];
          const uint32_t blk = static_cast<uint32_t>(bn1 * sc.kc * 2);
          for (int ch = 0; ch < sc.nchunk; ++ch, ++nch) {
            const uint32_t off = g2 ? 3 * (sc.b_off + ch * blk) : (sc.b_off + ch * blk);
            s_cb[L][nch] = umma_desc_lo(smem_u32(smem_b) + off, 16) | (sc.kc == 64 ? 0x80000000u : 0u)
                           | ((sc.drop_last && ch == sc.nchunk - 1) ? 0x40000000u : 0u);
          }
        }
        mbar_wait(&s_bfull, 0);
        const int nstages = p.stages;   // even (host)
        const uint32_t a_step16 = static_cast<uint32_t>(p.a_stage_bytes) >> 4;
        const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem_a), 16);
        const bool no_mma = (p.dbg & 2) != 0;
        int stage = L;
        uint32_t phase = 0;
        for (int it = 0; t_begin + it < t_end; ++it) {
          const int buf = it & 1;
          const uint32_t acc_phase = (it >> 1) & 1;
          const uint32_t d_addr = tmem_base + static_cast<uint32_t>(buf * BN);
          const int c0 = (((it * nch) & 1) == L) ? 0 : 1;   // my first chunk of this tile
          mbar_wait(&s_accempty[buf], acc_phase ^ 1u);       // epilogue has drained this accumulator
          if (c0 != 0) mbar_wait(&s_init[buf], acc_phase);   // the other lane's chunk 0 has initialised it
          tc_fence_after();
#pragma unroll 1
          for (int c = c0; c < nch; c += 2) {
            mbar_wait(&s_afull[stage], phase);
            const uint32_t cb = s_cb[L][c];
            const uint32_t b_lo = cb & 0x3fffffffu;
            const uint32_t kind = cb >> 30;
            const uint32_t a_lo = a_lo0 + stage * a_step16;
            const uint32_t accumulate = c == 0 ? 0u : 1u;
            if (no_mma) {
            } else if (g2) {
              if (kind == 2) halo_issue_chunk_g2<64, 32>(d_addr, a_lo, b_lo, 3 * b_tap16, idesc1, idesc, accumulate);
              else if (kind == 0) halo_issue_chunk_g2<32, 32>(d_addr, a_lo, b_lo, 3 * b_tap16, idesc1, idesc, accumulate);
              else if (kind == 3) halo_issue_chunk_g2<64, 32, 3>(d_addr, a_lo, b_lo, 3 * b_tap16, idesc1, idesc, accumulate);
              else halo_issue_chunk_g2<32, 32, 1>(d_addr, a_lo, b_lo, 3 * b_tap16, idesc1, idesc, accumulate);
            } else {
              if (kind == 2) halo_issue_chunk<64>(d_addr, a_lo, b_lo, b_tap16, idesc, accumulate);
              else if (kind == 0) halo_issue_chunk<32>(d_addr, a_lo, b_lo, b_tap16, idesc, accumulate);
              else if (kind == 3) halo_issue_chunk<64, 3>(d_addr, a_lo, b_lo, b_tap16, idesc, accumulate);
              else halo_issue_chunk<32, 1>(d_addr, a_lo, b_lo, b_tap16, idesc, accumulate);
            }
            if (c == 0) umma_commit(&s_init[buf]);
            umma_commit(&s_aempty[stage]);
            stage += 2;
            if (stage >= nstages) { stage -= nstages; phase ^= 1u; }
          }
          umma_commit(&s_accfull[buf]);
        }
      }
    } else if (L < nlanes && elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
      const uint32_t idesc_p48 = umma_idesc_bf16(128, 48, 0, 0), idesc_p32 = umma_idesc_bf16(128, 32, 0, 0);
      const uint32_t idesc1 = umma_idesc_bf16(128, bn1, 0, 0);
      const uint32_t b_tap16 = static_cast<uint32_t>(p.b_tap_bytes) >> 4;
      const bool g2 = p.G == 2;
      int nch = 0;
      for (int s = 0; s < p.nsrc; ++s) {
        const HaloSrc sc = p.src[s];
        const uint32_t blk = static_cast<uint32_t>(bn1 * sc.kc * 2);
        for (int ch = 0; ch < sc.nchunk; ++ch, ++nch) {
          const uint32_t off = g2 ? 3 * (sc.b_off + ch * blk) : (sc.b_off + ch * blk);
          s_cb[L][nch] = umma_desc_lo(smem_u32(smem_b) + off, 16) | (sc.kc == 64 ? 0x80000000u : 0u)   // bit 31: 64-wide
                         | ((sc.drop_last && ch == sc.nchunk - 1) ? 0x40000000u : 0u);                  // bit 30: last K step is all zero
        }
      }
      mbar_wait(&s_bfull, 0);
      const int nstages = p.stages;
      const uint32_t a_step16 = static_cast<uint32_t>(p.a_stage_bytes) >> 4;
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem_a), 16);
      int stage = 0;
      uint32_t phase = 0;
      auto skip = [&](int n) {   // advance the ring position over n chunks (the other lane's tile)
        stage += n;
        while (stage >= nstages) { stage -= nstages; phase ^= 1u; }
      };
      if (L == 1) skip(nch);
      const bool dbg = (p.dbg & 1) && blockIdx.x == 0 && blockIdx.y == 0 && L == 0;
      const bool no_mma = (p.dbg & 2) != 0;
      long long c_acc = 0, c_data = 0, c_issue = 0, c_commit = 0, n_chunks = 0, n_tiles = 0;
      const long long c_start = dbg ? clock64() : 0;
      for (int it = L; t_begin + it < t_end; it += nlanes) {
        const int buf = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        long long t0 = dbg ? clock64() : 0;
        mbar_wait(&s_accempty[buf], acc_phase ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        if (dbg) { const long long t1 = clock64(); c_acc += t1 - t0; }
        const uint32_t d_addr = tmem_base + static_cast<uint32_t>(buf * BN);
        uint32_t accumulate = 0;
        uint32_t a_lo = a_lo0 + stage * a_step16;
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
          if (dbg) t0 = clock64();
          mbar_wait(&s_afull[stage], phase);   // TMA data: the mbarrier acquire is all the MMA needs
          long long t1 = 0;
          if (dbg) { t1 = clock64(); c_data += t1 - t0; }
          const uint32_t cb = s_cb[L][c];
          const uint32_t b_lo = cb & 0x3fffffffu;
          const uint32_t kind = cb >> 30;   // 0: 32 wide, 1: 32 wide minus one K step, 2: 64 wide, 3: 64 wide minus one
          if (no_mma) {
            // debug: no MMAs
          } else if (p.pair24) {
            halo_issue_chunk_pair24(d_addr, a_lo, b_lo, b_tap16, idesc_p48, idesc_p32, accumulate);
          } else if (g2) {
            if (kind == 2) halo_issue_chunk_g2<64, 32>(d_addr, a_lo, b_lo, 3 * b_tap16, idesc1, idesc, accumulate);
            else if (kind == 0) halo_issue_chunk_g2<32, 32>(d_addr, a_lo, b_lo, 3 * b_tap16, idesc1, idesc, accumulate);
            else if (kind == 3) halo_issue_chunk_g2<64, 32, 3>(d_addr, a_lo, b_lo, 3 * b_tap16, idesc1, idesc, accumulate);
            else halo_issue_chunk_g2<32, 32, 1>(d_addr, a_lo, b_lo, 3 * b_tap16, idesc1, idesc, accumulate);
          } else {
            if (kind == 2) halo_issue_chunk<64>(d_addr, a_lo, b_lo, b_tap16, idesc, accumulate);
            else if (kind == 0) halo_issue_chunk<32>(d_addr, a_lo, b_lo, b_tap16, idesc, accumulate);
            else if (kind == 3) halo_issue_chunk<64, 3>(d_addr, a_lo, b_lo, b_tap16, idesc, accumulate);
            else halo_issue_chunk<32, 1>(d_addr, a_lo, b_lo, b_tap16, idesc, accumulate);
          }
          accumulate = 1;
          long long t2 = 0;
          if (dbg) { t2 = clock64(); c_issue += t2 - t1; }
          umma_commit(&s_aempty[stage]);
          if (dbg) { c_commit += clock64() - t2; ++n_chunks; }
          a_lo += a_step16;
          if (++stage == nstages) { stage = 0; phase ^= 1u; a_lo = a_lo0; }
        }
        umma_commit(&s_accfull[buf]);
        if (nlanes == 2) skip(nch);
        ++n_tiles;
      }
      if (dbg) {
        g_halo_dbg[0] += c_acc; g_halo_dbg[1] += c_data; g_halo_dbg[2] += c_issue; g_halo_dbg[3] += c_commit;
        g_halo_dbg[4] += n_tiles; g_halo_dbg[5] += n_chunks; g_halo_dbg[6] += clock64() - c_start;
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue warps 2..9 (TMEM lane quarter = warp % 4)
    if constexpr (BWD) {   // host guarantees statistics, BN = 32 / 64 and one CTA per SM
      if constexpr (P == 2) {
        if (BN == 32) halo_epilogue<32, 2, true>(p, tmem_base, t_begin, t_end, ntile, warp, lane, s_bias, s_part, s_chunk, s_accfull, s_accempty, s_bw, smem_y, s_yfull, s_yempty);
        else halo_epilogue<64, 2, true>(p, tmem_base, t_begin, t_end, ntile, warp, lane, s_bias, s_part, s_chunk, s_accfull, s_accempty, s_bw, smem_y, s_yfull, s_yempty);
      } else {
        halo_epilogue<64, 4, true>(p, tmem_base, t_begin, t_end, ntile, warp, lane, s_bias, s_part, s_chunk, s_accfull, s_accempty, s_bw, smem_y, s_yfull, s_yempty);
      }
    } else if constexpr (P == 2) {
      if (p.stat_sum != nullptr && BN == 32) halo_epilogue<32, 2>(p, tmem_base, t_begin, t_end, ntile, warp, lane, s_bias, s_part, s_chunk, s_accfull, s_accempty, s_bw, smem_y, s_yfull, s_yempty);
      else if (p.stat_sum != nullptr && BN == 64) halo_epilogue<64, 2>(p, tmem_base, t_begin, t_end, ntile, warp, lane, s_bias, s_part, s_chunk, s_accfull, s_accempty, s_bw, smem_y, s_yfull, s_yempty);
      else halo_epilogue<0, 2>(p, tmem_base, t_begin, t_end, ntile, warp, lane, s_bias, s_part, s_chunk, s_accfull, s_accempty, s_bw, smem_y, s_yfull, s_yempty);
    } else {
      if (p.stat_sum != nullptr && BN == 64) halo_epilogue<64, 4>(p, tmem_base, t_begin, t_end, ntile, warp, lane, s_bias, s_part, s_chunk, s_accfull, s_accempty, s_bw, smem_y, s_yfull, s_yempty);
      else halo_epilogue<0, 4>(p, tmem_base, t_begin, t_end, ntile, warp, lane, s_bias, s_part, s_chunk, s_accfull, s_accempty, s_bw, smem_y, s_yfull, s_yempty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ================================================================================================ weight gradient
// dW[dh,dw][ci][co] = sum_p x[p + (dh,dw)][ci] * dy[p][co].  With p' = p + (dh,0) the sum becomes
// sum_p' x[p' + (0,dw)][ci] * dy[p' - (dh,0)][co]: the three horizontal taps are the SAME x rows shifted by one pixel
// (stacked along M through the leading-dimension offset = one pixel), and the three vertical taps are the SAME dy rows
// shifted by one tile row (stacked along N, leading-dimension offset = 8 pixels).  One 128-pixel tile therefore costs
// 8 MMAs of shape M = 128 ([dw=-1|0|+1|unused] x 32 ci), N = 3 x b_kc ([dh=+1|0|-1] x co), K = 16 pixels, instead of
// 24 narrow ones: an SS MMA pays 32 cycles of shared-memory bandwidth for its A slice whatever N is, so wide N is what
// makes the 24/48-channel layers cheap.
struct WgradHaloParams {
  CUtensorMap amap;       // x source, box (32, 10, 16, 1), 64B swizzle: column halo only
  CUtensorMap bmap;       // dy, box (b_kc, 8, 18, 1): row halo only
  int32_t a_C;            // padded channels of the source
  int32_t b_kc;           // dy channels per CTA (32: 64B swizzle, 64: 128B swizzle)
  int32_t NN, n_tiles, tmem_cols, stages;
  int32_t tiles_w, tiles_h, n_ptiles, splits;
  int32_t a_bytes, b_bytes, stage_bytes;
  int32_t n_rows, ld_k, k0;
  float* dw_acc;
};

__global__ void __launch_bounds__(128) wgrad_halo_kernel(const __grid_constant__ WgradHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_full[kHaloMaxStages], s_empty[kHaloMaxStages];
  __shared__ uint64_t s_accum;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // chunk (source channels) fastest: CTAs that run together read the same dy tiles -> L2 hits
  const int chunk = blockIdx.x, split = blockIdx.y, ntile = blockIdx.z;
  const int per = (p.n_ptiles + p.splits - 1) / p.splits;
  const int pt_begin = split * per;
  const int pt_end = min(p.n_ptiles, pt_begin + per);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 1); }
    mbar_init(&s_accum, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&s_tmem, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  pdl_trigger();
  pdl_wait();

  if (pt_end > pt_begin) {
    if (warp == 0 && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = pt_begin; pt < pt_end; ++pt) {
        const int w0 = (pt % p.tiles_w) * 8;
        const int h0 = ((pt / p.tiles_w) % p.tiles_h) * 16;
        const int n = pt / (p.tiles_w * p.tiles_h);
        mbar_wait(&s_empty[stage], phase ^ 1u);
        uint8_t* a_dst = smem + stage * p.stage_bytes;
        uint8_t* b_dst = a_dst + p.a_bytes;
        mbar_arrive_expect_tx(&s_full[stage], static_cast<uint32_t>(kHaloW * 16 * 64 + 8 * kHaloH * p.b_kc * 2));
        tma_load_4d(a_dst, &p.amap, &s_full[stage], chunk * 32, w0 - 1, h0, n);
        tma_load_4d(b_dst, &p.bmap, &s_full[stage], ntile * p.b_kc, w0, h0 - 1, n);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    } else if (warp == 1 && elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, p.NN, 1, 1);
      const uint32_t b_rowb = p.b_kc * 2u;
      const uint32_t a_hi = umma_desc_hi(kHaloW * 64, 4u);                      // K step of 8 pixels = one halo row
      const uint32_t b_hi = umma_desc_hi(8 * b_rowb, p.b_kc == 64 ? 2u : 4u);   // K step of 8 pixels = one tile row
      const uint32_t b_j16 = b_rowb;  // 16 pixel rows of the dy tile, in 16-byte units
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = pt_begin; pt < pt_end; ++pt) {
        mbar_wait(&s_full[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + stage * p.stage_bytes);
        const uint32_t a_lo = umma_desc_lo(a_base, 64);                          // next M atom = next pixel (dw + 1)
        const uint32_t b_lo = umma_desc_lo(a_base + p.a_bytes, 8 * b_rowb);      // next N atom = next tile row (dh - 1)
        const uint32_t first = (pt > pt_begin) ? 1u : 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {    // 16 pixels = tile rows 2j, 2j+1
          umma_bf16_lohi(tmem_base, a_lo + (2 * j * kHaloW) * 4, a_hi, b_lo + j * b_j16, b_hi, idesc, j > 0 ? 1u : first);
        }
        umma_commit(&s_empty[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      umma_commit(&s_accum);
    }
    __syncwarp();
    mbar_wait(&s_accum, 0);
    tc_fence_after();
    __syncwarp();

    // warp = horizontal tap (dw + 1), lane = channel inside the chunk; column group a = 0,1,2 <-> dh = 1 - a
    const int ci = chunk * 32 + lane;
    const bool valid = (warp < 3) && (ci < p.a_C);
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    for (int a = 0; a < 3; ++a) {
      const int tap = (2 - a) * 3 + warp;
      for (int c = 0; c < p.b_kc; c += 16) {
        float v[16];
        tmem_ld16(taddr + a * p.b_kc + c, v);
        if (valid) {
          float* dst = p.dw_acc + (static_cast<int64_t>(tap) * p.n_rows + ntile * p.b_kc + c) * p.ld_k + p.k0 + ci;
#pragma unroll
          for (int i = 0; i < 16; ++i) atomicAdd(dst + static_cast<int64_t>(i) * p.ld_k, v[i]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}


// ---- all concat sources in one launch ------------------------------------------------------------------------------
// The per-source kernel above is bound by TMA box rows (~4 cycles per 64-byte row per SM, 160 rows of x + 144 rows of
// dy per 128-pixel tile), not by its 8 MMAs.  Here one CTA owns a GROUP of up to five 32-channel chunks (any sources)
// and loads the dy box once per tile for all of them: 144 + 160 k rows for k chunks instead of 304 k.  Each chunk has
// its own TMEM accumulator (NN columns), so a group is limited by 512 TMEM columns.
constexpr int kWgMaxChunks = 48;
struct WgradMultiParams {
  CUtensorMap amap[MTBC_MAX_VIEWS];
  CUtensorMap bmap;
  int16_t c_view[kWgMaxChunks], c_ch[kWgMaxChunks], c_ac[kWgMaxChunks];   // source, first channel, channels of the source
  int32_t c_col[kWgMaxChunks];                                             // first column in dw_acc rows
  int32_t nchunks, group;
  int32_t b_kc, NN, n_tiles, tmem_cols, stages;
  int32_t tiles_w, tiles_h, n_ptiles, splits;
  int32_t a_bytes, b_bytes, stage_bytes;
  int32_t n_rows, ld_k;
  float* dw_acc;
};

__global__ void __launch_bounds__(128) wgrad_halo_multi_kernel(const __grid_constant__ WgradMultiParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_full[kHaloMaxStages], s_empty[kHaloMaxStages];
  __shared__ uint64_t s_accum;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int grp = blockIdx.x, split = blockIdx.y, ntile = blockIdx.z;
  const int c_begin = grp * p.group;
  const int nck = min(p.group, p.nchunks - c_begin);
  const int per = (p.n_ptiles + p.splits - 1) / p.splits;
  const int pt_begin = split * per;
  const int pt_end = min(p.n_ptiles, pt_begin + per);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 1); }
    mbar_init(&s_accum, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&s_tmem, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  pdl_trigger();
  pdl_wait();

  if (pt_end > pt_begin && nck > 0) {
    if (warp != 1 && elect_one()) {
      // Three producer lanes (warps 0, 2, 3): a CTA's TMA box loads are served about one 64-byte row per 4 cycles when
      // one lane issues them all; halving the CTAs per SM doubled this kernel's time, i.e. the limit is per issuing
      // stream, so the boxes of a stage are spread over three lanes.  Warp 0 owns the barrier bookkeeping
      // (arrive + expect_tx for the whole stage) and the dy box; the x boxes are dealt round-robin.
      const int pw = warp == 0 ? 0 : warp - 1;   // 0, 1, 2
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t bytes = static_cast<uint32_t>(8 * kHaloH * p.b_kc * 2 + nck * (kHaloW * 16 * 64));
      for (int pt = pt_begin; pt < pt_end; ++pt) {
        const int w0 = (pt % p.tiles_w) * 8;
        const int h0 = ((pt / p.tiles_w) % p.tiles_h) * 16;
        const int n = pt / (p.tiles_w * p.tiles_h);
        mbar_wait(&s_empty[stage], phase ^ 1u);
        uint8_t* b_dst = smem + stage * p.stage_bytes;
        if (pw == 0) {
          mbar_arrive_expect_tx(&s_full[stage], bytes);
          tma_load_4d(b_dst, &p.bmap, &s_full[stage], ntile * p.b_kc, w0, h0 - 1, n);
        }
        for (int c = pw; c < nck; c += 3)
          tma_load_4d(b_dst + p.b_bytes + c * p.a_bytes, &p.amap[p.c_view[c_begin + c]], &s_full[stage],
                      p.c_ch[c_begin + c], w0 - 1, h0, n);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    } else if (warp == 1 && elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, p.NN, 1, 1);
      const uint32_t b_rowb = p.b_kc * 2u;
      const uint32_t a_hi = umma_desc_hi(kHaloW * 64, 4u);                      // K step of 8 pixels = one halo row
      const uint32_t b_hi = umma_desc_hi(8 * b_rowb, p.b_kc == 64 ? 2u : 4u);   // K step of 8 pixels = one tile row
      const uint32_t b_j16 = b_rowb;  // 16 pixel rows of the dy tile, in 16-byte units
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = pt_begin; pt < pt_end; ++pt) {
        mbar_wait(&s_full[stage], phase);
        tc_fence_after();
        const uint32_t base = smem_u32(smem + stage * p.stage_bytes);
        const uint32_t b_lo = umma_desc_lo(base, 8 * b_rowb);                    // next N atom = next tile row (dh - 1)
        const uint32_t first = (pt > pt_begin) ? 1u : 0u;
        for (int c = 0; c < nck; ++c) {
          const uint32_t a_lo = umma_desc_lo(base + p.b_bytes + c * p.a_bytes, 64);   // next M atom = next pixel (dw + 1)
          const uint32_t d = tmem_base + static_cast<uint32_t>(c * p.NN);
#pragma unroll
          for (int j = 0; j < 8; ++j)    // 16 pixels = tile rows 2j, 2j+1
            umma_bf16_lohi(d, a_lo + (2 * j * kHaloW) * 4, a_hi, b_lo + j * b_j16, b_hi, idesc, j > 0 ? 1u : first);
        }
        umma_commit(&s_empty[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      umma_commit(&s_accum);
    }
    __syncwarp();
    mbar_wait(&s_accum, 0);
    tc_fence_after();
    __syncwarp();

    // warp = horizontal tap (dw + 1), lane = channel inside the chunk; column group a = 0,1,2 <-> dh = 1 - a
    for (int c = 0; c < nck; ++c) {
      const int ci = p.c_ch[c_begin + c] + lane;
      const bool valid = (warp < 3) && (ci < p.c_ac[c_begin + c]);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(c * p.NN);
      for (int a = 0; a < 3; ++a) {
        const int tap = (2 - a) * 3 + warp;
        for (int cc = 0; cc < p.b_kc; cc += 16) {
          float v[16];
          tmem_ld16(taddr + a * p.b_kc + cc, v);
          if (valid) {
            float* dst = p.dw_acc + (static_cast<int64_t>(tap) * p.n_rows + ntile * p.b_kc + cc) * p.ld_k +
                         p.c_col[c_begin + c] + lane;
#pragma unroll
            for (int i = 0; i < 16; ++i) atomicAdd(dst + static_cast<int64_t>(i) * p.ld_k, v[i]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ================================================================================================ host side
static int tmem_cols_pow2(int n) { int c = 32; while (c < n) c <<= 1; return c; }
static bool halo_disabled() { const char* e = getenv("MTBC_NO_HALO"); return e && e[0] == '1'; }
static int sm_count() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
  return n;
}

struct ConvHaloOp : public OpBase {
  ConvHaloParams p;
  dim3 grid;
  int smem_bytes;
  double flops;
  int ctas_per_sm = 1;
  int epi_parts = 2;
  int launch(cudaStream_t st) override {
    // two CTAs per SM: the other CTA's MMA lane already feeds the pipe, one lane each (and no extra warp: registers)
    if (p.bwd_y != nullptr) {
      if (epi_parts == 4) launch_pdl(conv_halo_kernel<1, 4, true>, grid, dim3(96 + 512), smem_bytes, st, p);
      else launch_pdl(conv_halo_kernel<1, 2, true>, grid, dim3(96 + 256), smem_bytes, st, p);
    } else if (epi_parts == 4) launch_pdl(conv_halo_kernel<1, 4>, grid, dim3(96 + 512), smem_bytes, st, p);
    else if (ctas_per_sm == 2) launch_pdl(conv_halo_kernel<2, 2>, grid, dim3(64 + 256), smem_bytes, st, p);
    else launch_pdl(conv_halo_kernel<1, 2>, grid, dim3(96 + 256), smem_bytes, st, p);
    return check_launch("conv_halo_kernel");
  }
  double op_flops() const override { return flops; }
};

struct WgradHaloOp : public OpBase {
  WgradHaloParams p;
  dim3 grid;
  int smem_bytes;
  double flops;
  int launch(cudaStream_t st) override {
    launch_pdl(wgrad_halo_kernel, grid, dim3(128), smem_bytes, st, p);
    return check_launch("wgrad_halo_kernel");
  }
  double op_flops() const override { return flops; }
};

int conv_halo_try_create(const mtbc_conv_gemm_desc* d, OpBase** out) {
  if (halo_disabled()) return 1;
  if (d->epi_mode != 0 || d->nviews < 1 || d->nviews > MTBC_MAX_VIEWS || d->nseg != 9 * d->nviews) return 1;
  if (d->nouts < 0 || d->nouts > MTBC_MAX_VIEWS) return set_error(MTBC_ERR_INVALID, "conv_gemm: bad nouts");
  if (d->nouts > 0) {
    if (d->bias || d->stat_sum) return set_error(MTBC_ERR_INVALID, "conv_gemm: routed outputs take no bias / statistics");
    int col = 0;
    for (int i = 0; i < d->nouts; ++i) {
      const mtbc_out_slice& o = d->outs[i];
      if (!o.ptr || o.col0 != col || o.ncols <= 0 || o.ncols % 32 != 0 || o.out_C % 8 != 0 || o.out_C > o.ncols ||
          o.out_C + 31 < o.ncols)
        return set_error(MTBC_ERR_INVALID, "conv_gemm: output slice %d malformed", i);
      col += o.ncols;
    }
    if (col != d->ncols) return set_error(MTBC_ERR_INVALID, "conv_gemm: output slices do not tile ncols");
  }
  if (d->H % 16 != 0 || d->W % 8 != 0 || d->w_ntaps != 9 || d->ncols % 32 != 0) return 1;
  if (d->stat_fold != 0 && (d->stat_fold < 0 || d->nouts != 0 || d->stat_C % d->stat_fold != 0))
    return set_error(MTBC_ERR_INVALID, "conv_gemm: pixel-pair view: single output, stat_C %% stat_fold == 0");
  // every view must appear with the full 3x3 stencil, tap index = (dh+1)*3 + (dw+1), one weight column offset
  int wk0[MTBC_MAX_VIEWS];
  int seen[MTBC_MAX_VIEWS];
  for (int v = 0; v < d->nviews; ++v) { wk0[v] = -1; seen[v] = 0; }
  for (int s = 0; s < d->nseg; ++s) {
    const mtbc_gemm_seg& g = d->seg[s];
    if (g.view < 0 || g.view >= d->nviews || g.dh < -1 || g.dh > 1 || g.dw < -1 || g.dw > 1) return 1;
    if (g.wtap != (g.dh + 1) * 3 + (g.dw + 1)) return 1;
    if (wk0[g.view] >= 0 && wk0[g.view] != g.wk0) return 1;
    wk0[g.view] = g.wk0;
    seen[g.view] |= 1 << g.wtap;
  }
  int kused = 0;
  for (int v = 0; v < d->nviews; ++v) {
    if (seen[v] != 0x1FF) return 1;
    const mtbc_act_view& a = d->views[v];
    if (a.W != d->W || a.H != d->H || a.N != d->N || a.C % 8 != 0) return 1;
    kused += (a.C + 31) & ~31;   // K is padded to 32 per source (TMA zero-fills the channels a dense tensor lacks)
  }
  // N tile such that the layer's weights stay resident in shared memory next to at least two halo stages
  int kc_any = 32;
  for (int v = 0; v < d->nviews; ++v) if (((d->views[v].C + 31) & ~31) % 64 == 0 && wk0[v] % 64 == 0) kc_any = 64;
  const int a_stage = ((kHaloRows * kc_any * 2) + 1023) & ~1023;
  int BN = 0;
  for (int bn = 16; bn <= 256 && bn <= d->ncols; bn += 16) {
    if (d->ncols % bn != 0) continue;
    const int wbytes = 9 * bn * kused * 2;
    // single-output layers keep the historical 120 KB cap (more halo stages); routed outputs (fused data gradients)
    // prefer wide N tiles: the dy halo is then read once per tile instead of once per source
    const char* cap_env = getenv("MTBC_HALO_WCAP_KB");
    const int cap = d->nouts > 0 ? 200 * 1024 - 2 * a_stage : (cap_env ? atoi(cap_env) : 120) * 1024;
    if (wbytes <= cap) BN = bn;
  }
  if (BN < 32 && BN != d->ncols) return 1;
  // Wide layers whose weights only fit as narrow N tiles are better served by the generic kernel (weights streamed,
  // one wide N tile): e.g. 192 -> 192 @32x32 ran 6 N tiles of 32 columns here (290 TF/s) against ~980 TF/s there.
  {
    const char* w_env = getenv("MTBC_HALO_WIDE_NCOLS");
    const int wide = w_env ? atoi(w_env) : 96;
    const char* wb_env = getenv("MTBC_HALO_WIDE_BN");
    const int wide_bn = wb_env ? atoi(wb_env) : 64;
    if (d->ncols >= wide && BN < wide_bn && d->nouts == 0) return 1;
    if (d->ncols >= wide && BN < 64 && d->nouts > 0) return set_error(MTBC_ERR_INVALID, "conv_gemm: fused data gradient would need N tiles of %d columns", BN);
  }

  // G = 2 (two output rows per accumulator row) for single-N-tile 32-column layers on planes with H % 32 == 0: the
  // 24-channel full-resolution convolutions, whose N = 32 MMAs cost the same cycles as N = 64 ones
  const char* g_env = getenv("MTBC_HALO_G");
  int G = 1;
  if (BN == 32 && d->H % 32 == 0 && d->nouts <= 1 && !(g_env && g_env[0] == '1')) G = 2;

  ConvHaloOp* op = new ConvHaloOp();
  ConvHaloParams& p = op->p;
  memset(&p, 0, sizeof(p));
  p.nsrc = d->nviews;
  p.G = G; p.bn1 = BN;
  { const char* lr = getenv("MTBC_HALO_LATE_RELEASE"); p.late_release = (lr && lr[0] == '1') ? 1 : 0; }
  { const char* dg = getenv("MTBC_HALO_DBG"); p.dbg = dg ? atoi(dg) : 0; }
  { const char* ln = getenv("MTBC_HALO_LANES"); p.lanes = (ln && ln[0] == '1') ? 1 : 2; }
  const int halo_h = 16 * G + 2;
  int kcmax = 32, b_off = 0;
  bool use32 = false, use64 = false;
  // G = 2 doubles the halo tile: keep 32-channel chunks so that at least three stages fit next to the weights
  const bool allow64 = (G == 1) || (9 * BN * kused * 2 <= 64 * 1024);
  for (int v = 0; v < d->nviews; ++v) {
    const mtbc_act_view& a = d->views[v];
    const int aC = (a.C + 31) & ~31;
    const int kc = (allow64 && aC % 64 == 0 && wk0[v] % 64 == 0) ? 64 : 32;
    if (kc > kcmax) kcmax = kc;
    (kc == 64 ? use64 : use32) = true;
    p.src[v].kc = (int16_t)kc; p.src[v].nchunk = (int16_t)(aC / kc); p.src[v].wk0 = wk0[v]; p.src[v].b_off = b_off;
    {
      const char* k_env = getenv("MTBC_HALO_SKIP_ZERO_K");
      p.src[v].drop_last = (aC - a.C >= 16 && !(k_env && k_env[0] == '0')) ? 1 : 0;
    }
    b_off += aC * BN * 2;
    if (wk0[v] % 32 != 0 || wk0[v] + aC > d->w_ktot) { delete op; return set_error(MTBC_ERR_INVALID, "conv_halo: weight columns out of range"); }
    int rc = encode_act(&p.amap[v], a, kc, kHaloW, halo_h, 1);
    if (rc) { delete op; return rc; }
  }
  p.b_tap_bytes = b_off;
  p.b_total_bytes = 9 * b_off;
  if (use32) { int rc = encode_w(&p.wmap[0], d->wpack, d->w_ktot, d->ncols, 9, 32, BN); if (rc) { delete op; return rc; } }
  if (use64) { int rc = encode_w(&p.wmap[1], d->wpack, d->w_ktot, d->ncols, 9, 64, BN); if (rc) { delete op; return rc; } }
  p.W = d->W; p.H = d->H; p.N = d->N;
  p.tiles_w = d->W / 8; p.tiles_h = d->H / (16 * G);
  p.n_mtiles = p.tiles_w * p.tiles_h * d->N;
  p.BN = G * BN;
  p.tmem_cols = tmem_cols_pow2(2 * G * BN);
  p.a_stage_bytes = ((kHaloW * halo_h * kcmax * 2) + 1023) & ~1023;
  int total_chunks = 0;
  for (int v = 0; v < d->nviews; ++v) total_chunks += p.src[v].nchunk;
  if (total_chunks > 64) { delete op; return 1; }   // chunk schedule table of the MMA lane (s_cb)
  int stages = (200 * 1024 - p.b_total_bytes) / p.a_stage_bytes;
  if (stages > kHaloMaxStages) stages = kHaloMaxStages;
  // with small resident weights keep the footprint below half an SM so two CTAs overlap each other's epilogues
  const int half = 100 * 1024;
  if (p.b_total_bytes + 3 * p.a_stage_bytes <= half) {
    int s2 = (half - p.b_total_bytes) / p.a_stage_bytes;
    if (s2 < stages) stages = s2;
  }
  if (stages < 2) { delete op; return 1; }
  p.stages = stages;
  op->smem_bytes = p.b_total_bytes + stages * p.a_stage_bytes + 1024;
  int ctas_per_sm = ((op->smem_bytes + 6 * 1024) * 2 <= 227 * 1024 && 2 * p.tmem_cols <= 512) ? 2 : 1;
  const char* c_env = getenv("MTBC_HALO_CTAS");
  // forward G = 2 layers keep 64 statistics accumulators per epilogue thread: the 2-CTA variant (96 registers) spills
  // them, one CTA per SM with a deeper halo ring measured 4 % faster (3.24 -> 3.11 ms over the 35 forward convs)
  // The same holds for G = 1 layers with 64 statistics columns: conv_1_0.conv_0 ([24] -> 48 @128^2) ran 69 us as two CTAs
  // per SM and 34 us as one (tools/profile_plan.py with MTBC_HALO_CTAS=1, round 2).
  const bool stats_heavy = d->stat_sum != nullptr && G * BN >= 64 && !getenv("MTBC_HALO_STATS2CTA");
  // (fused InstanceNorm backward statistics always run one CTA per SM: the 96-register variant spills their epilogue)
  const bool one_cta = (c_env && c_env[0] == '1') || d->bwd_y != nullptr ||
                       ((G == 2 || stats_heavy) && d->stat_sum != nullptr && !(c_env && c_env[0] == '2'));
  if (one_cta && ctas_per_sm == 2) {
    ctas_per_sm = 1;
    int st1 = (200 * 1024 - p.b_total_bytes) / p.a_stage_bytes;
    if (st1 > kHaloMaxStages) st1 = kHaloMaxStages;
    p.stages = st1;
    op->smem_bytes = p.b_total_bytes + st1 * p.a_stage_bytes + 1024;
  }
  op->ctas_per_sm = ctas_per_sm;
  {
    const char* e_env = getenv("MTBC_HALO_EPI");
    // (pixel-pair view of a 24 -> 24 layer: G = 1, 64 columns with statistics: the 8-warp epilogue walks 32 columns
    //  per warp and tile; MTBC_HALO_EPI4_G1=1 tries the same for every G = 1 statistics layer of 64 columns)
    const char* e1_env = getenv("MTBC_HALO_EPI4_G1");
    const bool g1_64 = G == 1 && p.BN == 64 && d->stat_sum != nullptr &&
                       (d->stat_fold > 0 || (e1_env && e1_env[0] == '1'));
    const bool want4 = ((G == 2 && p.BN == 64) || g1_64) && !(e_env && e_env[0] == '2');
    if (want4) {   // 16 epilogue warps, one CTA per SM (576 threads)
      if (ctas_per_sm == 2) {
        int st1 = (200 * 1024 - p.b_total_bytes) / p.a_stage_bytes;
        if (st1 > kHaloMaxStages) st1 = kHaloMaxStages;
        p.stages = st1;
        op->smem_bytes = p.b_total_bytes + st1 * p.a_stage_bytes + 1024;
        ctas_per_sm = 1;
        op->ctas_per_sm = 1;
      }
      op->epi_parts = 4;
    }
  }
  // Two MMA lanes only when the ring holds a whole number of lane PAIRS of tiles (stages % (2 * chunks per tile) == 0):
  // each lane then revisits only ring slots whose previous fill it consumed itself, so its one-bit mbarrier parity can
  // never be a lap ahead of the barrier (a lane that skips the other lane's chunks could otherwise pass a wait on a
  // slot whose previous fill has not even landed).  Otherwise, and with two CTAs per SM, one lane.
  const char* sp_env = getenv("MTBC_HALO_SPLIT");
  // (not in deterministic mode: two lanes adding into ONE accumulator race for the order of their fp32 additions,
  //  which made multi-source layers differ in the last bit from run to run -- tools/diag_det.py)
  const bool det = (current_mode() & MODE_DETERMINISTIC) != 0;
  if (ctas_per_sm == 1 && p.lanes == 2 && total_chunks >= 2 && p.stages >= 4 && !(sp_env && sp_env[0] == '0') && !det) {
    // split-chunk lanes: every tile's chunks alternate between the two lanes (see the kernel); needs an even ring
    p.lanes = 3;
    if (p.stages & 1) {
      p.stages -= 1;
      op->smem_bytes = p.b_total_bytes + p.stages * p.a_stage_bytes + 1024;
    }
  } else if (ctas_per_sm == 2) {
    p.lanes = 1;
  } else if (p.lanes == 2) {
    const int pair = 2 * total_chunks;
    const int s2 = (p.stages / pair) * pair;
    if (s2 >= pair) {
      p.stages = s2;
      op->smem_bytes = p.b_total_bytes + s2 * p.a_stage_bytes + 1024;
    } else {
      p.lanes = 1;
    }
  }
  {
    // zero blocks of the pair operand are not issued (halo_issue_chunk_pair24); MTBC_PAIR_SKIP=0 issues all 27
    const char* ps_env = getenv("MTBC_PAIR_SKIP");
    p.pair24 = (d->stat_fold == 24 && d->nviews == 1 && d->views[0].C == 48 && d->ncols == 64 && BN == 64 && G == 1 &&
                p.src[0].kc == 64 && p.src[0].nchunk == 1 && p.lanes != 3 && !(ps_env && ps_env[0] == '0')) ? 1 : 0;
  }
  int gx = sm_count() * ctas_per_sm;
  if (gx > p.n_mtiles) gx = p.n_mtiles;
  p.stat_C = d->stat_C;
  p.stat_fold = d->stat_fold;
  if (d->nouts > 0) {
    p.nouts = d->nouts;
    for (int i = 0; i < d->nouts; ++i) {
      p.outs[i].ptr = reinterpret_cast<__nv_bfloat16*>(d->outs[i].ptr);
      p.outs[i].out_C = d->outs[i].out_C; p.outs[i].col0 = d->outs[i].col0;
      p.outs[i].col_end = d->outs[i].col0 + d->outs[i].ncols; p.outs[i].accumulate = d->outs[i].accumulate;
    }
  } else {
    p.nouts = 1;
    p.outs[0].ptr = reinterpret_cast<__nv_bfloat16*>(d->out);
    p.outs[0].out_C = d->out_C; p.outs[0].col0 = 0; p.outs[0].col_end = d->ncols; p.outs[0].accumulate = d->accumulate;
  }
  p.bias = d->bias; p.stat_sum = d->stat_sum; p.stat_sq = d->stat_sq;
  if (d->bwd_y != nullptr) {
    // fused InstanceNorm backward statistics: served by the register-resident statistics epilogue only
    if (d->nouts != 0 || d->accumulate || d->bias || !d->stat_sum || !d->stat_sq || !d->bwd_mean || !d->bwd_rstd ||
        (p.BN != 32 && p.BN != 64) || d->stat_C != d->out_C) {
      delete op;
      return set_error(MTBC_ERR_INVALID, "conv_gemm: fused InstanceNorm backward statistics need a single-output, "
                       "non-accumulating data gradient with N tiles of 32 or 64 columns (got %d)", p.BN);
    }
    p.bwd_y = reinterpret_cast<const __nv_bfloat16*>(d->bwd_y);
    p.bwd_mean = d->bwd_mean; p.bwd_rstd = d->bwd_rstd; p.bwd_gamma = d->bwd_gamma; p.bwd_beta = d->bwd_beta;
    p.bwd_slope = d->bwd_slope;
    // y tile ring: three tiles of 8 x 16 G pixels x bn1 channels behind the halo ring (the halo ring gives up stages
    // two at a time, which keeps every lane-parity rule above intact)
    // (dense tile, pixel pitch 2 * out_C bytes: 16-byte shared-memory reads of 8 neighbouring pixels are conflict free
    //  for 24 / 40 / 56 channels, two-way for 48, 4- / 8-way for 32 / 64.  Measured at B = 32 (tools/diag_fused_dgrad.py):
    //  24 channels @256^2 66 -> 89 us for a 40 us reduction pass saved; 48 channels @128^2 34 -> 57 us for 21 us saved:
    //  only the conflict-free pitches are served)
    // (pixel-pair view of a 24-channel tensor: 96-byte rows, two-way conflicts on the y reads, and half the halo rows)
    if (((d->out_C / 8) % 2 == 0 && !(d->stat_fold > 0 && d->out_C == 48)) || d->ncols != BN) {
      delete op;
      return set_error(MTBC_ERR_INVALID, "conv_gemm: fused InstanceNorm backward statistics: channel pitch %d not served", d->out_C);
    }
    p.y_rowb = d->out_C * 2;
    p.y_stage_bytes = (8 * 16 * G * p.y_rowb + 127) & ~127;
    p.y_stages = d->stat_fold > 0 ? 2 : 3;   // (pair view: 72 KB of paired weights; two y slots leave four halo stages)
    const int ybytes = p.y_stages * p.y_stage_bytes;
    while (p.stages > 2 && p.b_total_bytes + p.stages * p.a_stage_bytes + ybytes + 1024 > 212 * 1024) p.stages -= 2;
    if (p.b_total_bytes + p.stages * p.a_stage_bytes + ybytes + 1024 > 212 * 1024 || total_chunks != 1) {
      delete op;
      return set_error(MTBC_ERR_INVALID, "conv_gemm: fused InstanceNorm backward statistics: no room for the y ring");
    }
    op->smem_bytes = p.b_total_bytes + p.stages * p.a_stage_bytes + ybytes + 1024;
    int rc = encode_rows(&p.ymap, d->bwd_y, d->out_C, d->W, d->H, d->N, 8, 16 * G);
    if (rc) { delete op; return rc; }
  }
  op->grid = dim3(gx, d->ncols / BN, 1);
  op->flops = 2.0 * double(d->N) * d->H * d->W * double(d->ncols) * kused * 9.0;
  cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_halo_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_halo_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024);
  if (e == cudaSuccess && d->bwd_y) e = cudaFuncSetAttribute(conv_halo_kernel<1, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024);
  if (e == cudaSuccess && d->bwd_y) e = cudaFuncSetAttribute(conv_halo_kernel<1, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024);
  if (e != cudaSuccess) { delete op; return set_error(MTBC_ERR_CUDA, "cudaFuncSetAttribute(conv_halo): %s", cudaGetErrorString(e)); }
  *out = op;
  return 0;
}

int wgrad_halo_try_create(const mtbc_wgrad_desc* d, OpBase** out) {
  if (halo_disabled()) return 1;
  if (d->a_nviews != 1 || d->b_nviews != 1 || d->ntaps != 9) return 1;
  if (d->H % 16 != 0 || d->W % 8 != 0) return 1;
  for (int t = 0; t < 9; ++t) {
    const mtbc_wgrad_tap& tp = d->taps[t];
    if (tp.a_view != 0 || tp.b_view != 0 || tp.a_dh != t / 3 - 1 || tp.a_dw != t % 3 - 1) return 1;
  }
  const mtbc_act_view& a = d->a_views[0];
  const mtbc_act_view& b = d->b_views[0];
  if (a.C % 8 != 0 || b.C % 8 != 0 || a.W != d->W || a.H != d->H || a.N != d->N) return 1;
  const int aCp = (a.C + 31) & ~31, bCp = (b.C + 31) & ~31;   // GEMM extents; TMA zero-fills what a dense tensor lacks
  const int b_kc = (bCp % 64 == 0) ? 64 : 32;
  WgradHaloOp* op = new WgradHaloOp();
  WgradHaloParams& p = op->p;
  memset(&p, 0, sizeof(p));
  int rc = encode_act(&p.amap, a, 32, kHaloW, 16, 1);
  if (rc) { delete op; return rc; }
  rc = encode_act(&p.bmap, b, b_kc, 8, kHaloH, 1);
  if (rc) { delete op; return rc; }
  p.a_C = a.C; p.b_kc = b_kc; p.NN = 3 * b_kc; p.n_tiles = bCp / b_kc;
  p.tmem_cols = tmem_cols_pow2(p.NN);
  p.tiles_w = d->W / 8; p.tiles_h = d->H / 16;
  p.n_ptiles = p.tiles_w * p.tiles_h * d->N;
  p.a_bytes = ((kHaloW * 16 * 64) + 1023) & ~1023;            // 10 KB
  p.b_bytes = ((8 * kHaloH * b_kc * 2) + 1023) & ~1023;       // 9 / 18 KB
  p.stage_bytes = p.a_bytes + p.b_bytes;
  int stages = (96 * 1024) / p.stage_bytes;  // <= half an SM: two CTAs hide each other's epilogue / start-up
  if (stages > kHaloMaxStages) stages = kHaloMaxStages;
  if (stages < 2) stages = 2;
  p.stages = stages;
  op->smem_bytes = stages * p.stage_bytes + 1024;
  p.n_rows = d->n_rows; p.ld_k = d->ld_k; p.k0 = d->k0; p.dw_acc = d->dw_acc;
  const int chunks = aCp / 32;
  const int base = chunks * p.n_tiles;
  int splits = d->splits;
  if (splits <= 0) {
    splits = (2 * sm_count() + base - 1) / base;
    int maxs = (p.n_ptiles + 3) / 4; if (maxs < 1) maxs = 1;
    if (splits > maxs) splits = maxs;
  }
  if (splits > p.n_ptiles) splits = p.n_ptiles;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  p.splits = splits;
  op->grid = dim3(chunks, splits, p.n_tiles);
  op->flops = 2.0 * double(d->N) * d->H * d->W * double(a.C) * double(b.C) * 9.0;
  cudaError_t e = cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (e != cudaSuccess) { delete op; return set_error(MTBC_ERR_CUDA, "cudaFuncSetAttribute(wgrad_halo): %s", cudaGetErrorString(e)); }
  *out = op;
  return 0;
}

struct WgradMultiOp : public OpBase {
  WgradMultiParams p;
  dim3 grid;
  int smem_bytes;
  double flops;
  int launch(cudaStream_t st) override {
    launch_pdl(wgrad_halo_multi_kernel, grid, dim3(128), smem_bytes, st, p);
    return check_launch("wgrad_halo_multi_kernel");
  }
  double op_flops() const override { return flops; }
};

int wgrad_halo_multi_create(const mtbc_wgrad_multi_desc* d, OpBase** out) {
  if (d->nsrc < 1 || d->nsrc > MTBC_MAX_VIEWS) return set_error(MTBC_ERR_INVALID, "wgrad_multi: nsrc");
  if (halo_disabled() || d->H % 16 != 0 || d->W % 8 != 0)
    return set_error(MTBC_ERR_INVALID, "wgrad_multi: needs H %% 16 == 0 and W %% 8 == 0");
  const mtbc_act_view& b = d->dy;
  if (b.C % 8 != 0 || b.W != d->W || b.H != d->H || b.N != d->N) return set_error(MTBC_ERR_INVALID, "wgrad_multi: dy view");
  const int bCp = (b.C + 31) & ~31;
  const int b_kc = (bCp % 64 == 0) ? 64 : 32;
  WgradMultiOp* op = new WgradMultiOp();
  WgradMultiParams& p = op->p;
  memset(&p, 0, sizeof(p));
  int rc = encode_act(&p.bmap, b, b_kc, 8, kHaloH, 1);
  if (rc) { delete op; return rc; }
  int nch = 0;
  double ksum = 0;
  for (int v = 0; v < d->nsrc; ++v) {
    const mtbc_act_view& a = d->x[v];
    if (a.C % 8 != 0 || a.W != d->W || a.H != d->H || a.N != d->N) { delete op; return set_error(MTBC_ERR_INVALID, "wgrad_multi: x view %d", v); }
    rc = encode_act(&p.amap[v], a, 32, kHaloW, 16, 1);
    if (rc) { delete op; return rc; }
    const int aCp = (a.C + 31) & ~31;
    for (int ch = 0; ch < aCp; ch += 32) {
      if (nch >= kWgMaxChunks) { delete op; return set_error(MTBC_ERR_INVALID, "wgrad_multi: more than %d chunks", kWgMaxChunks); }
      p.c_view[nch] = (int16_t)v; p.c_ch[nch] = (int16_t)ch; p.c_ac[nch] = (int16_t)a.C; p.c_col[nch] = d->k0[v] + ch;
      ++nch;
    }
    ksum += a.C;
  }
  p.nchunks = nch;
  p.b_kc = b_kc; p.NN = 3 * b_kc; p.n_tiles = bCp / b_kc;
  // accumulators per CTA: 5 fit TMEM at NN = 96; at NN = 192 two would need all 512 columns and leave one CTA per SM
  // (measured slower than one chunk per CTA with two CTAs per SM), so wide dy tiles keep one chunk per CTA
  const int gmax = p.NN == 96 ? 5 : 1;
  const int ngroups = (nch + gmax - 1) / gmax;
  p.group = (nch + ngroups - 1) / ngroups;
  p.tmem_cols = tmem_cols_pow2(p.group * p.NN);
  p.tiles_w = d->W / 8; p.tiles_h = d->H / 16;
  p.n_ptiles = p.tiles_w * p.tiles_h * d->N;
  p.a_bytes = ((kHaloW * 16 * 64) + 1023) & ~1023;            // 10 KB
  p.b_bytes = ((8 * kHaloH * b_kc * 2) + 1023) & ~1023;       // 9 / 18 KB
  p.stage_bytes = p.b_bytes + p.group * p.a_bytes;
  // two CTAs per SM when two 2-stage rings fit, else one CTA with a deeper ring
  int ctas = 2;
  int stages = (100 * 1024) / p.stage_bytes;
  if (stages < 2) { ctas = 1; stages = (200 * 1024) / p.stage_bytes; }
  if (stages > kHaloMaxStages) stages = kHaloMaxStages;
  if (stages < 2 || ctas * p.tmem_cols > 512) {
    if (ctas == 2) { ctas = 1; stages = (200 * 1024) / p.stage_bytes; if (stages > kHaloMaxStages) stages = kHaloMaxStages; }
    if (stages < 2) { delete op; return set_error(MTBC_ERR_INVALID, "wgrad_multi: stage does not fit shared memory"); }
  }
  p.stages = stages;
  op->smem_bytes = stages * p.stage_bytes + 1024;
  p.n_rows = d->n_rows; p.ld_k = d->ld_k; p.dw_acc = d->dw_acc;
  const int base = ngroups * p.n_tiles;
  int splits = d->splits;
  if (splits <= 0) {
    splits = (ctas * sm_count() + base - 1) / base;
    int maxs = (p.n_ptiles + 3) / 4; if (maxs < 1) maxs = 1;
    if (splits > maxs) splits = maxs;
  }
  { const char* dv = getenv("MTBC_WGRAD_SPLIT_DIV"); if (dv && atoi(dv) > 1) splits = (splits + atoi(dv) - 1) / atoi(dv); }
  if (splits > p.n_ptiles) splits = p.n_ptiles;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  p.splits = splits;
  op->grid = dim3(ngroups, splits, p.n_tiles);
  op->flops = 2.0 * double(d->N) * d->H * d->W * ksum * double(b.C) * 9.0;
  cudaError_t e = cudaFuncSetAttribute(wgrad_halo_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (e != cudaSuccess) { delete op; return set_error(MTBC_ERR_CUDA, "cudaFuncSetAttribute(wgrad_multi): %s", cudaGetErrorString(e)); }
  *out = op;
  return 0;
}

}  // namespace mtbc

// Debug helper (not part of the documented ABI surface used by the product path): copy / reset the cycle breakdown.
extern "C" int mtbc_debug_halo_times(long long* out8, int reset) {
  long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (out8 && cudaMemcpyFromSymbol(out8, mtbc::g_halo_dbg, sizeof(z)) != cudaSuccess) return -1;
  if (reset && cudaMemcpyToSymbol(mtbc::g_halo_dbg, z, sizeof(z)) != cudaSuccess) return -1;
  return 0;
}
