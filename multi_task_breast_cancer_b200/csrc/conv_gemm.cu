// Implicit-GEMM convolution kernels for sm_100a: TMA (tiled mode, zero OOB fill = conv padding) -> swizzled smem ->
// tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> tcgen05.ld epilogue.
//
//  * conv_gemm_kernel : forward / data-gradient.  A = 128 output pixels x KC channels per (tap, source, chunk)
//    K-block (K-major, 64B/128B swizzle), B = BN x KC slice of the packed weights (K-major).  A skip concatenation is
//    just more (tap, source) segments: the concat tensor is never materialised (MTUNetPlusPlus.py:107-118).
//  * wgrad_kernel     : weight gradient.  Reduction runs over pixels, so both operands are MN-major views of the very
//    same NHWC boxes; accumulators for up to 9 taps live side by side in TMEM; split over pixel ranges with fp32 atomics.
#include "ptx.cuh"
#include "internal.h"
#include "act_io.cuh"
#include <cuda.h>
#include <vector>
#include <algorithm>
#include <string.h>
#include <type_traits>
#include <stdlib.h>

namespace mtbc {

// ------------------------------------------------------------------------------------------------ kernel params
struct SegDev {
  int16_t view, dh, dw, wtap;
  int32_t wk0;
  int16_t nchunk, kc;
};  // 16 B

struct ConvGemmParams {
  CUtensorMap amap[MTBC_MAX_VIEWS];
  CUtensorMap wmap[2];  // [0]: 32-channel boxes / 64B swizzle, [1]: 64-channel boxes / 128B swizzle
                        // (fp32 operands: [0] = the 32-float / 128B-swizzle box of wpack, [1] = the same of wpack_lo)
  SegDev seg[MTBC_MAX_SEGS];
  int32_t nseg;
  int32_t TW, TH, TN, tiles_w, tiles_h, n_mtiles, n_ntiles;
  int32_t W, H, N;
  int32_t BN, tmem_cols, stages, a_stage_bytes, stage_bytes;
  int32_t epi_mode, out_C, up_k, up_cp, accumulate, stat_C, bias_len;
  void* out;
  const float* bias;
  float* stat_sum;
  float* stat_sq;
};

constexpr int kMaxStages = 8;
constexpr int kMaxBias = 1024;

// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue (TMEM lane quarter =
// warp % 4).  Two TMEM accumulators: the epilogue of work item i (TMEM -> registers -> bf16 global stores, the
// dominant cost of the transposed convolutions, whose GEMM is a single K block) overlaps the loads and MMAs of item
// i+1.  Work item = (pixel tile, N tile), N tile fastest so neighbouring CTAs share the A tile in L2.
//
// PREC 0: bf16 operands (kind::f16).  PREC 1: fp32 operands read as TF32 (kind::tf32): a 32-channel chunk is a
// 128-byte row (the geometry of the bf16 64-channel chunk), four K = 8 instructions per chunk.  PREC 3: 3xTF32 -- the
// four extra warps 10..13 split every landed A tile in place into its TF32 rounding and, in a second buffer, the
// remainder; weights arrive pre-split (wpack, wpack_lo); three products per K step give fp32-grade results.
template <int PREC>
__global__ void __launch_bounds__(PREC == 3 ? 448 : 320) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  using OutT = typename std::conditional<PREC == 0, __nv_bfloat16, float>::type;
  constexpr int ES = PREC == 0 ? 2 : 4;          // operand element size
  constexpr int NT = PREC == 3 ? 448 : 320;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_full[kMaxStages];
  __shared__ uint64_t s_split[kMaxStages];       // PREC 3: "A tile of this stage has been split"
  __shared__ uint64_t s_empty[kMaxStages];
  __shared__ uint64_t s_accfull[2], s_accempty[2];
  __shared__ uint32_t s_tmem;
  __shared__ float s_stat[2][256];
  __shared__ __align__(16) float s_bias[kMaxBias];   // whole bias vector (epi_mode 0: ncols, pixel shuffle: up_cp)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int BN = p.BN;
  const int n_items = p.n_mtiles * p.n_ntiles;

  for (int i = tid; i < 512; i += NT) (&s_stat[0][0])[i] = 0.f;
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 1);
      mbar_init(&s_split[s], 4);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(&s_accfull[b], 1); mbar_init(&s_accempty[b], 8); }
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
  // programmatic dependent launch: nothing above read global memory (ptx.cuh: pdl_wait)
  pdl_trigger();
  pdl_wait();
  if (p.bias != nullptr) {   // uniform
    for (int i = tid; i < p.bias_len; i += NT) s_bias[i] = p.bias[i];
    __syncthreads();
  }

  if (warp == 0) {
    if (elect_one()) {
      // -------------------------------------------------------------- TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int mt = item / p.n_ntiles, ntile = item - mt * p.n_ntiles;
        const int w0 = (mt % p.tiles_w) * p.TW;
        const int h0 = ((mt / p.tiles_w) % p.tiles_h) * p.TH;
        const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.TN;
        for (int s = 0; s < p.nseg; ++s) {
          const SegDev sg = p.seg[s];
          const CUtensorMap* am = &p.amap[sg.view];
          const CUtensorMap* wm = &p.wmap[(PREC == 0 && sg.kc == 64) ? 1 : 0];
          const uint32_t bytes = static_cast<uint32_t>((128 + (PREC == 3 ? 2 : 1) * BN) * sg.kc * ES);
          for (int ch = 0; ch < sg.nchunk; ++ch) {
            mbar_wait(&s_empty[stage], phase ^ 1u);
            uint8_t* a_dst = smem + stage * p.stage_bytes;
            uint8_t* b_dst = a_dst + (PREC == 3 ? 2 : 1) * p.a_stage_bytes;
            mbar_arrive_expect_tx(&s_full[stage], bytes);
            tma_load_4d(a_dst, am, &s_full[stage], ch * sg.kc, w0 + sg.dw, h0 + sg.dh, n0);
            tma_load_3d(b_dst, wm, &s_full[stage], sg.wk0 + ch * sg.kc, ntile * BN, sg.wtap);
            if (PREC == 3)
              tma_load_3d(b_dst + BN * sg.kc * ES, &p.wmap[1], &s_full[stage], sg.wk0 + ch * sg.kc, ntile * BN, sg.wtap);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // -------------------------------------------------------------- MMA issuer (one thread)
      const uint32_t idesc = PREC == 0 ? umma_idesc_bf16(128, BN, 0, 0) : umma_idesc_tf32(128, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&s_accempty[buf], acc_phase ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_addr = tmem_base + static_cast<uint32_t>(buf * BN);
        uint32_t accumulate = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const SegDev sg = p.seg[s];
          const uint32_t hi = umma_desc_hi(8u * sg.kc * ES, sg.kc * ES == 128 ? 2u : 4u);
          for (int ch = 0; ch < sg.nchunk; ++ch) {
            mbar_wait(PREC == 3 ? &s_split[stage] : &s_full[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + stage * p.stage_bytes);
            const uint32_t a_lo = umma_desc_lo(a_addr, 16);
            const uint32_t b_lo = umma_desc_lo(a_addr + (PREC == 3 ? 2 : 1) * p.a_stage_bytes, 16);
            // low descriptor words advance by compile-time steps: keeps the issuing lane at the tensor pipe's floor
            if (PREC == 0) {
              umma_bf16_lohi(d_addr, a_lo, hi, b_lo, hi, idesc, accumulate);
              umma_bf16_lohi(d_addr, a_lo + 2, hi, b_lo + 2, hi, idesc, 1u);
              if (sg.kc == 64) {
                umma_bf16_lohi(d_addr, a_lo + 4, hi, b_lo + 4, hi, idesc, 1u);
                umma_bf16_lohi(d_addr, a_lo + 6, hi, b_lo + 6, hi, idesc, 1u);
              }
            } else {
              // 32 fp32 channels = one 128-byte row = four K = 8 steps of 32 bytes
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_tf32_lohi(d_addr, a_lo + 2 * k, hi, b_lo + 2 * k, hi, idesc, (k == 0) ? accumulate : 1u);
              if (PREC == 3) {
                // remainder terms: a_lo * w_hi and a_hi * w_lo (a_lo * w_lo is below fp32 resolution)
                const uint32_t a2 = umma_desc_lo(a_addr + p.a_stage_bytes, 16);
                const uint32_t b2 = umma_desc_lo(a_addr + 2 * p.a_stage_bytes + BN * sg.kc * ES, 16);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32_lohi(d_addr, a2 + 2 * k, hi, b_lo + 2 * k, hi, idesc, 1u);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32_lohi(d_addr, a_lo + 2 * k, hi, b2 + 2 * k, hi, idesc, 1u);
              }
            }
            accumulate = 1;
            umma_commit(&s_empty[stage]);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
        umma_commit(&s_accfull[buf]);
      }
    }
  } else if (PREC == 3 && warp >= 10) {
    // ---------------------------------------------------------------- 3xTF32 splitter: 4 warps, 128 threads.  In place
    // a -> tf32_rn(a); remainder a - tf32_rn(a) (itself exactly representable in fp32) into the second A buffer at
    // the same (swizzled) offset; then make the generic-proxy writes visible to the tensor core's async proxy.
    int stage = 0;
    uint32_t phase = 0;
    const int st = tid - 320;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      for (int s = 0; s < p.nseg; ++s) {
        const int nchunk = p.seg[s].nchunk;
        for (int ch = 0; ch < nchunk; ++ch) {
          mbar_wait(&s_full[stage], phase);
          float4* a_hi = reinterpret_cast<float4*>(smem + stage * p.stage_bytes);
          float4* a_lo = reinterpret_cast<float4*>(smem + stage * p.stage_bytes + p.a_stage_bytes);
          const int n4 = p.a_stage_bytes >> 4;
          for (int i = st; i < n4; i += 128) {
            const float4 v = a_hi[i];
            float4 h, l;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(*reinterpret_cast<uint32_t*>(&h.x)) : "f"(v.x));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(*reinterpret_cast<uint32_t*>(&h.y)) : "f"(v.y));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(*reinterpret_cast<uint32_t*>(&h.z)) : "f"(v.z));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(*reinterpret_cast<uint32_t*>(&h.w)) : "f"(v.w));
            l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
            a_hi[i] = h;
            a_lo[i] = l;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_split[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 2 && warp < 10) {
    // ---------------------------------------------------------------- epilogue: 8 warps, two per TMEM lane quarter,
    // each pair splitting the N tile's 16-column chunks in halves.  Short dependent chain per chunk (these warps are
    // alone on their schedulers): running output pointers instead of divisions, bias from shared memory, TMEM loads
    // issued in pairs, 32-byte stores, accumulation as fire-and-forget bf16 reductions.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int TW = p.TW, TH = p.TH, TN = p.TN, tiles_w = p.tiles_w, tiles_h = p.tiles_h, n_ntiles = p.n_ntiles;
    const int tw = row % TW, th = (row / TW) % TH, tn = row / (TW * TH);
    const int etid = (warp - 2) * 32 + lane;
    const bool do_stats = (p.stat_sum != nullptr);
    const bool has_bias = (p.bias != nullptr);
    const bool acc = p.accumulate != 0;
    const bool wide = (p.out_C % 16) == 0;
    const int epi_mode = p.epi_mode, up_k = p.up_k, up_cp = p.up_cp, out_C = p.out_C;
    const int H = p.H, W = p.W, N = p.N;
    const int64_t Wo = static_cast<int64_t>(W) * up_k;
    OutT* const out = static_cast<OutT*>(p.out);
    const int nchunks = BN >> 4;
    const int k_begin = half ? (nchunks + 1) >> 1 : 0;
    const int k_end = half ? nchunks : (nchunks + 1) >> 1;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int mt = item / n_ntiles, ntile = item - mt * n_ntiles;
      const int w = (mt % tiles_w) * TW + tw;
      const int h = ((mt / tiles_w) % tiles_h) * TH + th;
      const int n0 = (mt / (tiles_w * tiles_h)) * TN;
      const int n = n0 + tn;
      const bool valid = (n < N) && (h < H) && (w < W);
      // destination of this thread's first chunk, then running (q, bcol) for the pixel shuffle
      const int col0 = ntile * BN + k_begin * 16;
      OutT* pixbase;
      int bcol, qi = 0, qj = 0;
      if (epi_mode == 0) {
        pixbase = out + ((static_cast<int64_t>(n) * H + h) * W + w) * out_C;
        bcol = col0;
      } else {
        pixbase = out + ((static_cast<int64_t>(n) * H * up_k + static_cast<int64_t>(h) * up_k) * Wo + static_cast<int64_t>(w) * up_k) * out_C;
        const int qq = col0 / up_cp;
        bcol = col0 - qq * up_cp;
        qi = qq / up_k;
        qj = qq - qi * up_k;
      }
      mbar_wait(&s_accfull[buf], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * BN);

      auto do_chunk = [&](int k, uint32_t (&r)[16]) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
        if (has_bias) {
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + bcol);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 bb = b4[i];
            v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
          }
        }
        if (do_stats) {
          float sv[16], sq[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            sv[i] = valid ? v[i] : 0.f;
            sq[i] = sv[i] * sv[i];
          }
          const float cs = warp_colsum16(sv, lane);
          const float cq = warp_colsum16(sq, lane);
          if ((lane & 1) == 0) {
            const int cc = k * 16 + col16_of_lane(lane);
            atomicAdd(&s_stat[0][cc], cs);
            atomicAdd(&s_stat[1][cc], cq);
          }
        }
        if (valid) {
          const int left = out_C - bcol;   // columns of this chunk that exist in the (possibly dense) destination
          emit16_n<OutT>(pixbase + (static_cast<int64_t>(qi) * Wo + qj) * out_C + bcol, v, acc,
                         left >= 16 ? 16 : (left >= 8 ? 8 : 0), wide);
        }
        bcol += 16;
        if (epi_mode != 0 && bcol == up_cp) {   // next sub-pixel of the transposed convolution
          bcol = 0;
          if (++qj == up_k) { qj = 0; ++qi; }
        }
      };
#pragma unroll 1
      for (int k = k_begin; k < k_end; k += 2) {
        const bool two = (k + 1 < k_end);   // warp uniform
        uint32_t r[2][16];
        tmem_ld16_nowait(taddr + k * 16, r[0]);
        if (two) tmem_ld16_nowait(taddr + k * 16 + 16, r[1]);
        tmem_wait_ld();
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (u == 0 || two) do_chunk(k + u, r[u]);
      }
      // accumulator fully read: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_accempty[buf]);
      if (do_stats) {
        // TN == 1 is enforced by the host when statistics are fused: the whole tile belongs to sample n0.
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int i = etid; i < BN; i += 256) {
          if (ntile * BN + i < p.stat_C) {   // GEMM pad columns of a dense tensor have no channel
            const int64_t o = static_cast<int64_t>(n0) * p.stat_C + ntile * BN + i;
            atomicAdd(p.stat_sum + o, s_stat[0][i]);
            atomicAdd(p.stat_sq + o, s_stat[1][i]);
          }
          s_stat[0][i] = 0.f;
          s_stat[1][i] = 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
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

// ------------------------------------------------------------------------------------------------ weight gradient
struct WgradParams {
  CUtensorMap amap[4];
  CUtensorMap bmap[4];
  int16_t t_aview[9], t_adh[9], t_adw[9], t_bview[9];
  int32_t ntaps, T;       // taps in total / per CTA
  int32_t a_C, a_kc, a_boxes;  // channels of A source, channels per box, boxes per 128-channel M block
  int32_t b_kc, b_boxes;  // boxes per BN tile
  int32_t BN, n_tiles, tmem_cols;
  int32_t TW, TH, TN, tiles_w, tiles_h, n_ptiles, splits;
  int32_t a_stage_bytes, b_stage_bytes, a_box_bytes, b_box_bytes;
  int32_t n_rows, ld_k, k0;
  int32_t a_stages, b_stages;  // ring depths, sized by the host to fill shared memory (bytes in flight, not MMA rate,
                               // bound this kernel: with 3 + 2 stages the 48 -> 48 transposed-conv gradient ran at 2.3 TB/s)
  float* dw_acc;
};

constexpr int kWgMaxStages = 8;

__global__ void __launch_bounds__(128) wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_afull[kWgMaxStages], s_aempty[kWgMaxStages];
  __shared__ uint64_t s_bfull[kWgMaxStages], s_bempty[kWgMaxStages];
  __shared__ uint64_t s_accum;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  const int kWgAStages = p.a_stages, kWgBStages = p.b_stages;
  uint8_t* smem_b = smem + kWgAStages * p.a_stage_bytes;

  const int split = blockIdx.x;
  const int mblk = blockIdx.y;
  const int tapgroup = blockIdx.z / p.n_tiles;
  const int ntile = blockIdx.z % p.n_tiles;
  const int BN = p.BN;
  const int T = p.T;

  // pixel tiles of this split
  const int per = (p.n_ptiles + p.splits - 1) / p.splits;
  const int pt_begin = split * per;
  const int pt_end = min(p.n_ptiles, pt_begin + per);
  const int n_my = max(0, pt_end - pt_begin);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kWgAStages; ++s) { mbar_init(&s_afull[s], 1); mbar_init(&s_aempty[s], 1); }
    for (int s = 0; s < kWgBStages; ++s) { mbar_init(&s_bfull[s], 1); mbar_init(&s_bempty[s], 1); }
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

  // number of A boxes that actually exist for this M block
  const int a_c0 = mblk * 128;
  const int a_nbox = min(p.a_boxes, (p.a_C - a_c0 + p.a_kc - 1) / p.a_kc);

  if (n_my > 0) {
    if (warp == 0 && elect_one()) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int pt = pt_begin; pt < pt_end; ++pt) {
        const int w0 = (pt % p.tiles_w) * p.TW;
        const int h0 = ((pt / p.tiles_w) % p.tiles_h) * p.TH;
        const int n0 = (pt / (p.tiles_w * p.tiles_h)) * p.TN;
        // B tile(s) of this pixel tile.  A tap group shares one b_view only if all its taps do; otherwise (convT)
        // the host sets T so that each CTA's taps are loaded as separate B stages (see below).
        for (int t = 0; t < T; ++t) {
          const int tap = tapgroup * T + t;
          const bool need_b = (t == 0) || (p.t_bview[tap] != p.t_bview[tap - 1]);
          if (need_b) {
            mbar_wait(&s_bempty[sb], pb ^ 1u);
            mbar_arrive_expect_tx(&s_bfull[sb], static_cast<uint32_t>(p.b_boxes * p.b_box_bytes));
            const CUtensorMap* bm = &p.bmap[p.t_bview[tap]];
            for (int b = 0; b < p.b_boxes; ++b)
              tma_load_4d(smem_b + sb * p.b_stage_bytes + b * p.b_box_bytes, bm, &s_bfull[sb],
                          ntile * BN + b * p.b_kc, w0, h0, n0);
            if (++sb == kWgBStages) { sb = 0; pb ^= 1u; }
          }
          // the transposed conv reads the SAME x tile for its k*k taps: load it once per pixel tile
          const bool need_a = (t == 0) || p.t_aview[tap] != p.t_aview[tap - 1] || p.t_adh[tap] != p.t_adh[tap - 1] ||
                              p.t_adw[tap] != p.t_adw[tap - 1];
          if (need_a) {
            mbar_wait(&s_aempty[sa], pa ^ 1u);
            mbar_arrive_expect_tx(&s_afull[sa], static_cast<uint32_t>(a_nbox * p.a_box_bytes));
            const CUtensorMap* am = &p.amap[p.t_aview[tap]];
            for (int b = 0; b < a_nbox; ++b)
              tma_load_4d(smem_a + sa * p.a_stage_bytes + b * p.a_box_bytes, am, &s_afull[sa], a_c0 + b * p.a_kc,
                          w0 + p.t_adw[tap], h0 + p.t_adh[tap], n0);
            if (++sa == kWgAStages) { sa = 0; pa ^= 1u; }
          }
        }
      }
    } else if (warp == 1 && elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, BN, 1, 1);
      const uint32_t a_hi = umma_desc_hi(8u * p.a_kc * 2u, p.a_kc == 64 ? 2u : 4u);
      const uint32_t b_hi = umma_desc_hi(8u * p.b_kc * 2u, p.b_kc == 64 ? 2u : 4u);
      const uint32_t a_k16 = 2u * p.a_kc, b_k16 = 2u * p.b_kc;  // 16 pixel rows, in 16-byte units
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int cur_b = -1, cur_a = -1;
      for (int pt = pt_begin; pt < pt_end; ++pt) {
        for (int t = 0; t < T; ++t) {
          const int tap = tapgroup * T + t;
          const bool need_b = (t == 0) || (p.t_bview[tap] != p.t_bview[tap - 1]);
          if (need_b) {
            if (cur_b >= 0) {  // release the previous B stage: all MMAs reading it have been issued
              umma_commit(&s_bempty[cur_b]);
            }
            mbar_wait(&s_bfull[sb], pb);
            cur_b = sb;
            if (++sb == kWgBStages) { sb = 0; pb ^= 1u; }
          }
          const bool need_a = (t == 0) || p.t_aview[tap] != p.t_aview[tap - 1] || p.t_adh[tap] != p.t_adh[tap - 1] ||
                              p.t_adw[tap] != p.t_adw[tap - 1];
          if (need_a) {
            if (cur_a >= 0) umma_commit(&s_aempty[cur_a]);   // all MMAs reading the previous A stage have been issued
            mbar_wait(&s_afull[sa], pa);
            cur_a = sa;
            if (++sa == kWgAStages) { sa = 0; pa ^= 1u; }
          }
          tc_fence_after();
          const uint32_t a_lo = umma_desc_lo(smem_u32(smem_a + cur_a * p.a_stage_bytes), p.a_box_bytes);
          const uint32_t b_lo = umma_desc_lo(smem_u32(smem_b + cur_b * p.b_stage_bytes), p.b_box_bytes);
          const uint32_t d_addr = tmem_base + static_cast<uint32_t>(t * BN);
#pragma unroll
          for (int k = 0; k < 8; ++k) {  // 128 pixels / 16
            umma_bf16_lohi(d_addr, a_lo + k * a_k16, a_hi, b_lo + k * b_k16, b_hi, idesc, (pt > pt_begin || k > 0) ? 1u : 0u);
          }
        }
      }
      umma_commit(&s_accum);
    }
    __syncwarp();
    mbar_wait(&s_accum, 0);
    tc_fence_after();
    __syncwarp();

    const int m = warp * 32 + lane;
    const int ci = a_c0 + m;
    const bool valid = ci < p.a_C;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    for (int t = 0; t < T; ++t) {
      const int tap = tapgroup * T + t;
      for (int c = 0; c < BN; c += 16) {
        float v[16];
        tmem_ld16(taddr + t * BN + c, v);
        if (valid) {
          float* dst = p.dw_acc + (static_cast<int64_t>(tap) * p.n_rows + ntile * BN + c) * p.ld_k + p.k0 + ci;
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


// ------------------------------------------------------------------------------------------------ fp32 weight gradient
// TF32 parity modes: dW[tap][co][k0 + ci] += sum_pixels A[pixel + (dh, dw)][ci] * B[pixel][co] in plain fp32 on the CUDA
// cores (exact products).  32 (ci) x 32 (co) tile per block, 32 pixels per shared-memory step, 2 x 2 outputs per
// thread, the pixel range split over blockIdx.z with fp32 atomics at the end.  Not a fast kernel: a checker-grade one.
struct WgradF32Params {
  const float* a[4];
  const float* b[4];
  int64_t a_sW[4], a_sH[4], a_sN[4], b_sW[4], b_sH[4], b_sN[4];
  int16_t t_aview[9], t_adh[9], t_adw[9], t_bview[9];
  int32_t ntaps, splits, aC, bC, W, H, N, n_rows, ld_k, k0;
  float* dw;
};

__global__ void __launch_bounds__(256) wgrad_f32_kernel(const __grid_constant__ WgradF32Params p) {
  __shared__ float xs[32][33], ds[32][33];
  const int tap = blockIdx.z % p.ntaps, split = blockIdx.z / p.ntaps;
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  const int64_t P = static_cast<int64_t>(p.N) * p.H * p.W;
  int64_t per = (P + p.splits - 1) / p.splits;
  per = (per + 31) / 32 * 32;
  const int64_t p_begin = split * per, p_end = p_begin + per < P ? p_begin + per : P;
  const float* A = p.a[p.t_aview[tap]];
  const float* B = p.b[p.t_bview[tap]];
  const int av = p.t_aview[tap], bv = p.t_bview[tap], dh = p.t_adh[tap], dw_ = p.t_adw[tap];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int64_t p0 = p_begin; p0 < p_end; p0 += 32) {
    for (int e = threadIdx.x; e < 1024; e += 256) {
      const int px = e >> 5, c = e & 31;
      const int64_t pp = p0 + px;
      float xv = 0.f, dv = 0.f;
      if (pp < p_end) {
        const int w = static_cast<int>(pp % p.W);
        const int h = static_cast<int>((pp / p.W) % p.H);
        const int64_t n = pp / (static_cast<int64_t>(p.W) * p.H);
        const int ha = h + dh, wa = w + dw_;
        if (ha >= 0 && ha < p.H && wa >= 0 && wa < p.W && ci0 + c < p.aC)
          xv = A[n * p.a_sN[av] + ha * p.a_sH[av] + wa * p.a_sW[av] + ci0 + c];
        if (co0 + c < p.bC) dv = B[n * p.b_sN[bv] + h * p.b_sH[bv] + w * p.b_sW[bv] + co0 + c];
      }
      xs[px][c] = xv;
      ds[px][c] = dv;
    }
    __syncthreads();
#pragma unroll 8
    for (int q = 0; q < 32; ++q) {
      const float x0 = xs[q][2 * tx], x1 = xs[q][2 * tx + 1], d0 = ds[q][2 * ty], d1 = ds[q][2 * ty + 1];
      acc[0][0] = fmaf(x0, d0, acc[0][0]); acc[0][1] = fmaf(x0, d1, acc[0][1]);
      acc[1][0] = fmaf(x1, d0, acc[1][0]); acc[1][1] = fmaf(x1, d1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int ci = ci0 + 2 * tx + i, co = co0 + 2 * ty + j;
      if (ci < p.aC && co < p.bC)
        atomicAdd(p.dw + (static_cast<int64_t>(tap) * p.n_rows + co) * p.ld_k + p.k0 + ci, acc[i][j]);
    }
}

struct WgradF32Op : public OpBase {
  WgradF32Params p;
  dim3 grid;
  double flops;
  int launch(cudaStream_t st) override {
    wgrad_f32_kernel<<<grid, 256, 0, st>>>(p);
    return check_launch("wgrad_f32_kernel");
  }
  double op_flops() const override { return flops; }
};

static int wgrad_f32_create(const mtbc_wgrad_desc* d, OpBase** out) {
  if (d->a_nviews < 1 || d->a_nviews > 4 || d->b_nviews < 1 || d->b_nviews > 4 || d->ntaps < 1 || d->ntaps > 9)
    return set_error(MTBC_ERR_INVALID, "wgrad(fp32): bad view/tap counts");
  WgradF32Op* op = new WgradF32Op();
  WgradF32Params& p = op->p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < d->a_nviews; ++i) {
    p.a[i] = static_cast<const float*>(d->a_views[i].ptr);
    p.a_sW[i] = d->a_views[i].sW; p.a_sH[i] = d->a_views[i].sH; p.a_sN[i] = d->a_views[i].sN;
  }
  for (int i = 0; i < d->b_nviews; ++i) {
    p.b[i] = static_cast<const float*>(d->b_views[i].ptr);
    p.b_sW[i] = d->b_views[i].sW; p.b_sH[i] = d->b_views[i].sH; p.b_sN[i] = d->b_views[i].sN;
  }
  for (int t = 0; t < d->ntaps; ++t) {
    const mtbc_wgrad_tap& tp = d->taps[t];
    if (tp.a_view < 0 || tp.a_view >= d->a_nviews || tp.b_view < 0 || tp.b_view >= d->b_nviews) { delete op; return set_error(MTBC_ERR_INVALID, "wgrad(fp32): tap %d view out of range", t); }
    p.t_aview[t] = (int16_t)tp.a_view; p.t_adh[t] = (int16_t)tp.a_dh; p.t_adw[t] = (int16_t)tp.a_dw; p.t_bview[t] = (int16_t)tp.b_view;
  }
  p.ntaps = d->ntaps; p.aC = d->a_views[0].C; p.bC = d->b_views[0].C;
  p.W = d->W; p.H = d->H; p.N = d->N; p.n_rows = d->n_rows; p.ld_k = d->ld_k; p.k0 = d->k0; p.dw = d->dw_acc;
  const int tiles = ((p.aC + 31) / 32) * ((p.bC + 31) / 32) * d->ntaps;
  const int64_t P = static_cast<int64_t>(d->N) * d->H * d->W;
  int splits = d->splits > 0 ? d->splits : (148 * 4 + tiles - 1) / tiles;
  const int64_t maxs = (P + 255) / 256;
  if (splits > maxs) splits = static_cast<int>(maxs);
  if (splits < 1) splits = 1;
  if (static_cast<int64_t>(splits) * d->ntaps > 65535) splits = 65535 / d->ntaps;
  p.splits = splits;
  op->grid = dim3((p.aC + 31) / 32, (p.bC + 31) / 32, d->ntaps * splits);
  op->flops = 2.0 * double(P) * double(p.aC) * double(p.bC) * d->ntaps;
  *out = op;
  return 0;
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode() {
  static PFN_tmapEncodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(f);
  }
  return fn;
}
bool tensor_map_available() { return get_encode() != nullptr; }

// bf16 NHWC view -> 4D tensor map with box (kc, TW, TH, TN)
int encode_act(CUtensorMap* m, const mtbc_act_view& v, int kc, int TW, int TH, int TN, int fp32) {
  PFN_tmapEncodeTiled enc = get_encode();
  if (!enc) return set_error(MTBC_ERR_NO_DEVICE, "cuTensorMapEncodeTiled not available");
  if (v.C % 8 != 0 || (reinterpret_cast<uintptr_t>(v.ptr) & 15) != 0)
    return set_error(MTBC_ERR_INVALID, "activation view: C %% 8 != 0 or pointer not 16B aligned");
  const cuuint64_t es = fp32 ? 4 : 2;
  cuuint64_t gdim[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N};
  cuuint64_t gstr[3] = {(cuuint64_t)v.sW * es, (cuuint64_t)v.sH * es, (cuuint64_t)v.sN * es};
  cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TN};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                   const_cast<void*>(v.ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   kc * es == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(MTBC_ERR_CUDA, "cuTensorMapEncodeTiled(act) failed: %d", (int)r);
  return 0;
}
// Dense bf16 NHWC tensor seen as whole image-row segments: dims (W * C / 2 words, H, N), box (box_w * C / 2 words, box_h,
// 1), no swizzle -- one TMA row per image row of the box instead of one per pixel, for tiles that are read back with
// ordinary shared-memory loads (pixel pitch C * 2 bytes inside the tile).
int encode_rows(CUtensorMap* m, const void* ptr, int C, int W, int H, int N, int box_w, int box_h) {
  PFN_tmapEncodeTiled enc = get_encode();
  if (!enc) return set_error(MTBC_ERR_NO_DEVICE, "cuTensorMapEncodeTiled not available");
  if (C % 8 != 0 || (reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || box_w * C / 2 > 256 || box_h > 256)
    return set_error(MTBC_ERR_INVALID, "row view: C %% 8 != 0, pointer not 16B aligned or box too large");
  cuuint64_t gdim[3] = {(cuuint64_t)W * C / 2, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstr[2] = {(cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[3] = {(cuuint32_t)(box_w * C / 2), (cuuint32_t)box_h, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(MTBC_ERR_CUDA, "cuTensorMapEncodeTiled(rows) failed: %d", (int)r);
  return 0;
}
int encode_w(CUtensorMap* m, const void* w, int ktot, int nrows, int ntaps, int kc, int BN, int fp32) {
  PFN_tmapEncodeTiled enc = get_encode();
  if (!enc) return set_error(MTBC_ERR_NO_DEVICE, "cuTensorMapEncodeTiled not available");
  const cuuint64_t es = fp32 ? 4 : 2;
  cuuint64_t gdim[3] = {(cuuint64_t)ktot, (cuuint64_t)nrows, (cuuint64_t)ntaps};
  cuuint64_t gstr[2] = {(cuuint64_t)ktot * es, (cuuint64_t)ktot * es * nrows};
  cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)BN, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                   const_cast<void*>(w), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   kc * es == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(MTBC_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
  return 0;
}

// 128-pixel tile shape for an (N, H, W) row space
static void pick_tile(int W, int H, int N, int* TW, int* TH, int* TN) {
  int tw = W >= 16 ? 16 : W;
  int th = 128 / tw;
  if (th > H) th = H;
  int tn = 128 / (tw * th);
  *TW = tw; *TH = th; *TN = tn;
}
static bool pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }
static int tmem_cols_for(int n) { int c = 32; while (c < n) c <<= 1; return c; }
static int pick_bn(int ncols, int maxbn, int multiple) {
  int best = 0;
  for (int bn = multiple; bn <= maxbn && bn <= ncols; bn += multiple)
    if (ncols % bn == 0 && bn % 16 == 0) best = bn;
  return best;
}

struct ConvGemmOp : public OpBase {
  ConvGemmParams p;
  dim3 grid;
  int smem_bytes;
  int prec = 0;
  double flops;
  int launch(cudaStream_t st) override {
    if (prec == 0) launch_pdl(conv_gemm_kernel<0>, grid, dim3(320), smem_bytes, st, p);
    else if (prec == 1) launch_pdl(conv_gemm_kernel<1>, grid, dim3(320), smem_bytes, st, p);
    else launch_pdl(conv_gemm_kernel<3>, grid, dim3(448), smem_bytes, st, p);
    return check_launch("conv_gemm_kernel");
  }
  double op_flops() const override { return flops; }
};

struct WgradOp : public OpBase {
  WgradParams p;
  dim3 grid;
  int smem_bytes;
  double flops;
  int launch(cudaStream_t st) override {
    launch_pdl(wgrad_kernel, grid, dim3(128), smem_bytes, st, p);
    return check_launch("wgrad_kernel");
  }
  double op_flops() const override { return flops; }
};

int conv_gemm_create(const mtbc_conv_gemm_desc* d, OpBase** out) {
  if (!d || !out) return set_error(MTBC_ERR_INVALID, "null argument");
  if (d->dtype != 0 && d->dtype != 1 && d->dtype != 3) return set_error(MTBC_ERR_INVALID, "conv_gemm: dtype %d", d->dtype);
  const int fp32 = d->dtype != 0, es = fp32 ? 4 : 2, x3 = d->dtype == 3;
  if (x3 && !d->wpack_lo) return set_error(MTBC_ERR_INVALID, "conv_gemm: 3xTF32 needs wpack_lo");
  if (!fp32) {
    int rc = conv_halo_try_create(d, out);  // halo-tile kernel for 3x3 convs on large planes (conv_halo.cu)
    if (rc <= 0) return rc;                  // 0 = created, < 0 = error, > 0 = not eligible -> generic kernel below
  }
  if (d->bwd_y != nullptr)
    return set_error(MTBC_ERR_INVALID, "conv_gemm: fused InstanceNorm backward statistics need a halo-eligible 3x3 shape");
  if (d->nouts > 0) return set_error(MTBC_ERR_INVALID, "conv_gemm: routed outputs need a halo-eligible 3x3 shape");
  if (d->stat_fold != 0) return set_error(MTBC_ERR_INVALID, "conv_gemm: the pixel-pair view needs a halo-eligible 3x3 shape (bf16)");
  if (d->nviews < 1 || d->nviews > MTBC_MAX_VIEWS || d->nseg < 1 || d->nseg > MTBC_MAX_SEGS)
    return set_error(MTBC_ERR_INVALID, "conv_gemm: bad nviews/nseg (%d, %d)", d->nviews, d->nseg);
  if (d->ncols % 32 != 0) return set_error(MTBC_ERR_INVALID, "conv_gemm: ncols %% 32 != 0");
  ConvGemmOp* op = new ConvGemmOp();
  ConvGemmParams& p = op->p;
  memset(&p, 0, sizeof(p));
  int TW, TH, TN;
  pick_tile(d->W, d->H, d->N, &TW, &TH, &TN);
  if (!pow2(d->W) && (d->W % TW)) { delete op; return set_error(MTBC_ERR_INVALID, "conv_gemm: W=%d not tileable", d->W); }
  if (d->W % TW || d->H % TH) { delete op; return set_error(MTBC_ERR_INVALID, "conv_gemm: H,W (%d,%d) not multiples of the tile (%d,%d)", d->H, d->W, TH, TW); }
  p.TW = TW; p.TH = TH; p.TN = TN;
  p.tiles_w = d->W / TW; p.tiles_h = d->H / TH;
  const int tiles_n = (d->N + TN - 1) / TN;
  p.W = d->W; p.H = d->H; p.N = d->N;
  // Transposed-conv forward (pixel-shuffle epilogue) is bound by its epilogue and by HBM writes, not by MMAs: N tiles of
  // 128 leave TMEM and shared memory for two CTAs per SM, i.e. 16 epilogue warps instead of 8 (MTBC_CONVT_BN=256 restores
  // the single 256-wide tile).
  const char* bn_env = getenv("MTBC_CONVT_BN");
  const int bn_cap = ((d->epi_mode == 1 && !(bn_env && atoi(bn_env) == 256)) || x3) ? 128 : 256;
  const int BN = pick_bn(d->ncols, bn_cap, 16);
  if (BN == 0) { delete op; return set_error(MTBC_ERR_INVALID, "conv_gemm: no N tile for ncols=%d", d->ncols); }
  p.BN = BN;
  p.tmem_cols = tmem_cols_for(2 * BN);
  int kcmax = 32;
  int kc_of_view[MTBC_MAX_VIEWS];
  for (int i = 0; i < d->nviews; ++i) {
    const mtbc_act_view& v = d->views[i];
    if (v.W != d->W || v.H != d->H || v.N != d->N) { delete op; return set_error(MTBC_ERR_INVALID, "conv_gemm: view %d geometry mismatch", i); }
    kc_of_view[i] = (!fp32 && ((v.C + 31) & ~31) % 64 == 0) ? 64 : 32;   // 128-byte rows either way
    if (kc_of_view[i] > kcmax) kcmax = kc_of_view[i];
    int rc = encode_act(&p.amap[i], v, kc_of_view[i], TW, TH, TN, fp32);
    if (rc) { delete op; return rc; }
  }
  bool use32 = false, use64 = false;
  double ksum = 0;
  for (int s = 0; s < d->nseg; ++s) {
    const mtbc_gemm_seg& g = d->seg[s];
    if (g.view < 0 || g.view >= d->nviews || g.wtap < 0 || g.wtap >= d->w_ntaps) { delete op; return set_error(MTBC_ERR_INVALID, "conv_gemm: segment %d out of range", s); }
    const int kc = kc_of_view[g.view];
    const int C = (d->views[g.view].C + 31) & ~31;   // K extent per source, padded (TMA zero-fills missing channels)
    if (g.wk0 % 32 != 0 || g.wk0 + C > d->w_ktot) { delete op; return set_error(MTBC_ERR_INVALID, "conv_gemm: segment %d weight columns out of range", s); }
    if (kc == 64 && g.wk0 % 64 != 0) { delete op; return set_error(MTBC_ERR_INVALID, "conv_gemm: 64-wide source at a non-64-aligned K offset"); }
    p.seg[s].view = (int16_t)g.view; p.seg[s].dh = (int16_t)g.dh; p.seg[s].dw = (int16_t)g.dw;
    p.seg[s].wtap = (int16_t)g.wtap; p.seg[s].wk0 = g.wk0; p.seg[s].kc = (int16_t)kc; p.seg[s].nchunk = (int16_t)(C / kc);
    (kc == 64 ? use64 : use32) = true;
    ksum += C;
  }
  p.nseg = d->nseg;
  if (use32) { int rc = encode_w(&p.wmap[0], d->wpack, d->w_ktot, d->ncols, d->w_ntaps, 32, BN, fp32); if (rc) { delete op; return rc; } }
  if (use64) { int rc = encode_w(&p.wmap[1], d->wpack, d->w_ktot, d->ncols, d->w_ntaps, 64, BN, 0); if (rc) { delete op; return rc; } }
  if (x3) { int rc = encode_w(&p.wmap[1], d->wpack_lo, d->w_ktot, d->ncols, d->w_ntaps, 32, BN, 1); if (rc) { delete op; return rc; } }
  op->prec = d->dtype;
  p.a_stage_bytes = 128 * kcmax * es;
  p.stage_bytes = (x3 ? 2 : 1) * (128 + BN) * kcmax * es;
  p.stage_bytes = (p.stage_bytes + 1023) & ~1023;
  // small stages and <= 256 TMEM columns: two persistent CTAs per SM (they hide each other's pipeline bubbles)
  const bool two = (p.stage_bytes * 3 <= 100 * 1024) && (2 * p.tmem_cols <= 512);
  const int budget = two ? 100 * 1024 : 200 * 1024;
  int stages = budget / p.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) stages = 2;
  p.stages = stages;
  op->smem_bytes = stages * p.stage_bytes + 1024;
  p.epi_mode = d->epi_mode; p.out_C = d->out_C; p.up_k = d->up_k > 0 ? d->up_k : 1; p.up_cp = d->up_cp > 0 ? d->up_cp : d->ncols;
  p.accumulate = d->accumulate; p.stat_C = d->stat_C;
  p.out = d->out;
  p.bias = d->bias; p.stat_sum = d->stat_sum; p.stat_sq = d->stat_sq;
  p.bias_len = d->epi_mode == 1 ? p.up_cp : d->ncols;
  if (d->bias && p.bias_len > kMaxBias) { delete op; return set_error(MTBC_ERR_INVALID, "conv_gemm: bias longer than %d", kMaxBias); }
  if (d->stat_sum && TN != 1) { delete op; return set_error(MTBC_ERR_INVALID, "conv_gemm: fused statistics need >= 128 pixels per sample (TN=%d)", TN); }
  if (d->epi_mode == 1 && (p.up_cp % 16 != 0)) { delete op; return set_error(MTBC_ERR_INVALID, "conv_gemm: up_cp %% 16"); }
  p.n_mtiles = p.tiles_w * p.tiles_h * tiles_n;
  p.n_ntiles = d->ncols / BN;
  {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    int gx = sms * (two ? 2 : 1);
    if (gx > p.n_mtiles * p.n_ntiles) gx = p.n_mtiles * p.n_ntiles;
    op->grid = dim3(gx, 1, 1);
  }
  op->flops = 2.0 * double(d->N) * d->H * d->W * double(d->ncols) * ksum;
  cudaError_t e = d->dtype == 0 ? cudaFuncSetAttribute(conv_gemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)
                  : d->dtype == 1 ? cudaFuncSetAttribute(conv_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)
                                  : cudaFuncSetAttribute(conv_gemm_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (e != cudaSuccess) { delete op; return set_error(MTBC_ERR_CUDA, "cudaFuncSetAttribute(conv_gemm): %s", cudaGetErrorString(e)); }
  *out = op;
  return 0;
}

int wgrad_create(const mtbc_wgrad_desc* d, OpBase** out) {
  if (!d || !out) return set_error(MTBC_ERR_INVALID, "null argument");
  if (d->dtype == 1) return wgrad_f32_create(d, out);
  if (d->dtype != 0) return set_error(MTBC_ERR_INVALID, "wgrad: dtype %d", d->dtype);
  {
    int rc = wgrad_halo_try_create(d, out);
    if (rc <= 0) return rc;
  }
  if (d->a_nviews < 1 || d->a_nviews > 4 || d->b_nviews < 1 || d->b_nviews > 4 || d->ntaps < 1 || d->ntaps > 9)
    return set_error(MTBC_ERR_INVALID, "wgrad: bad view/tap counts");
  WgradOp* op = new WgradOp();
  WgradParams& p = op->p;
  memset(&p, 0, sizeof(p));
  int TW, TH, TN;
  pick_tile(d->W, d->H, d->N, &TW, &TH, &TN);
  if (d->W % TW || d->H % TH) { delete op; return set_error(MTBC_ERR_INVALID, "wgrad: H,W not multiples of the tile"); }
  p.TW = TW; p.TH = TH; p.TN = TN; p.tiles_w = d->W / TW; p.tiles_h = d->H / TH;
  const int tiles_n = (d->N + TN - 1) / TN;
  p.n_ptiles = p.tiles_w * p.tiles_h * tiles_n;
  const int aC = d->a_views[0].C, bC = d->b_views[0].C;
  for (int i = 0; i < d->a_nviews; ++i) if (d->a_views[i].C != aC) { delete op; return set_error(MTBC_ERR_INVALID, "wgrad: A views differ in C"); }
  for (int i = 0; i < d->b_nviews; ++i) if (d->b_views[i].C != bC) { delete op; return set_error(MTBC_ERR_INVALID, "wgrad: B views differ in C"); }
  if (aC % 8 != 0 || bC % 8 != 0) { delete op; return set_error(MTBC_ERR_INVALID, "wgrad: C %% 8 != 0"); }
  const int aCp = (aC + 31) & ~31, bCp = (bC + 31) & ~31;   // GEMM extents of dense tensors
  p.a_C = aC;
  p.a_kc = (aCp % 64 == 0) ? 64 : 32;
  p.b_kc = (bCp % 64 == 0) ? 64 : 32;
  p.a_boxes = 128 / p.a_kc;
  const int BN = pick_bn(bCp, 128, p.b_kc);
  if (BN == 0) { delete op; return set_error(MTBC_ERR_INVALID, "wgrad: no N tile for C=%d", bC); }
  p.BN = BN; p.n_tiles = bCp / BN; p.b_boxes = BN / p.b_kc;
  int T = 1;
  for (int t = 1; t <= d->ntaps; ++t) if (d->ntaps % t == 0 && t * BN <= 512) T = t;
  p.T = T; p.ntaps = d->ntaps;
  p.tmem_cols = tmem_cols_for(T * BN);
  for (int i = 0; i < d->a_nviews; ++i) { int rc = encode_act(&p.amap[i], d->a_views[i], p.a_kc, TW, TH, TN, 0); if (rc) { delete op; return rc; } }
  for (int i = 0; i < d->b_nviews; ++i) { int rc = encode_act(&p.bmap[i], d->b_views[i], p.b_kc, TW, TH, TN, 0); if (rc) { delete op; return rc; } }
  for (int t = 0; t < d->ntaps; ++t) {
    const mtbc_wgrad_tap& tp = d->taps[t];
    if (tp.a_view < 0 || tp.a_view >= d->a_nviews || tp.b_view < 0 || tp.b_view >= d->b_nviews) { delete op; return set_error(MTBC_ERR_INVALID, "wgrad: tap %d view out of range", t); }
    p.t_aview[t] = (int16_t)tp.a_view; p.t_adh[t] = (int16_t)tp.a_dh; p.t_adw[t] = (int16_t)tp.a_dw; p.t_bview[t] = (int16_t)tp.b_view;
  }
  p.a_box_bytes = 128 * p.a_kc * 2; p.b_box_bytes = 128 * p.b_kc * 2;
  {
    const int a_nbox_max = std::min(p.a_boxes, (aC + p.a_kc - 1) / p.a_kc);   // boxes of the fullest M block
    p.a_stage_bytes = a_nbox_max * p.a_box_bytes;
  }
  p.b_stage_bytes = p.b_boxes * p.b_box_bytes;   // BN * 256 B
  // Ring depths: one CTA per SM, shared memory filled with whichever operand is (re)loaded more often per pixel tile
  // (transposed conv: one x tile, k*k dy tiles; 3x3 conv on small planes: nine shifted x tiles, one dy tile).
  {
    int a_loads = 0, b_loads = 0;
    for (int t = 0; t < T; ++t) {
      if (t == 0 || p.t_bview[t] != p.t_bview[t - 1]) ++b_loads;
      if (t == 0 || p.t_aview[t] != p.t_aview[t - 1] || p.t_adh[t] != p.t_adh[t - 1] || p.t_adw[t] != p.t_adw[t - 1]) ++a_loads;
    }
    const char* env = getenv("MTBC_WGRAD_SMEM_KB");
    const int budget = (env ? atoi(env) : 200) * 1024;
    int as = 2, bs = 2;
    for (;;) {
      const bool a_first = a_loads * bs >= b_loads * as;   // loads per stage: deepen the busier ring first
      const bool can_a = as < kWgMaxStages && (as + 1) * p.a_stage_bytes + bs * p.b_stage_bytes + 1024 <= budget;
      const bool can_b = bs < kWgMaxStages && as * p.a_stage_bytes + (bs + 1) * p.b_stage_bytes + 1024 <= budget;
      if (a_first && can_a) ++as;
      else if (!a_first && can_b) ++bs;
      else if (can_a && as < 2 * a_loads + 1) ++as;
      else if (can_b && bs < 2 * b_loads + 1) ++bs;
      else break;
    }
    p.a_stages = as; p.b_stages = bs;
  }
  op->smem_bytes = p.a_stages * p.a_stage_bytes + p.b_stages * p.b_stage_bytes + 1024;
  p.n_rows = d->n_rows; p.ld_k = d->ld_k; p.k0 = d->k0; p.dw_acc = d->dw_acc;
  const int mblocks = (aC + 127) / 128;
  const int base_ctas = mblocks * (d->ntaps / T) * p.n_tiles;
  int splits = d->splits;
  if (splits <= 0) {
    const int per_sm = op->smem_bytes <= 110 * 1024 ? 2 : 1;
    splits = (per_sm * 148 + base_ctas - 1) / base_ctas;
    // keep at least 4 pixel tiles per split so the TMEM/setup cost amortises
    int maxs = (p.n_ptiles + 3) / 4; if (maxs < 1) maxs = 1;
    if (splits > maxs) splits = maxs;
  }
  if (splits > p.n_ptiles) splits = p.n_ptiles;
  if (splits < 1) splits = 1;
  p.splits = splits;
  op->grid = dim3(splits, mblocks, (d->ntaps / T) * p.n_tiles);
  op->flops = 2.0 * double(d->N) * d->H * d->W * double(aC) * double(bC) * d->ntaps;
  cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (e != cudaSuccess) { delete op; return set_error(MTBC_ERR_CUDA, "cudaFuncSetAttribute(wgrad): %s", cudaGetErrorString(e)); }
  *out = op;
  return 0;
}

}  // namespace mtbc
