// Fused backward of ConvTranspose2d(kernel = stride = 2) for sm_100a: data gradient, weight gradient and bias gradient
// in ONE pass over the output gradient.
//
// The three launches it replaces (conv_gemm_kernel data gradient, wgrad_kernel, channel_sum) each stream the same
// (N, 2H, 2W, Cout) gradient from HBM: 700 MB for the 48 -> 48 @128^2 layers of U-Net++ at B = 32 (MONAI UpCat,
// MTUNetPlusPlus.py:107-118; MTnnUNet.py:96-100), 166 us together.  Here a persistent CTA loads, per tile of 128 input
// pixels, the x tile and the four sub-lattice tiles dy_q (q = 2 i + j: dy[n, 2h+i, 2w+j, :]) ONCE (TMA, 128-byte
// swizzled rows = pixels) and the tensor core reads every dy_q tile twice out of shared memory:
//   data gradient    D'[pixel][ci]   = sum_q dy_q[pixel][co] * wd[q][ci][co]     (dy_q as the K-major M x K operand)
//   weight gradient  D_q[ci][co]    += sum_pixel x[pixel][ci] * dy_q[pixel][co]  (x, dy_q as MN-major operands, K = pixels)
// The x tile has channels to spare (Cin < 64: TMA zero-fills the rest of the 64-channel box): one warp writes 1.0 into
// channel `Cin` of every landed tile, so row Cin of D_q is sum_pixel dy_q -- the bias gradient costs no instruction.
// D_q (4 x 64 TMEM columns) stays resident for the CTA's whole tile range and is added to the pack-layout accumulator
// once; D' is double buffered and drained per tile by four epilogue warps (bf16 store, or red.add for accumulation).
// Shared memory is a ring of 16 KB slots, five per tile in the order dy_0 .. dy_3, x -- the order in which the MMA lane
// lets go of them (dy_q after the weight-gradient MMAs of tap q, x after tap 3) -- so the producer refills a slot as soon
// as its last reader has been issued and ~11 loads stay in flight (two whole 80 KB stages left the loads of the next
// tile waiting for the last MMA of the previous one: 100 us per 48 -> 48 @128^2 layer at B = 32).
#include "ptx.cuh"
#include "internal.h"
#include <string.h>

namespace mtbc {

struct ConvTBwdParams {
  CUtensorMap xmap;       // x (N, H, W, Cin): box (64, TW, TH, 1)
  CUtensorMap dymap[4];   // sub-lattices of dy: box (64, TW, TH, 1)
  CUtensorMap wmap;       // wd [4][rows][ld]: box (64, NW, 1)
  int32_t W, H, N, TW, TH, tiles_w, tiles_h, n_tiles;
  int32_t ci, co, NW;     // true channel counts; NW = data-gradient GEMM columns (Cin rounded up to 16)
  int32_t kd;             // K steps of the data gradient: ceil(Cout / 16)
  int32_t slots, w_tap_bytes, w_bytes;   // ring of 16 KB slots, then the resident data-gradient weights
  int32_t n_rows, ld_k;
  float* dw_acc;
  float* dbias;
  __nv_bfloat16* dx;
  int32_t dx_C, accumulate;
};

constexpr int kCtBox = 128 * 128;   // one 128-pixel x 64-channel bf16 tile
constexpr int kCtMaxSlots = 12;
constexpr int kCtDCol = 256;        // first TMEM column of the data-gradient accumulators (after 4 x 64 of D_q)

__global__ void __launch_bounds__(256, 1) convT_bwd_kernel(const __grid_constant__ ConvTBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_full[kCtMaxSlots], s_fixed[kCtMaxSlots], s_empty[kCtMaxSlots];
  __shared__ uint64_t s_dfull[2], s_dempty[2], s_wfull, s_accum;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_s = smem;                            // slot ring
  uint8_t* smem_w = smem + p.slots * kCtBox;         // resident data-gradient weights: [4 taps][NW rows x 128 B]
  // (the weights come last: the M = 128 weight-gradient MMA reads a second, meaningless 64-channel box behind the x slot,
  //  which must still be shared memory when x sits in the last slot)
  const int S = p.slots;
  const int t_begin = static_cast<int>(static_cast<int64_t>(p.n_tiles) * blockIdx.x / gridDim.x);
  const int t_end = static_cast<int>(static_cast<int64_t>(p.n_tiles) * (blockIdx.x + 1) / gridDim.x);
  const int NW = p.NW;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_fixed[s], 1); mbar_init(&s_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&s_dfull[b], 1); mbar_init(&s_dempty[b], 4); }
    mbar_init(&s_wfull, 1);
    mbar_init(&s_accum, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&s_tmem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  pdl_trigger();
  pdl_wait();

  if (t_begin < t_end) {
    if (warp == 0) {
      if (elect_one()) {
        // ---------------------------------------------------------------- producer
        mbar_arrive_expect_tx(&s_wfull, static_cast<uint32_t>(4 * NW * 128));
        for (int q = 0; q < 4; ++q) tma_load_3d(smem_w + q * p.w_tap_bytes, &p.wmap, &s_wfull, 0, 0, q);
        int slot = 0;
        uint32_t phase = 0;
        for (int t = t_begin; t < t_end; ++t) {
          const int w0 = (t % p.tiles_w) * p.TW;
          const int h0 = ((t / p.tiles_w) % p.tiles_h) * p.TH;
          const int n = t / (p.tiles_w * p.tiles_h);
          for (int j = 0; j < 5; ++j) {   // dy_0 .. dy_3, x
            mbar_wait(&s_empty[slot], phase ^ 1u);
            mbar_arrive_expect_tx(&s_full[slot], static_cast<uint32_t>(kCtBox));
            tma_load_4d(smem_s + slot * kCtBox, j < 4 ? &p.dymap[j] : &p.xmap, &s_full[slot], 0, w0, h0, n);
            if (++slot == S) { slot = 0; phase ^= 1u; }
          }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        // ---------------------------------------------------------------- MMA lane
        const uint32_t idesc_w = umma_idesc_bf16(128, 64, 1, 1);   // D_q[ci][co]: both operands MN-major, K = pixels
        const uint32_t idesc_d = umma_idesc_bf16(128, NW, 0, 0);   // D'[pixel][ci]: both operands K-major, K = co
        const uint32_t hi = umma_desc_hi(1024u, 2u);               // 8 rows of 128 bytes, 128-byte swizzle
        mbar_wait(&s_wfull, 0);
        int slot = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int t = t_begin; t < t_end; ++t, ++it) {
          const int buf = it & 1;
          const uint32_t acc_phase = (it >> 1) & 1;
          mbar_wait(&s_dempty[buf], acc_phase ^ 1u);  // epilogue has drained this data-gradient accumulator
          tc_fence_after();
          const uint32_t d_dx = tmem_base + kCtDCol + static_cast<uint32_t>(buf * NW);
          // data gradient first: its epilogue (global stores) overlaps the weight-gradient MMAs of the same tile
          int sq[4];
          for (int q = 0; q < 4; ++q) {
            sq[q] = slot;
            mbar_wait(&s_fixed[slot], phase);   // landed (s_full) and passed on by the ones-channel warp
            tc_fence_after();
            const uint32_t a_lo = umma_desc_lo(smem_u32(smem_s + slot * kCtBox), 16);
            const uint32_t b_lo = umma_desc_lo(smem_u32(smem_w + q * p.w_tap_bytes), 16);
            for (int k = 0; k < p.kd; ++k)
              umma_bf16_lohi(d_dx, a_lo + 2 * k, hi, b_lo + 2 * k, hi, idesc_d, (q | k) ? 1u : 0u);
            if (++slot == S) { slot = 0; phase ^= 1u; }
          }
          umma_commit(&s_dfull[buf]);
          const int sx = slot;
          mbar_wait(&s_fixed[sx], phase);   // x tile landed and its ones channel is written
          tc_fence_after();
          if (++slot == S) { slot = 0; phase ^= 1u; }
          const uint32_t x_lo = umma_desc_lo(smem_u32(smem_s + sx * kCtBox), kCtBox);
          for (int q = 0; q < 4; ++q) {
            const uint32_t b_lo = umma_desc_lo(smem_u32(smem_s + sq[q] * kCtBox), kCtBox);
            const uint32_t d_w = tmem_base + static_cast<uint32_t>(q * 64);
#pragma unroll
            for (int k = 0; k < 8; ++k)   // 128 pixels = 8 K steps of 16 rows (2048 bytes = 128 descriptor units)
              umma_bf16_lohi(d_w, x_lo + k * 128, hi, b_lo + k * 128, hi, idesc_w, (it > 0 || k > 0) ? 1u : 0u);
            umma_commit(&s_empty[sq[q]]);   // every MMA that reads dy_q has been issued
          }
          umma_commit(&s_empty[sx]);
        }
        umma_commit(&s_accum);
      }
    } else if (warp == 2) {
      // ------------------------------------------------------------------ ones channel of every landed x tile
      // Every slot passes through this warp (s_full -> s_fixed), so that s_fixed completes once per lap of the ring like
      // the other barriers and one phase bit serves them all; only the x slots (every fifth) are written to.
      const int chunk = p.ci >> 3, sub = p.ci & 7;
      int slot = 0;
      uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        for (int j = 0; j < 5; ++j) {
          mbar_wait(&s_full[slot], phase);
          if (j == 4) {
            uint8_t* xb = smem_s + slot * kCtBox;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = lane + 32 * i;
              *reinterpret_cast<uint16_t*>(xb + r * 128 + ((chunk ^ (r & 7)) << 4) + sub * 2) = 0x3F80;   // bf16 1.0
            }
            fence_proxy_async_smem();
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_fixed[slot]);
          if (++slot == S) { slot = 0; phase ^= 1u; }
        }
      }
    } else if (warp >= 4) {
      // ------------------------------------------------------------------ epilogue warps (TMEM lane quarter = warp % 4)
      const int q4 = warp & 3;
      const int m = q4 * 32 + lane;
      const int th = m / p.TW, tw = m - th * p.TW;
      const uint32_t tq = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);
      const bool wide = (p.dx_C % 16) == 0;
      int it = 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const int buf = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const int w0 = (t % p.tiles_w) * p.TW;
        const int h0 = ((t / p.tiles_w) % p.tiles_h) * p.TH;
        const int n = t / (p.tiles_w * p.tiles_h);
        __nv_bfloat16* dst = p.dx + ((static_cast<int64_t>(n) * p.H + h0 + th) * p.W + w0 + tw) * p.dx_C;
        mbar_wait(&s_dfull[buf], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tq + kCtDCol + static_cast<uint32_t>(buf * NW);
        for (int c = 0; c < NW; c += 32) {
          const bool two = c + 16 < NW;   // warp uniform
          uint32_t r[2][16];
          tmem_ld16_nowait(taddr + c, r[0]);
          if (two) tmem_ld16_nowait(taddr + c + 16, r[1]);
          tmem_wait_ld();
          if (c + 32 >= NW) {   // last columns are in registers: hand the accumulator back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_dempty[buf]);
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (u == 1 && !two) break;
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[u][i]);
            const int left = p.dx_C - (c + 16 * u);
            emit_bf16x16_n(dst + c + 16 * u, v, p.accumulate != 0, left >= 16 ? 16 : (left >= 8 ? 8 : 0), wide);
          }
        }
      }
      // weight and bias gradients of this CTA's tile range
      mbar_wait(&s_accum, 0);
      tc_fence_after();
      if (q4 * 32 <= p.ci) {   // warp uniform: tcgen05.ld is a whole-warp instruction
        for (int q = 0; q < 4; ++q)
          for (int c = 0; c < 64; c += 16) {
            if (c >= p.co) break;
            float v[16];
            tmem_ld16(tq + q * 64 + c, v);
            if (m > p.ci) continue;
            if (m < p.ci) {
              float* o = p.dw_acc + (static_cast<int64_t>(q) * p.n_rows + c) * p.ld_k + m;
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (c + i < p.co) atomicAdd(o + static_cast<int64_t>(i) * p.ld_k, v[i]);
            } else if (p.dbias != nullptr) {   // m == ci: the ones channel
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (c + i < p.co) atomicAdd(p.dbias + c + i, v[i]);
            }
          }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

struct ConvTBwdOp : public OpBase {
  ConvTBwdParams p;
  dim3 grid;
  int smem_bytes;
  double flops;
  int launch(cudaStream_t st) override {
    launch_pdl(convT_bwd_kernel, grid, dim3(256), smem_bytes, st, p);
    return check_launch("convT_bwd_kernel");
  }
  double op_flops() const override { return flops; }
};

int convT_bwd_create(const mtbc_convT_bwd_desc* d, OpBase** out) {
  if (!d || !out) return set_error(MTBC_ERR_INVALID, "null argument");
  const mtbc_act_view& x = d->x;
  const int ci = x.C, co = d->Cout;
  // shapes served: Cin < 64 (a spare channel for the ones column), Cout <= 64, planes tiled by 16 x 8 pixel boxes
  if (ci % 8 != 0 || ci <= 0 || ci >= 64 || co <= 0 || co > 64 || co % 8 != 0)
    return set_error(MTBC_ERR_INVALID, "convT_bwd: needs Cin %% 8 == 0, Cin < 64, Cout <= 64 (got %d, %d)", ci, co);
  if (x.W % 16 != 0 || x.H % 8 != 0) return set_error(MTBC_ERR_INVALID, "convT_bwd: plane %dx%d not tiled by 16x8", x.H, x.W);
  if (!d->wd || !d->dw_acc || !d->dx || d->wd_ld < co || d->wd_ld % 8 != 0 || d->wd_rows < ((ci + 15) & ~15) || d->dx_C < ci ||
      d->dx_C % 8 != 0)
    return set_error(MTBC_ERR_INVALID, "convT_bwd: malformed operands");
  for (int q = 0; q < 4; ++q)
    if (d->dy[q].W != x.W || d->dy[q].H != x.H || d->dy[q].N != x.N || d->dy[q].C < co)
      return set_error(MTBC_ERR_INVALID, "convT_bwd: sub-lattice %d does not match the input plane", q);
  ConvTBwdOp* op = new ConvTBwdOp();
  ConvTBwdParams& p = op->p;
  memset(&p, 0, sizeof(p));
  p.W = x.W; p.H = x.H; p.N = x.N; p.TW = 16; p.TH = 8;
  p.tiles_w = x.W / 16; p.tiles_h = x.H / 8;
  p.n_tiles = p.tiles_w * p.tiles_h * x.N;
  p.ci = ci; p.co = co; p.NW = (ci + 15) & ~15; p.kd = (co + 15) / 16;
  p.w_tap_bytes = (p.NW * 128 + 1023) & ~1023;
  p.w_bytes = 4 * p.w_tap_bytes;
  p.slots = (212 * 1024 - 1024 - p.w_bytes) / kCtBox;
  if (p.slots > kCtMaxSlots) p.slots = kCtMaxSlots;
  if (p.slots < 6) { delete op; return set_error(MTBC_ERR_INVALID, "convT_bwd: no room for the slot ring"); }
  p.n_rows = d->n_rows; p.ld_k = d->ld_k;
  p.dw_acc = d->dw_acc; p.dbias = d->dbias;
  p.dx = reinterpret_cast<__nv_bfloat16*>(d->dx); p.dx_C = d->dx_C; p.accumulate = d->accumulate;
  if (d->n_rows < co || d->ld_k < ci) { delete op; return set_error(MTBC_ERR_INVALID, "convT_bwd: accumulator too small"); }
  int rc = encode_act(&p.xmap, x, 64, 16, 8, 1);
  for (int q = 0; q < 4 && !rc; ++q) rc = encode_act(&p.dymap[q], d->dy[q], 64, 16, 8, 1);
  if (!rc) rc = encode_w(&p.wmap, d->wd, d->wd_ld, d->wd_rows, 4, 64, p.NW);
  if (rc) { delete op; return rc; }
  op->smem_bytes = p.w_bytes + p.slots * kCtBox + 1024;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  op->grid = dim3(p.n_tiles < sms ? p.n_tiles : sms);
  op->flops = 2.0 * 2.0 * double(x.N) * x.H * x.W * 4.0 * 64.0 * 64.0;
  cudaError_t e = cudaFuncSetAttribute(convT_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, op->smem_bytes);
  if (e != cudaSuccess) { delete op; return set_error(MTBC_ERR_CUDA, "cudaFuncSetAttribute(convT_bwd): %s", cudaGetErrorString(e)); }
  *out = op;
  return 0;
}

}  // namespace mtbc
