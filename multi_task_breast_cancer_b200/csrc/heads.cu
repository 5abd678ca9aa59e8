// Mask heads (1x1 projection, algebraically composed deep-supervision heads) and classification heads
// (global-average-pool + 2 FC, flatten + 2 FC).  All CUDA-core: M = batch is far below a tensor tile and the 1x1
// projections are pure HBM streams.
#include "ptx.cuh"
#include "internal.h"
#include "act_io.cuh"

namespace mtbc {

// ------------------------------------------------------------------------------------------------ 1x1 head
template <typename T>
__global__ void __launch_bounds__(256) head1x1_fwd_kernel(const T* __restrict__ a, int64_t npix, int Cp,
                                                          int C, const float* __restrict__ w,
                                                          const float* __restrict__ b, float* __restrict__ logits) {
  pdl_trigger();   // programmatic dependent launch (ptx.cuh)
  pdl_wait();

  extern __shared__ float s_w[];  // [Cp]
  for (int i = threadIdx.x; i < Cp; i += 256) s_w[i] = i < C ? w[i] : 0.f;
  __syncthreads();
  const float bias = b ? b[0] : 0.f;
  const int cvec = Cp / 8;
  for (int64_t pix = blockIdx.x * 256ll + threadIdx.x; pix < npix; pix += gridDim.x * 256ll) {
    const T* src = a + pix * Cp;
    float acc = bias;
    for (int v = 0; v < cvec; ++v) {
      const V8 u = load8<T>(src + v * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc = fmaf(u.f[k], s_w[v * 8 + k], acc);
    }
    logits[pix] = acc;
  }
}

// dA[pix][c] (+)= dl[pix]*w[c];  dw[c] += sum_pix dl[pix]*a[pix][c];  db += sum_pix dl[pix]
// Four independent 16-byte vectors per thread and iteration (loads issued before use), block reduction through a
// parked-partials table instead of contended shared-memory atomics (see stream_pipe.cu).
template <typename T>
__global__ void __launch_bounds__(256) head1x1_bwd_kernel(const T* __restrict__ a,
                                                          const float* __restrict__ dl, int64_t npix, int Cp, int C,
                                                          const float* __restrict__ w, T* __restrict__ dA,
                                                          int accumulate, float* __restrict__ dw,
                                                          float* __restrict__ db) {
  pdl_trigger();   // programmatic dependent launch (ptx.cuh)
  pdl_wait();

  __shared__ float s_stage[9 * 256];   // [k][thread]: 8 channel partials + the bias partial
  const int cvec = Cp / 8;
  const int bd = blockDim.x;
  const int64_t total = npix * cvec;
  const int64_t start = static_cast<int64_t>(blockIdx.x) * bd + threadIdx.x, stride = static_cast<int64_t>(gridDim.x) * bd;
  const int v = static_cast<int>(start % cvec);
  float wv[8], acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, accb = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) wv[k] = (v * 8 + k) < C ? w[v * 8 + k] : 0.f;
  constexpr int U = 4;
  for (int64_t i0 = start; i0 < total; i0 += U * stride) {
    Raw8<T> u[U], q[U];
    float g[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int64_t i = i0 + j * stride;
      if (i < total) {
        u[j].load(a + i * 8);
        g[j] = dl[i / cvec];
        if (accumulate) q[j].load(dA + i * 8);
      }
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int64_t i = i0 + j * stride;
      if (i >= total) break;
      const V8 uv = u[j].unpack();
      V8 o;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc[k] = fmaf(g[j], uv.f[k], acc[k]);
        o.f[k] = g[j] * wv[k];
      }
      if (accumulate) {
        const V8 qv = q[j].unpack();
#pragma unroll
        for (int k = 0; k < 8; ++k) o.f[k] += qv.f[k];
      }
      store8<T>(dA + i * 8, o);
      if (v == 0) accb += g[j];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) s_stage[k * 256 + threadIdx.x] = acc[k];
  s_stage[8 * 256 + threadIdx.x] = accb;
  __syncthreads();
  // thread's channel group is (blockIdx.x * bd + threadIdx.x) % cvec = threadIdx.x % cvec (bd is a multiple of cvec)
  for (int ch = threadIdx.x; ch < C; ch += bd) {
    const float* row = s_stage + (ch & 7) * 256;
    float sum = 0.f;
    for (int r = ch >> 3; r < bd; r += cvec) sum += row[r];
    atomicAdd(dw + ch, sum);
  }
  if (threadIdx.x == 0) {
    float sum = 0.f;
    for (int r = 0; r < bd; r += cvec) sum += s_stage[8 * 256 + r];
    atomicAdd(db, sum);
  }
}

// ------------------------------------------------------------------------------------------------ composed DS head
__global__ void dshead_compose_kernel(const float* __restrict__ wt, const float* __restrict__ bt,
                                      const float* __restrict__ w1, const float* __restrict__ b1, int C, int kk,
                                      float* __restrict__ wc, float* __restrict__ bc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C * kk) {
    const int ci = i / kk, q = i % kk;
    // four independent chains, eight loads in flight: a single chain of C dependent load + FMA steps made this 37 us
    // for C = 128, k = 8 (latency, not bytes)
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const float* wp = wt + static_cast<int64_t>(ci) * C * kk + q;
    int co = 0;
#pragma unroll 2
    for (; co + 4 <= C; co += 4) {
      a0 = fmaf(wp[static_cast<int64_t>(co) * kk], w1[co], a0);
      a1 = fmaf(wp[static_cast<int64_t>(co + 1) * kk], w1[co + 1], a1);
      a2 = fmaf(wp[static_cast<int64_t>(co + 2) * kk], w1[co + 2], a2);
      a3 = fmaf(wp[static_cast<int64_t>(co + 3) * kk], w1[co + 3], a3);
    }
    for (; co < C; ++co) a0 = fmaf(wp[static_cast<int64_t>(co) * kk], w1[co], a0);
    wc[i] = (a0 + a1) + (a2 + a3);
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x >= blockDim.x - 32) {   // last warp of the last block: the composed bias
    const int lane = threadIdx.x & 31;       // (one thread walking C dependent loads was most of this launch's time)
    float a = 0.f;
    for (int co = lane; co < C; co += 32) a = fmaf(bt[co], w1[co], a);
    a = warp_sum(a);
    if (lane == 0) bc[0] = a + b1[0];
  }
}

// Composed deep-supervision head, round-2 kernels.  Tiles of 64 pixels per block step, activations / dl / the composed
// weights in shared memory (fp32), 4 threads per pixel.  Round 1 processed 256 / k^2 pixels per step (4 at k = 8) and
// funnelled every block's C * k^2 weight-gradient partials through fp32 atomics on the same 8 192 addresses: 1.19 ms for
// the k = 8 head of nnU-Net at B = 32 @256^2 against ~0.03 ms of CUDA-core work (tools/profile_plan.py nnunet).  Now
// every block keeps its weight-gradient partials in registers over all of its tiles and writes them ONCE to its own row
// of a partials buffer; mtbc_dshead_decompose adds the rows.
constexpr int kDsTP = 64;      // pixels per tile

__device__ __forceinline__ int ds_logit_index(int64_t pix, int q, int H, int W, int k, int64_t* out) {
  const int w = static_cast<int>(pix % W);
  const int h = static_cast<int>((pix / W) % H);
  const int64_t n = pix / (static_cast<int64_t>(W) * H);
  *out = (n * H * k + static_cast<int64_t>(h) * k + q / k) * (static_cast<int64_t>(W) * k) + static_cast<int64_t>(w) * k + q % k;
  return 0;
}

template <typename T>
__device__ __forceinline__ void ds_load_tile(const T* __restrict__ a, int64_t pix0, int64_t npix, int Cp, int C, int CA,
                                             float* __restrict__ s_a) {
  const int cvec = C >> 3;
  for (int i = threadIdx.x; i < kDsTP * cvec; i += 256) {
    const int px = i / cvec, v = i - px * cvec;
    const int64_t pix = pix0 + px;
    V8 x;
    if (pix < npix) x = load8<T>(a + pix * Cp + v * 8);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x.f[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s_a[px * CA + v * 8 + j] = x.f[j];
  }
}

// logits[n, k*h+i, k*w+j] = sum_ci a[n,h,w,ci]*wc[ci][q] + bc.  Thread (pixel p = tid / 4, part = tid % 4) owns the
// outputs q = part + 4 j: its reads of the weight row are conflict free and the four threads of a pixel together write
// whole k-float output rows.
template <typename T, int KK>
__global__ void __launch_bounds__(256) dshead_fwd_kernel(const T* __restrict__ a, int N, int H, int W, int Cp, int C,
                                                         int k, const float* __restrict__ wc,
                                                         const float* __restrict__ bc, float* __restrict__ logits) {
  extern __shared__ float sm[];
  constexpr int PQ = KK >= 4 ? KK / 4 : 1;
  const int CA = C + 1;
  float* s_wc = sm;             // [C][KK]
  float* s_a = sm + C * KK;     // [64][C + 1]
  for (int i = threadIdx.x; i < C * KK; i += 256) s_wc[i] = wc[i];
  const float bias = bc[0];
  const int64_t npix = static_cast<int64_t>(N) * H * W;
  const int64_t ntiles = (npix + kDsTP - 1) / kDsTP;
  const int p = threadIdx.x >> 2, part = threadIdx.x & 3;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    ds_load_tile<T>(a, tile * kDsTP, npix, Cp, C, CA, s_a);
    __syncthreads();
    float acc[PQ];
#pragma unroll
    for (int j = 0; j < PQ; ++j) acc[j] = bias;
    const float* ar = s_a + p * CA;
    for (int ci = 0; ci < C; ++ci) {
      const float av = ar[ci];
      const float* wr = s_wc + ci * KK + part;
#pragma unroll
      for (int j = 0; j < PQ; ++j) acc[j] = fmaf(av, wr[4 * j], acc[j]);
    }
    const int64_t pix = tile * kDsTP + p;
    if (pix < npix) {
#pragma unroll
      for (int j = 0; j < PQ; ++j) {
        const int q = part + 4 * j;
        if (q < KK) {
          int64_t o;
          ds_logit_index(pix, q, H, W, k, &o);
          logits[o] = acc[j];
        }
      }
    }
  }
}

// dA[pix][ci] (+)= sum_q dl[pix][q] * wc[ci][q];  partial weight gradient of this block: dwc_part[block][ci][q] =
// sum over the block's pixels of a[pix][ci] * dl[pix][q];  dbc_part[block] = sum of dl.
template <typename T, int KK>
__global__ void __launch_bounds__(256) dshead_bwd_kernel(const T* __restrict__ a, const float* __restrict__ dl, int N,
                                                         int H, int W, int Cp, int C, int k,
                                                         const float* __restrict__ wc, T* __restrict__ dA,
                                                         int accumulate, float* __restrict__ dwc_part,
                                                         float* __restrict__ dbc_part) {
  extern __shared__ float sm[];
  constexpr int KP = KK + 1, KW = KK + 4;
  const int CA = C + 1, cvec = C >> 3, cpvec = Cp >> 3;
  float* s_wc = sm;                        // [C][KK + 4]: the row of channel ci starts (ci / 8) % 4 words late
  float* s_a = s_wc + C * KW;              // [64][C + 1]
  float* s_dl = s_a + kDsTP * CA;          // [64][KK + 1]
  __shared__ float s_red[8];
  for (int i = threadIdx.x; i < C * KK; i += 256) {
    const int ci = i / KK, q = i - ci * KK;
    s_wc[ci * KW + q + ((ci >> 3) & 3)] = wc[i];
  }
  const int64_t npix = static_cast<int64_t>(N) * H * W;
  const int64_t ntiles = (npix + kDsTP - 1) / kDsTP;
  const int p = threadIdx.x >> 2, part = threadIdx.x & 3;
  const int nout = C * KK;
  const int q_w = threadIdx.x % KK;        // 256 % KK == 0: the thread's weight-gradient column is fixed
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  float accb = 0.f;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    ds_load_tile<T>(a, tile * kDsTP, npix, Cp, C, CA, s_a);
    for (int i = threadIdx.x; i < kDsTP * KK; i += 256) {
      const int px = i / KK, q = i - px * KK;
      const int64_t pix = tile * kDsTP + px;
      float g = 0.f;
      if (pix < npix) {
        int64_t o;
        ds_logit_index(pix, q, H, W, k, &o);
        g = dl[o];
      }
      s_dl[px * KP + q] = g;
      accb += g;
    }
    __syncthreads();
    // ---- data gradient: thread (pixel p, channel vectors v = part, part + 4, ...), 8 channels at a time; the skew
    // puts the four threads of a pixel on banks q + 0 / 1 / 2 / 3, threads of other pixels read the same words
    const int64_t pix = tile * kDsTP + p;
    const float* dr = s_dl + p * KP;
    for (int v = part; v < cpvec; v += 4) {
      V8 g;
#pragma unroll
      for (int j = 0; j < 8; ++j) g.f[j] = 0.f;
      if (v < cvec) {
        const float* wr = s_wc + v * 8 * KW + part;
        for (int q = 0; q < KK; ++q) {
          const float dv = dr[q];
#pragma unroll
          for (int j = 0; j < 8; ++j) g.f[j] = fmaf(dv, wr[j * KW + q], g.f[j]);
        }
      }
      if (pix < npix) {
        T* dst = dA + pix * Cp + v * 8;
        if (accumulate) {
          const V8 old = load8<T>(dst);
#pragma unroll
          for (int j = 0; j < 8; ++j) g.f[j] += old.f[j];
        }
        store8<T>(dst, g);
      }
    }
    // ---- weight gradient partials: output o = tid + 256 j  ->  (ci = o / KK, q = q_w)
    if constexpr (KK >= 64) {
      // k = 8 (C * 64 outputs, up to 32 per thread): pixel loop outside, the dl value is read once for all of the
      // thread's channels; ci = tid / KK + (256 / KK) j.  144 -> 117 us for the k = 8 head of nnU-Net at B = 32 @256^2.
      const float* ap = s_a + threadIdx.x / KK;
#pragma unroll 2
      for (int px = 0; px < kDsTP; ++px) {
        const float d = s_dl[px * KP + q_w];
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (threadIdx.x + 256 * j < nout) acc[j] = fmaf(ap[px * CA + (256 / KK) * j], d, acc[j]);
      }
    } else {
      // k = 2, 4 (few outputs per thread): output loop outside, so that the accumulators that do not exist cost
      // nothing (the pixel-outer form walks all 32 predicated slots per pixel: 108 -> 554 us for the k = 2 head)
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int o = threadIdx.x + 256 * j;
        if (o < nout) {
          const int ci = o / KK;
          float sacc = acc[j];
#pragma unroll 8
          for (int px = 0; px < kDsTP; ++px) sacc = fmaf(s_a[px * CA + ci], s_dl[px * KP + q_w], sacc);
          acc[j] = sacc;
        }
      }
    }
  }
  float* mine = dwc_part + static_cast<int64_t>(blockIdx.x) * nout;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int o = threadIdx.x + 256 * j;
    if (o < nout) mine[o] = acc[j];
  }
  accb = warp_sum(accb);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = accb;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    dbc_part[blockIdx.x] = t;
  }
}

// dwc_red[e] = sum over the blocks' partial rows; dbc_red = sum of the blocks' bias partials
__global__ void __launch_bounds__(256) dshead_reduce_kernel(const float* __restrict__ part, const float* __restrict__ bpart,
                                                            int nparts, int n, float* __restrict__ red,
                                                            float* __restrict__ bred) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e < n) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;   // (independent chains: the loads of a row are what takes time)
    int r = 0;
#pragma unroll 2
    for (; r + 4 <= nparts; r += 4) {
      s0 += part[static_cast<int64_t>(r) * n + e];
      s1 += part[static_cast<int64_t>(r + 1) * n + e];
      s2 += part[static_cast<int64_t>(r + 2) * n + e];
      s3 += part[static_cast<int64_t>(r + 3) * n + e];
    }
    for (; r < nparts; ++r) s0 += part[static_cast<int64_t>(r) * n + e];
    red[e] = (s0 + s1) + (s2 + s3);
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x >= 224) {   // last warp of the (partly idle) last block
    const int lane = threadIdx.x & 31;
    float b = 0.f;
    for (int r = lane; r < nparts; r += 32) b += bpart[r];
    b = warp_sum(b);
    if (lane == 0) bred[0] = b;
  }
}

// dwt[ci][co][q] += dwc[ci][q] * w1[co]
__global__ void dshead_dwt_kernel(const float* __restrict__ dwc, const float* __restrict__ w1, int C, int kk,
                                  float* __restrict__ dwt) {
  const int64_t total = static_cast<int64_t>(C) * C * kk;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int q = static_cast<int>(i % kk);
    const int co = static_cast<int>((i / kk) % C);
    const int ci = static_cast<int>(i / (static_cast<int64_t>(kk) * C));
    dwt[i] += dwc[ci * kk + q] * w1[co];
  }
}
// one block per output channel co of the transposed conv:
// dw1[co] += gb * bt[co] + sum_{ci,q} dwc[ci][q] * wt[ci][co][q];  dbt[co] += gb * w1[co];  db1 += gb
__global__ void __launch_bounds__(256) dshead_dw1_kernel(const float* __restrict__ dwc, const float* __restrict__ dbc,
                                                         const float* __restrict__ wt, const float* __restrict__ bt,
                                                         const float* __restrict__ w1, int C, int kk,
                                                         float* __restrict__ dbt, float* __restrict__ dw1,
                                                         float* __restrict__ db1) {
  __shared__ float s_red[8];
  const int co = blockIdx.x;
  float s = 0.f;
  for (int i = threadIdx.x; i < C * kk; i += 256) {
    const int ci = i / kk, q = i - ci * kk;
    s = fmaf(dwc[i], wt[(static_cast<int64_t>(ci) * C + co) * kk + q], s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    const float gb = dbc[0];
    dw1[co] += t + gb * bt[co];
    dbt[co] += gb * w1[co];
    if (co == 0) db1[0] += gb;
  }
}

// ------------------------------------------------------------------------------------------------ GAP + FC head
// gap[n][c] = mean over the plane.  grid (N, ceil(Cp / 64)): 8 channel vectors (16 B each) x 32 pixel lanes per block.
template <typename T>
__global__ void __launch_bounds__(256) gap_kernel(const T* __restrict__ a, int HW, int Cp, int F,
                                                  float* __restrict__ gap) {
  __shared__ float s_red[32][65];
  const int n = blockIdx.x, vec = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c0 = blockIdx.y * 64 + vec * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < Cp) {
    const T* src = a + static_cast<int64_t>(n) * HW * Cp + c0;
    for (int p = pl; p < HW; p += 32) {
      const V8 u = load8<T>(src + static_cast<int64_t>(p) * Cp);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += u.f[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) s_red[pl][vec * 8 + k] = acc[k];
  __syncthreads();
  if (threadIdx.x < 64) {
    const int c = blockIdx.y * 64 + threadIdx.x;
    float s = 0.f;
#pragma unroll 8
    for (int q = 0; q < 32; ++q) s += s_red[q][threadIdx.x];
    if (c < F) gap[static_cast<int64_t>(n) * F + c] = s / static_cast<float>(HW);
  }
}

// hidden[n][j] = relu(W1[j] . gap[n] + b1[j]).  grid (N, ceil(Hd / 32)): 8 warps x 4 rows per block, 128-bit loads with
// all of a row's loads in flight (one block per sample with a row loop per warp was a 180 us chain of dependent loads).
__global__ void __launch_bounds__(256) fc1_fwd_kernel(const float* __restrict__ gap, int F, const float* __restrict__ w1,
                                                      const float* __restrict__ b1, int Hd, float* __restrict__ hidden) {
  extern __shared__ __align__(16) float s_gap[];   // [F]
  const int n = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = threadIdx.x; c < F; c += 256) s_gap[c] = gap[static_cast<int64_t>(n) * F + c];
  __syncthreads();
  const int F4 = F >> 2;   // F % 4 == 0 is checked by the host
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int j = blockIdx.y * 32 + warp * 4 + r;
    if (j >= Hd) break;
    const float4* wr = reinterpret_cast<const float4*>(w1 + static_cast<int64_t>(j) * F);
    float s = 0.f;
#pragma unroll 4
    for (int c = lane; c < F4; c += 32) {
      const float4 wv = __ldg(wr + c);
      const float4 gv = *reinterpret_cast<const float4*>(s_gap + 4 * c);
      s = fmaf(wv.x, gv.x, s); s = fmaf(wv.y, gv.y, s); s = fmaf(wv.z, gv.z, s); s = fmaf(wv.w, gv.w, s);
    }
    s = warp_sum(s);
    if (lane == 0) hidden[static_cast<int64_t>(n) * Hd + j] = fmaxf(s + b1[j], 0.f);
  }
}
// logits[n][k] = W2[k] . hidden[n] + b2[k]; one block per sample
__global__ void __launch_bounds__(256) fc2_fwd_kernel(const float* __restrict__ hidden, int Hd,
                                                      const float* __restrict__ w2, const float* __restrict__ b2, int K,
                                                      float* __restrict__ logits) {
  const int n = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < K; k += 8) {
    const float* wr = w2 + static_cast<int64_t>(k) * Hd;
    float s = 0.f;
    for (int j = lane; j < Hd; j += 32) s = fmaf(wr[j], hidden[static_cast<int64_t>(n) * Hd + j], s);
    s = warp_sum(s);
    if (lane == 0) logits[static_cast<int64_t>(n) * K + k] = s + b2[k];
  }
}

// per sample: dh (written over hidden), dgap[n][c] = (1/HW) sum_j w1[j][c] dh[j], dw2/db2 via atomics
__global__ void __launch_bounds__(256) gap_fc_bwd_sample_kernel(const float* __restrict__ dl, int HW, int F,
                                                                const float* __restrict__ w1, int Hd,
                                                                const float* __restrict__ w2, int K,
                                                                float* __restrict__ hidden, float* __restrict__ dgap,
                                                                float* __restrict__ dw2, float* __restrict__ db2) {
  extern __shared__ float sm[];
  float* s_dh = sm;        // [Hd]
  float* s_dl = sm + Hd;   // [K]
  const int n = blockIdx.x;
  for (int k = threadIdx.x; k < K; k += 256) {
    const float g = dl[static_cast<int64_t>(n) * K + k];
    s_dl[k] = g;
    atomicAdd(db2 + k, g);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < Hd; j += 256) {
    const float h = hidden[static_cast<int64_t>(n) * Hd + j];
    float s = 0.f;
    for (int k = 0; k < K; ++k) {
      s = fmaf(w2[static_cast<int64_t>(k) * Hd + j], s_dl[k], s);
      atomicAdd(dw2 + static_cast<int64_t>(k) * Hd + j, s_dl[k] * h);
    }
    s = h > 0.f ? s : 0.f;
    s_dh[j] = s;
    hidden[static_cast<int64_t>(n) * Hd + j] = s;
  }
  __syncthreads();
  const float inv = 1.f / static_cast<float>(HW);
  for (int c = threadIdx.x; c < F; c += 256) {
    float s = 0.f;
    for (int j = 0; j < Hd; ++j) s = fmaf(w1[static_cast<int64_t>(j) * F + c], s_dh[j], s);
    dgap[static_cast<int64_t>(n) * F + c] = s * inv;
  }
}

// dA[n][p][c] (+)= dgap[n][c] for every pixel: one 16-byte vector per thread, whole grid
template <typename T>
__global__ void __launch_bounds__(256) gap_bcast_kernel(const float* __restrict__ dgap, int HW, int Cp, int F,
                                                        T* __restrict__ dA, int accumulate, int64_t total) {
  const int cvec = Cp / 8;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int v = static_cast<int>(i % cvec);
    const int64_t n = i / (static_cast<int64_t>(cvec) * HW);
    V8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.f[k] = (v * 8 + k) < F ? dgap[n * F + v * 8 + k] : 0.f;
    if (accumulate) {
      const V8 q = load8<T>(dA + i * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) o.f[k] += q.f[k];
    }
    store8<T>(dA + i * 8, o);
  }
}
// dw1[j][c] += sum_n dh[n][j]*gap[n][c] ; db1[j] += sum_n dh[n][j]
__global__ void __launch_bounds__(256) fc1_wgrad_kernel(const float* __restrict__ dh, const float* __restrict__ x,
                                                        int N, int Hd, int F, float* __restrict__ dw1,
                                                        float* __restrict__ db1) {
  const int64_t total = static_cast<int64_t>(Hd) * F;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int j = static_cast<int>(i / F), c = static_cast<int>(i % F);
    float s = 0.f;
    for (int n = 0; n < N; ++n) s = fmaf(dh[static_cast<int64_t>(n) * Hd + j], x[static_cast<int64_t>(n) * F + c], s);
    dw1[i] += s;
    if (c == 0) {
      float b = 0.f;
      for (int n = 0; n < N; ++n) b += dh[static_cast<int64_t>(n) * Hd + j];
      db1[j] += b;
    }
  }
}

// ------------------------------------------------------------------------------------------------ flatten + FC head
// hidden_pre[n][j] += sum_{k in slice} w1[j][k]*a[n][k]   with k = c*HW + hw (NCHW flatten order of the reference)
template <typename T>
__global__ void __launch_bounds__(256) flat_fc1_kernel(const T* __restrict__ a, int N, int HW, int Cp,
                                                       int C, const float* __restrict__ w1, int Hd, int slices,
                                                       float* __restrict__ hidden_pre) {
  const int j = blockIdx.x, sl = blockIdx.y;
  const int64_t F = static_cast<int64_t>(C) * HW;
  const int64_t k0 = F * sl / slices, k1 = F * (sl + 1) / slices;
  __shared__ float s_red[8];
  for (int n0 = 0; n0 < N; n0 += 8) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t k = k0 + threadIdx.x; k < k1; k += 256) {
      const float wv = w1[static_cast<int64_t>(j) * F + k];
      const int c = static_cast<int>(k / HW), hw = static_cast<int>(k % HW);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (n0 + i < N) acc[i] = fmaf(wv, ld1<T>(a + (static_cast<int64_t>(n0 + i) * HW + hw) * Cp + c), acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s = warp_sum(acc[i]);
      __syncthreads();
      if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
      __syncthreads();
      if (threadIdx.x == 0 && n0 + i < N) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += s_red[w];
        atomicAdd(hidden_pre + static_cast<int64_t>(n0 + i) * Hd + j, t);
      }
    }
  }
}
__global__ void flat_fc2_kernel(float* __restrict__ hidden, const float* __restrict__ b1, int N, int Hd,
                                const float* __restrict__ w2, const float* __restrict__ b2, int K,
                                float* __restrict__ logits) {
  extern __shared__ float s_h[];  // [Hd]
  const int n = blockIdx.x;
  for (int j = threadIdx.x; j < Hd; j += blockDim.x) {
    const float h = fmaxf(hidden[static_cast<int64_t>(n) * Hd + j] + b1[j], 0.f);
    hidden[static_cast<int64_t>(n) * Hd + j] = h;
    s_h[j] = h;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < K; k += blockDim.x / 32) {
    float s = 0.f;
    for (int j = lane; j < Hd; j += 32) s = fmaf(w2[static_cast<int64_t>(k) * Hd + j], s_h[j], s);
    s = warp_sum(s);
    if (lane == 0) logits[static_cast<int64_t>(n) * K + k] = s + b2[k];
  }
}
// dh[n][j] = relu'(hidden) * sum_k w2[k][j]*dl[n][k] ; dw2, db2, db1 accumulate
__global__ void flat_fc_bwd_small_kernel(const float* __restrict__ dl, int N, int Hd, const float* __restrict__ w2,
                                         int K, const float* __restrict__ hidden, float* __restrict__ dh,
                                         float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ db2) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < Hd) {
    float bsum = 0.f;
    for (int n = 0; n < N; ++n) {
      const float h = hidden[static_cast<int64_t>(n) * Hd + j];
      float s = 0.f;
      for (int k = 0; k < K; ++k) s = fmaf(w2[static_cast<int64_t>(k) * Hd + j], dl[static_cast<int64_t>(n) * K + k], s);
      s = h > 0.f ? s : 0.f;
      dh[static_cast<int64_t>(n) * Hd + j] = s;
      bsum += s;
    }
    db1[j] += bsum;
    for (int k = 0; k < K; ++k) {
      float s = 0.f;
      for (int n = 0; n < N; ++n) s = fmaf(dl[static_cast<int64_t>(n) * K + k], hidden[static_cast<int64_t>(n) * Hd + j], s);
      dw2[static_cast<int64_t>(k) * Hd + j] += s;
    }
  }
  if (j < K) {
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += dl[static_cast<int64_t>(n) * K + j];
    db2[j] += s;
  }
}
// one thread per flattened column k: single pass over w1 (read) and dw1 (write)
template <typename T>
__global__ void __launch_bounds__(256) flat_fc_bwd_big_kernel(const T* __restrict__ a, int N, int HW,
                                                              int Cp, int C, const float* __restrict__ w1, int Hd,
                                                              const float* __restrict__ dh,
                                                              T* __restrict__ dA, int accumulate,
                                                              float* __restrict__ dw1) {
  extern __shared__ float s_dh[];  // [N][Hd]
  for (int i = threadIdx.x; i < N * Hd; i += 256) s_dh[i] = dh[i];
  __syncthreads();
  const int64_t F = static_cast<int64_t>(C) * HW;
  const int64_t k = blockIdx.x * 256ll + threadIdx.x;
  if (k >= F) return;
  const int c = static_cast<int>(k / HW), hw = static_cast<int>(k % HW);
  for (int n0 = 0; n0 < N; n0 += 8) {
    float av[8], da[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      av[i] = (n0 + i < N) ? ld1<T>(a + (static_cast<int64_t>(n0 + i) * HW + hw) * Cp + c) : 0.f;
      da[i] = 0.f;
    }
    for (int j = 0; j < Hd; ++j) {
      const float wv = w1[static_cast<int64_t>(j) * F + k];
      float g = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (n0 + i < N) {
          const float d = s_dh[(n0 + i) * Hd + j];
          da[i] = fmaf(d, wv, da[i]);
          g = fmaf(d, av[i], g);
        }
      }
      dw1[static_cast<int64_t>(j) * F + k] += g;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (n0 + i < N) {
        T* d = dA + (static_cast<int64_t>(n0 + i) * HW + hw) * Cp + c;
        float o = da[i];
        if (accumulate) o += ld1<T>(d);
        st1<T>(d, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ row softmax
// nn.Softmax(dim=1) on the (N, K) class logits of nnUNetClassifier (nnUNet_classifier.py:110,165-166).  K is the class
// count (<= 32): one thread per sample, the row lives in registers; N * K * 4 bytes, latency bound by construction.
__global__ void __launch_bounds__(128) softmax_rows_fwd_kernel(const float* __restrict__ x, int N, int K,
                                                               float* __restrict__ p) {
  const int n = blockIdx.x * 128 + threadIdx.x;
  if (n >= N) return;
  const float* r = x + static_cast<int64_t>(n) * K;
  float m = r[0];
  for (int k = 1; k < K; ++k) m = fmaxf(m, r[k]);
  float e[32];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    if (k < K) { e[k] = expf(r[k] - m); s += e[k]; }
  }
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    if (k < K) p[static_cast<int64_t>(n) * K + k] = e[k] / s;
  }
}
// dx_k = p_k * (dp_k - sum_j dp_j p_j)
__global__ void __launch_bounds__(128) softmax_rows_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dp,
                                                               int N, int K, float* __restrict__ dx) {
  const int n = blockIdx.x * 128 + threadIdx.x;
  if (n >= N) return;
  const float* pr = p + static_cast<int64_t>(n) * K;
  const float* gr = dp + static_cast<int64_t>(n) * K;
  float dot = 0.f;
  for (int k = 0; k < K; ++k) dot = fmaf(gr[k], pr[k], dot);
  for (int k = 0; k < K; ++k) dx[static_cast<int64_t>(n) * K + k] = pr[k] * (gr[k] - dot);
}

}  // namespace mtbc

using namespace mtbc;
// k*k = 4 / 16 / 64 (k = 2 / 4 / 8) are compiled; C a multiple of 8 with C * k * k <= 8192 (32 accumulators per thread)
template <typename T, int KK>
static void dshead_fwd_launch(int g, size_t smem, cudaStream_t st, const void* a, int N, int H, int W, int Cp, int C, int k,
                              const float* wc, const float* bc, float* logits) {
  if (smem > 48 * 1024) cudaFuncSetAttribute(dshead_fwd_kernel<T, KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  dshead_fwd_kernel<T, KK><<<g, 256, smem, st>>>(static_cast<const T*>(a), N, H, W, Cp, C, k, wc, bc, logits);
}
template <typename T, int KK>
static void dshead_bwd_launch(int g, size_t smem, cudaStream_t st, const void* a, const float* dl, int N, int H, int W,
                              int Cp, int C, int k, const float* wc, void* dA, int accumulate, float* dwc_part,
                              float* dbc_part) {
  if (smem > 48 * 1024) cudaFuncSetAttribute(dshead_bwd_kernel<T, KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  dshead_bwd_kernel<T, KK><<<g, 256, smem, st>>>(static_cast<const T*>(a), dl, N, H, W, Cp, C, k, wc, static_cast<T*>(dA),
                                                 accumulate, dwc_part, dbc_part);
}
#define MTBC_DISPATCH_KK(kk, fn, ...)                                   \
  do {                                                                  \
    if (mtbc::current_mode() & mtbc::MODE_ACT_FP32) {                   \
      if ((kk) == 4) fn<float, 4>(__VA_ARGS__);                         \
      else if ((kk) == 16) fn<float, 16>(__VA_ARGS__);                  \
      else fn<float, 64>(__VA_ARGS__);                                  \
    } else {                                                            \
      if ((kk) == 4) fn<__nv_bfloat16, 4>(__VA_ARGS__);                 \
      else if ((kk) == 16) fn<__nv_bfloat16, 16>(__VA_ARGS__);          \
      else fn<__nv_bfloat16, 64>(__VA_ARGS__);                          \
    }                                                                   \
  } while (0)

static int dshead_check(int32_t Cp, int32_t C, int32_t k) {
  const int kk = k * k;
  if (kk != 4 && kk != 16 && kk != 64) return set_error(MTBC_ERR_INVALID, "dshead: k must be 2, 4 or 8");
  if (C % 8 != 0 || Cp % 8 != 0 || C > Cp) return set_error(MTBC_ERR_INVALID, "dshead: C %% 8 != 0 or C > Cp");
  if (C * kk > 8192) return set_error(MTBC_ERR_INVALID, "dshead: C * k * k > 8192");
  return MTBC_OK;
}
static int dshead_grid(int32_t N, int32_t H, int32_t W) {
  const int64_t ntiles = cdiv(static_cast<int64_t>(N) * H * W, kDsTP);
  return static_cast<int>(ntiles < 148 * 2 ? (ntiles < 1 ? 1 : ntiles) : 148 * 2);
}
#define ST(s) static_cast<cudaStream_t>(s)
#define TP(p) static_cast<T*>(p)
#define CTP(p) static_cast<const T*>(p)

extern "C" {

int mtbc_head1x1_fwd(const void* a, int64_t npix, int32_t Cp, int32_t C, const float* w, const float* b,
                     float* logits, void* stream) {
  int g = cdiv(npix, 256); if (g > 148 * 8) g = 148 * 8;
  MTBC_DISPATCH_ACT((launch_pdl(head1x1_fwd_kernel<T>, dim3(g), dim3(256), Cp * sizeof(float), ST(stream), CTP(a), npix, Cp, C, w, b, logits)));
  return check_launch("head1x1_fwd");
}
int mtbc_head1x1_bwd(const void* a, const float* dlogits, int64_t npix, int32_t Cp, int32_t C, const float* w,
                     void* dA, int32_t accumulate, float* dw, float* db, void* stream) {
  const int cvec = Cp / 8;
  if (Cp % 8 != 0 || cvec > 256) return set_error(MTBC_ERR_INVALID, "head1x1_bwd: Cp %% 8 != 0 or Cp > 2048");
  const int bd = (256 / cvec) * cvec;   // block size multiple of the channel-group count: a thread's group is invariant
  int g = cdiv(npix * cvec, bd * 8); if (g > 148 * 8) g = 148 * 8; if (g < 1) g = 1;
  MTBC_DISPATCH_ACT((launch_pdl(head1x1_bwd_kernel<T>, dim3(g), dim3(bd), 0, ST(stream), CTP(a), dlogits, npix, Cp, C, w, TP(dA), accumulate, dw, db)));
  return check_launch("head1x1_bwd");
}
int mtbc_dshead_compose(const float* wt, const float* bt, const float* w1, const float* b1, int32_t C, int32_t k,
                        float* wc, float* bc, void* stream) {
  const int kk = k * k;
  dshead_compose_kernel<<<cdiv(C * kk, 128), 128, 0, ST(stream)>>>(wt, bt, w1, b1, C, kk, wc, bc);
  return check_launch("dshead_compose");
}
int mtbc_dshead_fwd(const void* a, int32_t N, int32_t H, int32_t W, int32_t Cp, int32_t C, int32_t k, const float* wc,
                    const float* bc, float* logits, void* stream) {
  if (int rc = dshead_check(Cp, C, k)) return rc;
  const int kk = k * k;
  const int64_t ntiles = cdiv(static_cast<int64_t>(N) * H * W, kDsTP);
  const int g = static_cast<int>(ntiles < 148 * 4 ? (ntiles < 1 ? 1 : ntiles) : 148 * 4);
  const size_t smem = (static_cast<size_t>(C) * kk + static_cast<size_t>(kDsTP) * (C + 1)) * sizeof(float);
  MTBC_DISPATCH_KK(kk, dshead_fwd_launch, g, smem, ST(stream), a, N, H, W, Cp, C, k, wc, bc, logits);
  return check_launch("dshead_fwd");
}
int mtbc_dshead_bwd_parts(int32_t N, int32_t H, int32_t W) { return dshead_grid(N, H, W); }
int mtbc_dshead_bwd(const void* a, const float* dlogits, int32_t N, int32_t H, int32_t W, int32_t Cp, int32_t C,
                    int32_t k, const float* wc, void* dA, int32_t accumulate, float* dwc_part, float* dbc_part,
                    int32_t nparts, void* stream) {
  if (int rc = dshead_check(Cp, C, k)) return rc;
  const int kk = k * k;
  const int g = dshead_grid(N, H, W);
  if (nparts != g) return set_error(MTBC_ERR_INVALID, "dshead_bwd: nparts %d != mtbc_dshead_bwd_parts() = %d", nparts, g);
  const size_t smem = (static_cast<size_t>(C) * (kk + 4) + static_cast<size_t>(kDsTP) * (C + 1) +
                       static_cast<size_t>(kDsTP) * (kk + 1)) * sizeof(float);
  MTBC_DISPATCH_KK(kk, dshead_bwd_launch, g, smem, ST(stream), a, dlogits, N, H, W, Cp, C, k, wc, dA, accumulate, dwc_part,
                   dbc_part);
  return check_launch("dshead_bwd");
}
int mtbc_dshead_decompose(float* dwc_part, float* dbc_part, int32_t nparts, const float* wt, const float* bt,
                          const float* w1, int32_t C, int32_t k, float* dwt, float* dbt, float* dw1, float* db1,
                          void* stream) {
  const int kk = k * k, n = C * kk;
  float* dwc = dwc_part + static_cast<int64_t>(nparts) * n;   // row nparts of both buffers receives the totals
  float* dbc = dbc_part + nparts;
  dshead_reduce_kernel<<<cdiv(n, 256), 256, 0, ST(stream)>>>(dwc_part, dbc_part, nparts, n, dwc, dbc);
  const int64_t total = static_cast<int64_t>(C) * n;
  int g = cdiv(total, 256); if (g > 148 * 4) g = 148 * 4;
  dshead_dwt_kernel<<<g, 256, 0, ST(stream)>>>(dwc, w1, C, kk, dwt);
  dshead_dw1_kernel<<<C, 256, 0, ST(stream)>>>(dwc, dbc, wt, bt, w1, C, kk, dbt, dw1, db1);
  return check_launch("dshead_decompose");
}

int mtbc_gap_fc_fwd(const void* a, int32_t N, int32_t HW, int32_t Cp, int32_t F, const float* w1, const float* b1,
                    int32_t Hd, const float* w2, const float* b2, int32_t K, float* gap, float* hidden, float* logits,
                    void* stream) {
  if (Cp % 8 != 0 || F > Cp) return set_error(MTBC_ERR_INVALID, "gap_fc_fwd: Cp %% 8 != 0 or F > Cp");
  MTBC_DISPATCH_ACT((gap_kernel<T><<<dim3(N, cdiv(Cp, 64)), 256, 0, ST(stream)>>>(CTP(a), HW, Cp, F, gap)));
  int rc = check_launch("gap");
  if (rc) return rc;
  if (F % 4 != 0) return set_error(MTBC_ERR_INVALID, "gap_fc_fwd: F %% 4 != 0");
  fc1_fwd_kernel<<<dim3(N, cdiv(Hd, 32)), 256, F * sizeof(float), ST(stream)>>>(gap, F, w1, b1, Hd, hidden);
  rc = check_launch("fc1_fwd");
  if (rc) return rc;
  fc2_fwd_kernel<<<N, 256, 0, ST(stream)>>>(hidden, Hd, w2, b2, K, logits);
  return check_launch("gap_fc_fwd");
}
int mtbc_gap_fc_bwd(const float* dlogits, int32_t N, int32_t HW, int32_t Cp, int32_t F, const float* w1, int32_t Hd,
                    const float* w2, int32_t K, const float* gap, float* hidden, void* dA, int32_t accumulate,
                    float* dw1, float* db1, float* dw2, float* db2, float* dgap, void* stream) {
  if (!dgap) return set_error(MTBC_ERR_INVALID, "gap_fc_bwd: dgap scratch [N][F] is required");
  gap_fc_bwd_sample_kernel<<<N, 256, (Hd + K) * sizeof(float), ST(stream)>>>(dlogits, HW, F, w1, Hd, w2, K, hidden,
                                                                             dgap, dw2, db2);
  int rc = check_launch("gap_fc_bwd_sample");
  if (rc) return rc;
  int g = cdiv(static_cast<int64_t>(Hd) * F, 256); if (g > 148 * 4) g = 148 * 4;
  fc1_wgrad_kernel<<<g, 256, 0, ST(stream)>>>(hidden, gap, N, Hd, F, dw1, db1);
  rc = check_launch("fc1_wgrad");
  if (rc) return rc;
  const int64_t total = static_cast<int64_t>(N) * HW * (Cp / 8);
  int gb = cdiv(total, 256); if (gb > 148 * 8) gb = 148 * 8;
  MTBC_DISPATCH_ACT((gap_bcast_kernel<T><<<gb, 256, 0, ST(stream)>>>(dgap, HW, Cp, F, TP(dA), accumulate, total)));
  return check_launch("gap_bcast");
}

int mtbc_flat_fc_fwd(const void* a, int32_t N, int32_t HW, int32_t Cp, int32_t C, const float* w1, const float* b1,
                     int32_t Hd, const float* w2, const float* b2, int32_t K, float* hidden, float* logits,
                     void* stream) {
  cudaError_t e = cudaMemsetAsync(hidden, 0, static_cast<size_t>(N) * Hd * sizeof(float), ST(stream));
  if (e != cudaSuccess) return set_error(MTBC_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
  const int slices = 4;
  MTBC_DISPATCH_ACT((flat_fc1_kernel<T><<<dim3(Hd, slices), 256, 0, ST(stream)>>>(CTP(a), N, HW, Cp, C, w1, Hd, slices, hidden)));
  int rc = check_launch("flat_fc1");
  if (rc) return rc;
  flat_fc2_kernel<<<N, 128, Hd * sizeof(float), ST(stream)>>>(hidden, b1, N, Hd, w2, b2, K, logits);
  return check_launch("flat_fc2");
}
int mtbc_flat_fc_bwd(const void* a, const float* dlogits, int32_t N, int32_t HW, int32_t Cp, int32_t C,
                     const float* w1, int32_t Hd, const float* w2, int32_t K, const float* hidden, void* dA,
                     int32_t accumulate, float* dw1, float* db1, float* dw2, float* db2, float* scratch, void* stream) {
  flat_fc_bwd_small_kernel<<<cdiv(Hd > K ? Hd : K, 128), 128, 0, ST(stream)>>>(dlogits, N, Hd, w2, K, hidden, scratch,
                                                                             db1, dw2, db2);
  int rc = check_launch("flat_fc_bwd_small");
  if (rc) return rc;
  const int64_t F = static_cast<int64_t>(C) * HW;
  const size_t smem = static_cast<size_t>(N) * Hd * sizeof(float);
  if (smem > 48 * 1024) {
    cudaFuncSetAttribute(flat_fc_bwd_big_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(flat_fc_bwd_big_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  }
  MTBC_DISPATCH_ACT((flat_fc_bwd_big_kernel<T><<<cdiv(F, 256), 256, smem, ST(stream)>>>(CTP(a), N, HW, Cp, C, w1, Hd, scratch, TP(dA),
                                                                                         accumulate, dw1)));
  return check_launch("flat_fc_bwd_big");
}

int mtbc_softmax_rows_fwd(const float* logits, int32_t N, int32_t K, float* probs, void* stream) {
  if (K < 1 || K > 32) return set_error(MTBC_ERR_INVALID, "softmax_rows_fwd: K must be in [1, 32]");
  softmax_rows_fwd_kernel<<<cdiv(N, 128), 128, 0, ST(stream)>>>(logits, N, K, probs);
  return check_launch("softmax_rows_fwd");
}
int mtbc_softmax_rows_bwd(const float* probs, const float* dprobs, int32_t N, int32_t K, float* dlogits, void* stream) {
  if (K < 1 || K > 32) return set_error(MTBC_ERR_INVALID, "softmax_rows_bwd: K must be in [1, 32]");
  softmax_rows_bwd_kernel<<<cdiv(N, 128), 128, 0, ST(stream)>>>(probs, dprobs, N, K, dlogits);
  return check_launch("softmax_rows_bwd");
}

}  // extern "C"
