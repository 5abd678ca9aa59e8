// Memory-bound kernels on bf16 NHWC activations (channel count Cp, multiple of 32): InstanceNorm statistics, the fused
// normalise + affine + LeakyReLU (+ 2x2 max-pool) pass and its two-pass backward, max-pool / nearest-upsample
// backward, per-channel sums.  One thread moves 8 channels (128 bit) per access; a thread's channel group is invariant
// over its grid-stride loop so per-channel constants live in registers.
#include "ptx.cuh"
#include "internal.h"
#include "act_io.cuh"
#include <string>

namespace mtbc {

static int gcd_i(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }
// grid.x such that (grid.x * block) % cvec == 0 (channel group invariant per thread) and ~target blocks overall.
static int pick_gx(int64_t work_items, int block, int cvec, int n_outer, int per_thread) {
  const int m = cvec / gcd_i(cvec, block);
  int64_t want = (work_items + static_cast<int64_t>(block) * per_thread - 1) / (static_cast<int64_t>(block) * per_thread);
  const int64_t cap = (148 * 8 + n_outer - 1) / n_outer;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  int64_t gx = (want + m - 1) / m * m;
  return static_cast<int>(gx);
}

// ------------------------------------------------------------------------------------------------ statistics
template <typename T>
__global__ void __launch_bounds__(256) in_stats_kernel(const T* __restrict__ y, int64_t HW, int Cp,
                                                       float* __restrict__ ssum, float* __restrict__ ssq) {
  extern __shared__ float s_acc[];  // [2][Cp]
  const int n = blockIdx.y, cvec = Cp / 8;
  for (int i = threadIdx.x; i < 2 * Cp; i += 256) s_acc[i] = 0.f;
  __syncthreads();
  const int64_t total = HW * cvec;
  const int64_t start = blockIdx.x * 256ll + threadIdx.x, stride = gridDim.x * 256ll;
  const int v = static_cast<int>(start % cvec);
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const T* base = y + static_cast<int64_t>(n) * HW * Cp;
  for (int64_t i = start; i < total; i += stride) {
    const V8 x = load8<T>(base + i * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] += x.f[k]; q[k] = fmaf(x.f[k], x.f[k], q[k]); }
  }
  if (start < total) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { atomicAdd(&s_acc[v * 8 + k], a[k]); atomicAdd(&s_acc[Cp + v * 8 + k], q[k]); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cp; i += 256) {
    atomicAdd(ssum + static_cast<int64_t>(n) * Cp + i, s_acc[i]);
    atomicAdd(ssq + static_cast<int64_t>(n) * Cp + i, s_acc[Cp + i]);
  }
}


// Order-independent variant (MODE_DETERMINISTIC): no floating-point atomics anywhere.  Every thread's partial is parked
// in a [k][thread] table, one thread per channel adds its rows in a fixed order, the block's sums go to
// partial[n][block][2][Cp], and the LAST block of a sample to arrive (integer ticket) adds the blocks' partials in block
// order.  Which block is last varies from run to run; the order of the additions does not, so the statistics -- and with
// them every bf16 rounding downstream -- are bit-identical across runs.  blockDim.x is a multiple of Cp / 8.
template <typename T>
__global__ void __launch_bounds__(256) in_stats_det_kernel(const T* __restrict__ y, int64_t HW, int Cp,
                                                           float* __restrict__ partial, int* __restrict__ ticket,
                                                           float* __restrict__ ssum, float* __restrict__ ssq) {
  __shared__ float s_tab[16 * 256];
  __shared__ int s_last;
  const int n = blockIdx.y, cvec = Cp / 8, bd = blockDim.x;
  const int64_t total = HW * cvec;
  const int64_t start = static_cast<int64_t>(blockIdx.x) * bd + threadIdx.x, stride = static_cast<int64_t>(gridDim.x) * bd;
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const T* base = y + static_cast<int64_t>(n) * HW * Cp;
  for (int64_t i = start; i < total; i += stride) {
    const V8 x = load8<T>(base + i * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] += x.f[k]; q[k] = fmaf(x.f[k], x.f[k], q[k]); }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) { s_tab[k * 256 + threadIdx.x] = a[k]; s_tab[(8 + k) * 256 + threadIdx.x] = q[k]; }
  __syncthreads();
  float* mine = partial + (static_cast<int64_t>(n) * gridDim.x + blockIdx.x) * 2 * Cp;
  for (int ch = threadIdx.x; ch < Cp; ch += bd) {   // thread t's channel group is t % cvec (bd % cvec == 0)
    float s0 = 0.f, s1 = 0.f;
    for (int r = ch >> 3; r < bd; r += cvec) { s0 += s_tab[(ch & 7) * 256 + r]; s1 += s_tab[(8 + (ch & 7)) * 256 + r]; }
    mine[ch] = s0;
    mine[Cp + ch] = s1;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket + n, 1) == static_cast<int>(gridDim.x) - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float* all = partial + static_cast<int64_t>(n) * gridDim.x * 2 * Cp;
  for (int ch = threadIdx.x; ch < Cp; ch += bd) {
    float s0 = 0.f, s1 = 0.f;
    for (unsigned b = 0; b < gridDim.x; ++b) { s0 += all[static_cast<int64_t>(b) * 2 * Cp + ch]; s1 += all[static_cast<int64_t>(b) * 2 * Cp + Cp + ch]; }
    ssum[static_cast<int64_t>(n) * Cp + ch] = s0;
    ssq[static_cast<int64_t>(n) * Cp + ch] = s1;
  }
  if (threadIdx.x == 0) ticket[n] = 0;   // ready for the next launch (CUDA-graph replay)
}
// Blocks per SAMPLE: a function of the plane only, never of the batch size, so that a sample's statistics are the same
// bits whether it is processed in a batch of 16 or of 32 (data-parallel runs reproduce the single-GPU run exactly).
static int stats_det_blocks(int64_t HW, int cvec, int N) {
  (void)N;
  int64_t want = (HW * cvec + 256 * 16 - 1) / (256 * 16);
  if (want > 32) want = 32;
  if (want < 1) want = 1;
  return static_cast<int>(want);
}

// ------------------------------------------------------------------------------------------------ forward apply
struct NormConst {
  float sc[8], sh[8];
};
__device__ __forceinline__ NormConst norm_consts(const float* ssum, const float* ssq, const float* gamma,
                                                 const float* beta, int n, int Cp, int v, float inv_hw, float eps,
                                                 float* mean_out, float* rstd_out, bool write) {
  NormConst c;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = v * 8 + k;
    const float m = ssum[static_cast<int64_t>(n) * Cp + ch] * inv_hw;
    float var = ssq[static_cast<int64_t>(n) * Cp + ch] * inv_hw - m * m;
    var = fmaxf(var, 0.f);
    const float r = rsqrtf(var + eps);
    const float g = gamma ? gamma[ch] : 1.f;
    const float b = beta ? beta[ch] : 0.f;
    c.sc[k] = r * g;
    c.sh[k] = b - m * r * g;
    if (write) { mean_out[static_cast<int64_t>(n) * Cp + ch] = m; rstd_out[static_cast<int64_t>(n) * Cp + ch] = r; }
  }
  return c;
}

template <typename T>
__global__ void __launch_bounds__(256) in_apply_kernel(const T* __restrict__ y, int64_t HW, int Cp,
                                                       const float* __restrict__ ssum, const float* __restrict__ ssq,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float eps, float slope, T* __restrict__ a,
                                                       float* __restrict__ mean, float* __restrict__ rstd) {
  pdl_trigger();   // programmatic dependent launch (ptx.cuh): wait before the first global access
  pdl_wait();
  const int n = blockIdx.y, cvec = Cp / 8;
  const int64_t total = HW * cvec;
  const int64_t start = blockIdx.x * 256ll + threadIdx.x, stride = gridDim.x * 256ll;
  const int v = static_cast<int>(start % cvec);
  const NormConst c = norm_consts(ssum, ssq, gamma, beta, n, Cp, v, 1.f / static_cast<float>(HW), eps, mean, rstd,
                                  start < cvec);
  const T* src = y + static_cast<int64_t>(n) * HW * Cp;
  T* dst = a + static_cast<int64_t>(n) * HW * Cp;
  for (int64_t i = start; i < total; i += 2 * stride) {
    const bool two = (i + stride) < total;
    const V8 x0 = load8<T>(src + i * 8);
    V8 x1;
    if (two) x1 = load8<T>(src + (i + stride) * 8);
    V8 o0, o1;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float z0 = fmaf(x0.f[k], c.sc[k], c.sh[k]);
      o0.f[k] = z0 > 0.f ? z0 : z0 * slope;
    }
    store8<T>(dst + i * 8, o0);
    if (two) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float z1 = fmaf(x1.f[k], c.sc[k], c.sh[k]);
        o1.f[k] = z1 > 0.f ? z1 : z1 * slope;
      }
      store8<T>(dst + (i + stride) * 8, o1);
    }
  }
}

// Same, plus the 2x2/2 max-pooled tensor; one work item = one pooled pixel x 8 channels.
template <typename T>
__global__ void __launch_bounds__(256) in_apply_pool_kernel(const T* __restrict__ y, int H, int W, int Cp,
                                                            const float* __restrict__ ssum,
                                                            const float* __restrict__ ssq,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, float slope,
                                                            T* __restrict__ a,
                                                            T* __restrict__ pooled,
                                                            float* __restrict__ mean, float* __restrict__ rstd) {
  pdl_trigger();   // programmatic dependent launch (ptx.cuh): wait before the first global access
  pdl_wait();
  const int n = blockIdx.y, cvec = Cp / 8;
  const int Hp = H / 2, Wp = W / 2;
  const int64_t HW = static_cast<int64_t>(H) * W;
  const int64_t total = static_cast<int64_t>(Hp) * Wp * cvec;
  const int64_t start = blockIdx.x * 256ll + threadIdx.x, stride = gridDim.x * 256ll;
  const int v = static_cast<int>(start % cvec);
  const NormConst c = norm_consts(ssum, ssq, gamma, beta, n, Cp, v, 1.f / static_cast<float>(HW), eps, mean, rstd,
                                  start < cvec);
  const T* src = y + static_cast<int64_t>(n) * HW * Cp + v * 8;
  T* dst = a + static_cast<int64_t>(n) * HW * Cp + v * 8;
  T* pdst = pooled + static_cast<int64_t>(n) * Hp * Wp * Cp + v * 8;
  for (int64_t i = start; i < total; i += stride) {
    const int64_t pp = i / cvec;
    const int ph = static_cast<int>(pp / Wp), pw = static_cast<int>(pp - static_cast<int64_t>(ph) * Wp);
    const int64_t o00 = (static_cast<int64_t>(2 * ph) * W + 2 * pw) * Cp;
    const int64_t offs[4] = {o00, o00 + Cp, o00 + static_cast<int64_t>(W) * Cp, o00 + static_cast<int64_t>(W) * Cp + Cp};
    V8 x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = load8<T>(src + offs[j]);
    V8 mx;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float z = fmaf(x[j].f[k], c.sc[k], c.sh[k]);
        // as it reads back from storage, so the pooled value equals the stored activation bit for bit
        const float o = as_stored<T>(z > 0.f ? z : z * slope);
        x[j].f[k] = o;
        mx.f[k] = (j == 0) ? o : fmaxf(mx.f[k], o);
      }
      store8<T>(dst + offs[j], x[j]);
    }
    store8<T>(pdst + pp * Cp, mx);
  }
}

// ------------------------------------------------------------------------------------------------ backward
struct BwdConst {
  float mean[8], rstd[8], g[8], b[8];
};
__device__ __forceinline__ BwdConst bwd_consts(const float* mean, const float* rstd, const float* gamma,
                                               const float* beta, int n, int Cp, int v) {
  BwdConst c;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = v * 8 + k;
    c.mean[k] = mean[static_cast<int64_t>(n) * Cp + ch];
    c.rstd[k] = rstd[static_cast<int64_t>(n) * Cp + ch];
    c.g[k] = gamma ? gamma[ch] : 1.f;
    c.b[k] = beta ? beta[ch] : 0.f;
  }
  return c;
}

template <typename T>
__global__ void __launch_bounds__(256) in_bwd_reduce_kernel(const T* __restrict__ dA,
                                                            const T* __restrict__ y, int64_t HW, int Cp,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float slope,
                                                            float* __restrict__ s1, float* __restrict__ s2) {
  pdl_trigger();   // programmatic dependent launch (ptx.cuh): wait before the first global access
  pdl_wait();
  extern __shared__ float s_acc[];  // [2][Cp]
  const int n = blockIdx.y, cvec = Cp / 8;
  for (int i = threadIdx.x; i < 2 * Cp; i += 256) s_acc[i] = 0.f;
  __syncthreads();
  const int64_t total = HW * cvec;
  const int64_t start = blockIdx.x * 256ll + threadIdx.x, stride = gridDim.x * 256ll;
  const int v = static_cast<int>(start % cvec);
  const BwdConst c = bwd_consts(mean, rstd, gamma, beta, n, Cp, v);
  float a1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, a2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const T* gsrc = dA + static_cast<int64_t>(n) * HW * Cp;
  const T* ysrc = y + static_cast<int64_t>(n) * HW * Cp;
  for (int64_t i = start; i < total; i += stride) {
    const V8 g = load8<T>(gsrc + i * 8);
    const V8 x = load8<T>(ysrc + i * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xh = (x.f[k] - c.mean[k]) * c.rstd[k];
      const float z = fmaf(xh, c.g[k], c.b[k]);
      const float gg = z > 0.f ? g.f[k] : g.f[k] * slope;
      a1[k] += gg;
      a2[k] = fmaf(gg, xh, a2[k]);
    }
  }
  // combine lanes that share a channel group before touching shared memory
  if ((32 % cvec) == 0) {
    for (int o = cvec; o < 32; o <<= 1) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        a1[k] += __shfl_xor_sync(0xffffffffu, a1[k], o);
        a2[k] += __shfl_xor_sync(0xffffffffu, a2[k], o);
      }
    }
    if ((threadIdx.x & 31) < cvec) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { atomicAdd(&s_acc[v * 8 + k], a1[k]); atomicAdd(&s_acc[Cp + v * 8 + k], a2[k]); }
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) { atomicAdd(&s_acc[v * 8 + k], a1[k]); atomicAdd(&s_acc[Cp + v * 8 + k], a2[k]); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cp; i += 256) {
    atomicAdd(s1 + static_cast<int64_t>(n) * Cp + i, s_acc[i]);
    atomicAdd(s2 + static_cast<int64_t>(n) * Cp + i, s_acc[Cp + i]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) in_bwd_apply_kernel(const T* __restrict__ dA,
                                                           const T* __restrict__ y, int64_t HW, int Cp,
                                                           const float* __restrict__ mean,
                                                           const float* __restrict__ rstd,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float slope,
                                                           const float* __restrict__ s1, const float* __restrict__ s2,
                                                           T* __restrict__ dy) {
  pdl_trigger();   // programmatic dependent launch (ptx.cuh): wait before the first global access
  pdl_wait();
  const int n = blockIdx.y, cvec = Cp / 8;
  const int64_t total = HW * cvec;
  const int64_t start = blockIdx.x * 256ll + threadIdx.x, stride = gridDim.x * 256ll;
  const int v = static_cast<int>(start % cvec);
  const BwdConst c = bwd_consts(mean, rstd, gamma, beta, n, Cp, v);
  float m1[8], m2[8], rg[8];
  const float inv_hw = 1.f / static_cast<float>(HW);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    m1[k] = s1[static_cast<int64_t>(n) * Cp + v * 8 + k] * inv_hw;
    m2[k] = s2[static_cast<int64_t>(n) * Cp + v * 8 + k] * inv_hw;
    rg[k] = c.rstd[k] * c.g[k];
  }
  const T* gsrc = dA + static_cast<int64_t>(n) * HW * Cp;
  const T* ysrc = y + static_cast<int64_t>(n) * HW * Cp;
  T* dst = dy + static_cast<int64_t>(n) * HW * Cp;
  for (int64_t i = start; i < total; i += stride) {
    const V8 g = load8<T>(gsrc + i * 8);
    const V8 x = load8<T>(ysrc + i * 8);
    V8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xh = (x.f[k] - c.mean[k]) * c.rstd[k];
      const float z = fmaf(xh, c.g[k], c.b[k]);
      const float gg = z > 0.f ? g.f[k] : g.f[k] * slope;
      o.f[k] = rg[k] * (gg - m1[k] - xh * m2[k]);
    }
    store8<T>(dst + i * 8, o);
  }
}

// dgamma[c] += sum_n s2[n][c], dbeta[c] += sum_n s1[n][c]
__global__ void in_affine_grad_kernel(const float* __restrict__ s1, const float* __restrict__ s2, int N, int Cp,
                                      int C_true, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C_true) return;
  float a = 0.f, b = 0.f;
  for (int n = 0; n < N; ++n) { a += s2[static_cast<int64_t>(n) * Cp + c]; b += s1[static_cast<int64_t>(n) * Cp + c]; }
  dgamma[c] += a;
  dbeta[c] += b;
}

// ------------------------------------------------------------------------------------------------ pooling / upsample
template <typename T>
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const T* __restrict__ a,
                                                           const T* __restrict__ dP, int H, int W, int Cp,
                                                           T* __restrict__ dA, int accumulate,
                                                           int64_t total) {
  pdl_trigger();   // programmatic dependent launch (ptx.cuh): wait before the first global access
  pdl_wait();
  const int cvec = Cp / 8, Hp = H / 2, Wp = W / 2;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int v = static_cast<int>(i % cvec);
    const int64_t pp = i / cvec;  // (n, ph, pw)
    const int pw = static_cast<int>(pp % Wp);
    const int ph = static_cast<int>((pp / Wp) % Hp);
    const int64_t n = pp / (static_cast<int64_t>(Wp) * Hp);
    const int64_t o00 = ((n * H + 2 * ph) * W + 2 * pw) * Cp + v * 8;
    const int64_t offs[4] = {o00, o00 + Cp, o00 + static_cast<int64_t>(W) * Cp, o00 + static_cast<int64_t>(W) * Cp + Cp};
    V8 x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = load8<T>(a + offs[j]);
    const V8 g = load8<T>(dP + pp * Cp + v * 8);
    V8 o[4];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int arg = 0;
      float best = x[0].f[k];
#pragma unroll
      for (int j = 1; j < 4; ++j)
        if (x[j].f[k] > best) { best = x[j].f[k]; arg = j; }
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j].f[k] = (j == arg) ? g.f[k] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (accumulate) {
        const V8 old = load8<T>(dA + offs[j]);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[j].f[k] += old.f[k];
      }
      store8<T>(dA + offs[j], o[j]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) upsample2_fwd_kernel(const T* __restrict__ x, int H, int W, int Cp,
                                                            T* __restrict__ y, int64_t total) {
  const int cvec = Cp / 8;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int v = static_cast<int>(i % cvec);
    const int64_t pp = i / cvec;  // (n, h, w) of the input
    const int w = static_cast<int>(pp % W);
    const int h = static_cast<int>((pp / W) % H);
    const int64_t n = pp / (static_cast<int64_t>(W) * H);
    const V8 u = load8<T>(x + pp * Cp + v * 8);
    const int64_t o00 = ((n * 2 * H + 2 * h) * (2 * W) + 2 * w) * Cp + v * 8;
    store8<T>(y + o00, u);
    store8<T>(y + o00 + Cp, u);
    store8<T>(y + o00 + static_cast<int64_t>(2 * W) * Cp, u);
    store8<T>(y + o00 + static_cast<int64_t>(2 * W) * Cp + Cp, u);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) upsample2_bwd_kernel(const T* __restrict__ dy, int H, int W, int Cp,
                                                            T* __restrict__ dx, int accumulate,
                                                            int64_t total) {
  const int cvec = Cp / 8;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int v = static_cast<int>(i % cvec);
    const int64_t pp = i / cvec;
    const int w = static_cast<int>(pp % W);
    const int h = static_cast<int>((pp / W) % H);
    const int64_t n = pp / (static_cast<int64_t>(W) * H);
    const int64_t o00 = ((n * 2 * H + 2 * h) * (2 * W) + 2 * w) * Cp + v * 8;
    const V8 a = load8<T>(dy + o00), b = load8<T>(dy + o00 + Cp), c = load8<T>(dy + o00 + static_cast<int64_t>(2 * W) * Cp),
             d = load8<T>(dy + o00 + static_cast<int64_t>(2 * W) * Cp + Cp);
    V8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.f[k] = (a.f[k] + b.f[k]) + (c.f[k] + d.f[k]);
    if (accumulate) {
      const V8 old = load8<T>(dx + pp * Cp + v * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) o.f[k] += old.f[k];
    }
    store8<T>(dx + pp * Cp + v * 8, o);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) channel_sum_kernel(const T* __restrict__ t, int64_t npix, int Cp,
                                                          int C_true, float* __restrict__ out) {
  extern __shared__ float s_acc[];  // [Cp]
  const int cvec = Cp / 8;
  for (int i = threadIdx.x; i < Cp; i += 256) s_acc[i] = 0.f;
  __syncthreads();
  const int64_t total = npix * cvec;
  const int64_t start = blockIdx.x * 256ll + threadIdx.x, stride = gridDim.x * 256ll;
  const int v = static_cast<int>(start % cvec);
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t i = start; i < total; i += stride) {
    const V8 x = load8<T>(t + i * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] += x.f[k];
  }
  if (start < total) {
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&s_acc[v * 8 + k], a[k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C_true; i += 256) atomicAdd(out + i, s_acc[i]);
}

}  // namespace mtbc

using namespace mtbc;

#define ST(s) static_cast<cudaStream_t>(s)
#define TP(p) static_cast<T*>(p)
#define CTP(p) static_cast<const T*>(p)
// bf16 only: the bulk-copy pipelined kernels of stream_pipe.cu serve the product path's large tensors
static inline bool use_pipe(int64_t N, int64_t HW, int Cp) {
  return !(current_mode() & MODE_ACT_FP32) && pipe_eligible(N, HW, Cp);
}

extern "C" {

int mtbc_in_stats(const void* y, int32_t N, int32_t HW, int32_t Cp, float* stat_sum, float* stat_sq, void* stream) {
  if (Cp % 8) return set_error(MTBC_ERR_INVALID, "in_stats: Cp %% 8");
  const int cvec = Cp / 8;
  const int gx = pick_gx(static_cast<int64_t>(HW) * cvec, 256, cvec, N, 8);
  MTBC_DISPATCH_ACT((in_stats_kernel<T><<<dim3(gx, N), 256, 2 * Cp * sizeof(float), ST(stream)>>>(CTP(y), HW, Cp, stat_sum, stat_sq)));
  return check_launch("in_stats");
}

int64_t mtbc_query_workspace_bytes(const char* op, int32_t N, int64_t HW, int32_t Cp) {
  if (op == nullptr) return -1;
  const std::string name(op);
  if (name == "in_stats_det") {
    if (Cp % 8 || Cp <= 0 || N <= 0) return -1;
    const int gx = stats_det_blocks(HW, Cp / 8, N);
    return static_cast<int64_t>(N) * gx * 2 * Cp * sizeof(float) + static_cast<int64_t>(N) * sizeof(int) + 256;
  }
  return 0;   // every other entry point works in the buffers it is handed
}

int mtbc_in_stats_det(const void* y, int32_t N, int32_t HW, int32_t Cp, float* stat_sum, float* stat_sq,
                      void* workspace, void* stream) {
  if (Cp % 8 || Cp / 8 > 256) return set_error(MTBC_ERR_INVALID, "in_stats_det: Cp %% 8 != 0 or Cp > 2048");
  if (!workspace) return set_error(MTBC_ERR_INVALID, "in_stats_det: workspace of mtbc_query_workspace_bytes(\"in_stats_det\") bytes, zeroed once, is required");
  const int cvec = Cp / 8;
  const int bd = (256 / cvec) * cvec;
  const int gx = stats_det_blocks(HW, cvec, N);
  int* ticket = static_cast<int*>(workspace);                       // [N], zero between launches
  float* partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + ((static_cast<size_t>(N) * sizeof(int) + 255) & ~size_t(255)));
  MTBC_DISPATCH_ACT((in_stats_det_kernel<T><<<dim3(gx, N), bd, 0, ST(stream)>>>(CTP(y), HW, Cp, partial, ticket, stat_sum, stat_sq)));
  return check_launch("in_stats_det");
}

int mtbc_in_apply(const void* y, int32_t N, int32_t H, int32_t W, int32_t Cp, const float* stat_sum,
                  const float* stat_sq, const float* gamma, const float* beta, int32_t C_true, float eps, float slope,
                  void* a, void* pooled, float* mean, float* rstd, void* stream) {
  (void)C_true;
  if (Cp % 8) return set_error(MTBC_ERR_INVALID, "in_apply: Cp %% 8");
  const int cvec = Cp / 8;
  const int64_t HW = static_cast<int64_t>(H) * W;
  if (pooled) {
    if ((H & 1) || (W & 1)) return set_error(MTBC_ERR_INVALID, "in_apply: pooled output needs even H, W");
    const int gx = pick_gx(HW / 4 * cvec, 256, cvec, N, 2);
    MTBC_DISPATCH_ACT((launch_pdl(in_apply_pool_kernel<T>, dim3(gx, N), dim3(256), 0, ST(stream), CTP(y), H, W, Cp, stat_sum, stat_sq, gamma, beta, eps,
                                                                                     slope, TP(a), TP(pooled), mean, rstd)));
  } else {
    if (use_pipe(N, HW, Cp))
      return in_apply_pipe(y, N, HW, Cp, stat_sum, stat_sq, gamma, beta, eps, slope, a, mean, rstd, ST(stream));
    const int gx = pick_gx(HW * cvec, 256, cvec, N, 4);
    MTBC_DISPATCH_ACT((launch_pdl(in_apply_kernel<T>, dim3(gx, N), dim3(256), 0, ST(stream), CTP(y), HW, Cp, stat_sum, stat_sq, gamma, beta, eps, slope,
                                                                                TP(a), mean, rstd)));
  }
  return check_launch("in_apply");
}

int mtbc_in_bwd_reduce(const void* dA, const void* y, int32_t N, int32_t HW, int32_t Cp, const float* mean,
                       const float* rstd, const float* gamma, const float* beta, float slope, float* s1, float* s2,
                       void* stream) {
  if (use_pipe(N, HW, Cp))
    return in_bwd_reduce_pipe(dA, y, N, HW, Cp, mean, rstd, gamma, beta, slope, s1, s2, ST(stream));
  const int cvec = Cp / 8;
  const int gx = pick_gx(static_cast<int64_t>(HW) * cvec, 256, cvec, N, 8);
  MTBC_DISPATCH_ACT((launch_pdl(in_bwd_reduce_kernel<T>, dim3(gx, N), dim3(256), 2 * Cp * sizeof(float), ST(stream), CTP(dA), CTP(y), HW, Cp, mean, rstd,
                                                                                                         gamma, beta, slope, s1, s2)));
  return check_launch("in_bwd_reduce");
}

int mtbc_in_bwd_apply(const void* dA, const void* y, int32_t N, int32_t HW, int32_t Cp, const float* mean,
                      const float* rstd, const float* gamma, const float* beta, float slope, const float* s1,
                      const float* s2, void* dy, float* dgamma, float* dbeta, int32_t C_true, void* stream) {
  if (use_pipe(N, HW, Cp))
    return in_bwd_apply_pipe(dA, y, N, HW, Cp, mean, rstd, gamma, beta, slope, s1, s2, dy, dgamma, dbeta, C_true,
                             ST(stream));
  const int cvec = Cp / 8;
  const int gx = pick_gx(static_cast<int64_t>(HW) * cvec, 256, cvec, N, 4);
  MTBC_DISPATCH_ACT((launch_pdl(in_bwd_apply_kernel<T>, dim3(gx, N), dim3(256), 0, ST(stream), CTP(dA), CTP(y), HW, Cp, mean, rstd, gamma, beta, slope, s1,
                                                                                  s2, TP(dy))));
  int rc = check_launch("in_bwd_apply");
  if (rc) return rc;
  if (dgamma && dbeta) {
    in_affine_grad_kernel<<<cdiv(C_true, 128), 128, 0, ST(stream)>>>(s1, s2, N, Cp, C_true, dgamma, dbeta);
    rc = check_launch("in_affine_grad");
  }
  return rc;
}

int mtbc_in_bwd(const void* dA, const void* y, int32_t N, int32_t HW, int32_t Cp, const float* mean, const float* rstd,
                const float* gamma, const float* beta, float slope, float* s1, float* s2, void* dy, float* dgamma,
                float* dbeta, int32_t C_true, int32_t* counters, void* stream) {
  if (counters != nullptr && !(current_mode() & MODE_ACT_FP32) && in_bwd_fused_eligible(N, HW, Cp))
    return in_bwd_fused(dA, y, N, HW, Cp, mean, rstd, gamma, beta, slope, s1, s2, dy, dgamma, dbeta, C_true, counters,
                        ST(stream));
  int rc = mtbc_in_bwd_reduce(dA, y, N, HW, Cp, mean, rstd, gamma, beta, slope, s1, s2, stream);
  if (rc) return rc;
  return mtbc_in_bwd_apply(dA, y, N, HW, Cp, mean, rstd, gamma, beta, slope, s1, s2, dy, dgamma, dbeta, C_true, stream);
}

int mtbc_maxpool2_bwd(const void* a, const void* dP, int32_t N, int32_t H, int32_t W, int32_t Cp, void* dA,
                      int32_t accumulate, void* stream) {
  const int64_t total = static_cast<int64_t>(N) * (H / 2) * (W / 2) * (Cp / 8);
  int g = cdiv(total, 256); if (g > 148 * 8) g = 148 * 8;
  MTBC_DISPATCH_ACT((launch_pdl(maxpool2_bwd_kernel<T>, dim3(g), dim3(256), 0, ST(stream), CTP(a), CTP(dP), H, W, Cp, TP(dA), accumulate, total)));
  return check_launch("maxpool2_bwd");
}
int mtbc_upsample2_fwd(const void* x, int32_t N, int32_t H, int32_t W, int32_t Cp, void* y, void* stream) {
  const int64_t total = static_cast<int64_t>(N) * H * W * (Cp / 8);
  int g = cdiv(total, 256); if (g > 148 * 8) g = 148 * 8;
  MTBC_DISPATCH_ACT((upsample2_fwd_kernel<T><<<g, 256, 0, ST(stream)>>>(CTP(x), H, W, Cp, TP(y), total)));
  return check_launch("upsample2_fwd");
}
int mtbc_upsample2_bwd(const void* dy, int32_t N, int32_t H, int32_t W, int32_t Cp, void* dx, int32_t accumulate,
                       void* stream) {
  const int64_t total = static_cast<int64_t>(N) * H * W * (Cp / 8);
  int g = cdiv(total, 256); if (g > 148 * 8) g = 148 * 8;
  MTBC_DISPATCH_ACT((upsample2_bwd_kernel<T><<<g, 256, 0, ST(stream)>>>(CTP(dy), H, W, Cp, TP(dx), accumulate, total)));
  return check_launch("upsample2_bwd");
}
int mtbc_channel_sum(const void* t, int64_t npix, int32_t Cp, int32_t C_true, float* out, int32_t add, void* stream) {
  if (!add) {
    cudaError_t e = cudaMemsetAsync(out, 0, C_true * sizeof(float), ST(stream));
    if (e != cudaSuccess) return set_error(MTBC_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
  }
  if (use_pipe(1, npix, Cp)) return channel_sum_pipe(t, npix, Cp, C_true, out, ST(stream));
  const int cvec = Cp / 8;
  const int gx = pick_gx(npix * cvec, 256, cvec, 1, 16);
  MTBC_DISPATCH_ACT((channel_sum_kernel<T><<<gx, 256, Cp * sizeof(float), ST(stream)>>>(CTP(t), npix, Cp, C_true, out)));
  return check_launch("channel_sum");
}

}  // extern "C"
