// Kernels only the BatchNorm / residual / dropout sibling backbone needs (reference src/models/segmentation/
// ResidualUNet.py: BatchNorm2d :40,55,110-133, F.dropout(p=0.2) :61,139-145, path + residual :69,155, stride-2 3x3
// convs :115-131).  All memory bound, 8 channels (one 16-byte bf16 vector, two for fp32) per thread.
//
// BatchNorm2d reuses the InstanceNorm kernels: mtbc_in_apply / mtbc_in_bwd_* take per-(sample, channel) sums; pooling
// those sums over the batch first (mtbc_bn_pool_*) turns the very same passes into batch normalisation.
#include "ptx.cuh"
#include "internal.h"
#include "act_io.cuh"

namespace mtbc {

template <typename T>
__global__ void __launch_bounds__(256) add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out,
                                                  int64_t n8) {
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n8; i += gridDim.x * 256ll) {
    const V8 x = load8<T>(a + i * 8), y = load8<T>(b + i * 8);
    V8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.f[k] = x.f[k] + y.f[k];
    store8<T>(out + i * 8, o);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) accumulate_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t n8,
                                                         int accumulate) {
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n8; i += gridDim.x * 256ll) {
    V8 o = load8<T>(src + i * 8);
    if (accumulate) {
      const V8 d = load8<T>(dst + i * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) o.f[k] += d.f[k];
    }
    store8<T>(dst + i * 8, o);
  }
}

// Counter-based generator (splitmix64 finaliser over (seed, draw counter, layer, element)): no state to carry, the same
// (seed, counter) always reproduces the same mask, a CUDA-graph replay advances it through the device-side counter.
__device__ __forceinline__ uint32_t mix32(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return static_cast<uint32_t>(z >> 32);
}

// out = x * keep / (1 - p); keep (one byte per element) is drawn here, or read when `external` (parity tests hand the
// oracle's mask over).  F.dropout(path, p=0.2) is called with its default training=True: it is active in eval mode too.
template <typename T>
__global__ void __launch_bounds__(256) dropout_fwd_kernel(const T* __restrict__ x, T* __restrict__ out,
                                                          uint8_t* __restrict__ mask, int64_t n8, float scale,
                                                          uint32_t thresh, uint64_t seed, const int32_t* __restrict__ counter,
                                                          uint32_t layer, int external) {
  const uint64_t base = seed ^ (static_cast<uint64_t>(counter ? counter[0] : 0) << 32) ^ (static_cast<uint64_t>(layer) << 20);
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n8; i += gridDim.x * 256ll) {
    const V8 v = load8<T>(x + i * 8);
    uint8_t m[8];
    if (external) {
      const uint2 u = *reinterpret_cast<const uint2*>(mask + i * 8);
#pragma unroll
      for (int k = 0; k < 4; ++k) { m[k] = (u.x >> (8 * k)) & 0xFF; m[4 + k] = (u.y >> (8 * k)) & 0xFF; }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        m[k] = mix32(base + 0x9E3779B97F4A7C15ull * static_cast<uint64_t>(i * 8 + k + 1)) >= thresh ? 1 : 0;
      uint2 u;
      u.x = m[0] | (m[1] << 8) | (m[2] << 16) | (m[3] << 24);
      u.y = m[4] | (m[5] << 8) | (m[6] << 16) | (m[7] << 24);
      *reinterpret_cast<uint2*>(mask + i * 8) = u;
    }
    V8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.f[k] = m[k] ? v.f[k] * scale : 0.f;
    store8<T>(out + i * 8, o);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) dropout_bwd_kernel(const T* __restrict__ g, const uint8_t* __restrict__ mask,
                                                          T* __restrict__ out, int64_t n8, float scale, int accumulate) {
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n8; i += gridDim.x * 256ll) {
    const V8 v = load8<T>(g + i * 8);
    const uint2 u = *reinterpret_cast<const uint2*>(mask + i * 8);
    V8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t w = k < 4 ? u.x : u.y;
      o.f[k] = ((w >> (8 * (k & 3))) & 0xFF) ? v.f[k] * scale : 0.f;
    }
    if (accumulate) {
      const V8 d = load8<T>(out + i * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) o.f[k] += d.f[k];
    }
    store8<T>(out + i * 8, o);
  }
}

// out (N, 2H, 2W, Cp) = dy placed on the even lattice, zero elsewhere: the data gradient of a stride-2 convolution is
// the stride-1 data gradient of this tensor, so the stride-1 tensor-core kernels serve it unchanged.
template <typename T>
__global__ void __launch_bounds__(256) zero_stuff2_kernel(const T* __restrict__ dy, int H, int W, int Cp,
                                                          T* __restrict__ out, int64_t total) {
  const int cvec = Cp / 8;
  V8 z;
#pragma unroll
  for (int k = 0; k < 8; ++k) z.f[k] = 0.f;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int v = static_cast<int>(i % cvec);
    const int64_t pp = i / cvec;   // (n, h, w) of dy
    const int w = static_cast<int>(pp % W);
    const int h = static_cast<int>((pp / W) % H);
    const int64_t n = pp / (static_cast<int64_t>(W) * H);
    const V8 u = load8<T>(dy + pp * Cp + v * 8);
    const int64_t o00 = ((n * 2 * H + 2 * h) * (2 * W) + 2 * w) * Cp + v * 8;
    store8<T>(out + o00, u);
    store8<T>(out + o00 + Cp, z);
    store8<T>(out + o00 + static_cast<int64_t>(2 * W) * Cp, z);
    store8<T>(out + o00 + static_cast<int64_t>(2 * W) * Cp + Cp, z);
  }
}

// One thread per channel.  training: every sample's row of (sum, sum of squares) becomes the batch average of the rows,
// so mtbc_in_apply normalises with the batch statistics; running_mean / running_var get nn.BatchNorm2d's update
// (momentum, unbiased variance) and num_batches_tracked advances.  eval: the rows are synthesised from the running
// statistics (sum = mean * HW, sum of squares = (var + mean^2) * HW).
__global__ void bn_pool_fwd_kernel(float* __restrict__ ssum, float* __restrict__ ssq, int N, int Cp, int C, float HW,
                                   int training, float momentum, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, long long* __restrict__ nbt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cp) return;
  float s, q;
  if (training) {
    float ts = 0.f, tq = 0.f;
    for (int n = 0; n < N; ++n) { ts += ssum[static_cast<int64_t>(n) * Cp + c]; tq += ssq[static_cast<int64_t>(n) * Cp + c]; }
    s = ts / static_cast<float>(N);
    q = tq / static_cast<float>(N);
    if (c < C && running_mean != nullptr) {
      const float M = static_cast<float>(N) * HW;
      const float mean = ts / M;
      const float var = fmaxf(tq / M - mean * mean, 0.f);
      const float unbiased = M > 1.f ? var * M / (M - 1.f) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
      if (c == 0 && nbt != nullptr) nbt[0] += 1;
    }
  } else {
    const float m = c < C ? running_mean[c] : 0.f, v = c < C ? running_var[c] : 1.f;
    s = m * HW;
    q = (v + m * m) * HW;
  }
  for (int n = 0; n < N; ++n) { ssum[static_cast<int64_t>(n) * Cp + c] = s; ssq[static_cast<int64_t>(n) * Cp + c] = q; }
}

// backward: the two per-(sample, channel) sums of the InstanceNorm backward, pooled over the batch.  In eval mode the
// statistics are constants: both sums vanish from the gradient (dx = rstd * gamma * g).
__global__ void bn_pool_bwd_kernel(float* __restrict__ s1, float* __restrict__ s2, int N, int Cp, int C, int training,
                                   float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cp) return;
  float a = 0.f, b = 0.f;
  for (int n = 0; n < N; ++n) { a += s1[static_cast<int64_t>(n) * Cp + c]; b += s2[static_cast<int64_t>(n) * Cp + c]; }
  if (c < C && dgamma != nullptr) { dgamma[c] += b; dbeta[c] += a; }   // affine gradients: the un-pooled totals
  if (training) {
    a /= static_cast<float>(N);
    b /= static_cast<float>(N);
  } else {
    a = 0.f;
    b = 0.f;
  }
  for (int n = 0; n < N; ++n) { s1[static_cast<int64_t>(n) * Cp + c] = a; s2[static_cast<int64_t>(n) * Cp + c] = b; }
}

}  // namespace mtbc

using namespace mtbc;
#define ST(s) static_cast<cudaStream_t>(s)
#define TP(p) static_cast<T*>(p)
#define CTP(p) static_cast<const T*>(p)
static int grid_for8(int64_t n8) {
  int64_t g = (n8 + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

extern "C" {

int mtbc_add(const void* a, const void* b, void* out, int64_t n, void* stream) {
  if (n % 8) return set_error(MTBC_ERR_INVALID, "add: element count %% 8");
  MTBC_DISPATCH_ACT((add_kernel<T><<<grid_for8(n / 8), 256, 0, ST(stream)>>>(CTP(a), CTP(b), TP(out), n / 8)));
  return check_launch("add");
}
int mtbc_accumulate(const void* src, void* dst, int64_t n, int32_t accumulate, void* stream) {
  if (n % 8) return set_error(MTBC_ERR_INVALID, "accumulate: element count %% 8");
  MTBC_DISPATCH_ACT((accumulate_kernel<T><<<grid_for8(n / 8), 256, 0, ST(stream)>>>(CTP(src), TP(dst), n / 8, accumulate)));
  return check_launch("accumulate");
}
int mtbc_dropout_fwd(const void* x, void* out, uint8_t* mask, int64_t n, float p, uint64_t seed, const int32_t* counter,
                     int32_t layer, int32_t external_mask, void* stream) {
  if (n % 8 || !(p >= 0.f && p < 1.f)) return set_error(MTBC_ERR_INVALID, "dropout_fwd: n %% 8 != 0 or p outside [0, 1)");
  const uint32_t thresh = static_cast<uint32_t>(static_cast<double>(p) * 4294967296.0);
  MTBC_DISPATCH_ACT((dropout_fwd_kernel<T><<<grid_for8(n / 8), 256, 0, ST(stream)>>>(CTP(x), TP(out), mask, n / 8, 1.f / (1.f - p),
                                                                                     thresh, seed, counter, static_cast<uint32_t>(layer),
                                                                                     external_mask)));
  return check_launch("dropout_fwd");
}
int mtbc_dropout_bwd(const void* g, const uint8_t* mask, void* out, int64_t n, float p, int32_t accumulate, void* stream) {
  if (n % 8 || !(p >= 0.f && p < 1.f)) return set_error(MTBC_ERR_INVALID, "dropout_bwd: n %% 8 != 0 or p outside [0, 1)");
  MTBC_DISPATCH_ACT((dropout_bwd_kernel<T><<<grid_for8(n / 8), 256, 0, ST(stream)>>>(CTP(g), mask, TP(out), n / 8, 1.f / (1.f - p), accumulate)));
  return check_launch("dropout_bwd");
}
int mtbc_zero_stuff2(const void* dy, int32_t N, int32_t H, int32_t W, int32_t Cp, void* out, void* stream) {
  if (Cp % 8) return set_error(MTBC_ERR_INVALID, "zero_stuff2: Cp %% 8");
  const int64_t total = static_cast<int64_t>(N) * H * W * (Cp / 8);
  MTBC_DISPATCH_ACT((zero_stuff2_kernel<T><<<grid_for8(total), 256, 0, ST(stream)>>>(CTP(dy), H, W, Cp, TP(out), total)));
  return check_launch("zero_stuff2");
}
int mtbc_bn_pool_fwd(float* stat_sum, float* stat_sq, int32_t N, int32_t Cp, int32_t C, int64_t HW, int32_t training,
                     float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked, void* stream) {
  if (!training && (!running_mean || !running_var)) return set_error(MTBC_ERR_INVALID, "bn_pool_fwd: eval mode needs running statistics");
  bn_pool_fwd_kernel<<<cdiv(Cp, 128), 128, 0, ST(stream)>>>(stat_sum, stat_sq, N, Cp, C, static_cast<float>(HW), training, momentum,
                                                          running_mean, running_var,
                                                          reinterpret_cast<long long*>(num_batches_tracked));
  return check_launch("bn_pool_fwd");
}
int mtbc_bn_pool_bwd(float* s1, float* s2, int32_t N, int32_t Cp, int32_t C, int32_t training, float* dgamma,
                     float* dbeta, void* stream) {
  bn_pool_bwd_kernel<<<cdiv(Cp, 128), 128, 0, ST(stream)>>>(s1, s2, N, Cp, C, training, dgamma, dbeta);
  return check_launch("bn_pool_bwd");
}

}  // extern "C"
