// Dice + focal multi-task objective, prediction refinement, hard-Dice counts and the fused Adam step.
#include "ptx.cuh"
#include "internal.h"

namespace mtbc {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// ------------------------------------------------------------------------------------------------ Dice
__global__ void __launch_bounds__(256) dice_sums_kernel(const float* __restrict__ logits,
                                                        const float* __restrict__ target, int64_t HW,
                                                        float* __restrict__ sums) {
  pdl_trigger();   // programmatic dependent launch (ptx.cuh)
  pdl_wait();

  const int n = blockIdx.y;
  const float4* x4 = reinterpret_cast<const float4*>(logits + static_cast<int64_t>(n) * HW);
  const float4* t4 = reinterpret_cast<const float4*>(target + static_cast<int64_t>(n) * HW);
  // 128-bit loads only when every sample's base address is 16-byte aligned (HW % 4 == 0 and aligned tensors); e.g.
  // 125 x 125 masks take the scalar loop below for the whole plane instead of faulting on a misaligned float4
  const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(target)) & 15) == 0;
  const int64_t nv = vec ? HW / 4 : 0;
  float sI = 0.f, sG = 0.f, sP = 0.f;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < nv; i += gridDim.x * 256ll) {
    const float4 x = x4[i], t = t4[i];
    const float p0 = sigmoidf_(x.x), p1 = sigmoidf_(x.y), p2 = sigmoidf_(x.z), p3 = sigmoidf_(x.w);
    sI += t.x * p0 + t.y * p1 + t.z * p2 + t.w * p3;
    sG += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
    sP += p0 * p0 + p1 * p1 + p2 * p2 + p3 * p3;
  }
  if (!vec) {  // scalar path, grid-strided like the vector one
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < HW; i += gridDim.x * 256ll) {
      const float x = logits[static_cast<int64_t>(n) * HW + i], t = target[static_cast<int64_t>(n) * HW + i];
      const float p = sigmoidf_(x);
      sI += t * p; sG += t * t; sP += p * p;
    }
  }
  __shared__ float s_red[3][8];
  sI = warp_sum(sI); sG = warp_sum(sG); sP = warp_sum(sP);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_red[0][warp] = sI; s_red[1][warp] = sG; s_red[2][warp] = sP; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_red[threadIdx.x][w];
    atomicAdd(sums + static_cast<int64_t>(n) * 3 + threadIdx.x, t);
  }
}

__global__ void dice_finalize_kernel(const float* __restrict__ sums, int N, float* __restrict__ loss) {
  pdl_trigger();   // programmatic dependent launch (ptx.cuh)
  pdl_wait();

  float s = 0.f;
  for (int n = threadIdx.x; n < N; n += 32) {
    const float I = sums[n * 3], G = sums[n * 3 + 1], P = sums[n * 3 + 2];
    s += 1.f - (2.f * I + 1.f) / (G + P + 1.f);
  }
  s = warp_sum(s);
  if (threadIdx.x == 0) loss[0] = s / static_cast<float>(N);
}

__global__ void __launch_bounds__(256) dice_bwd_kernel(const float* __restrict__ logits,
                                                       const float* __restrict__ target, int64_t HW, int N,
                                                       const float* __restrict__ sums, const float* __restrict__ gscale,
                                                       float gmul, float* __restrict__ dlogits) {
  pdl_trigger();   // programmatic dependent launch (ptx.cuh)
  pdl_wait();

  const int n = blockIdx.y;
  const float I = sums[n * 3], G = sums[n * 3 + 1], P = sums[n * 3 + 2];
  const float D = G + P + 1.f, num = 2.f * I + 1.f;
  const float gs = (gscale ? gscale[0] : 1.f) * gmul / static_cast<float>(N);
  const float c1 = -2.f * gs / D;          // coefficient of t
  const float c2 = 2.f * gs * num / (D * D);  // coefficient of p
  const float* x = logits + static_cast<int64_t>(n) * HW;
  const float* t = target + static_cast<int64_t>(n) * HW;
  float* d = dlogits + static_cast<int64_t>(n) * HW;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < HW; i += gridDim.x * 256ll) {
    const float p = sigmoidf_(x[i]);
    d[i] = (c1 * t[i] + c2 * p) * p * (1.f - p);
  }
}

// ------------------------------------------------------------------------------------------------ focal
__device__ __forceinline__ void focal_sample(const float* x, const float* t, int K, float alpha, float gamma,
                                             float* fl, float* dfl_dce, float* lse_out, float* tsum_out) {
  float mx = x[0];
  for (int k = 1; k < K; ++k) mx = fmaxf(mx, x[k]);
  float se = 0.f;
  for (int k = 0; k < K; ++k) se += expf(x[k] - mx);
  const float lse = mx + logf(se);
  float ce = 0.f, ts = 0.f;
  for (int k = 0; k < K; ++k) { ce -= t[k] * (x[k] - lse); ts += t[k]; }
  const float pt = expf(-ce);
  const float om = 1.f - pt;
  const float pw = powf(fmaxf(om, 0.f), gamma);
  *fl = alpha * pw * ce;
  // d/dce [ (1-pt)^g * ce ] = (1-pt)^g + g*(1-pt)^(g-1)*pt*ce
  const float pwm1 = (gamma == 2.f) ? om : powf(fmaxf(om, 0.f), gamma - 1.f);
  *dfl_dce = alpha * (pw + gamma * pwm1 * pt * ce);
  *lse_out = lse;
  *tsum_out = ts;
}

__global__ void focal_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ target, int N, int K,
                                 float alpha, float gamma, float* __restrict__ loss) {
  pdl_trigger();   // programmatic dependent launch (ptx.cuh)
  pdl_wait();

  float s = 0.f;
  for (int n = threadIdx.x; n < N; n += 32) {
    float fl, d, lse, ts;
    focal_sample(logits + static_cast<int64_t>(n) * K, target + static_cast<int64_t>(n) * K, K, alpha, gamma, &fl, &d,
                 &lse, &ts);
    s += fl;
  }
  s = warp_sum(s);
  if (threadIdx.x == 0) loss[0] = s / static_cast<float>(N);
}
__global__ void focal_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ target, int N, int K,
                                 float alpha, float gamma, const float* __restrict__ gscale, float gmul,
                                 float* __restrict__ dlogits) {
  pdl_trigger();   // programmatic dependent launch (ptx.cuh)
  pdl_wait();

  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* x = logits + static_cast<int64_t>(n) * K;
  const float* t = target + static_cast<int64_t>(n) * K;
  float fl, d, lse, ts;
  focal_sample(x, t, K, alpha, gamma, &fl, &d, &lse, &ts);
  const float gs = (gscale ? gscale[0] : 1.f) * gmul / static_cast<float>(N);
  for (int k = 0; k < K; ++k) {
    const float sm = expf(x[k] - lse);
    dlogits[static_cast<int64_t>(n) * K + k] = gs * d * (sm * ts - t[k]);
  }
}

__global__ void multitask_loss_kernel(const float* __restrict__ dice, int nheads, int inv_w,
                                      const float* __restrict__ focal, float alpha_mix, float* __restrict__ out) {
  pdl_trigger();   // programmatic dependent launch (ptx.cuh)
  pdl_wait();

  float seg = 0.f;
  for (int j = 0; j < nheads; ++j) seg += inv_w ? dice[j] / static_cast<float>(j + 1) : dice[j];
  const float cls = focal[0];
  out[0] = alpha_mix * seg + (1.f - alpha_mix) * cls;
  out[1] = seg;
  out[2] = cls;
  out[3] = (isnan(seg) || isnan(cls)) ? 1.f : 0.f;
}

// ------------------------------------------------------------------------------------------------ refinement
__global__ void __launch_bounds__(256) refine_count_kernel(const float* __restrict__ mask_logits, int64_t HW,
                                                           int32_t* __restrict__ count) {
  pdl_trigger();   // programmatic dependent launch (ptx.cuh)
  pdl_wait();

  const int n = blockIdx.y;
  const float* x = mask_logits + static_cast<int64_t>(n) * HW;
  int c = 0;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < HW; i += gridDim.x * 256ll) c += x[i] > 0.f ? 1 : 0;
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count + n, c);
}
__global__ void __launch_bounds__(256) refine_apply_kernel(const float* __restrict__ mask_logits,
                                                           const float* __restrict__ class_logits, int64_t HW, int K,
                                                           int normal_id, int seg_by_class, int class_by_seg,
                                                           int pixel_threshold, const int32_t* __restrict__ count,
                                                           uint8_t* __restrict__ mask_out,
                                                           int32_t* __restrict__ class_out) {
  pdl_trigger();   // programmatic dependent launch (ptx.cuh)
  pdl_wait();

  const int n = blockIdx.y;
  int cls = 0;
  float best = class_logits[static_cast<int64_t>(n) * K];
  for (int k = 1; k < K; ++k) {
    const float v = class_logits[static_cast<int64_t>(n) * K + k];
    if (v > best) { best = v; cls = k; }
  }
  const int cnt = count[n];
  const bool kill = (pixel_threshold > 0 && cnt <= pixel_threshold) || (seg_by_class && cls == normal_id);
  if (blockIdx.x == 0 && threadIdx.x == 0) class_out[n] = (class_by_seg && cnt == 0) ? normal_id : cls;
  const float* x = mask_logits + static_cast<int64_t>(n) * HW;
  uint8_t* m = mask_out + static_cast<int64_t>(n) * HW;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < HW; i += gridDim.x * 256ll)
    m[i] = (!kill && x[i] > 0.f) ? 1 : 0;
}

__global__ void __launch_bounds__(256) hard_dice_counts_kernel(const float* __restrict__ logits,
                                                               const float* __restrict__ target, int64_t n,
                                                               unsigned long long* __restrict__ out) {
  int tp = 0, fp = 0, fn = 0;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
    const bool s = logits[i] > 0.f, g = target[i] != 0.f;
    tp += (s && g); fp += (s && !g); fn += (!s && g);
  }
  tp = __reduce_add_sync(0xffffffffu, tp);
  fp = __reduce_add_sync(0xffffffffu, fp);
  fn = __reduce_add_sync(0xffffffffu, fn);
  if ((threadIdx.x & 31) == 0) {
    if (tp) atomicAdd(out, static_cast<unsigned long long>(tp));
    if (fp) atomicAdd(out + 1, static_cast<unsigned long long>(fp));
    if (fn) atomicAdd(out + 2, static_cast<unsigned long long>(fn));
  }
}

// Per-sample confusion counts of a refined uint8 mask against a {0,1} fp32 target: out[n][4] = tp, fp, fn, tn.
// grid = (blocks per sample, N); 16 mask bytes + 4 x 16 target bytes per thread and iteration.
__global__ void __launch_bounds__(256) confusion_counts_kernel(const uint8_t* __restrict__ mask,
                                                               const float* __restrict__ target, int64_t HW,
                                                               unsigned long long* __restrict__ out) {
  const int n = blockIdx.y;
  const uint8_t* m = mask + n * HW;
  const float* t = target + n * HW;
  int tp = 0, fp = 0, fn = 0, tn = 0;
  const bool vec = (HW % 16 == 0) && ((reinterpret_cast<uintptr_t>(m) & 15) == 0) && ((reinterpret_cast<uintptr_t>(t) & 15) == 0);
  if (vec) {
    for (int64_t i = (blockIdx.x * 256ll + threadIdx.x) * 16; i < HW; i += gridDim.x * 256ll * 16) {
      const uint4 mv = *reinterpret_cast<const uint4*>(m + i);
      const uint32_t mw[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 tv = *reinterpret_cast<const float4*>(t + i + 4 * q);
        const float tt[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const bool s = ((mw[q] >> (8 * b)) & 0xffu) != 0, g = tt[b] != 0.f;
          tp += (s && g); fp += (s && !g); fn += (!s && g); tn += (!s && !g);
        }
      }
    }
  } else {
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < HW; i += gridDim.x * 256ll) {
      const bool s = m[i] != 0, g = t[i] != 0.f;
      tp += (s && g); fp += (s && !g); fn += (!s && g); tn += (!s && !g);
    }
  }
  tp = __reduce_add_sync(0xffffffffu, tp);
  fp = __reduce_add_sync(0xffffffffu, fp);
  fn = __reduce_add_sync(0xffffffffu, fn);
  tn = __reduce_add_sync(0xffffffffu, tn);
  if ((threadIdx.x & 31) == 0) {
    unsigned long long* o = out + 4 * n;
    if (tp) atomicAdd(o, static_cast<unsigned long long>(tp));
    if (fp) atomicAdd(o + 1, static_cast<unsigned long long>(fp));
    if (fn) atomicAdd(o + 2, static_cast<unsigned long long>(fn));
    if (tn) atomicAdd(o + 3, static_cast<unsigned long long>(tn));
  }
}

// One step of the epoch bookkeeping the reference does with .item() calls (training_multitask.py:99,108-109,146-152):
//   acc[0..2] += total / seg / cls loss, acc[3] += NaN flag, acc[4] += hard Dice of this batch
//   (metrics.py:255-267: 1 or 0 when the batch has no foreground), acc[5] += 1; confusion[gt][pred] += 1 per sample
//   (pred = argmax of the class logits, gt = argmax of the one-hot label; first maximum on ties).
// counts (tp, fp, fn of mtbc_hard_dice_counts) is zeroed afterwards so the next step accumulates from 0.
__global__ void metrics_accumulate_kernel(const float* __restrict__ loss4, unsigned long long* __restrict__ counts,
                                          const float* __restrict__ class_logits, const float* __restrict__ onehot,
                                          int B, int K, double* __restrict__ acc,
                                          unsigned long long* __restrict__ confusion) {
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int p = 0, g = 0;
    for (int k = 1; k < K; ++k) {
      if (class_logits[b * K + k] > class_logits[b * K + p]) p = k;
      if (onehot[b * K + k] > onehot[b * K + g]) g = k;
    }
    atomicAdd(confusion + g * K + p, 1ull);
  }
  if (threadIdx.x == 0) {
    acc[0] += static_cast<double>(loss4[0]);
    acc[1] += static_cast<double>(loss4[1]);
    acc[2] += static_cast<double>(loss4[2]);
    acc[3] += static_cast<double>(loss4[3]);
    const double tp = static_cast<double>(counts[0]), fp = static_cast<double>(counts[1]), fn = static_cast<double>(counts[2]);
    double dice;
    if (tp + fn == 0.0) dice = (tp + fp == 0.0) ? 1.0 : 0.0;
    else dice = 2.0 * tp / (2.0 * tp + fp + fn);
    acc[4] += dice;
    acc[5] += 1.0;
    counts[0] = 0; counts[1] = 0; counts[2] = 0;
  }
}

// ------------------------------------------------------------------------------------------------ Adam
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
                                                   float b1, float b2, float eps, float gscale, float bc1, float bc2s) {
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
    const float gg = g[i] * gscale;
    const float mm = b1 * m[i] + (1.f - b1) * gg;
    const float vv = b2 * v[i] + (1.f - b2) * gg * gg;
    m[i] = mm;
    v[i] = vv;
    const float denom = sqrtf(vv) / bc2s + eps;
    p[i] -= (lr / bc1) * (mm / denom);
  }
}

// Graph-friendly variant: step count and learning rate live in device memory so a captured launch stays valid.
__global__ void __launch_bounds__(256) adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                       const float* __restrict__ lr_dev, float b1, float b2, float eps,
                                                       float gscale, const int32_t* __restrict__ step_dev) {
  pdl_trigger();   // programmatic dependent launch (ptx.cuh)
  pdl_wait();

  const float step = static_cast<float>(step_dev[0]);
  const float bc1 = 1.f - powf(b1, step);
  const float bc2s = sqrtf(1.f - powf(b2, step));
  const float lr = lr_dev[0];
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
    const float gg = g[i] * gscale;
    const float mm = b1 * m[i] + (1.f - b1) * gg;
    const float vv = b2 * v[i] + (1.f - b2) * gg * gg;
    m[i] = mm;
    v[i] = vv;
    const float denom = sqrtf(vv) / bc2s + eps;
    p[i] -= (lr / bc1) * (mm / denom);
  }
}
__global__ void increment_i32_kernel(int32_t* p) { p[0] += 1; }

// Hausdorff distance exactly as the reference computes it (utils/metrics.py:236-252): scipy's directed_hausdorff on the
// two (H, W) boolean images, i.e. each image ROW is one point of {0,1}^W, so the distance between two rows is
// sqrt(Hamming distance).  Integer work: rows are packed to bit words in shared memory (one ballot per 32 pixels),
// every thread owns rows of u and scans all rows of v with XOR + popcount.  grid = (2 directions, N);
// out[n] = {max_i min_j d2(seg_i, gt_j), max_i min_j d2(gt_i, seg_j), #seg pixels, #gt pixels}.
__global__ void __launch_bounds__(256) row_hausdorff_kernel(const uint8_t* __restrict__ mask,
                                                            const float* __restrict__ target, int H, int W,
                                                            int* __restrict__ out) {
  extern __shared__ uint32_t s_bits[];  // [2][H][Wd]
  __shared__ int s_max, s_cnt[2];
  const int n = blockIdx.y, dir = blockIdx.x;
  const int Wd = (W + 31) / 32;
  const int64_t HW = static_cast<int64_t>(H) * W;
  const uint8_t* m = mask + n * HW;
  const float* t = target + n * HW;
  uint32_t* sb = s_bits;
  uint32_t* gb = s_bits + H * Wd;
  if (threadIdx.x == 0) { s_max = 0; s_cnt[0] = 0; s_cnt[1] = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int c0 = 0, c1 = 0;
  for (int idx = warp; idx < H * Wd; idx += 8) {
    const int r = idx / Wd, w = idx - r * Wd;
    const int col = w * 32 + lane;
    bool sv = false, gv = false;
    if (col < W) {
      sv = m[static_cast<int64_t>(r) * W + col] != 0;
      gv = t[static_cast<int64_t>(r) * W + col] != 0.f;
    }
    const uint32_t sw = __ballot_sync(0xffffffffu, sv), gw = __ballot_sync(0xffffffffu, gv);
    if (lane == 0) { sb[idx] = sw; gb[idx] = gw; c0 += __popc(sw); c1 += __popc(gw); }
  }
  if (lane == 0) { atomicAdd(&s_cnt[0], c0); atomicAdd(&s_cnt[1], c1); }
  __syncthreads();
  const uint32_t* u = dir == 0 ? sb : gb;
  const uint32_t* v = dir == 0 ? gb : sb;
  int worst = 0;
  for (int i = threadIdx.x; i < H; i += 256) {
    uint32_t ur[32];
#pragma unroll
    for (int w = 0; w < 32; ++w) ur[w] = w < Wd ? u[i * Wd + w] : 0u;
    int best = 0x7fffffff;
    for (int j = 0; j < H; ++j) {
      int d = 0;
#pragma unroll
      for (int w = 0; w < 32; ++w)
        if (w < Wd) d += __popc(ur[w] ^ v[j * Wd + w]);
      best = min(best, d);
    }
    worst = max(worst, best);
  }
  atomicMax(&s_max, worst);
  __syncthreads();
  if (threadIdx.x == 0) {
    out[n * 4 + dir] = s_max;
    if (dir == 0) { out[n * 4 + 2] = s_cnt[0]; out[n * 4 + 3] = s_cnt[1]; }
  }
}

}  // namespace mtbc

using namespace mtbc;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" {

int mtbc_dice_sums(const float* logits, const float* target, int32_t N, int64_t HW, float* sums, void* stream) {
  int gx = cdiv(HW / 4 + 1, 256 * 4); if (gx > (148 * 8) / N + 1) gx = (148 * 8) / N + 1; if (gx < 1) gx = 1;
  launch_pdl(dice_sums_kernel, dim3(gx, N), dim3(256), 0, ST(stream), logits, target, HW, sums);
  return check_launch("dice_sums");
}
int mtbc_dice_finalize(const float* sums, int32_t N, float* loss, void* stream) {
  launch_pdl(dice_finalize_kernel, dim3(1), dim3(32), 0, ST(stream), sums, N, loss);
  return check_launch("dice_finalize");
}
int mtbc_dice_bwd(const float* logits, const float* target, int32_t N, int64_t HW, const float* sums,
                  const float* gscale, float gmul, float* dlogits, void* stream) {
  int gx = cdiv(HW, 256 * 4); if (gx > (148 * 8) / N + 1) gx = (148 * 8) / N + 1; if (gx < 1) gx = 1;
  launch_pdl(dice_bwd_kernel, dim3(gx, N), dim3(256), 0, ST(stream), logits, target, HW, N, sums, gscale, gmul, dlogits);
  return check_launch("dice_bwd");
}
int mtbc_focal_fwd(const float* logits, const float* target, int32_t N, int32_t K, float alpha, float gamma,
                   float* loss, void* stream) {
  launch_pdl(focal_fwd_kernel, dim3(1), dim3(32), 0, ST(stream), logits, target, N, K, alpha, gamma, loss);
  return check_launch("focal_fwd");
}
int mtbc_focal_bwd(const float* logits, const float* target, int32_t N, int32_t K, float alpha, float gamma,
                   const float* gscale, float gmul, float* dlogits, void* stream) {
  launch_pdl(focal_bwd_kernel, dim3(cdiv(N, 64)), dim3(64), 0, ST(stream), logits, target, N, K, alpha, gamma, gscale, gmul, dlogits);
  return check_launch("focal_bwd");
}
int mtbc_multitask_loss(const float* dice_losses, int32_t nheads, int32_t inversely_weighted, const float* focal,
                        float alpha_mix, float* out, void* stream) {
  launch_pdl(multitask_loss_kernel, dim3(1), dim3(1), 0, ST(stream), dice_losses, nheads, inversely_weighted, focal, alpha_mix, out);
  return check_launch("multitask_loss");
}

int mtbc_refine_predictions(const float* mask_logits, const float* class_logits, int32_t N, int64_t HW, int32_t K,
                            int32_t normal_id, int32_t seg_by_class, int32_t class_by_seg, int32_t pixel_threshold,
                            uint8_t* mask_out, int32_t* class_out, int32_t* count_out, void* stream) {
  cudaError_t e = cudaMemsetAsync(count_out, 0, N * sizeof(int32_t), ST(stream));
  if (e != cudaSuccess) return set_error(MTBC_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
  int gx = cdiv(HW, 256 * 8); if (gx > (148 * 8) / N + 1) gx = (148 * 8) / N + 1; if (gx < 1) gx = 1;
  launch_pdl(refine_count_kernel, dim3(gx, N), dim3(256), 0, ST(stream), mask_logits, HW, count_out);
  int rc = check_launch("refine_count");
  if (rc) return rc;
  launch_pdl(refine_apply_kernel, dim3(gx, N), dim3(256), 0, ST(stream), mask_logits, class_logits, HW, K, normal_id, seg_by_class,
                                                          class_by_seg, pixel_threshold, count_out, mask_out, class_out);
  return check_launch("refine_apply");
}
int mtbc_hard_dice_counts(const float* logits, const float* target, int64_t n, long long* out, void* stream) {
  int g = cdiv(n, 256 * 8); if (g > 148 * 8) g = 148 * 8; if (g < 1) g = 1;
  hard_dice_counts_kernel<<<g, 256, 0, ST(stream)>>>(logits, target, n, reinterpret_cast<unsigned long long*>(out));
  return check_launch("hard_dice_counts");
}

int mtbc_confusion_counts(const uint8_t* mask, const float* target, int32_t N, int64_t HW, long long* out, void* stream) {
  if (N <= 0 || HW <= 0) return 0;
  int gx = cdiv(HW, 256 * 16 * 4); if (gx < 1) gx = 1; if (gx > 64) gx = 64;
  confusion_counts_kernel<<<dim3(gx, N), 256, 0, ST(stream)>>>(mask, target, HW, reinterpret_cast<unsigned long long*>(out));
  return check_launch("confusion_counts");
}
int mtbc_row_hausdorff(const uint8_t* mask, const float* target, int32_t N, int32_t H, int32_t W, int32_t* out,
                       void* stream) {
  if (N <= 0) return 0;
  if (H <= 0 || W <= 0 || W > 1024) return set_error(MTBC_ERR_INVALID, "row_hausdorff: W must be in [1, 1024]");
  const size_t smem = 2 * static_cast<size_t>(H) * ((W + 31) / 32) * sizeof(uint32_t);
  if (smem > 200 * 1024) return set_error(MTBC_ERR_INVALID, "row_hausdorff: the packed images do not fit shared memory");
  { static bool done = false; if (smem > 48 * 1024 && !done) { cudaFuncSetAttribute(row_hausdorff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); done = true; } }
  row_hausdorff_kernel<<<dim3(2, N), 256, smem, ST(stream)>>>(mask, target, H, W, out);
  return check_launch("row_hausdorff");
}
int mtbc_metrics_accumulate(const float* loss4, long long* counts, const float* class_logits, const float* onehot,
                            int32_t B, int32_t K, double* acc, long long* confusion, void* stream) {
  metrics_accumulate_kernel<<<1, 128, 0, ST(stream)>>>(loss4, reinterpret_cast<unsigned long long*>(counts), class_logits,
                                                       onehot, B, K, acc, reinterpret_cast<unsigned long long*>(confusion));
  return check_launch("metrics_accumulate");
}

int mtbc_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                   float beta1, float beta2, float eps, float grad_scale, int32_t step, void* stream) {
  if (n <= 0) return 0;
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
  const float bc2s = sqrtf(1.f - powf(beta2, static_cast<float>(step)));
  int g = cdiv(n, 256 * 4); if (g > 148 * 8) g = 148 * 8;
  adam_kernel<<<g, 256, 0, ST(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, grad_scale, bc1, bc2s);
  return check_launch("adam_step");
}

int mtbc_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                       const float* lr_dev, float beta1, float beta2, float eps, float grad_scale,
                       const int32_t* step_dev, void* stream) {
  if (n <= 0) return 0;
  int g = cdiv(n, 256 * 4); if (g > 148 * 8) g = 148 * 8;
  launch_pdl(adam_dev_kernel, dim3(g), dim3(256), 0, ST(stream), param, grad, exp_avg, exp_avg_sq, n, lr_dev, beta1, beta2, eps, grad_scale, step_dev);
  return check_launch("adam_step_dev");
}
int mtbc_increment_i32(int32_t* p, void* stream) {
  increment_i32_kernel<<<1, 1, 0, ST(stream)>>>(p);
  return check_launch("increment_i32");
}

}  // extern "C"
