// Parameter (fp32, PyTorch layout) <-> GEMM operand (bf16, padded, K-major) packing, the Cin<=4 first layer, and small
// utilities.  All memory bound / tiny.
#include "ptx.cuh"
#include "internal.h"
#include "act_io.cuh"
#include <vector>

namespace mtbc {

// ------------------------------------------------------------------------------------------------ conv weights
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int ksz, int c_begin,
                                        int c_count, __nv_bfloat16* __restrict__ wf, int wf_rows, int wf_ld, int wf_k0,
                                        __nv_bfloat16* __restrict__ wd, int wd_rows, int wd_ld) {
  const int taps = ksz * ksz;
  const int64_t total = static_cast<int64_t>(Cout) * c_count * taps;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int tap = static_cast<int>(i % taps);
    const int cl = static_cast<int>((i / taps) % c_count);
    const int co = static_cast<int>(i / (static_cast<int64_t>(taps) * c_count));
    const float v = w[(static_cast<int64_t>(co) * Cin + c_begin + cl) * taps + tap];
    const __nv_bfloat16 b = __float2bfloat16(v);
    if (wf) wf[(static_cast<int64_t>(tap) * wf_rows + co) * wf_ld + wf_k0 + cl] = b;
    if (wd) wd[(static_cast<int64_t>(taps - 1 - tap) * wd_rows + cl) * wd_ld + co] = b;
  }
}

__global__ void pack_convT_weight_kernel(const float* __restrict__ w, int Cin, int Cout, int k, int cp,
                                         __nv_bfloat16* __restrict__ wf, int wf_ld, __nv_bfloat16* __restrict__ wd,
                                         int wd_rows, int wd_ld) {
  const int kk = k * k;
  const int64_t total = static_cast<int64_t>(Cin) * Cout * kk;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int q = static_cast<int>(i % kk);
    const int co = static_cast<int>((i / kk) % Cout);
    const int ci = static_cast<int>(i / (static_cast<int64_t>(kk) * Cout));
    const __nv_bfloat16 b = __float2bfloat16(w[i]);
    if (wf) wf[(static_cast<int64_t>(q) * cp + co) * wf_ld + ci] = b;
    if (wd) wd[(static_cast<int64_t>(q) * wd_rows + ci) * wd_ld + co] = b;
  }
}

__global__ void unpack_conv_wgrad_kernel(const float* __restrict__ acc, int rows, int ld, int k0,
                                         float* __restrict__ grad, int Cout, int Cin, int ksz, int c_begin,
                                         int c_count, int add) {
  const int taps = ksz * ksz;
  const int64_t total = static_cast<int64_t>(Cout) * c_count * taps;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int tap = static_cast<int>(i % taps);
    const int cl = static_cast<int>((i / taps) % c_count);
    const int co = static_cast<int>(i / (static_cast<int64_t>(taps) * c_count));
    const float v = acc[(static_cast<int64_t>(tap) * rows + co) * ld + k0 + cl];
    float* g = grad + (static_cast<int64_t>(co) * Cin + c_begin + cl) * taps + tap;
    *g = add ? (*g + v) : v;
  }
}

__global__ void unpack_convT_wgrad_kernel(const float* __restrict__ acc, int rows, int ld, float* __restrict__ grad,
                                          int Cin, int Cout, int k, int add) {
  const int kk = k * k;
  const int cp = rows / kk;
  const int64_t total = static_cast<int64_t>(Cin) * Cout * kk;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int q = static_cast<int>(i % kk);
    const int co = static_cast<int>((i / kk) % Cout);
    const int ci = static_cast<int>(i / (static_cast<int64_t>(kk) * Cout));
    const float v = acc[(static_cast<int64_t>(q) * cp + co) * ld + ci];
    grad[i] = add ? (grad[i] + v) : v;
  }
}


// ------------------------------------------------------------------------------------------------ batched jobs
// One launch for all the small parameter-side jobs of a step (padded copies of bias/gamma/beta vectors, weight packs,
// weight-gradient unpacks): a U-Net++ step has ~250 of them and as separate launches they cost more in launch gaps
// than in work.  Block b handles kChunk consecutive elements of job chunk_job[b].
constexpr int kChunk = 2048;

struct ParamJobDev {
  int32_t kind;
  int32_t i[11];
  const void* src;
  void* dst0;
  void* dst1;
};

// Packed-operand element types: bf16 (product path), or fp32 holding TF32 values (parity mode): part 1 = the weight
// rounded to nearest TF32 (cvt.rna), part 2 = the TF32 rounding of what that left over (second term of the 3xTF32 split).
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
template <typename WT>
__device__ __forceinline__ void put_w(WT* p, float v, int part);
template <>
__device__ __forceinline__ void put_w<__nv_bfloat16>(__nv_bfloat16* p, float v, int) { *p = __float2bfloat16(v); }
template <>
__device__ __forceinline__ void put_w<float>(float* p, float v, int part) {
  const float hi = tf32_rn(v);
  *p = part == 2 ? tf32_rn(v - hi) : hi;
}

// 32 (co) x 32 (ci) x TAPS weight tile through shared memory.  TAPS is a compile-time constant (1 or 9 on this path)
// so the index divisions become multiply-shifts: with run-time `taps` the two divisions per element made the 15 M
// parameter pack an instruction-bound 134 us launch (0.9 TB/s).
template <int TAPS, typename WT>
__device__ __forceinline__ void pack_conv_tile(const ParamJobDev& j, int local, float (*s_tile)[32 * 9 + 1]) {
  const int Cout = j.i[0], Cin = j.i[1], c_begin = j.i[3], c_count = j.i[4];
  const int tiles_cl = (c_count + 31) >> 5;
  const int co0 = (local / tiles_cl) << 5, cl0 = (local % tiles_cl) << 5;
  constexpr int roww = 32 * TAPS;
  const float* w = static_cast<const float*>(j.src);
  WT* wf = static_cast<WT*>(j.dst0);
  WT* wd = static_cast<WT*>(j.dst1);
  const int part = j.i[10];
  for (int idx = threadIdx.x; idx < 32 * roww; idx += 256) {
    const int r = idx / roww, off = idx - r * roww;
    const int co = co0 + r, cl = cl0 + off / TAPS;
    s_tile[r][off] = (co < Cout && cl < c_count)
                         ? w[(static_cast<int64_t>(co) * Cin + c_begin + cl0) * TAPS + off] : 0.f;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < TAPS * 1024; idx += 256) {
    const int lo = idx & 31, mid = (idx >> 5) & 31, tap = idx >> 10;
    if (wf) {   // lanes run over ci
      const int cl = cl0 + lo, co = co0 + mid;
      if (cl < c_count && co < Cout)
        put_w<WT>(wf + (static_cast<int64_t>(tap) * j.i[5] + co) * j.i[6] + j.i[7] + cl, s_tile[mid][lo * TAPS + tap], part);
    }
    if (wd) {   // lanes run over co
      const int co = co0 + lo, cl = cl0 + mid;
      if (cl < c_count && co < Cout)
        put_w<WT>(wd + (static_cast<int64_t>(TAPS - 1 - tap) * j.i[8] + cl) * j.i[9] + co, s_tile[lo][mid * TAPS + tap], part);
    }
  }
}

template <int TAPS>
__device__ __forceinline__ void unpack_conv_tile(const ParamJobDev& j, int local, float (*s_tile)[32 * 9 + 1]) {
  const int rows = j.i[0], ld = j.i[1], k0 = j.i[2], Cout = j.i[3], Cin = j.i[4], c_begin = j.i[6], c_count = j.i[7],
            add = j.i[8];
  const int tiles_cl = (c_count + 31) >> 5;
  const int co0 = (local / tiles_cl) << 5, cl0 = (local % tiles_cl) << 5;
  constexpr int roww = 32 * TAPS;
  const float* acc = static_cast<const float*>(j.src);
  float* grad = static_cast<float*>(j.dst0);
  for (int idx = threadIdx.x; idx < TAPS * 1024; idx += 256) {
    const int lo = idx & 31, mid = (idx >> 5) & 31, tap = idx >> 10;
    const int cl = cl0 + lo, co = co0 + mid;
    s_tile[mid][lo * TAPS + tap] =
        (cl < c_count && co < Cout) ? acc[(static_cast<int64_t>(tap) * rows + co) * ld + k0 + cl] : 0.f;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 32 * roww; idx += 256) {
    const int r = idx / roww, off = idx - r * roww;
    const int co = co0 + r, cl = cl0 + off / TAPS;
    if (co < Cout && cl < c_count) {
      float* g = grad + (static_cast<int64_t>(co) * Cin + c_begin + cl0) * TAPS + off;
      *g = add ? (*g + s_tile[r][off]) : s_tile[r][off];
    }
  }
}

__global__ void __launch_bounds__(256) param_jobs_kernel(const ParamJobDev* __restrict__ jobs,
                                                         const int32_t* __restrict__ chunk_job,
                                                         const int32_t* __restrict__ chunk_first) {
  __shared__ float s_tile[32][32 * 9 + 1];   // conv weight tiles: [co][ci * taps + tap] (+1: conflict-free columns)
  const ParamJobDev j = jobs[chunk_job[blockIdx.x]];
  const int64_t begin = static_cast<int64_t>(blockIdx.x - chunk_first[blockIdx.x]) * kChunk;
  if (j.kind == MTBC_JOB_COPY_F32) {
    const int64_t n = j.i[0];
    const float* src = static_cast<const float*>(j.src);
    float* dst = static_cast<float*>(j.dst0);
    for (int64_t e = begin + threadIdx.x; e < begin + kChunk && e < n; e += 256) dst[e] = src[e];
  } else if (j.kind == MTBC_JOB_PACK_CONV) {
    // i: Cout, Cin, ksz, c_begin, c_count, wf_rows, wf_ld, wf_k0, wd_rows, wd_ld.  One block = one 32 (co) x 32 (ci)
    // tile with all taps, transposed through shared memory: the parameter is read in rows of 32*taps contiguous
    // floats, wf is written in runs of 32 ci (64 B) and wd in runs of 32 co (64 B).  (Element-wise, the wd writes were
    // 2-byte scatters with a stride of a whole weight row: 0.27 ms per step for 15 M parameters.)
    const int local = blockIdx.x - chunk_first[blockIdx.x];
    if (j.i[10] == 0) {
      switch (j.i[2]) {
        case 1: pack_conv_tile<1, __nv_bfloat16>(j, local, s_tile); break;
        case 2: pack_conv_tile<4, __nv_bfloat16>(j, local, s_tile); break;
        default: pack_conv_tile<9, __nv_bfloat16>(j, local, s_tile); break;
      }
    } else {   // fp32 operands holding TF32 values (i[10] = 1: rounded weight, 2: rounded remainder)
      switch (j.i[2]) {
        case 1: pack_conv_tile<1, float>(j, local, s_tile); break;
        case 2: pack_conv_tile<4, float>(j, local, s_tile); break;
        default: pack_conv_tile<9, float>(j, local, s_tile); break;
      }
    }
  } else if (j.kind == MTBC_JOB_PACK_CONVT) {
    // i: Cin, Cout, k, cp, wf_ld, wd_rows, wd_ld ; element order = parameter order (ci, co, q)
    const int Cout = j.i[1], kk = j.i[2] * j.i[2], cp = j.i[3];
    const int64_t total = static_cast<int64_t>(j.i[0]) * Cout * kk;
    const float* w = static_cast<const float*>(j.src);
    const int part = j.i[10];
    for (int64_t e = begin + threadIdx.x; e < begin + kChunk && e < total; e += 256) {
      const int q = static_cast<int>(e % kk);
      const int co = static_cast<int>((e / kk) % Cout);
      const int ci = static_cast<int>(e / (static_cast<int64_t>(kk) * Cout));
      const int64_t of = (static_cast<int64_t>(q) * cp + co) * j.i[4] + ci, od = (static_cast<int64_t>(q) * j.i[5] + ci) * j.i[6] + co;
      if (part == 0) {
        if (j.dst0) put_w<__nv_bfloat16>(static_cast<__nv_bfloat16*>(j.dst0) + of, w[e], 0);
        if (j.dst1) put_w<__nv_bfloat16>(static_cast<__nv_bfloat16*>(j.dst1) + od, w[e], 0);
      } else {
        if (j.dst0) put_w<float>(static_cast<float*>(j.dst0) + of, w[e], part);
        if (j.dst1) put_w<float>(static_cast<float*>(j.dst1) + od, w[e], part);
      }
    }
  } else if (j.kind == MTBC_JOB_PACK_CONV_PAIR) {
    // i: Cout, Cin, c_begin, c_count, rows, ld, k0, n0, Ks, Np, dgrad (include/mtbc.h).  One element per
    // (tap row, pair offset, input parity, output parity, co, ci): the non-zero elements of the pixel-pair operand.
    const int Cout = j.i[0], Cin = j.i[1], c_begin = j.i[2], c_count = j.i[3], rows = j.i[4], ld = j.i[5], k0 = j.i[6],
              n0 = j.i[7], Ks = j.i[8], Np = j.i[9], dgrad = j.i[10];
    const int64_t per_tap = 4ll * Cout * c_count;
    const int64_t total = 9 * per_tap;
    const float* w = static_cast<const float*>(j.src);
    __nv_bfloat16* wp = static_cast<__nv_bfloat16*>(j.dst0);
    for (int64_t e = begin + threadIdx.x; e < begin + kChunk && e < total; e += 256) {
      const int tapp = static_cast<int>(e / per_tap);
      int r = static_cast<int>(e - tapp * per_tap);
      const int ci = r % c_count; r /= c_count;
      const int co = r % Cout; r /= Cout;
      const int par = r & 1, op = r >> 1;
      const int dh = tapp / 3 - 1, dq = tapp % 3 - 1;
      const int dw = 2 * dq + par - op;
      if (dw < -1 || dw > 1) continue;
      const int kh = dgrad ? 1 - dh : dh + 1, kw = dgrad ? 1 - dw : dw + 1;
      const float v = w[((static_cast<int64_t>(co) * Cin + c_begin + ci) * 3 + kh) * 3 + kw];
      const int n = dgrad ? ci : co, k = dgrad ? co : ci;
      wp[(static_cast<int64_t>(tapp) * rows + n0 + op * Np + n) * ld + k0 + par * Ks + k] = __float2bfloat16(v);
    }
  } else if (j.kind == MTBC_JOB_UNPACK_CONV) {
    // i: rows, ld, k0, Cout, Cin, ksz, c_begin, c_count, add.  Same 32 x 32 x taps tile, the other way round.
    const int local = blockIdx.x - chunk_first[blockIdx.x];
    switch (j.i[5]) {
      case 1: unpack_conv_tile<1>(j, local, s_tile); break;
      case 2: unpack_conv_tile<4>(j, local, s_tile); break;
      default: unpack_conv_tile<9>(j, local, s_tile); break;
    }
  } else if (j.kind == MTBC_JOB_UNPACK_CONVT) {
    // i: rows, ld, Cin, Cout, k, add ; element order = parameter order (ci, co, q)
    const int rows = j.i[0], ld = j.i[1], Cout = j.i[3], kk = j.i[4] * j.i[4], add = j.i[5];
    const int cp = rows / kk;
    const int64_t total = static_cast<int64_t>(j.i[2]) * Cout * kk;
    const float* acc = static_cast<const float*>(j.src);
    float* grad = static_cast<float*>(j.dst0);
    for (int64_t e = begin + threadIdx.x; e < begin + kChunk && e < total; e += 256) {
      const int q = static_cast<int>(e % kk);
      const int co = static_cast<int>((e / kk) % Cout);
      const int ci = static_cast<int>(e / (static_cast<int64_t>(kk) * Cout));
      const float v = acc[(static_cast<int64_t>(q) * cp + co) * ld + ci];
      grad[e] = add ? (grad[e] + v) : v;
    }
  }
}

struct ParamJobsOp : public OpBase {
  ParamJobDev* d_jobs = nullptr;
  int32_t* d_chunk_job = nullptr;
  int32_t* d_chunk_first = nullptr;
  int nblocks = 0;
  ~ParamJobsOp() override {
    cudaFree(d_jobs);
    cudaFree(d_chunk_job);
    cudaFree(d_chunk_first);
  }
  int launch(cudaStream_t st) override {
    if (nblocks == 0) return 0;
    param_jobs_kernel<<<nblocks, 256, 0, st>>>(d_jobs, d_chunk_job, d_chunk_first);
    return check_launch("param_jobs_kernel");
  }
};

static int64_t job_elems(const mtbc_param_job& j) {
  switch (j.kind) {
    case MTBC_JOB_COPY_F32: return j.i[0];
    case MTBC_JOB_PACK_CONV: return static_cast<int64_t>(j.i[0]) * j.i[4] * j.i[2] * j.i[2];
    case MTBC_JOB_PACK_CONVT: return static_cast<int64_t>(j.i[0]) * j.i[1] * j.i[2] * j.i[2];
    case MTBC_JOB_UNPACK_CONV: return static_cast<int64_t>(j.i[3]) * j.i[7] * j.i[5] * j.i[5];
    case MTBC_JOB_UNPACK_CONVT: return static_cast<int64_t>(j.i[2]) * j.i[3] * j.i[4] * j.i[4];
    case MTBC_JOB_PACK_CONV_PAIR: return 36ll * j.i[0] * j.i[3];
    default: return -1;
  }
}

int param_jobs_create(const mtbc_param_job* jobs, int n, OpBase** out) {
  if (n < 0 || (n > 0 && !jobs) || !out) return set_error(MTBC_ERR_INVALID, "param_jobs: bad arguments");
  std::vector<ParamJobDev> h(n);
  std::vector<int32_t> cj, cf;
  for (int a = 0; a < n; ++a) {
    const int64_t e = job_elems(jobs[a]);
    const bool is_pack = jobs[a].kind == MTBC_JOB_PACK_CONV || jobs[a].kind == MTBC_JOB_PACK_CONVT;
    if (e < 0 || !jobs[a].src || !(jobs[a].dst0 || (is_pack && jobs[a].dst1)))
      return set_error(MTBC_ERR_INVALID, "param_jobs: job %d is malformed", a);
    if (jobs[a].kind == MTBC_JOB_PACK_CONV_PAIR) {
      const int32_t* i = jobs[a].i;   // Cout, Cin, c_begin, c_count, rows, ld, k0, n0, Ks, Np, dgrad
      const int nn = i[10] ? i[3] : i[0], kk = i[10] ? i[0] : i[3];
      if (i[0] < 1 || i[3] < 1 || i[2] < 0 || i[2] + i[3] > i[1] || kk > i[8] || nn > i[9] || i[6] < 0 || i[7] < 0 ||
          i[6] + i[8] + kk > i[5] || i[7] + i[9] + nn > i[4])
        return set_error(MTBC_ERR_INVALID, "param_jobs: job %d: pair pack does not fit its operand", a);
    }
    h[a].kind = jobs[a].kind;
    for (int k = 0; k < 11; ++k) h[a].i[k] = jobs[a].i[k];
    h[a].src = jobs[a].src; h[a].dst0 = jobs[a].dst0; h[a].dst1 = jobs[a].dst1;
    const int first = static_cast<int>(cj.size());
    int64_t nchunks = (e + kChunk - 1) / kChunk;
    if (jobs[a].kind == MTBC_JOB_PACK_CONV || jobs[a].kind == MTBC_JOB_UNPACK_CONV) {
      const bool pk = jobs[a].kind == MTBC_JOB_PACK_CONV;
      const int cout = pk ? jobs[a].i[0] : jobs[a].i[3], ccount = pk ? jobs[a].i[4] : jobs[a].i[7];
      const int ksz = pk ? jobs[a].i[2] : jobs[a].i[5];
      if (ksz < 1 || ksz > 3) return set_error(MTBC_ERR_INVALID, "param_jobs: job %d kernel size %d", a, ksz);
      nchunks = static_cast<int64_t>((cout + 31) / 32) * ((ccount + 31) / 32);   // one block per 32 x 32 x taps tile
    }
    for (int64_t c = 0; c < nchunks; ++c) { cj.push_back(a); cf.push_back(first); }
  }
  ParamJobsOp* op = new ParamJobsOp();
  op->nblocks = static_cast<int>(cj.size());
  if (op->nblocks > 0) {
    cudaError_t e1 = cudaMalloc(&op->d_jobs, sizeof(ParamJobDev) * n);
    cudaError_t e2 = cudaMalloc(&op->d_chunk_job, sizeof(int32_t) * cj.size());
    cudaError_t e3 = cudaMalloc(&op->d_chunk_first, sizeof(int32_t) * cf.size());
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { delete op; return set_error(MTBC_ERR_CUDA, "param_jobs: cudaMalloc failed"); }
    // synchronous copies: creation is a set-up time call (never inside a capture)
    cudaMemcpy(op->d_jobs, h.data(), sizeof(ParamJobDev) * n, cudaMemcpyHostToDevice);
    cudaMemcpy(op->d_chunk_job, cj.data(), sizeof(int32_t) * cj.size(), cudaMemcpyHostToDevice);
    cudaError_t e4 = cudaMemcpy(op->d_chunk_first, cf.data(), sizeof(int32_t) * cf.size(), cudaMemcpyHostToDevice);
    if (e4 != cudaSuccess) { delete op; return set_error(MTBC_ERR_CUDA, "param_jobs: upload failed: %s", cudaGetErrorString(e4)); }
  }
  *out = op;
  return 0;
}

// ------------------------------------------------------------------------------------------------ first layer
// xs[n][ci][r*3+s] = sum over output pixels (h,w) of x[n,ci,h+r-1,w+s-1] (zero padded): the nine shifted plane sums
// from which the per-(n,channel) mean of the conv output follows linearly.  grid = N*Cin blocks of 256 threads.
// Two steps: (1) grid (planes, slices): each block adds its slice's five partial sums (total, first / last row,
// first / last column) into xs[plane][0..4] (zeroed by the caller); (2) one thread per plane turns them into the nine
// shifted sums in place.  (One block per plane was a 57 us launch for 32 planes of 256 x 256.)
__global__ void __launch_bounds__(256) conv_first_shift_sums_kernel(const float* __restrict__ x, int H, int W,
                                                                    float* __restrict__ xs) {
  const float* plane = x + static_cast<int64_t>(blockIdx.x) * H * W;
  const int HW = H * W;
  const int per = (HW + gridDim.y - 1) / gridDim.y;
  const int i0 = blockIdx.y * per, i1 = min(HW, i0 + per);
  float tot = 0.f, r0 = 0.f, rl = 0.f, c0 = 0.f, cl = 0.f;
  for (int i = i0 + threadIdx.x; i < i1; i += 256) {
    const float v = plane[i];
    const int h = i / W, w = i - h * W;
    tot += v;
    if (h == 0) r0 += v;
    if (h == H - 1) rl += v;
    if (w == 0) c0 += v;
    if (w == W - 1) cl += v;
  }
  __shared__ float red[5][8];
  float vals[5] = {tot, r0, rl, c0, cl};
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    float v = vals[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    float t = 0.f;
    for (int j = 0; j < 8; ++j) t += red[threadIdx.x][j];
    atomicAdd(xs + blockIdx.x * 9 + threadIdx.x, t);
  }
}
__global__ void conv_first_shift_finish_kernel(const float* __restrict__ x, int H, int W, int planes,
                                               float* __restrict__ xs) {
  const int pl = blockIdx.x * blockDim.x + threadIdx.x;
  if (pl >= planes) return;
  const float* plane = x + static_cast<int64_t>(pl) * H * W;
  float t[5];
  for (int k = 0; k < 5; ++k) t[k] = xs[pl * 9 + k];
  const float x00 = plane[0], x0l = plane[W - 1], xl0 = plane[(H - 1) * W], xll = plane[(H - 1) * W + W - 1];
  // tap (r,s) reads x[h+r-1][w+s-1]: r=0 never touches the last row, r=2 never the first; same for columns
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) {
      float v = t[0];
      if (r == 0) v -= t[2];
      if (r == 2) v -= t[1];
      if (s == 0) v -= t[4];
      if (s == 2) v -= t[3];
      if (r == 0 && s == 0) v += xll;
      if (r == 0 && s == 2) v += xl0;
      if (r == 2 && s == 0) v += x0l;
      if (r == 2 && s == 2) v += x00;
      xs[pl * 9 + r * 3 + s] = v;
    }
}

// One thread per output pixel, 128 consecutive pixels of one sample per block; weights staged in smem.
// xs != nullptr: the per-(n,channel) mean of the conv output (exact, from the shifted plane sums) is subtracted before
// the bf16 rounding and the statistics.  InstanceNorm is shift invariant, so everything downstream is unchanged, but
// the raw 0..255 image (BUSI_dataset.py:102, no normalisation) puts a DC level of many standard deviations on this
// layer's output: stored centred, bf16 spends its 8 mantissa bits on the signal and sum(y^2) does not cancel.
// CIN is a compile-time constant so the 9*CIN input taps live in registers (a run-time Cin puts them in local memory
// and every FMA pays a local load); weights are staged k-major [k][Cp] so 4 output channels come from one 128-bit
// shared-memory broadcast load.
// PPT pixels per thread (same sample, 128 pixels apart so every access stays coalesced): the InstanceNorm partial sums
// stay in registers over the PPT pixels and are combined once (16 shuffles per 16 columns and quantity), not per pixel:
// at one pixel per thread the 64 shuffles + shared-memory atomics per pixel, not the 100 MB store, bounded the kernel
// (105 us at 32 x 256 x 256 x 24 against a 26 us write floor).  NCH = 16-column chunks (Cs / 16) when PPT > 1.
template <typename T, int CIN, int NCH = 0, int PPT = 1>
__global__ void __launch_bounds__(128) conv_first_fwd_kernel(const float* __restrict__ x, int N, int H, int W,
                                                             const float* __restrict__ w, const float* __restrict__ bias,
                                                             int Cout, T* __restrict__ y, int Cp,
                                                             float* __restrict__ stat_sum, float* __restrict__ stat_sq,
                                                             const float* __restrict__ xs) {
  extern __shared__ __align__(16) float sm[];
  constexpr int K = CIN * 9;
  const int Cs = (Cp + 15) & ~15;     // staging pitch: whole 16-column chunks even when the tensor pitch is dense
  float* s_w = sm;                    // [K][Cs], zero padded
  float* s_b = s_w + Cs * K;          // [Cs]
  float* s_st = s_b + Cs;             // [2][Cs]
  for (int i = threadIdx.x; i < Cs * K; i += 128) {
    const int c = i / K, k = i - c * K;
    s_w[k * Cs + c] = c < Cout ? w[i] : 0.f;
  }
  for (int i = threadIdx.x; i < Cs; i += 128) {
    s_b[i] = (bias != nullptr && i < Cout) ? bias[i] : 0.f;
    s_st[i] = 0.f;
    s_st[Cs + i] = 0.f;
  }
  __syncthreads();
  const int64_t HW = static_cast<int64_t>(H) * W;
  const int64_t pix0 = blockIdx.x * (128ll * PPT) + threadIdx.x;  // within the whole batch
  const int n = static_cast<int>(pix0 / HW);
  const int lane = threadIdx.x & 31;
  if (xs != nullptr) {
    // the block's pixels belong to one sample (HW % (128 * PPT) == 0): replace the bias by -(mean of w * x)
    const float inv = 1.f / static_cast<float>(HW);
    for (int c = threadIdx.x; c < Cs; c += 128) {
      float m = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) m = fmaf(s_w[k * Cs + c], xs[static_cast<int64_t>(n) * K + k], m);
      s_b[c] = -m * inv;
    }
    __syncthreads();
  }
  const bool wide = (Cp % 16) == 0;
  constexpr int NACC = NCH > 0 ? NCH * 16 : 1;
  float ps[NACC], pq[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { ps[i] = 0.f; pq[i] = 0.f; }
#pragma unroll 1
  for (int j = 0; j < PPT; ++j) {
    const int64_t pix = pix0 + j * 128ll;
    const int hw = static_cast<int>(pix - n * HW);
    const int h = hw / W, ww = hw - h * W;
    float xin[K];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int hh = h + r - 1, wc = ww + s - 1;
          xin[ci * 9 + r * 3 + s] =
              (hh >= 0 && hh < H && wc >= 0 && wc < W) ? __ldg(x + (static_cast<int64_t>(n) * CIN + ci) * HW + hh * W + wc) : 0.f;
        }
    T* dst = y + pix * Cp;
    auto chunk = [&](int c0, float (&v)[16]) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b4 = *reinterpret_cast<const float4*>(s_b + c0 + 4 * q);
        v[4 * q] = b4.x; v[4 * q + 1] = b4.y; v[4 * q + 2] = b4.z; v[4 * q + 3] = b4.w;
      }
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float xk = xin[k];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 w4 = *reinterpret_cast<const float4*>(s_w + k * Cs + c0 + 4 * q);
          v[4 * q] = fmaf(xk, w4.x, v[4 * q]); v[4 * q + 1] = fmaf(xk, w4.y, v[4 * q + 1]);
          v[4 * q + 2] = fmaf(xk, w4.z, v[4 * q + 2]); v[4 * q + 3] = fmaf(xk, w4.w, v[4 * q + 3]);
        }
      }
    };
    if constexpr (NCH > 0) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        float v[16];
        chunk(c * 16, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) { ps[c * 16 + i] += v[i]; pq[c * 16 + i] = fmaf(v[i], v[i], pq[c * 16 + i]); }
        const int left = Cp - c * 16;
        emit16_n<T>(dst + c * 16, v, false, left >= 16 ? 16 : 8, wide);
      }
    } else {
      for (int c0 = 0; c0 < Cs; c0 += 16) {
        float v[16];
        chunk(c0, v);
        if (stat_sum != nullptr) {
          float sq[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) sq[i] = v[i] * v[i];
          const float cs = warp_colsum16(v, lane);
          const float cq = warp_colsum16(sq, lane);
          if ((lane & 1) == 0) {
            const int cc = c0 + col16_of_lane(lane);
            atomicAdd(&s_st[cc], cs);
            atomicAdd(&s_st[Cs + cc], cq);
          }
        }
        const int left = Cp - c0;
        emit16_n<T>(dst + c0, v, false, left >= 16 ? 16 : 8, wide);
      }
    }
  }
  if constexpr (NCH > 0) {
    if (stat_sum != nullptr) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        float a[16], q[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { a[i] = ps[c * 16 + i]; q[i] = pq[c * 16 + i]; }
        const float cs = warp_colsum16(a, lane);
        const float cq = warp_colsum16(q, lane);
        if ((lane & 1) == 0) {
          const int cc = c * 16 + col16_of_lane(lane);
          atomicAdd(&s_st[cc], cs);
          atomicAdd(&s_st[Cs + cc], cq);
        }
      }
    }
  }
  if (stat_sum != nullptr) {
    __syncthreads();
    for (int i = threadIdx.x; i < Cp; i += 128) {
      atomicAdd(stat_sum + static_cast<int64_t>(n) * Cp + i, s_st[i]);
      atomicAdd(stat_sq + static_cast<int64_t>(n) * Cp + i, s_st[Cs + i]);
    }
  }
}

// dW[co][ci][r][s] += sum_pix x[n,ci,h+r-1,w+s-1] * dy[pix][co].  grid = (pixel blocks, Cin).  G = 4 or 8 consecutive
// threads share one pixel and each owns one 16-byte vector (8 output channels) of its dy row, so a warp reads whole
// contiguous pixel rows and dy crosses HBM exactly once per input channel (the previous version looped over the
// 8-channel groups OUTSIDE the pixel loop and fetched every 32-byte sector of the dense 48-byte rows up to three
// times: 289 MB of DRAM traffic for a 100 MB tensor, 139 us).  The nine taps x 8 channels accumulate in registers over
// the whole pixel range and are combined once: xor shuffles across the lanes of equal group, then shared-memory atomics.
template <int G, typename T>
__global__ void __launch_bounds__(256) conv_first_wgrad_kernel(const float* __restrict__ x, int N, int Cin, int H,
                                                               int W, const T* __restrict__ dy, int Cp,
                                                               int Cout, float* __restrict__ dw) {
  __shared__ float s_acc[9][64];
  const int ci = blockIdx.y;
  for (int i = threadIdx.x; i < 9 * 64; i += 256) (&s_acc[0][0])[i] = 0.f;
  __syncthreads();
  const int64_t HW = static_cast<int64_t>(H) * W;
  const int64_t total = static_cast<int64_t>(N) * HW;
  const int g = threadIdx.x & (G - 1);          // 8-channel group of this thread
  const int slot = threadIdx.x / G;             // pixel slot inside the block
  constexpr int PPB = 256 / G;                  // pixels per block and iteration
  const bool active = g * 8 < Cp;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
  if (active) {
#pragma unroll 2
    for (int64_t pix = blockIdx.x * static_cast<int64_t>(PPB) + slot; pix < total; pix += static_cast<int64_t>(gridDim.x) * PPB) {
      // 32-bit index arithmetic (the host checks N*H*W < 2^31): a 64-bit division per pixel costs more issue slots
      // than the 72 FMAs it feeds
      const uint32_t p32 = static_cast<uint32_t>(pix), hw32 = static_cast<uint32_t>(HW);
      const int n = static_cast<int>(p32 / hw32);
      const int hw = static_cast<int>(p32 - static_cast<uint32_t>(n) * hw32);
      const int h = hw / W, ww = hw - h * W;
      const float* xp = x + (static_cast<int64_t>(n) * Cin + ci) * HW;
      const V8 d0 = load8<T>(dy + pix * Cp + g * 8);
      float xs[9];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int hh = h + r - 1, wc = ww + s - 1;
          xs[r * 3 + s] = (hh >= 0 && hh < H && wc >= 0 && wc < W) ? __ldg(xp + hh * W + wc) : 0.f;
        }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[t][i] = fmaf(xs[t], d0.f[i], acc[t][i]);
      }
    }
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = acc[t][i];
#pragma unroll
      for (int o = 16; o >= G; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);   // lanes of equal group
      if (lane < G && active) atomicAdd(&s_acc[t][g * 8 + i], v);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * 64; i += 256) {
    const int t = i / 64, co = i % 64;
    if (co < Cout) atomicAdd(dw + (static_cast<int64_t>(co) * Cin + ci) * 9 + t, s_acc[t][co]);
  }
}

// Row-strip variant of the first-layer weight gradient (W % 64 == 0, H % 8 == 0, at most four 8-channel groups): the
// launch sits alone at the very end of the backward pass (nothing left to overlap it), so its 125 us were step time.
// A block owns 8 image rows of one sample: the 10 x rows it needs go to shared memory once, zero halo included (the
// per-pixel version spent half of its instructions on nine predicated, L1-served x loads per thread, with a quarter of
// the threads idle at 24 channels); thread (g = tid / 64, s = tid % 64) walks pixels s, s + 64, ... of every row with the
// 8 channels of group g: one 16-byte dy load, nine shared-memory loads and 72 FMAs per pixel.  The partials of a block
// are summed through shared memory (one pass, no shuffles, no contended atomics) and leave as 72 G global atomics.
constexpr int kFirstRB = 8;
template <typename T>
__global__ void __launch_bounds__(256) conv_first_wgrad_rows_kernel(const float* __restrict__ x, int N, int Cin, int H, int W,
                                                                    const T* __restrict__ dy, int Cp, int Cout,
                                                                    float* __restrict__ dw) {
  extern __shared__ float sm[];
  const int nthr = blockDim.x, NG = nthr >> 6;
  const int ci = blockIdx.y;
  const int strips = H / kFirstRB;
  const int n = blockIdx.x / strips, h0 = (blockIdx.x - n * strips) * kFirstRB;
  const int WP = W + 2;
  float* s_x = sm;                                   // [kFirstRB + 2][W + 2]
  const float* plane = x + (static_cast<int64_t>(n) * Cin + ci) * H * W;
  for (int i = threadIdx.x; i < (kFirstRB + 2) * WP; i += nthr) {
    const int r = i / WP, c = i - r * WP;
    const int hh = h0 - 1 + r, ww = c - 1;
    s_x[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(plane + hh * W + ww) : 0.f;
  }
  __syncthreads();
  const int g = threadIdx.x >> 6, sl = threadIdx.x & 63;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
  for (int r = 0; r < kFirstRB; ++r) {
    const T* drow = dy + ((static_cast<int64_t>(n) * H + h0 + r) * W) * Cp + g * 8;
    const float* xr = s_x + r * WP + sl;
#pragma unroll 2
    for (int w0 = 0; w0 < W; w0 += 64) {
      const V8 d0 = load8<T>(drow + static_cast<int64_t>(w0 + sl) * Cp);
      float xs[9];
#pragma unroll
      for (int rr = 0; rr < 3; ++rr)
#pragma unroll
        for (int ss = 0; ss < 3; ++ss) xs[rr * 3 + ss] = xr[rr * WP + w0 + ss];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[t][i] = fmaf(xs[t], d0.f[i], acc[t][i]);
    }
  }
  __syncthreads();                                   // s_x is dead: reuse the memory for the partials
  const int pitch = nthr + 1;
  float* s_p = sm;                                   // [72][nthr + 1]
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < 8; ++i) s_p[(t * 8 + i) * pitch + threadIdx.x] = acc[t][i];
  __syncthreads();
  for (int o = threadIdx.x; o < NG * 72; o += nthr) {
    const int g2 = o / 72, ti = o - g2 * 72;
    const int t = ti >> 3, co = g2 * 8 + (ti & 7);
    const float* row = s_p + ti * pitch + g2 * 64;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
    for (int k = 0; k < 64; k += 4) { a0 += row[k]; a1 += row[k + 1]; a2 += row[k + 2]; a3 += row[k + 3]; }
    if (co < Cout) atomicAdd(dw + (static_cast<int64_t>(co) * Cin + ci) * 9 + t, (a0 + a1) + (a2 + a3));
  }
}

// ------------------------------------------------------------------------------------------------ utilities
__global__ void fill_f32_kernel(float* p, int64_t n, float v) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    p[i] = v;
}
__global__ void f32_to_bf16_nhwc_kernel(const float* __restrict__ x, int N, int C, int H, int W,
                                        __nv_bfloat16* __restrict__ y, int Cp) {
  const int64_t HW = static_cast<int64_t>(H) * W;
  const int64_t total = static_cast<int64_t>(N) * HW * Cp;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cp);
    const int64_t pix = i / Cp;
    const int n = static_cast<int>(pix / HW);
    const int64_t hw = pix - n * HW;
    y[i] = __float2bfloat16(c < C ? x[(static_cast<int64_t>(n) * C + c) * HW + hw] : 0.f);
  }
}
__global__ void bf16_nhwc_to_f32_kernel(const __nv_bfloat16* __restrict__ x, int N, int C, int H, int W, int Cp,
                                        float* __restrict__ y) {
  const int64_t HW = static_cast<int64_t>(H) * W;
  const int64_t total = static_cast<int64_t>(N) * C * HW;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t hw = i % HW;
    const int c = static_cast<int>((i / HW) % C);
    const int n = static_cast<int>(i / (HW * C));
    y[i] = __bfloat162float(x[(n * HW + hw) * Cp + c]);
  }
}

static int grid_for(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace mtbc

using namespace mtbc;

extern "C" {

int mtbc_pack_conv_weight(const float* w, int32_t Cout, int32_t Cin, int32_t ksz, int32_t c_begin, int32_t c_count,
                          void* wf, int32_t wf_rows, int32_t wf_ld, int32_t wf_k0, void* wd, int32_t wd_rows,
                          int32_t wd_ld, void* stream) {
  if (!w || c_begin < 0 || c_begin + c_count > Cin) return set_error(MTBC_ERR_INVALID, "pack_conv_weight: bad slice");
  const int64_t total = static_cast<int64_t>(Cout) * c_count * ksz * ksz;
  pack_conv_weight_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, Cout, Cin, ksz, c_begin, c_count, static_cast<__nv_bfloat16*>(wf), wf_rows, wf_ld, wf_k0,
      static_cast<__nv_bfloat16*>(wd), wd_rows, wd_ld);
  return check_launch("pack_conv_weight");
}
int mtbc_pack_convT_weight(const float* w, int32_t Cin, int32_t Cout, int32_t k, int32_t cp, void* wf, int32_t wf_ld,
                           void* wd, int32_t wd_rows, int32_t wd_ld, void* stream) {
  if (!w) return set_error(MTBC_ERR_INVALID, "pack_convT_weight: null");
  const int64_t total = static_cast<int64_t>(Cin) * Cout * k * k;
  pack_convT_weight_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, Cin, Cout, k, cp, static_cast<__nv_bfloat16*>(wf), wf_ld, static_cast<__nv_bfloat16*>(wd), wd_rows, wd_ld);
  return check_launch("pack_convT_weight");
}
int mtbc_unpack_conv_wgrad(const float* acc, int32_t rows, int32_t ld, int32_t k0, float* grad, int32_t Cout,
                           int32_t Cin, int32_t ksz, int32_t c_begin, int32_t c_count, int32_t add, void* stream) {
  const int64_t total = static_cast<int64_t>(Cout) * c_count * ksz * ksz;
  unpack_conv_wgrad_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      acc, rows, ld, k0, grad, Cout, Cin, ksz, c_begin, c_count, add);
  return check_launch("unpack_conv_wgrad");
}
int mtbc_unpack_convT_wgrad(const float* acc, int32_t rows, int32_t ld, float* grad, int32_t Cin, int32_t Cout,
                            int32_t k, int32_t add, void* stream) {
  const int64_t total = static_cast<int64_t>(Cin) * Cout * k * k;
  unpack_convT_wgrad_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(acc, rows, ld, grad,
                                                                                                 Cin, Cout, k, add);
  return check_launch("unpack_convT_wgrad");
}

int mtbc_conv_first_fwd(const float* x, int32_t N, int32_t Cin, int32_t H, int32_t W, const float* w, const float* bias,
                        int32_t Cout, void* y, int32_t Cp, float* stat_sum, float* stat_sq, float* center_scratch,
                        void* stream) {
  if (Cin < 1 || Cin > 4 || Cp % 8 != 0 || Cout > Cp) return set_error(MTBC_ERR_INVALID, "conv_first_fwd: Cin must be 1..4");
  const int64_t HW = static_cast<int64_t>(H) * W;
  if (HW % 128 != 0) return set_error(MTBC_ERR_INVALID, "conv_first_fwd: H*W must be a multiple of 128");
  if (center_scratch != nullptr) {
    if (H < 2 || W < 2) return set_error(MTBC_ERR_INVALID, "conv_first_fwd: centring needs H, W >= 2");
    cudaStream_t st0 = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(center_scratch, 0, static_cast<size_t>(N) * Cin * 9 * sizeof(float), st0);
    if (e != cudaSuccess) return set_error(MTBC_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
    int slices = static_cast<int>(HW / 4096); if (slices < 1) slices = 1; if (slices > 32) slices = 32;
    if (current_mode() & MODE_DETERMINISTIC) slices = 1;   // one block per plane: no cross-block float atomics
    conv_first_shift_sums_kernel<<<dim3(N * Cin, slices), 256, 0, st0>>>(x, H, W, center_scratch);
    int rc = check_launch("conv_first_shift_sums");
    if (rc) return rc;
    conv_first_shift_finish_kernel<<<cdiv(N * Cin, 128), 128, 0, st0>>>(x, H, W, N * Cin, center_scratch);
    rc = check_launch("conv_first_shift_finish");
    if (rc) return rc;
  }
  const int Cs = (Cp + 15) & ~15;
  const int smem = (Cs * Cin * 9 + 3 * Cs) * sizeof(float);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = static_cast<int>(N * HW / 128);
  if (current_mode() & MODE_ACT_FP32) {   // TF32 parity mode: fp32 output tensor, generic kernel (speed is not its point)
    float* yf = static_cast<float*>(y);
    switch (Cin) {
      case 1: conv_first_fwd_kernel<float, 1><<<grid, 128, smem, st>>>(x, N, H, W, w, bias, Cout, yf, Cp, stat_sum, stat_sq, center_scratch); break;
      case 2: conv_first_fwd_kernel<float, 2><<<grid, 128, smem, st>>>(x, N, H, W, w, bias, Cout, yf, Cp, stat_sum, stat_sq, center_scratch); break;
      case 3: conv_first_fwd_kernel<float, 3><<<grid, 128, smem, st>>>(x, N, H, W, w, bias, Cout, yf, Cp, stat_sum, stat_sq, center_scratch); break;
      default: conv_first_fwd_kernel<float, 4><<<grid, 128, smem, st>>>(x, N, H, W, w, bias, Cout, yf, Cp, stat_sum, stat_sq, center_scratch); break;
    }
    return check_launch("conv_first_fwd");
  }
  __nv_bfloat16* yb = static_cast<__nv_bfloat16*>(y);
  // single-channel images on planes of whole 1024-pixel blocks (every benchmark shape): 8 pixels per thread
  if (Cin == 1 && Cs <= 32 && HW % 1024 == 0 && !getenv("MTBC_FIRST_PPT1")) {
    const int g8 = static_cast<int>(N * HW / 1024);
    if (Cs == 16) conv_first_fwd_kernel<__nv_bfloat16, 1, 1, 8><<<g8, 128, smem, st>>>(x, N, H, W, w, bias, Cout, yb, Cp, stat_sum, stat_sq, center_scratch);
    else conv_first_fwd_kernel<__nv_bfloat16, 1, 2, 8><<<g8, 128, smem, st>>>(x, N, H, W, w, bias, Cout, yb, Cp, stat_sum, stat_sq, center_scratch);
    return check_launch("conv_first_fwd");
  }
  switch (Cin) {
    case 1: conv_first_fwd_kernel<__nv_bfloat16, 1><<<grid, 128, smem, st>>>(x, N, H, W, w, bias, Cout, yb, Cp, stat_sum, stat_sq, center_scratch); break;
    case 2: conv_first_fwd_kernel<__nv_bfloat16, 2><<<grid, 128, smem, st>>>(x, N, H, W, w, bias, Cout, yb, Cp, stat_sum, stat_sq, center_scratch); break;
    case 3: conv_first_fwd_kernel<__nv_bfloat16, 3><<<grid, 128, smem, st>>>(x, N, H, W, w, bias, Cout, yb, Cp, stat_sum, stat_sq, center_scratch); break;
    default: conv_first_fwd_kernel<__nv_bfloat16, 4><<<grid, 128, smem, st>>>(x, N, H, W, w, bias, Cout, yb, Cp, stat_sum, stat_sq, center_scratch); break;
  }
  return check_launch("conv_first_fwd");
}
int mtbc_conv_first_wgrad(const float* x, int32_t N, int32_t Cin, int32_t H, int32_t W, const void* dy, int32_t Cp,
                          int32_t Cout, float* dw, void* stream) {
  if (Cin < 1 || Cin > 4 || Cout > 64 || Cp % 8 != 0) return set_error(MTBC_ERR_INVALID, "conv_first_wgrad: Cin 1..4, Cout <= 64");
  const int64_t total = static_cast<int64_t>(N) * H * W;
  if (total >= (1ll << 31)) return set_error(MTBC_ERR_INVALID, "conv_first_wgrad: N*H*W must be below 2^31");
  cudaStream_t st_rows = static_cast<cudaStream_t>(stream);
  if (W % 64 == 0 && H % kFirstRB == 0 && Cp <= 32 && !getenv("MTBC_FIRST_WGRAD_PIXELS")) {
    const int NG = Cp / 8, nthr = NG * 64;
    size_t smem = static_cast<size_t>(72) * (nthr + 1) * sizeof(float);
    const size_t sx = static_cast<size_t>(kFirstRB + 2) * (W + 2) * sizeof(float);
    if (sx > smem) smem = sx;
    if (smem <= 200 * 1024) {
      const dim3 grid(N * (H / kFirstRB), Cin);
      MTBC_DISPATCH_ACT(({
        if (smem > 48 * 1024) cudaFuncSetAttribute(conv_first_wgrad_rows_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        conv_first_wgrad_rows_kernel<T><<<grid, nthr, smem, st_rows>>>(x, N, Cin, H, W, static_cast<const T*>(dy), Cp, Cout, dw);
      }));
      return check_launch("conv_first_wgrad_rows");
    }
  }
  const int G = Cp <= 32 ? 4 : 8;
  int gx = static_cast<int>((total + (256 / G) * 16 - 1) / ((256 / G) * 16));
  if (gx > 148 * 4) gx = 148 * 4;
  if (gx < 1) gx = 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (G == 4)
    MTBC_DISPATCH_ACT((conv_first_wgrad_kernel<4, T><<<dim3(gx, Cin), 256, 0, st>>>(x, N, Cin, H, W, static_cast<const T*>(dy), Cp, Cout, dw)));
  else
    MTBC_DISPATCH_ACT((conv_first_wgrad_kernel<8, T><<<dim3(gx, Cin), 256, 0, st>>>(x, N, Cin, H, W, static_cast<const T*>(dy), Cp, Cout, dw)));
  return check_launch("conv_first_wgrad");
}

int mtbc_fill_f32(float* p, int64_t n, float v, void* stream) {
  if (n <= 0) return 0;
  fill_f32_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, n, v);
  return check_launch("fill_f32");
}
int mtbc_zero_bytes(void* p, int64_t nbytes, void* stream) {
  if (nbytes <= 0) return 0;
  cudaError_t e = cudaMemsetAsync(p, 0, static_cast<size_t>(nbytes), static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return set_error(MTBC_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
  return 0;
}
int mtbc_copy_f32(float* dst, const float* src, int64_t n, void* stream) {
  if (n <= 0) return 0;
  cudaError_t e = cudaMemcpyAsync(dst, src, static_cast<size_t>(n) * sizeof(float), cudaMemcpyDeviceToDevice,
                                  static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return set_error(MTBC_ERR_CUDA, "cudaMemcpyAsync: %s", cudaGetErrorString(e));
  return 0;
}
int mtbc_f32_to_bf16_nhwc(const float* x, int32_t N, int32_t C, int32_t H, int32_t W, void* y, int32_t Cp,
                          void* stream) {
  const int64_t total = static_cast<int64_t>(N) * H * W * Cp;
  f32_to_bf16_nhwc_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, N, C, H, W, static_cast<__nv_bfloat16*>(y), Cp);
  return check_launch("f32_to_bf16_nhwc");
}
int mtbc_bf16_nhwc_to_f32(const void* x, int32_t N, int32_t C, int32_t H, int32_t W, int32_t Cp, float* y,
                          void* stream) {
  const int64_t total = static_cast<int64_t>(N) * C * H * W;
  bf16_nhwc_to_f32_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), N, C, H, W, Cp, y);
  return check_launch("bf16_nhwc_to_f32");
}

}  // extern "C"
