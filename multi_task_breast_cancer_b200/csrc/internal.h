// Internal (non-ABI) declarations shared by the translation units of libmtbc.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include "../../include/mtbc.h"

namespace mtbc {

// printf-style error recording (thread local); returns `code` so callers can `return set_error(...)`.
int set_error(int code, const char* fmt, ...);
// cudaGetLastError() after a launch -> MTBC status (never synchronises).
int check_launch(const char* what);

// Per-thread execution mode of the element-type generic entry points (mtbc_set_mode): which dtype the NHWC activation
// pointers they receive hold, and whether reductions must be order independent.
enum { MODE_ACT_FP32 = 1, MODE_DETERMINISTIC = 2 };
int current_mode();

struct OpBase {
  virtual ~OpBase() {}
  virtual int launch(cudaStream_t st) = 0;
  virtual double op_flops() const { return 0.0; }
};

int conv_gemm_create(const mtbc_conv_gemm_desc* d, OpBase** out);
int wgrad_create(const mtbc_wgrad_desc* d, OpBase** out);
int param_jobs_create(const mtbc_param_job* jobs, int n, OpBase** out);
bool tensor_map_available();
// > 0: not eligible (caller falls back to the generic kernel), 0: created, < 0: error
int conv_halo_try_create(const mtbc_conv_gemm_desc* d, OpBase** out);
int wgrad_halo_try_create(const mtbc_wgrad_desc* d, OpBase** out);
int wgrad_halo_multi_create(const mtbc_wgrad_multi_desc* d, OpBase** out);
int convT_bwd_create(const mtbc_convT_bwd_desc* d, OpBase** out);
// tensor-map encoders (conv_gemm.cu): bf16 NHWC view with box (kc, bw, bh, bn); packed weights with box (kc, BN, 1)
int encode_act(CUtensorMap* m, const mtbc_act_view& v, int kc, int bw, int bh, int bn, int fp32 = 0);
int encode_w(CUtensorMap* m, const void* w, int ktot, int nrows, int ntaps, int kc, int BN, int fp32 = 0);
int encode_rows(CUtensorMap* m, const void* ptr, int C, int W, int H, int N, int box_w, int box_h);

// bulk-copy pipelined streaming kernels (stream_pipe.cu) for large tensors
bool pipe_eligible(int64_t N, int64_t HW, int Cp);
int in_apply_pipe(const void* y, int N, int64_t HW, int Cp, const float* ssum, const float* ssq, const float* gamma,
                  const float* beta, float eps, float slope, void* a, float* mean, float* rstd, cudaStream_t st);
int in_bwd_reduce_pipe(const void* dA, const void* y, int N, int64_t HW, int Cp, const float* mean, const float* rstd,
                       const float* gamma, const float* beta, float slope, float* s1, float* s2, cudaStream_t st);
int in_bwd_apply_pipe(const void* dA, const void* y, int N, int64_t HW, int Cp, const float* mean, const float* rstd,
                      const float* gamma, const float* beta, float slope, const float* s1, const float* s2, void* dy,
                      float* dgamma, float* dbeta, int C_true, cudaStream_t st);
bool in_bwd_fused_eligible(int64_t N, int64_t HW, int Cp);
int in_bwd_fused(const void* dA, const void* y, int N, int64_t HW, int Cp, const float* mean, const float* rstd,
                 const float* gamma, const float* beta, float slope, float* s1, float* s2, void* dy, float* dgamma,
                 float* dbeta, int C_true, int* counters, cudaStream_t st);
int channel_sum_pipe(const void* t, int64_t npix, int Cp, int C_true, float* out, cudaStream_t st);

// Launch with programmatic stream serialization (the kernel MUST call pdl_wait() before it touches anything another
// kernel wrote).  MTBC_PDL=0 turns the attribute off (plain stream order).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int cdiv(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace mtbc
