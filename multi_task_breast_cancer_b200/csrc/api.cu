// C-ABI glue: error reporting, opaque op handles, device check.
#include "internal.h"
#include <stdlib.h>
#include <stdarg.h>
#include <stdio.h>

namespace mtbc {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static thread_local int g_mode = 0;
int current_mode() { return g_mode; }

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MTBC_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(MTBC_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

}  // namespace mtbc

struct mtbc_op {
  mtbc::OpBase* impl;
};

extern "C" {

const char* mtbc_last_error(void) { return mtbc::g_err; }
int mtbc_abi_version(void) { return 4; }
#ifndef MTBC_BUILD_DIGEST
#define MTBC_BUILD_DIGEST "unstamped"
#endif
// SHA-256 of the sources + flags this binary was compiled from (build.py:_digest); _lib.load() refuses a mismatch.
const char* mtbc_build_digest(void) { return MTBC_BUILD_DIGEST; }

int mtbc_set_mode(int32_t flags) {
  if (flags & ~(mtbc::MODE_ACT_FP32 | mtbc::MODE_DETERMINISTIC)) return mtbc::set_error(MTBC_ERR_INVALID, "set_mode: unknown flag bits 0x%x", flags);
  mtbc::g_mode = flags;
  return 0;
}
int mtbc_get_mode(void) { return mtbc::g_mode; }

int mtbc_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return mtbc::set_error(MTBC_ERR_NO_DEVICE, "cudaGetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return mtbc::set_error(MTBC_ERR_NO_DEVICE, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return mtbc::set_error(MTBC_ERR_NO_DEVICE, "device %d is sm_%d%d, this library is sm_100a only", dev, prop.major, prop.minor);
  if (!mtbc::tensor_map_available())
    return mtbc::set_error(MTBC_ERR_NO_DEVICE, "driver entry point cuTensorMapEncodeTiled not found");
  return 0;
}

int mtbc_conv_gemm_create(const mtbc_conv_gemm_desc* d, mtbc_op** out) {
  mtbc::OpBase* b = nullptr;
  int rc = mtbc::conv_gemm_create(d, &b);
  if (rc) return rc;
  *out = new mtbc_op{b};
  return 0;
}
int mtbc_wgrad_create(const mtbc_wgrad_desc* d, mtbc_op** out) {
  mtbc::OpBase* b = nullptr;
  int rc = mtbc::wgrad_create(d, &b);
  if (rc) return rc;
  *out = new mtbc_op{b};
  return 0;
}
int mtbc_wgrad_multi_create(const mtbc_wgrad_multi_desc* d, mtbc_op** out) {
  if (!d || !out) return mtbc::set_error(MTBC_ERR_INVALID, "null argument");
  mtbc::OpBase* b = nullptr;
  int rc = mtbc::wgrad_halo_multi_create(d, &b);
  if (rc) return rc;
  *out = new mtbc_op{b};
  return 0;
}
int mtbc_convT_bwd_create(const mtbc_convT_bwd_desc* d, mtbc_op** out) {
  if (!d || !out) return mtbc::set_error(MTBC_ERR_INVALID, "null argument");
  mtbc::OpBase* b = nullptr;
  int rc = mtbc::convT_bwd_create(d, &b);
  if (rc) return rc;
  *out = new mtbc_op{b};
  return 0;
}
int mtbc_param_jobs_create(const mtbc_param_job* jobs, int32_t n, mtbc_op** out) {
  mtbc::OpBase* b = nullptr;
  int rc = mtbc::param_jobs_create(jobs, n, &b);
  if (rc) return rc;
  *out = new mtbc_op{b};
  return 0;
}
int mtbc_op_launch(mtbc_op* op, void* stream) {
  if (!op || !op->impl) return mtbc::set_error(MTBC_ERR_INVALID, "null op");
  return op->impl->launch(static_cast<cudaStream_t>(stream));
}
int mtbc_ops_launch(mtbc_op* const* ops, int32_t n, void* stream) {
  for (int i = 0; i < n; ++i) {
    int rc = mtbc_op_launch(ops[i], stream);
    if (rc) return rc;
  }
  return 0;
}
void mtbc_op_destroy(mtbc_op* op) {
  if (!op) return;
  delete op->impl;
  delete op;
}
double mtbc_op_flops(const mtbc_op* op) { return (op && op->impl) ? op->impl->op_flops() : 0.0; }

}  // extern "C"
