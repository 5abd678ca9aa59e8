// Element-type generic access to NHWC activations: the product path stores bf16 (8 channels = one 16-byte vector), the
// TF32 parity mode stores fp32 (8 channels = two 16-byte vectors).  Kernels are written once against V8 / load8 /
// store8 and instantiated for both.
#pragma once
#include "ptx.cuh"

namespace mtbc {

struct V8 {
  float f[8];
};

template <typename T>
__device__ __forceinline__ V8 load8(const T* p);
template <>
__device__ __forceinline__ V8 load8<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  V8 r;
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  r.f[0] = a.x; r.f[1] = a.y; r.f[2] = b.x; r.f[3] = b.y; r.f[4] = c.x; r.f[5] = c.y; r.f[6] = d.x; r.f[7] = d.y;
  return r;
}
template <>
__device__ __forceinline__ V8 load8<float>(const float* p) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  V8 r;
  r.f[0] = a.x; r.f[1] = a.y; r.f[2] = a.z; r.f[3] = a.w; r.f[4] = b.x; r.f[5] = b.y; r.f[6] = b.z; r.f[7] = b.w;
  return r;
}

template <typename T>
__device__ __forceinline__ void store8(T* p, const V8& v);
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const V8& v) {
  uint4 u;
  u.x = pack_bf16x2(v.f[0], v.f[1]); u.y = pack_bf16x2(v.f[2], v.f[3]);
  u.z = pack_bf16x2(v.f[4], v.f[5]); u.w = pack_bf16x2(v.f[6], v.f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
template <>
__device__ __forceinline__ void store8<float>(float* p, const V8& v) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v.f[0], v.f[1], v.f[2], v.f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v.f[4], v.f[5], v.f[6], v.f[7]);
}

// Raw (still packed) 8-channel vector: load several of these before use without paying 8 registers per channel group
template <typename T>
struct Raw8;
template <>
struct Raw8<__nv_bfloat16> {
  uint4 u;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { u = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ V8 unpack() const {
    V8 r;
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    r.f[0] = a.x; r.f[1] = a.y; r.f[2] = b.x; r.f[3] = b.y; r.f[4] = c.x; r.f[5] = c.y; r.f[6] = d.x; r.f[7] = d.y;
    return r;
  }
};
template <>
struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) { a = reinterpret_cast<const float4*>(p)[0]; b = reinterpret_cast<const float4*>(p)[1]; }
  __device__ __forceinline__ V8 unpack() const {
    V8 r;
    r.f[0] = a.x; r.f[1] = a.y; r.f[2] = a.z; r.f[3] = a.w; r.f[4] = b.x; r.f[5] = b.y; r.f[6] = b.z; r.f[7] = b.w;
    return r;
  }
};

// value as it will read back from storage (bf16 rounds; fp32 is exact)
template <typename T>
__device__ __forceinline__ float as_stored(float x);
template <>
__device__ __forceinline__ float as_stored<__nv_bfloat16>(float x) { return __bfloat162float(__float2bfloat16(x)); }
template <>
__device__ __forceinline__ float as_stored<float>(float x) { return x; }

template <typename T>
__device__ __forceinline__ float ld1(const T* p);
template <>
__device__ __forceinline__ float ld1<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <>
__device__ __forceinline__ float ld1<float>(const float* p) { return *p; }
template <typename T>
__device__ __forceinline__ void st1(T* p, float x);
template <>
__device__ __forceinline__ void st1<__nv_bfloat16>(__nv_bfloat16* p, float x) { *p = __float2bfloat16(x); }
template <>
__device__ __forceinline__ void st1<float>(float* p, float x) { *p = x; }

// 16 consecutive fp32 accumulator columns -> storage (plain store or accumulation), `nvalid` (0, 8 or 16) of them exist
template <typename T>
__device__ __forceinline__ void emit16_n(T* dst, const float (&v)[16], bool accumulate, int nvalid, bool wide);
template <>
__device__ __forceinline__ void emit16_n<__nv_bfloat16>(__nv_bfloat16* dst, const float (&v)[16], bool accumulate,
                                                        int nvalid, bool wide) {
  emit_bf16x16_n(dst, v, accumulate, nvalid, wide);
}
template <>
__device__ __forceinline__ void emit16_n<float>(float* dst, const float (&v)[16], bool accumulate, int nvalid, bool) {
  if (accumulate) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nvalid) atomicAdd(dst + i, v[i]);     // compiles to fire-and-forget RED.ADD.F32
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (4 * i < nvalid) reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
}

}  // namespace mtbc

// Dispatch a launcher body on the calling thread's activation dtype (mtbc_set_mode): `T` is bound inside `body`.
#define MTBC_DISPATCH_ACT(body)                         \
  do {                                                  \
    if (mtbc::current_mode() & mtbc::MODE_ACT_FP32) {   \
      using T = float;                                  \
      body;                                             \
    } else {                                            \
      using T = __nv_bfloat16;                          \
      body;                                             \
    }                                                   \
  } while (0)
