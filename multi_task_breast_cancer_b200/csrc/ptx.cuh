// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld).
// Everything here is hand-written PTX; no CUTLASS/CuTe types are used on the device side.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace mtbc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a fully converged warp.  ptxas only emits straight-line UTCHMMA / UTMALDG sequences (operands in uniform
// registers) when it can prove a single thread executes them; behind a plain `lane == 0` test every tcgen05.mma is
// wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop that costs ~45 cycles per instruction (tools/exp_mma_rate2.cu).
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
  return p != 0;
}

// ---------------------------------------------------------------- programmatic dependent launch
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization (internal.h: launch_pdl) may start while the
// previous kernel of its stream is still draining: pdl_wait() blocks until that kernel (and, transitively, everything
// before it) has completed and its writes are visible -- nothing produced by another kernel may be touched before it;
// pdl_trigger() lets the NEXT kernel of the stream begin its own prologue (it is called after tcgen05.alloc, so a
// dependent CTA can never take TMEM columns a CTA of this grid still has to allocate).  Both are no-ops in a plain launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ uint32_t mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// Bounded spin: a broken pipeline traps (-> launch error) instead of hanging the GPU box.
// (A non-blocking mbarrier.test_wait probe in front of the try_wait loop halves the latency of an already-completed
// wait, 235 -> 115 cycles, but made no difference to any kernel's duration: tools/diag_halo_dbg.py.)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) { __trap(); }  // ~2 s: pipeline is broken
    }
  }
}

// ---------------------------------------------------------------- TMA loads (tiled mode, zero OOB fill)
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {  // whole warp, ncols pow2 >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, with the two 64-bit descriptors passed as (lo, hi) halves.  The issuing thread is a single lane whose scalar
// instruction stream bounds the MMA issue rate (tools/exp_mma_rate.cu: ~160 cycles/MMA when the descriptors are rebuilt
// with 64-bit shifts every time, 45 cycles when only the low word is advanced), so callers keep `hi` constant per operand
// and step `lo` (start address >> 4) with compile-time offsets.
__device__ __forceinline__ void umma_bf16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                               uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::tf32: fp32 operands in shared memory, the tensor core reads the TF32 part (sign, 8 exponent, 10 mantissa bits).
__device__ __forceinline__ void umma_tf32_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                               uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint32_t umma_desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | ((layout_type & 7u) << 29);
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread t <-> TMEM lane base+t).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Same without the wait: issue several loads back to back, then one tmem_wait_ld() -- the epilogue warps are alone on
// their scheduler, so every exposed TMEM round trip is dead time.
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32-byte (256-bit) global store of 16 bf16: one full sector per thread, no partial-sector merges in L2.
__device__ __forceinline__ void st_global_256(void* p, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4,
                                              uint32_t w5, uint32_t w6, uint32_t w7) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w0), "r"(w1), "r"(w2), "r"(w3), "r"(w4),
               "r"(w5), "r"(w6), "r"(w7)
               : "memory");
}
// Fire-and-forget bf16 accumulation of 8 values (16 B) in L2: gradient accumulation for tensors with several
// consumers without the read-modify-write round trip in the epilogue.
__device__ __forceinline__ void red_add_bf16x8(void* p, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
  asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(w0), "r"(w1), "r"(w2), "r"(w3)
               : "memory");
}
// 16 fp32 -> bf16 -> global (store or accumulate), 32 B per thread.
__device__ __forceinline__ void emit_bf16x16(__nv_bfloat16* dst, const float (&v)[16], bool accumulate);

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (sm_100 "version 1"), see DESIGN.md section "UMMA descriptors".
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4   bits [46,48) version = 1
//   bits [61,64) layout: 0 none, 2 = 128B swizzle, 4 = 64B swizzle, 6 = 32B swizzle
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout_type & 7u) << 61;
  return d;
}
// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32, M = 128.
//   [4,6) c fmt (1 = f32)  [7,10) a fmt (1 = bf16)  [10,13) b fmt  [15] a major  [16] b major (0 = K, 1 = MN)
//   [17,23) N >> 3         [24,29) M >> 4
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn_major,
                                                             uint32_t b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
  d |= (a_mn_major & 1u) << 15;
  d |= (b_mn_major & 1u) << 16;
  d |= ((N >> 3) & 0x3Fu) << 17;
  d |= ((M >> 4) & 0x1Fu) << 24;
  return d;
}

// Instruction descriptor for kind::tf32 (a/b format 2 = TF32), fp32 accumulate, M = 128, both operands K-major.
__host__ __device__ __forceinline__ uint32_t umma_idesc_tf32(uint32_t M, uint32_t N) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 2u << 7;
  d |= 2u << 10;
  d |= ((N >> 3) & 0x3Fu) << 17;
  d |= ((M >> 4) & 0x1Fu) << 24;
  return d;
}

// ---------------------------------------------------------------- small helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Transposed butterfly: every lane holds 16 column values; afterwards lane l holds the 32-lane total of column
// col16_of_lane(l) (lanes l and l^1 hold the same column). 16 shuffles instead of 16*5.
__device__ __forceinline__ float warp_colsum16(const float (&v)[16], int lane) {
  float a8[8];
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float send = hi ? v[i] : v[i + 8];
      float keep = hi ? v[i + 8] : v[i];
      a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  float a4[4];
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float send = hi ? a8[i] : a8[i + 4];
      float keep = hi ? a8[i + 4] : a8[i];
      a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  float a2[2];
  {
    const bool hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float send = hi ? a4[i] : a4[i + 2];
      float keep = hi ? a4[i + 2] : a4[i];
      a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  float a1;
  {
    const bool hi = lane & 2;
    float send = hi ? a2[0] : a2[1];
    float keep = hi ? a2[1] : a2[0];
    a1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
  return a1;
}
__device__ __forceinline__ int col16_of_lane(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(t);
}
__device__ __forceinline__ void emit_bf16x16(__nv_bfloat16* dst, const float (&v)[16], bool accumulate) {
  const uint32_t w0 = pack_bf16x2(v[0], v[1]), w1 = pack_bf16x2(v[2], v[3]), w2 = pack_bf16x2(v[4], v[5]),
                 w3 = pack_bf16x2(v[6], v[7]), w4 = pack_bf16x2(v[8], v[9]), w5 = pack_bf16x2(v[10], v[11]),
                 w6 = pack_bf16x2(v[12], v[13]), w7 = pack_bf16x2(v[14], v[15]);
  if (accumulate) {
    red_add_bf16x8(dst, w0, w1, w2, w3);
    red_add_bf16x8(dst + 8, w4, w5, w6, w7);
  } else {
    st_global_256(dst, w0, w1, w2, w3, w4, w5, w6, w7);
  }
}

// Same for a destination whose channel pitch is only a multiple of 8 (dense 24 / 48-channel tensors): `nvalid` (0, 8 or
// 16) leading columns exist in the tensor, and the pixel address is 16-byte but not always 32-byte aligned, so plain
// stores go out as 128-bit halves unless `wide` (pitch % 16 == 0 channels, i.e. 32-byte aligned chunks).
__device__ __forceinline__ void st_global_128(void* p, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
  asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
}
__device__ __forceinline__ void emit_bf16x16_n(__nv_bfloat16* dst, const float (&v)[16], bool accumulate, int nvalid,
                                               bool wide) {
  if (nvalid <= 0) return;
  const uint32_t w0 = pack_bf16x2(v[0], v[1]), w1 = pack_bf16x2(v[2], v[3]), w2 = pack_bf16x2(v[4], v[5]),
                 w3 = pack_bf16x2(v[6], v[7]);
  if (nvalid >= 16) {
    const uint32_t w4 = pack_bf16x2(v[8], v[9]), w5 = pack_bf16x2(v[10], v[11]), w6 = pack_bf16x2(v[12], v[13]),
                   w7 = pack_bf16x2(v[14], v[15]);
    if (accumulate) {
      red_add_bf16x8(dst, w0, w1, w2, w3);
      red_add_bf16x8(dst + 8, w4, w5, w6, w7);
    } else if (wide) {
      st_global_256(dst, w0, w1, w2, w3, w4, w5, w6, w7);
    } else {
      st_global_128(dst, w0, w1, w2, w3);
      st_global_128(dst + 8, w4, w5, w6, w7);
    }
  } else {
    if (accumulate) red_add_bf16x8(dst, w0, w1, w2, w3);
    else st_global_128(dst, w0, w1, w2, w3);
  }
}

// Packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2 on sm_100): these streaming kernels need about half of the SM's issue
// slots at HBM speed, so halving the FMA / MUL / ADD instruction count is bandwidth (tools/bench_stream.py).
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
      "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
      "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
      "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
// bf16x2 word -> two fp32 (a bf16 is the upper half of an fp32: one shift, one mask)
__device__ __forceinline__ float2 bf2_to_f2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

}  // namespace mtbc
