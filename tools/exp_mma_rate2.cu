// Experiment (not product code): what bounds the tcgen05.mma rate for small N on sm_100a?
//   (a) SS operands, M = 64 vs 128 (is the A read from shared memory the limiter?)
//   (b) two issuing threads (warps 0 and 1) with their own accumulators (is single-thread issue the limiter?)
//   (c) A operand in TMEM (`[a_tmem]` form): rate when only B comes from shared memory
//   (d) tcgen05.cp 128x256b (smem -> TMEM) and tcgen05.shift rates
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp_mma_rate2 tools/exp_mma_rate2.cu
#include "../multi_task_breast_cancer_b200/csrc/ptx.cuh"
#include <cstdio>
#include <vector>
using namespace mtbc;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .b64 db;\n\t.reg .pred p;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(d), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint32_t lo, uint32_t hi) {
  asm volatile("{\n\t.reg .b64 d;\n\tmov.b64 d, {%1, %2};\n\ttcgen05.cp.cta_group::1.128x256b [%0], d;\n\t}" ::"r"(taddr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ void shift_down(uint32_t taddr) {
  asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(taddr) : "memory");
}

// mode 0: SS, one issuer; 1: SS, two issuers; 2: TS (A in TMEM); 3: cp only; 4: shift only; 5: per input row: 3 cp + 9 TS MMAs
template <int mode>
__global__ void __launch_bounds__(128) rate_kernel(int M, int N, int iters, long long* out, int kc = 64, int shift_rows = 0, int sbo_rows = 8, int a_mn = 0, int lbo = 16) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_bar[2];
  __shared__ uint32_t s_tmem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i % 7;
  if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&s_tmem, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const int nissue = (mode == 1) ? 2 : 1;
  if (warp < nissue && elect_one()) {
    const uint32_t rowb = kc * 2, layout = kc == 64 ? 2u : 4u;
    const uint32_t idesc = umma_idesc_bf16(M, N, a_mn, 0);
    const uint32_t a0 = smem_u32(smem) + warp * 16384, b0 = smem_u32(smem) + 48 * 1024;
    const uint32_t a_lo = umma_desc_lo(a0 + shift_rows * rowb, lbo), a_hi = umma_desc_hi(sbo_rows * rowb, layout);
    const uint32_t b_lo = umma_desc_lo(b0, 16), b_hi = umma_desc_hi(8 * rowb, layout);
    const uint32_t tm = s_tmem + warp * 256;   // accumulators: columns [0,128) of this issuer's half; A staging: [128, 256)
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 9) {
      if (mode == 5) { for (int c = 0; c < 3; ++c) cp_128x256b(tm + 320 + c * 8, a_lo + c * 8, a_hi); }
#pragma unroll
      for (int u = 0; u < 9; ++u) {
        if (mode <= 1) umma_bf16_lohi(tm + (u & 1) * 64, a_lo + (u & 3) * 2, a_hi, b_lo + (u & 3) * 2, b_hi, idesc, 1);
        else if (mode == 2) mma_ts(tm + (u & 1) * 64, tm + 320 + (u & 3) * 8, b_lo + (u & 3) * 2, b_hi, idesc, 1);
        else if (mode == 3) cp_128x256b(tm + 320 + (u & 3) * 8, a_lo + (u & 3) * 2, a_hi);
        else if (mode == 4) shift_down(tm + 320 + (u & 3) * 8);
        else mma_ts(tm + (u % 3) * 64, tm + 320 + (u / 3) * 8, b_lo + u * 2, b_hi, idesc, 1);
      }
    }
    umma_commit(&s_bar[warp]);
    mbar_wait(&s_bar[warp], 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && warp == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(s_tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 4500;
  struct C { int mode, M, N; const char* note; int kc = 64, shift = 0, sbo = 8, amn = 0, lbo = 16; };
  std::vector<C> cs = {
      {0, 128, 32, "SS kc=32 canonical", 32, 0, 8}, {0, 128, 32, "SS kc=32 halo view shift 11 sbo 10", 32, 11, 10},
      {0, 128, 32, "SS kc=64 halo view shift 11 sbo 10", 64, 11, 10}, {0, 128, 32, "SS kc=32 halo view shift 1 sbo 10", 32, 1, 10},
      {0, 128, 32, "SS kc=32 halo view shift 0 sbo 10", 32, 0, 10}, {0, 128, 64, "SS N=64 kc=64 halo shift 11", 64, 11, 10},
      {0, 128, 32, "SS MN-major A (wgrad halo) kc=32 lbo=64 sbo=10 rows", 32, 10, 10, 1, 64},
      {0, 128, 32, "SS MN-major A canonical kc=64 lbo=16384", 64, 0, 8, 1, 16384},
      {0, 128, 32, "SS kc=32 G=2 view: shift 11 sbo 20", 32, 11, 20}, {0, 128, 64, "SS N=64 kc=32 G=2 view: shift 11 sbo 20", 32, 11, 20},
      {0, 128, 64, "SS N=64 kc=32 shift 11 sbo 10", 32, 11, 10}, {0, 128, 64, "SS N=64 kc=32 canonical", 32, 0, 8},
      {0, 128, 64, "SS N=64 kc=32 G=2 view: shift 31 sbo 20", 32, 31, 20}, {0, 128, 32, "SS N=32 kc=32 G=2 view: shift 30 sbo 20", 32, 30, 20},
      {0, 128, 32, "SS 1 issuer"}, {0, 64, 32, "SS 1 issuer M=64"}, {0, 64, 64, "SS 1 issuer M=64"}, {0, 128, 16, "SS N=16"},
      {0, 128, 64, "SS 1 issuer"}, {0, 128, 96, "SS 1 issuer"}, {0, 128, 128, "SS"}, {0, 128, 192, "SS"},
      {1, 128, 32, "SS 2 issuers (per-issuer MMAs)"}, {1, 128, 64, "SS 2 issuers"},
      {2, 128, 32, "TS (A in TMEM)"}, {2, 128, 64, "TS"}, {2, 128, 16, "TS"}, {2, 128, 128, "TS"},
      {3, 128, 32, "tcgen05.cp 128x256b only"}, {4, 128, 32, "tcgen05.shift only"},
      {5, 128, 32, "3 cp + 9 TS MMA per group (cycles per MMA)"}, {5, 128, 64, "3 cp + 9 TS MMA per group"},
  };
  for (const C& c : cs) {
    switch (c.mode) {
#define RUN(m) case m: cudaFuncSetAttribute(rate_kernel<m>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024); rate_kernel<m><<<148, 128, 120 * 1024>>>(c.M, c.N, iters, d, c.kc, c.shift, c.sbo, c.amn, c.lbo); break;
      RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d M=%d N=%d: error %s\n", c.mode, c.M, c.N, cudaGetErrorString(e)); return 1; }
    long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    const double per = double(cyc) / iters;
    printf("mode=%d M=%3d N=%3d %-45s: %7.1f cycles/op  -> %6.0f MAC/clk/SM (per issuer)\n", c.mode, c.M, c.N, c.note, per,
           double(c.M) * c.N * 16 / per);
    fflush(stdout);
  }
  return 0;
}
