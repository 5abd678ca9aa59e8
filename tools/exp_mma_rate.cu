// Experiment (not product code): issue rate of tcgen05.mma kind::f16 (SS operands) as a function of M, N and the smem
// layout, one CTA per SM, operands resident in shared memory.  Decides how costly the small-N (24/48 output channel)
// layers are on the tensor pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp_mma_rate tools/exp_mma_rate.cu
#include "../multi_task_breast_cancer_b200/csrc/ptx.cuh"
#include <cstdio>
#include <vector>
using namespace mtbc;

__global__ void __launch_bounds__(128) rate_kernel(int M, int N, int kc, int iters, int a_shift_rows, int a_sbo_rows,
                                                   int a_mn, int nacc, int accum, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i % 7;
  if (tid == 0) { mbar_init(&s_bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&s_tmem, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (tid == 0) {
    const uint32_t rowb = kc * 2, layout = kc == 64 ? 2u : 4u;
    uint32_t idesc = umma_idesc_bf16(M, N, a_mn, 0);
    const uint32_t a0 = smem_u32(smem) + a_shift_rows * rowb, b0 = smem_u32(smem) + 48 * 1024;
    const uint64_t da0 = umma_smem_desc(a0, 16, a_sbo_rows * rowb, layout);
    const uint64_t db0 = umma_smem_desc(b0, 16, 8 * rowb, layout);
    const uint32_t a_lo = (uint32_t)da0, a_hi = (uint32_t)(da0 >> 32), b_lo = (uint32_t)db0, b_hi = (uint32_t)(db0 >> 32);
    const uint32_t tm = s_tmem;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        asm volatile(
            "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
            "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
            "setp.ne.b32 p, %6, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
            ::"r"(tm + (u & (nacc - 1)) * (512 / nacc)), "r"(a_lo + (u & 3) * 2), "r"(a_hi), "r"(b_lo + (u & 3) * 2), "r"(b_hi), "r"(idesc), "r"(accum)
            : "memory");
      }
    }
    umma_commit(&s_bar);
    mbar_wait(&s_bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(s_tmem, 512); }
}

int main() {
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  long long* d; cudaMalloc(&d, 8);
  const int iters = 4000;
  struct C { int M, N, kc, shift, sbo, amn; const char* note; int nacc = 2, accum = 1; };
  std::vector<C> cs = {
      {128, 32, 32, 0, 8, 0, "K-major canonical"}, {128, 32, 64, 0, 8, 0, "K-major canonical"},
      {128, 32, 32, 11, 10, 0, "K-major halo view"}, {128, 32, 64, 11, 10, 0, "K-major halo view"},
      {128, 64, 64, 0, 8, 0, ""}, {128, 64, 64, 11, 10, 0, "halo view"}, {128, 128, 64, 0, 8, 0, ""},
      {128, 256, 64, 0, 8, 0, ""}, {64, 32, 64, 0, 8, 0, "M=64"}, {64, 64, 64, 0, 8, 0, "M=64"}, {64, 256, 64, 0, 8, 0, "M=64"},
      {128, 16, 64, 0, 8, 0, ""}, {128, 32, 32, 10, 10, 1, "MN-major A, halo view, LBO=1 row (wgrad)"},
      {128, 128, 32, 10, 10, 1, "MN-major A halo (wgrad), N=128"},
  };
  cs.clear();
  for (int N : {32, 64, 128, 256})
    for (int nacc : {1, 2, 8})
      for (int accum : {0, 1}) {
        if (nacc * N > 512) continue;
        C c{128, N, 64, 0, 8, 0, "", nacc, accum};
        cs.push_back(c);
      }
  for (const C& c : cs) {
    for (int grid : {148}) {
      rate_kernel<<<grid, 128, 120 * 1024>>>(c.M, c.N, c.kc, iters, c.shift, c.sbo, c.amn, c.nacc, c.accum, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      const double per = double(cyc) / iters;
      printf("M=%3d N=%3d kc=%2d grid=%3d nacc=%d accumulate=%d %-10s: %7.1f cycles/MMA  -> %6.0f MAC/clk/SM\n", c.M, c.N, c.kc, grid, c.nacc, c.accum, c.note, per,
             double(c.M) * c.N * 16 / per);
    }
  }
  return 0;
}
