// Experiment (not product code): does a UMMA shared-memory descriptor whose start address is shifted by whole rows
// (not 1024B aligned) and whose 8-row-group stride (SBO) is not a multiple of 1024B still read a TMA-written 128B/64B
// swizzled tile correctly?  Decides whether "halo tile in smem + 9 shifted descriptor views" is possible for 3x3 convs.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_swizzle tools/exp_swizzle.cu && ./exp_swizzle
#include "../multi_task_breast_cancer_b200/csrc/ptx.cuh"
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

using namespace mtbc;

struct Cfg {
  int kc;        // 64 (SW128) or 32 (SW64)
  int r0;        // first row of the A view
  int gs;        // rows between consecutive 8-row groups
  int base_mode; // 0: base_offset = 0, 1: base_offset = (start >> 7) & 7
  int mn_major;  // 0: K-major A (fwd); 1: MN-major A (wgrad-like: rows are K)
};

__global__ void __launch_bounds__(128) exp_kernel(const __grid_constant__ CUtensorMap amap,
                                                  const __grid_constant__ CUtensorMap bmap, Cfg c, float* out, int rows) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_full, s_accum;
  __shared__ uint32_t s_tmem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rowbytes = c.kc * 2;
  uint8_t* a_smem = smem;                      // rows x rowbytes, TMA written with swizzle
  uint8_t* b_smem = smem + 64 * 1024;          // B: 64 x kc (K-major) identity-like or [K rows][N] for MN-major
  if (tid == 0) { mbar_init(&s_full, 1); mbar_init(&s_accum, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&s_tmem, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = s_tmem;
  const int N = 64;
  if (tid == 0) {
    // A: box (kc, 256 rows max) in up to two loads of 128 rows; B: one box
    const int nload = (rows + 127) / 128;
    mbar_arrive_expect_tx(&s_full, nload * 128 * rowbytes + (c.mn_major ? 16 * N * 2 : N * rowbytes));
    for (int i = 0; i < nload; ++i) {
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(smem_u32(a_smem + i * 128 * rowbytes)), "l"(reinterpret_cast<uint64_t>(&amap)), "r"(smem_u32(&s_full)),
                     "r"(0), "r"(i * 128) : "memory");
    }
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(b_smem)), "l"(reinterpret_cast<uint64_t>(&bmap)), "r"(smem_u32(&s_full)), "r"(0), "r"(0) : "memory");
    mbar_wait(&s_full, 0);
    tc_fence_after();
    const uint32_t layout = c.kc == 64 ? 2u : 4u;
    if (!c.mn_major) {
      // D[m][n] = sum_k A[row(m)][k] * B[n][k], row(m) = r0 + (m/8)*gs + m%8
      const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
      for (int k = 0; k < c.kc / 16; ++k) {
        const uint32_t a_addr = smem_u32(a_smem) + c.r0 * rowbytes + k * 32;
        uint64_t da = umma_smem_desc(a_addr, 16, c.gs * rowbytes, layout);
        if (c.base_mode) da |= static_cast<uint64_t>((a_addr >> 7) & 7u) << 49;
        const uint64_t db = umma_smem_desc(smem_u32(b_smem) + k * 32, 16, 8 * rowbytes, layout);
        umma_bf16(tmem, da, db, idesc, k > 0);
      }
    } else {
      // MN-major A: K = 16 pixel rows r0..r0+15 taken as two 8-row groups gs rows apart; M = kc channels (<=64 used)
      // D[m][n] = sum_{j<16} A[row(j)][m] * B[j][n], B = 16 x 64 MN-major tile (rows = K) loaded with 128B rows
      const uint32_t idesc = umma_idesc_bf16(128, N, 1, 1);
      const uint32_t a_addr = smem_u32(a_smem) + c.r0 * rowbytes;
      uint64_t da = umma_smem_desc(a_addr, 16 * 1024, c.gs * rowbytes, layout);
      if (c.base_mode) da |= static_cast<uint64_t>((a_addr >> 7) & 7u) << 49;
      const uint64_t db = umma_smem_desc(smem_u32(b_smem), 16 * 1024, 8 * 128, 2u);
      umma_bf16(tmem, da, db, idesc, 0);
    }
    umma_commit(&s_accum);
  }
  __syncwarp();
  mbar_wait(&s_accum, 0);
  tc_fence_after();
  __syncwarp();
  const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  for (int cc = 0; cc < N; cc += 16) {
    float v[16];
    tmem_ld16(taddr + cc, v);
    for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * N + cc + i] = v[i];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void make2d(PFN_enc enc, CUtensorMap* m, void* p, int cols, int rows, int boxc, int boxr, CUtensorMapSwizzle sw) {
  cuuint64_t gd[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gs[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)boxc, (cuuint32_t)boxr};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

int main() {
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  PFN_enc enc = (PFN_enc)f;
  const int R = 256;
  cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  std::vector<Cfg> cfgs;
  for (int kc : {64, 32})
    for (int mn : {0, 1})
      for (int bm : {0, 1})
        for (auto rg : std::vector<std::pair<int, int>>{{0, 8}, {1, 8}, {3, 8}, {8, 8}, {0, 10}, {11, 10}, {0, 18}, {19, 18}})
          cfgs.push_back(Cfg{kc, rg.first, rg.second, bm, mn});
  for (const Cfg& c : cfgs) {
    if (c.mn_major && c.r0 + 8 + c.gs > R) continue;
    const int cols = c.kc;
    std::vector<__nv_bfloat16> ha(R * cols), hb(64 * 64);
    std::vector<float> fa(R * cols);
    srand(1);
    for (int i = 0; i < R * cols; ++i) { float v = float((rand() % 2001) - 1000) / 1000.f; ha[i] = __float2bfloat16(v); fa[i] = __bfloat162float(ha[i]); }
    // K-major: B = [64 n][kc k] with B[n][k] = (n == k); MN-major: B = [16 k rows][64 n] with B[j][n] = (n == j) (selects row j)
    int brows, bcols;
    if (!c.mn_major) { brows = 64; bcols = c.kc; } else { brows = 16; bcols = 64; }
    std::vector<__nv_bfloat16> hb2(brows * bcols);
    for (int r = 0; r < brows; ++r) for (int k = 0; k < bcols; ++k) hb2[r * bcols + k] = __float2bfloat16(r == k ? 1.f : 0.f);
    __nv_bfloat16 *da, *db; float* dout;
    cudaMalloc(&da, R * cols * 2); cudaMalloc(&db, brows * bcols * 2); cudaMalloc(&dout, 128 * 64 * 4);
    cudaMemcpy(da, ha.data(), R * cols * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb2.data(), brows * bcols * 2, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0, 128 * 64 * 4);
    CUtensorMap am, bm;
    make2d(enc, &am, da, cols, R, cols, 128, c.kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
    if (!c.mn_major) make2d(enc, &bm, db, bcols, brows, bcols, brows, c.kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
    else make2d(enc, &bm, db, bcols, brows, bcols, brows, CU_TENSOR_MAP_SWIZZLE_128B);
    exp_kernel<<<1, 128, 100 * 1024>>>(am, bm, c, dout, R);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kc%d mn%d r0=%d gs=%d bm=%d: CUDA error %s\n", c.kc, c.mn_major, c.r0, c.gs, c.base_mode, cudaGetErrorString(e)); return 1; }
    std::vector<float> ho(128 * 64);
    cudaMemcpy(ho.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    int bad = 0, tot = 0;
    if (!c.mn_major) {
      for (int m = 0; m < 128; ++m) {
        const int row = c.r0 + (m / 8) * c.gs + (m % 8);
        if (row >= R) continue;
        for (int n = 0; n < c.kc; ++n) { ++tot; if (fabsf(ho[m * 64 + n] - fa[row * cols + n]) > 1e-3f) ++bad; }
      }
    } else {
      // D[m][n] = A[row(n)][m] for n < 16, row(j) = r0 + (j/8)*gs + j%8 ; m < kc
      for (int m = 0; m < c.kc; ++m)
        for (int n = 0; n < 16; ++n) {
          const int row = c.r0 + (n / 8) * c.gs + (n % 8);
          ++tot; if (fabsf(ho[m * 64 + n] - fa[row * cols + m]) > 1e-3f) ++bad;
        }
    }
    printf("kc=%d %s r0=%2d gs=%2d base_offset=%s : %s (%d/%d wrong)\n", c.kc, c.mn_major ? "MN-major" : "K-major ", c.r0, c.gs,
           c.base_mode ? "(addr>>7)&7" : "0", bad == 0 ? "OK" : "BAD", bad, tot);
    cudaFree(da); cudaFree(db); cudaFree(dout);
  }
  return 0;
}
