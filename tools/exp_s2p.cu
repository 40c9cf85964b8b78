// Experiment (not product code): "pair rows" for stride-2 convolutions.  Two horizontally adjacent NHWC
// pixels (contiguous in memory when the view covers the whole buffer pitch) are loaded as ONE operand row of
// 2*BK channels with the swizzle of that width.  Does a tcgen05 A descriptor with a K-slice offset of
// `half`*BK*2 bytes and a row shift address pixel 2*row + half?
// (First attempt: a box whose inner dimension is half the swizzle span - TMA pads such rows to the span,
//  the tile is not dense; all 100 cases mismatched.)
//
// Setup: src[R pixels][BK] fp16 (random) viewed as [R/2][2*BK], 2-D TMA loads {2*BK, 96 pairs}.
// MMA: D[128][BK] = A x I (B = identity, own layout), A start = base + shift*pairbytes + half*BK*2,
// SBO = sbo_rows*pairbytes.   Expected: D[m][:] == src[((m/8)*sbo_rows + (m%8) + shift)*2 + half][:].
//
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I yolosharp_b200/csrc -I include tools/exp_s2p.cu -o /tmp/exp_s2p
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_fp16.h>

#include "tc_ptx.cuh"

using namespace yb;

struct Args {
  CUtensorMap tmA, tmB;
  float* out;
  int BK, R, shift, sbo_rows, half;
  uint32_t layout_a, layout_b;
};

__global__ void __launch_bounds__(128, 1) exp_kernel(const __grid_constant__ Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t rowb = a.BK * 2;
  const uint32_t smA = base, smB = base + 64 * 1024;
  const uint32_t bar0 = smem_u32(&bars[0]), bar1 = smem_u32(&bars[1]);
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar0, a.R * rowb + a.BK * rowb);
    tma_load_2d(smA, &a.tmA, bar0, 0, 0);
    tma_load_2d(smA + 192 * rowb, &a.tmA, bar0, 0, 96);
    tma_load_2d(smB, &a.tmB, bar0, 0, 0);
    mbar_wait(bar0, 0);
    tc_fence_after();
    const uint32_t idesc = (1u << 4) | ((uint32_t)(a.BK >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t start = smA + a.shift * 2 * rowb + a.half * rowb;
    const uint64_t ad = umma_desc(start, (a.sbo_rows * 2 * rowb) >> 4, a.layout_a);
    const uint64_t bd = umma_desc(smB, (8 * rowb) >> 4, a.layout_b);
    for (int k = 0; k < a.BK / 16; k++) umma_f16(tmem, ad + 2 * k, bd + 2 * k, idesc, k != 0);
    umma_commit(bar1);
  }
  mbar_wait(bar1, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < a.BK; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; j++) a.out[row * a.BK + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeTiledFn encode = (EncodeTiledFn)fp;
  const int R = 384;  // two boxes of 192 pixels (box dimensions are limited to 256)
  CK(cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  for (int BK : {32, 16}) {
    std::vector<__half> src((size_t)R * BK), eye((size_t)BK * BK);
    unsigned s = 12345u + BK;
    for (auto& v : src) { s = s * 1664525u + 1013904223u; v = __float2half((float)((int)((s >> 16) % 4001) - 2000) / 16.0f); }
    for (int i = 0; i < BK; i++) for (int j = 0; j < BK; j++) eye[(size_t)i * BK + j] = __float2half(i == j ? 1.f : 0.f);
    __half *dsrc, *deye; float* dout;
    CK(cudaMalloc(&dsrc, src.size() * 2)); CK(cudaMalloc(&deye, eye.size() * 2)); CK(cudaMalloc(&dout, 128 * BK * 4));
    CK(cudaMemcpy(dsrc, src.data(), src.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(deye, eye.data(), eye.size() * 2, cudaMemcpyHostToDevice));
    Args a; memset(&a, 0, sizeof(a));
    const CUtensorMapSwizzle swz = BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : BK == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    const CUtensorMapSwizzle swz_a = BK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    a.layout_b = BK == 32 ? 4 : 6;
    a.layout_a = BK == 32 ? 2 : 4;
    cuuint64_t gd[2] = {(cuuint64_t)BK * 2, (cuuint64_t)R / 2}, gsa[1] = {(cuuint64_t)BK * 4}, gs[1] = {(cuuint64_t)BK * 2};
    cuuint32_t bx[2] = {(cuuint32_t)BK * 2, 96u}, es[2] = {1, 1};
    CUresult cr = encode(&a.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dsrc, gd, gsa, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz_a,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t gd2[2] = {(cuuint64_t)BK, (cuuint64_t)BK};
    cuuint32_t bx2[2] = {(cuuint32_t)BK, (cuuint32_t)BK};
    CUresult cr2 = encode(&a.tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, deye, gd2, gs, bx2, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr || cr2) { printf("encode failed %d %d\n", (int)cr, (int)cr2); return 1; }
    a.out = dout; a.BK = BK; a.R = R;
    std::vector<float> out((size_t)128 * BK);
    for (int sbo_rows : {8, 10, 18}) {
      for (int shift : {0, 1, 2, 3, 7, 8, 10, 11, 12, 20, 21, 22, 37}) {
        if ((15 * sbo_rows + 7 + shift) * 2 + 1 >= R) continue;
        for (int half : {0, 1}) {
          a.sbo_rows = sbo_rows; a.shift = shift; a.half = half;
          CK(cudaMemset(dout, 0, out.size() * 4));
          exp_kernel<<<1, 128, 66 * 1024 + BK * BK * 2 + 2048>>>(a);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("BK=%d sbo=%d shift=%d half=%d : kernel error %s\n", BK, sbo_rows, shift, half, cudaGetErrorString(e)); return 1; }
          CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
          int bad = 0, bad_rows = 0;
          for (int m = 0; m < 128; m++) {
            const int r = ((m / 8) * sbo_rows + (m % 8) + shift) * 2 + half;
            int rb = 0;
            for (int c = 0; c < BK; c++) if (out[(size_t)m * BK + c] != __half2float(src[(size_t)r * BK + c])) rb++;
            bad += rb; bad_rows += rb != 0;
          }
          printf("BK=%2d sbo_pairs=%2d shift=%2d half=%d : %s (%d bad values in %d rows)\n", BK, sbo_rows, shift, half,
                 bad ? "MISMATCH" : "ok", bad, bad_rows);
        }
      }
    }
    cudaFree(dsrc); cudaFree(deye); cudaFree(dout);
  }
  return 0;
}
