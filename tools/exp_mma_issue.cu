// Experiment (not product code): tcgen05.mma issue rate.  How many cycles does one issuing thread need
// per MMA (M=128, N, K=16), and does issuing from several warps (separate accumulators) scale?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I yolosharp_b200/csrc -I include tools/exp_mma_issue.cu -o /tmp/exp_mma_issue
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_fp16.h>

#include "tc_ptx.cuh"

using namespace yb;

// mode 0: descriptors recomputed per MMA like conv_tc_kernel (runtime k loop)
// mode 1: fully unrolled, descriptor low words precomputed, 4 MMAs per iteration
// mode 2: the WHOLE warp runs the loop (warp-uniform descriptors -> uniform datapath), elect.sync picks the issuing lane
__global__ void __launch_bounds__(160, 1) issue_kernel(int N, int iters, int n_issuers, int mode, long long* out, int lay = 2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + (base - smem_u32(smem)))[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; i++) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (mode == 2 && warp < n_issuers) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t smA = base, smB = base + 16 * 1024;
    const uint32_t d = tmem + warp * 128;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
      // lay 2 = SWIZZLE_128B rows of 128 B (SBO 1024 B), 4 = 64B rows (SBO 512 B), 6 = 32B rows (SBO 256 B)
      const uint32_t sbo = lay == 2 ? 64 : (lay == 4 ? 32 : 16);
      const uint64_t ad = umma_desc(smA + (it & 3) * 128, sbo, lay);
      const uint64_t bd = umma_desc(smB, sbo, lay);
#pragma unroll
      for (int k = 0; k < 4; k++) umma_f16_elect(d, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (it | k) != 0);
    }
    const long long t1 = clock64();
    umma_commit_elect(smem_u32(&bars[warp]));
    mbar_wait(smem_u32(&bars[warp]), 0);
    const long long t2 = clock64();
    if (lane == 0) { out[warp * 2] = t1 - t0; out[warp * 2 + 1] = t2 - t0; }
  } else if (warp < n_issuers && lane == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t smA = base, smB = base + 16 * 1024;
    const uint32_t d = tmem + warp * 128;  // own accumulator columns (N <= 128)
    const long long t0 = clock64();
    if (mode == 0) {
      for (int it = 0; it < iters; it++) {
        const uint64_t ad = umma_desc(smA + (it & 3) * 128, 64, 2);
        const uint64_t bd = umma_desc(smB, 64, 2);
        const int kk = 4;
        for (int k = 0; k < kk; k++) umma_f16(d, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (it | k) != 0);
      }
    } else {
      const uint64_t ad = umma_desc(smA, 64, 2), bd = umma_desc(smB, 64, 2);
      umma_f16(d, ad, bd, idesc, 0);
      for (int it = 0; it < iters; it++) {
        umma_f16(d, ad, bd, idesc, 1);
        umma_f16(d, ad + 2, bd + 2, idesc, 1);
        umma_f16(d, ad + 4, bd + 4, idesc, 1);
        umma_f16(d, ad + 6, bd + 6, idesc, 1);
      }
    }
    const long long t1 = clock64();
    umma_commit(smem_u32(&bars[warp]));
    mbar_wait(smem_u32(&bars[warp]), 0);
    const long long t2 = clock64();
    out[warp * 2] = t1 - t0;
    out[warp * 2 + 1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 64);
  cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000;
  for (int mode : {0, 1, 2})
    for (int N : {16, 64, 128})
      for (int ni : {1, 2, 4}) {
        issue_kernel<<<1, 160, 50 * 1024>>>(N, iters, ni, mode, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[8];
        cudaMemcpy(h, dout, 64, cudaMemcpyDeviceToHost);
        const double n_mma = 4.0 * iters;
        printf("mode=%d N=%3d issuers=%d : issue %.1f cyc/MMA, issue+drain %.1f cyc/MMA per issuer; aggregate %.1f cyc/MMA (ideal tensor %.1f)\n",
               mode, N, ni, h[0] / n_mma, h[1] / n_mma, h[1] / (n_mma * ni), 128.0 * N * 16 / 3868.0);
      }
  for (int lay : {2, 4, 6})
    for (int N : {16, 80, 160}) {
      issue_kernel<<<1, 160, 50 * 1024>>>(N, iters, 1, 2, dout, lay);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[8];
      cudaMemcpy(h, dout, 64, cudaMemcpyDeviceToHost);
      printf("layout=%d (rows of %d B) N=%3d one warp, elect: %.1f cyc/MMA (K-slices of one row per MMA: k*32 B offsets)\n", lay,
             lay == 2 ? 128 : (lay == 4 ? 64 : 32), N, h[1] / (4.0 * iters));
    }
  return 0;
}
