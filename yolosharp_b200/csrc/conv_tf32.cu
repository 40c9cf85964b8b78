// Tensor-core convolutions of the TRAINING step (fp32 storage, TF32 tcgen05 MMAs, fp32 accumulate in TMEM).
//
// Replaces, on the training path, what libtorch runs behind `Conv2d.forward` and `loss.backward()` for the Conv
// blocks of the reference (Modules/Convs.cs:44; Utils/Amp.cs:260-286 calls forward / backward / optimizer.step):
// forward, data gradient and weight gradient of a dense k x k convolution (k in {1, 3}, stride in {1, 2}), on the
// NHWC fp32 activations the training step keeps.  libtorch's own CUDA convolutions run TF32 tensor-core math by
// default (cudnn.allow_tf32), so TF32 products with fp32 accumulation are the reference's arithmetic class here;
// the fp32 CUDA-core kernels (yb_conv_forward_f32 / yb_conv_backward_*) stay as the parity twins these are tested
// against.
//
//   tf_conv_kernel   one implicit-GEMM kernel for forward AND data gradient.  M = 128 output positions (a BW x BH
//                    rectangle of one image, or 128 consecutive positions of the flattened batch for 1x1), N = output
//                    channels (tile <= 256), K = taps x input channels.  Both operands K-major: A = NHWC activations
//                    (one 4-D TMA box per tap and 32-channel slab, zero fill = padding, traversal stride = conv
//                    stride), B = weights re-packed per step as [tap][N][K] (3-D TMA box).  A launch is described
//                    by a TAP TABLE (box offset dh, dw + weight slab per tap), which covers
//                      forward          out(y,x) = sum_t in(y*s + kh - p, x*s + kw - p) W[kh][kw]
//                      dgrad, stride 1  dx(y,x)  = sum_t dz(y + p - kh, x + p - kw) W^T[kh][kw]
//                      dgrad, stride 2  four launches, one per output parity (py, px), each with the taps whose
//                                       (py + p - kh) and (px + p - kw) are even, writing every second position
//   tf_wgrad_kernel  dW[co][tap][ci] = sum_pixels dz[pix][co] * x[pix + tap][ci]: K = pixels, so both operands are
//                    MN-major - exactly the NHWC tiles TMA delivers (64 pixels x 32 channels, 128-byte rows, written
//                    with SWIZZLE_128B_ATOM_32B: the one shared-memory layout tcgen05 accepts for MN-major 32-bit
//                    operands, UMMA layout type SWIZZLE_128B_BASE32B).  One CTA
//                    owns 128 output channels x (taps x 32 nb) input channels in TMEM (<= 512 columns), walks its share
//                    of the pixel tiles (split-K over pixels) and stores a partial; a fixed-order fold sums the
//                    partials into the checkpoint layout (deterministic, no atomics).
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace yb {

typedef CUresult (*TfEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TfEncodeFn tf_encode_fn() {
  static TfEncodeFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return nullptr;
  }
  fn = (TfEncodeFn)p;
  return fn;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ uint64_t tf_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16, uint32_t layout) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)lbo16 << 16) | ((uint64_t)sbo16 << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)layout << 61);
}

constexpr int TF_MAX_STAGES = 8;
constexpr int TF_THREADS = 192;  // warp 0 TMA producer, warp 1 MMA issuer (+ TMEM alloc), warps 2-5 epilogue

// ------------------------------------------------------------------------------------------
// forward / data gradient
// ------------------------------------------------------------------------------------------
struct TfArgs {
  CUtensorMap tmA, tmB;
  float* out;
  const float* bias;                // [n_out] or nullptr
  long long o_img, o_row, o_pix;    // element strides of the output addressing
  long long o_off;
  int n_out;                        // valid output channels (columns >= n_out are not stored)
  int Ho, Wo, imgs;                 // extent of the output position grid the tiles cover
  int tiles_w, tiles_h, n_tiles, n_tile, total_tiles;
  int BW, BH;
  int in_stride;                    // A box origin = (w0 * in_stride + dw, h0 * in_stride + dh)
  int ntaps;
  int dh[9], dw[9], slab[9];
  int chunks, KK;                   // K slabs per tap, MMAs (K = 8) per slab
  int BK;
  int stages;
  uint32_t a_stride, b_stride, a_bytes, b_bytes;
  uint32_t layout, sbo16;
  uint32_t tmem_cols;
};

__global__ void __launch_bounds__(TF_THREADS, 2) tf_conv_kernel(const __grid_constant__ TfArgs a) {
  // the barrier / TMEM setup below overlaps the tail of the previous kernel (launch_pdl); pdl_wait() before any global access
  extern __shared__ __align__(1024) uint8_t tf_smem[];
  __shared__ __align__(8) uint64_t bars[2 * TF_MAX_STAGES + 4];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem0 = (smem_u32(tf_smem) + 1023u) & ~1023u;
  const uint32_t smemA = smem0, smemB = smem0 + a.stages * a.a_stride;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[TF_MAX_STAGES]);
  const uint32_t tfull0 = smem_u32(&bars[2 * TF_MAX_STAGES]), tempty0 = smem_u32(&bars[2 * TF_MAX_STAGES + 2]);
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int s = 0; s < 2; s++) { mbar_init(tfull0 + 8 * s, 1); mbar_init(tempty0 + 8 * s, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  pdl_trigger();
  const uint32_t tmem_base = tmem_slot;
  const int tiles_per_img = a.tiles_w * a.tiles_h;
  const int ksteps = a.ntaps * a.chunks;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&a.tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&a.tmB) : "memory");
      int st = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        const int mt = tile / a.n_tiles, nt = tile - mt * a.n_tiles;
        const int img = mt / tiles_per_img, r = mt - img * tiles_per_img;
        const int th = r / a.tiles_w, tw = r - th * a.tiles_w;
        const int wbase = tw * a.BW * a.in_stride, hbase = th * a.BH * a.in_stride;
        for (int t = 0; t < a.ntaps; t++)
          for (int ch = 0; ch < a.chunks; ch++) {
            mbar_wait(empty0 + 8 * st, ph ^ 1);
            mbar_arrive_expect_tx(full0 + 8 * st, a.a_bytes + a.b_bytes);
            tma_load_4d(smemA + st * a.a_stride, &a.tmA, full0 + 8 * st, ch * a.BK, wbase + a.dw[t], hbase + a.dh[t], img);
            tma_load_3d(smemB + st * a.b_stride, &a.tmB, full0 + 8 * st, ch * a.BK, nt * a.n_tile, a.slab[t]);
            if (++st == a.stages) { st = 0; ph ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D fp32 (bit 4), A / B format TF32 (2 at bits 7 and 10), K-major both, N >> 3 at 17, M >> 4 at 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.n_tile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int st = 0, li = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, li++) {
        const int acc = li & 1;
        mbar_wait(tempty0 + 8 * acc, ((uint32_t)(li >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * a.n_tile;
        uint32_t accf = 0;
        for (int ks = 0; ks < ksteps; ks++) {
          mbar_wait(full0 + 8 * st, ph);
          tc_fence_after();
          const uint32_t sa = smemA + st * a.a_stride, sb = smemB + st * a.b_stride;
          for (int k = 0; k < a.KK; k++) {  // 8 fp32 = 32 bytes per K step inside the swizzled row
            umma_tf32(d_tmem, tf_desc(sa + 32 * k, 1, a.sbo16, a.layout), tf_desc(sb + 32 * k, 1, a.sbo16, a.layout), idesc, accf);
            accf = 1;
          }
          umma_commit(empty0 + 8 * st);
          if (++st == a.stages) { st = 0; ph ^= 1; }
        }
        umma_commit(tfull0 + 8 * acc);
      }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    int li = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, li++) {
      const int acc = li & 1;
      const int mt = tile / a.n_tiles, nt = tile - mt * a.n_tiles;
      const int img = mt / tiles_per_img, r = mt - img * tiles_per_img;
      const int th = r / a.tiles_w, tw = r - th * a.tiles_w;
      const int hl = row / a.BW, wl = row - hl * a.BW;
      const int ho = th * a.BH + hl, wo = tw * a.BW + wl;
      const bool valid = hl < a.BH && ho < a.Ho && wo < a.Wo;
      const int n0 = nt * a.n_tile;
      float* orow = a.out + (long long)img * a.o_img + (long long)ho * a.o_row + (long long)wo * a.o_pix + a.o_off + n0;
      mbar_wait(tfull0 + 8 * acc, (uint32_t)(li >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * a.n_tile;
      for (int c0 = 0; c0 < a.n_tile; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int g = 0; g < 4; g++) {
            const int c = n0 + c0 + g * 4;
            if (c < a.n_out) {  // n_out is a multiple of 4
              float4 o = make_float4(__uint_as_float(v[g * 4]), __uint_as_float(v[g * 4 + 1]), __uint_as_float(v[g * 4 + 2]),
                                     __uint_as_float(v[g * 4 + 3]));
              if (a.bias) {
                const float4 b = *reinterpret_cast<const float4*>(a.bias + c);
                o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
              }
              *reinterpret_cast<float4*>(orow + c0 + g * 4) = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
}

// w (Cout, Cin, k, k) checkpoint layout -> wf [tap][Cout][Cin] (forward B operand: rows = output channels, K = Cin)
//                                        -> wb [tap][Cin][Cout] (dgrad B operand: rows = input channels, K = Cout)
__global__ void tf_pack_weights_kernel(const float* __restrict__ w, float* __restrict__ wf, float* __restrict__ wb, int Cout,
                                       int Cin, int taps) {
  const long long n = (long long)Cout * Cin * taps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps);
    const long long q = i / taps;
    const int ci = (int)(q % Cin), co = (int)(q / Cin);
    const float v = w[i];
    if (wf) wf[((size_t)t * Cout + co) * Cin + ci] = v;
    if (wb) wb[((size_t)t * Cin + ci) * Cout + co] = v;
  }
}

struct TfPackDesc { long long off, chunk0; int cout, cin, taps, pad_; };  // off: element offset in the flat buffers; chunk0: first block
// Every dense conv weight of a model in ONE launch (the native training step packs once per step, after the optimizer, instead
// of once per conv call: 158 launches of the kernel above in a YOLOv11s step).  wf / wb use the SAME element offsets as the
// checkpoint-layout tensors inside their flat buffer, so a layer's packed operands sit at WF + off and WB + off.
constexpr int TF_PACK_CHUNK = 4096;
__global__ void __launch_bounds__(256) tf_pack_all_kernel(const float* __restrict__ P, float* __restrict__ WF, float* __restrict__ WB,
                                                          const TfPackDesc* __restrict__ d, int nd) {
  int lo = 0, hi = nd - 1;  // last layer whose first chunk <= blockIdx.x
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (d[mid].chunk0 <= (long long)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const TfPackDesc L = d[lo];
  const long long n = (long long)L.cout * L.cin * L.taps;
  const long long i0 = ((long long)blockIdx.x - L.chunk0) * TF_PACK_CHUNK;
  const float* w = P + L.off;
  float* wf = WF + L.off;
  float* wb = WB + L.off;
  for (long long i = i0 + threadIdx.x; i < min(n, i0 + TF_PACK_CHUNK); i += 256) {
    const int t = (int)(i % L.taps);
    const long long q = i / L.taps;
    const int ci = (int)(q % L.cin), co = (int)(q / L.cin);
    const float v = w[i];
    wf[((size_t)t * L.cout + co) * L.cin + ci] = v;
    wb[((size_t)t * L.cin + ci) * L.cout + co] = v;
  }
}
long long tf_pack_chunks(int cout, int cin, int taps) { return ((long long)cout * cin * taps + TF_PACK_CHUNK - 1) / TF_PACK_CHUNK; }
int tf_pack_all(const float* P, float* WF, float* WB, const TfPackDesc* dev_descs, int nd, long long total_chunks, cudaStream_t s) {
  if (nd <= 0 || total_chunks <= 0) return YB_OK;
  tf_pack_all_kernel<<<(unsigned)total_chunks, 256, 0, s>>>(P, WF, WB, dev_descs, nd);
  YB_CUDA_CHECK(cudaGetLastError());
  return YB_OK;
}

static int tf_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

// One launch of tf_conv_kernel.  in: NHWC (N, Hi, Wi, Kc) fp32; wpk: [taps][Nc][Kc] fp32; out addressing given by strides.
struct TfLaunch {
  const float* in; int N, Hi, Wi, Kc;
  int in_pitch;      // elements between consecutive pixels of `in` (>= Kc: `in` may be a channel slice of a wider NHWC buffer)
  const float* wpk; int Nc, taps_total;
  float* out; long long o_img, o_row, o_pix, o_off;
  const float* bias;
  int Ho, Wo;        // output position grid
  int in_stride;
  int ntaps; int dh[9], dw[9], slab[9];
  bool flat;         // 1x1 stride 1, dense output: flatten the batch into one position dimension
};

static int tf_conv_launch(const TfLaunch& L, cudaStream_t s) {
  TfEncodeFn encode = tf_encode_fn();
  if (!encode) { set_error("cuTensorMapEncodeTiled entry point not found"); return YB_ERR_CUDA; }
  TfArgs a;
  memset(&a, 0, sizeof(a));
  a.out = L.out; a.bias = L.bias;
  a.o_img = L.o_img; a.o_row = L.o_row; a.o_pix = L.o_pix; a.o_off = L.o_off;
  a.n_out = L.Nc;
  a.in_stride = L.in_stride;
  a.ntaps = L.ntaps;
  for (int t = 0; t < L.ntaps; t++) { a.dh[t] = L.dh[t]; a.dw[t] = L.dw[t]; a.slab[t] = L.slab[t]; }
  a.BK = L.Kc >= 32 ? 32 : (L.Kc >= 16 ? 16 : 8);
  a.chunks = (L.Kc + a.BK - 1) / a.BK;
  a.KK = a.BK / 8;
  const uint32_t row_bytes = a.BK * 4;
  a.layout = a.BK == 32 ? 2 : (a.BK == 16 ? 4 : 6);
  a.sbo16 = (8 * row_bytes) >> 4;
  const CUtensorMapSwizzle swz = a.BK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : (a.BK == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  a.n_tile = std::min(256, (L.Nc + 15) / 16 * 16);
  a.n_tiles = (L.Nc + a.n_tile - 1) / a.n_tile;
  // (Tried: narrower N tiles for the 20 x 20 / 40 x 40 levels so that every SM gets a tile - 4.8 -> 5.1 ms over the step's
  //  177 launches: the extra activation re-reads and shorter MMAs cost more than the idle SMs, profiles/r2_ncu_tf_conv.txt.)
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t box[4], estr[4];
  if (L.flat) {
    const cuuint64_t npix = (cuuint64_t)L.N * L.Hi * L.Wi;
    gdim[0] = L.Kc; gdim[1] = npix; gdim[2] = 1; gdim[3] = 1;
    gstr[0] = (cuuint64_t)(L.in_pitch ? L.in_pitch : L.Kc) * 4; gstr[1] = gstr[0] * npix; gstr[2] = gstr[1];
    a.BW = 128; a.BH = 1;
    a.imgs = 1; a.Ho = 1; a.Wo = (int)npix;
    box[0] = a.BK; box[1] = 128; box[2] = 1; box[3] = 1;
    estr[0] = estr[1] = estr[2] = estr[3] = 1;
  } else {
    gdim[0] = L.Kc; gdim[1] = L.Wi; gdim[2] = L.Hi; gdim[3] = L.N;
    gstr[0] = (cuuint64_t)(L.in_pitch ? L.in_pitch : L.Kc) * 4; gstr[1] = gstr[0] * L.Wi; gstr[2] = gstr[1] * L.Hi;
    a.imgs = L.N; a.Ho = L.Ho; a.Wo = L.Wo;
    double best = -1;
    for (int bw = 1; bw <= std::min(L.Wo, 128); bw++) {
      const int bh = std::min(L.Ho, 128 / bw);
      if (bw * L.in_stride > 256 || bh * L.in_stride > 256) continue;
      const double tiles = (double)((L.Wo + bw - 1) / bw) * ((L.Ho + bh - 1) / bh);
      const double eff = (double)L.Wo * L.Ho / (tiles * 128.0);
      if (eff > best + 1e-9 || (eff > best - 1e-9 && bw > a.BW)) { best = eff; a.BW = bw; a.BH = bh; }
    }
    box[0] = a.BK; box[1] = a.BW * L.in_stride; box[2] = a.BH * L.in_stride; box[3] = 1;
    estr[0] = 1; estr[1] = L.in_stride; estr[2] = L.in_stride; estr[3] = 1;
  }
  CUresult cr = encode(&a.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(L.in), gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { set_error("tf32 conv: cuTensorMapEncodeTiled(A) failed with code " + std::to_string((int)cr)); return YB_ERR_CUDA; }
  {
    cuuint64_t bd[3] = {(cuuint64_t)L.Kc, (cuuint64_t)L.Nc, (cuuint64_t)L.taps_total};
    cuuint64_t bs[2] = {(cuuint64_t)L.Kc * 4, (cuuint64_t)L.Kc * 4 * L.Nc};
    cuuint32_t bb[3] = {(cuuint32_t)a.BK, (cuuint32_t)a.n_tile, 1};
    cuuint32_t be[3] = {1, 1, 1};
    cr = encode(&a.tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(L.wpk), bd, bs, bb, be, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("tf32 conv: cuTensorMapEncodeTiled(B) failed with code " + std::to_string((int)cr)); return YB_ERR_CUDA; }
  }
  a.tiles_w = (a.Wo + a.BW - 1) / a.BW;
  a.tiles_h = (a.Ho + a.BH - 1) / a.BH;
  a.total_tiles = a.imgs * a.tiles_w * a.tiles_h * a.n_tiles;
  a.a_bytes = (uint32_t)(a.BW * a.BH) * row_bytes;  // the box has BW*BH <= 128 rows; rows past it are never stored
  a.b_bytes = (uint32_t)a.n_tile * row_bytes;
  a.a_stride = (128 * row_bytes + 1023) / 1024 * 1024;
  a.b_stride = (a.n_tile * row_bytes + 1023) / 1024 * 1024;
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * a.n_tile)) cols <<= 1;
  a.tmem_cols = cols;
  // Two CTAs per SM when there are tiles for them and a CTA fits half an SM (<= 256 TMEM columns, a ring of >= 3 stages in
  // ~100 KiB): with one producer thread, one MMA thread and four epilogue warps per CTA the kernel is latency bound
  // (ncu: issue active 10 - 18 %, long-scoreboard / wait stalls, profiles/r2_ncu_tf_conv_big.txt); a second CTA fills
  // the other's load -> MMA -> epilogue bubbles, as in conv_tc_kernel.
  static const bool occ1 = getenv("YB_TF_OCC1") != nullptr;
  const size_t per_stage = (size_t)a.a_stride + a.b_stride;
  int occ = 1;
  if (!occ1 && cols <= 256 && (size_t)(100 * 1024) / per_stage >= 3 && a.total_tiles >= 2 * tf_num_sms()) occ = 2;
  a.stages = (int)std::min<size_t>(TF_MAX_STAGES, (size_t)((occ == 2 ? 100 : 190) * 1024) / per_stage);
  if (a.stages < 2) { set_error("tf32 conv: tile does not fit in shared memory"); return YB_ERR_SHAPE; }
  const size_t smem = (size_t)a.stages * (a.a_stride + a.b_stride) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    YB_CUDA_CHECK(cudaFuncSetAttribute(tf_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const int grid = std::min(a.total_tiles, occ * tf_num_sms());
  YB_CUDA_CHECK(launch_pdl(tf_conv_kernel, dim3(grid), dim3(TF_THREADS), smem, s, a));
  return 0;
}

size_t tf_conv_workspace_bytes(int N, int H, int W, int Cin, int Cout, int k, int stride);

static bool tf_shape_ok(int Cin, int Cout, int k, int stride, int pad) {
  return Cin % 8 == 0 && Cout % 8 == 0 && (k == 1 || k == 3) && (stride == 1 || stride == 2) && pad == k / 2;
}

int tf_conv_forward(const float* x, const float* w, const float* bias, int N, int H, int W, int Cin, int Cout, int k, int stride,
                    int pad, float* z, float* ws, size_t ws_bytes, cudaStream_t s, int x_pitch, const float* prepacked) {
  if (x_pitch && (x_pitch < Cin || x_pitch % 4 || ((uintptr_t)x & 15))) { set_error("tf32 conv: input view must be 16-byte aligned with a pitch multiple of 4"); return YB_ERR_SHAPE; }
  if (!tf_shape_ok(Cin, Cout, k, stride, pad)) { set_error("tf32 conv: channels must be multiples of 8, k in {1,3}, stride in {1,2}, pad = k/2"); return YB_ERR_SHAPE; }
  const size_t wn = (size_t)Cout * Cin * k * k;
  if (!prepacked) {  // prepacked: [tap][Cout][Cin], e.g. from tf_pack_all
    if (ws_bytes < wn * 4) { set_error("tf32 conv forward: workspace too small"); return YB_ERR_INVALID_ARG; }
    tf_pack_weights_kernel<<<(unsigned)std::min<size_t>((wn + 255) / 256, 1024), 256, 0, s>>>(w, ws, nullptr, Cout, Cin, k * k);
  }
  TfLaunch L;
  memset(&L, 0, sizeof(L));
  L.in = x; L.N = N; L.Hi = H; L.Wi = W; L.Kc = Cin; L.in_pitch = x_pitch;
  L.wpk = prepacked ? prepacked : ws; L.Nc = Cout; L.taps_total = k * k;
  L.Ho = (H + 2 * pad - k) / stride + 1; L.Wo = (W + 2 * pad - k) / stride + 1;
  L.out = z; L.o_pix = Cout; L.o_row = (long long)L.Wo * Cout; L.o_img = (long long)L.Ho * L.o_row; L.o_off = 0;
  L.bias = bias;
  L.in_stride = stride;
  L.ntaps = k * k;
  for (int t = 0; t < k * k; t++) { L.dh[t] = t / k - pad; L.dw[t] = t % k - pad; L.slab[t] = t; }
  L.flat = (k == 1 && stride == 1);
  return tf_conv_launch(L, s);
}

int tf_conv_backward_data(const float* dz, const float* w, int N, int H, int W, int Cin, int Cout, int k, int stride, int pad,
                          float* dx, float* ws, size_t ws_bytes, cudaStream_t s, const float* prepacked) {
  if (!tf_shape_ok(Cin, Cout, k, stride, pad) || (stride == 2 && ((H | W) & 1))) {
    set_error("tf32 dgrad: channels must be multiples of 8, k in {1,3}, stride in {1,2} (even size for stride 2), pad = k/2");
    return YB_ERR_SHAPE;
  }
  const size_t wn = (size_t)Cout * Cin * k * k;
  if (!prepacked) {  // prepacked: [tap][Cin][Cout]
    if (ws_bytes < wn * 4) { set_error("tf32 dgrad: workspace too small"); return YB_ERR_INVALID_ARG; }
    tf_pack_weights_kernel<<<(unsigned)std::min<size_t>((wn + 255) / 256, 1024), 256, 0, s>>>(w, nullptr, ws, Cout, Cin, k * k);
  }
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  TfLaunch L;
  memset(&L, 0, sizeof(L));
  L.in = dz; L.N = N; L.Hi = Ho; L.Wi = Wo; L.Kc = Cout;
  L.wpk = prepacked ? prepacked : ws; L.Nc = Cin; L.taps_total = k * k;
  L.out = dx; L.bias = nullptr;
  L.in_stride = 1;
  if (stride == 1) {
    L.Ho = H; L.Wo = W;
    L.o_pix = Cin; L.o_row = (long long)W * Cin; L.o_img = (long long)H * L.o_row; L.o_off = 0;
    L.ntaps = k * k;
    for (int t = 0; t < k * k; t++) { L.dh[t] = pad - t / k; L.dw[t] = pad - t % k; L.slab[t] = t; }
    L.flat = (k == 1);
    return tf_conv_launch(L, s);
  }
  // stride 2: dx(2a + py, 2b + px) = sum over taps with (py + pad - kh), (px + pad - kw) even of dz(a + (py+pad-kh)/2, b + ...)
  // k = 1 (pad 0): only parity (0,0) receives gradient; the other positions are zero.
  if (k == 1) YB_CUDA_CHECK(cudaMemsetAsync(dx, 0, (size_t)N * H * W * Cin * sizeof(float), s));
  for (int py = 0; py < 2; py++)
    for (int px = 0; px < 2; px++) {
      L.ntaps = 0;
      for (int kh = 0; kh < k; kh++)
        for (int kw = 0; kw < k; kw++) {
          const int eh = py + pad - kh, ew = px + pad - kw;
          if ((eh & 1) || (ew & 1)) continue;
          // arithmetic shift: eh in {-1..2} is even here, so eh / 2 is exact for 0 and 2; eh = -2 cannot occur (kh <= 2, pad = 1)
          L.dh[L.ntaps] = eh / 2; L.dw[L.ntaps] = ew / 2; L.slab[L.ntaps] = kh * k + kw;
          L.ntaps++;
        }
      if (L.ntaps == 0) continue;  // k = 1, odd parity: stays zero
      L.Ho = H / 2; L.Wo = W / 2;
      L.o_pix = 2LL * Cin; L.o_row = 2LL * W * Cin; L.o_img = (long long)H * W * Cin;
      L.o_off = ((long long)py * W + px) * Cin;
      L.flat = false;
      const int rc = tf_conv_launch(L, s);
      if (rc) return rc;
    }
  return 0;
}

// ------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------
constexpr int WG_PW = 8, WG_PH = 8;          // pixel tile of the dz grid (64 pixels = 8 MMAs of K = 8)
constexpr int WG_BLK = WG_PW * WG_PH * 128;  // bytes of one 32-channel block of a pixel tile (64 rows x 128 B)
constexpr int WG_A_STAGES = 2;

struct WgArgs {
  CUtensorMap tmDz, tmX;
  float* part;      // [split][co_pad][tap][ci_pad]
  int Cout, Cin, taps, ksz, stride, pad;
  int nb;           // 32-channel input blocks per CTA (N = nb * 32 columns per tap)
  int co_blocks;    // 32-channel output blocks loaded per CTA (<= 4)
  int co_tiles, ci_tiles, splits;
  int tpc, tap_groups;  // taps per CTA (3 = one kh row of a 3x3, 1 for 1x1) and groups of them
  int imgs, tiles_w, tiles_h, pix_tiles;
  int b_stages;
  int merge_kw;     // halo mode with one 32-channel input block: the three kw taps of a row as ONE N = 96 MMA
  int halo;         // 3x3 stride 1: ONE PH x (PW+2) input tile per pixel tile serves the three taps of a kh row (row-shifted windows)
  uint32_t xblk;    // bytes reserved per 32-channel block of an x stage (1 KiB aligned)
  int co_pad, ci_pad;
  uint32_t tmem_cols;
};

__global__ void __launch_bounds__(TF_THREADS, 1) tf_wgrad_kernel(const __grid_constant__ WgArgs a) {
  extern __shared__ __align__(1024) uint8_t wg_smem[];
  __shared__ __align__(8) uint64_t bars[2 * WG_A_STAGES + 2 * 16 + 1];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem0 = (smem_u32(wg_smem) + 1023u) & ~1023u;
  const uint32_t a_stride = 4 * WG_BLK, b_stride = (uint32_t)a.nb * a.xblk;
  const uint32_t smemA = smem0, smemB = smem0 + WG_A_STAGES * a_stride;
  const uint32_t afull = smem_u32(&bars[0]), aempty = smem_u32(&bars[WG_A_STAGES]);
  const uint32_t bfull = smem_u32(&bars[2 * WG_A_STAGES]), bempty = smem_u32(&bars[2 * WG_A_STAGES + 16]);
  const uint32_t done = smem_u32(&bars[2 * WG_A_STAGES + 32]);
  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_A_STAGES; s++) { mbar_init(afull + 8 * s, 1); mbar_init(aempty + 8 * s, 1); }
    for (int s = 0; s < a.b_stages; s++) { mbar_init(bfull + 8 * s, 1); mbar_init(bempty + 8 * s, 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // rows of the dz stages that are never loaded (Cout < 128) must not hold NaN bit patterns: 0 * NaN would poison nothing
  // (every accumulator row depends on its own A row only), but keep the tensor pipe away from denormal/NaN slow paths
  for (uint32_t i = threadIdx.x; i < WG_A_STAGES * a_stride / 16; i += TF_THREADS) st_shared_v4(smemA + i * 16, make_int4(0, 0, 0, 0));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  pdl_trigger();
  const uint32_t tmem_base = tmem_slot;

  // CTA -> (output-channel tile, input-channel tile, pixel split)
  const int split = blockIdx.x % a.splits;
  const int rest = blockIdx.x / a.splits;
  const int tg = rest % a.tap_groups, pair = rest / a.tap_groups;
  const int ci_t = pair % a.ci_tiles, co_t = pair / a.ci_tiles;
  const int tap0 = tg * a.tpc;
  const int tiles_per_img = a.tiles_w * a.tiles_h;
  const int my_tiles = split < a.pix_tiles ? (a.pix_tiles - split + a.splits - 1) / a.splits : 0;
  const int ncols = a.nb * 32;  // accumulator columns per tap

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&a.tmDz) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&a.tmX) : "memory");
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int i = 0; i < my_tiles; i++) {
        const int pt = split + i * a.splits;
        const int img = pt / tiles_per_img, r = pt - img * tiles_per_img;
        const int th = r / a.tiles_w, tw = r - th * a.tiles_w;
        const int w0 = tw * WG_PW, h0 = th * WG_PH;
        mbar_wait(aempty + 8 * sa, pa ^ 1);
        mbar_arrive_expect_tx(afull + 8 * sa, (uint32_t)a.co_blocks * WG_BLK);
        for (int b = 0; b < a.co_blocks; b++)
          tma_load_4d(smemA + sa * a_stride + b * WG_BLK, &a.tmDz, afull + 8 * sa, co_t * 128 + b * 32, w0, h0, img);
        if (++sa == WG_A_STAGES) { sa = 0; pa ^= 1; }
        if (a.halo) {
          // one box of PH x (PW + 2) input pixels per 32-channel block for the kh row of this CTA: tap (kh, kw) of output
          // row h reads its 8 pixels from rows h * (PW + 2) + kw .. + 7 of it
          const int kh = tap0 / 3;
          mbar_wait(bempty + 8 * sb, pb ^ 1);
          mbar_arrive_expect_tx(bfull + 8 * sb, (uint32_t)a.nb * (WG_PW + 2) * WG_PH * 128);
          for (int b = 0; b < a.nb; b++)
            tma_load_4d(smemB + sb * b_stride + b * a.xblk, &a.tmX, bfull + 8 * sb, (ci_t * a.nb + b) * 32, w0 - a.pad, h0 + kh - a.pad, img);
          if (++sb == a.b_stages) { sb = 0; pb ^= 1; }
        } else {
          for (int tt = 0; tt < a.tpc; tt++) {
            const int t = tap0 + tt;
            const int kh = t / a.ksz, kw = t - kh * a.ksz;
            mbar_wait(bempty + 8 * sb, pb ^ 1);
            mbar_arrive_expect_tx(bfull + 8 * sb, (uint32_t)a.nb * WG_BLK);
            for (int b = 0; b < a.nb; b++)
              tma_load_4d(smemB + sb * b_stride + b * a.xblk, &a.tmX, bfull + 8 * sb, (ci_t * a.nb + b) * 32,
                          w0 * a.stride + kw - a.pad, h0 * a.stride + kh - a.pad, img);
            if (++sb == a.b_stages) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // D fp32, A / B TF32, BOTH MN-major (bits 15, 16): operands are [pixel][channel] tiles, K runs over pixels
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(ncols >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      // MN-major 32-bit operands have ONE legal shared-memory layout: 128-byte rows (32 channels of one pixel) swizzled
      // in 32-byte units over groups of 4 rows (UMMA layout type 1 = SWIZZLE_128B_BASE32B, written by TMA's
      // SWIZZLE_128B_ATOM_32B).  LBO = next 32-channel block, SBO = next group of 4 pixels (K atoms of 4).
      const uint32_t lbo16 = WG_BLK >> 4, sbo16 = 512 >> 4;
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int i = 0; i < my_tiles; i++) {
        mbar_wait(afull + 8 * sa, pa);
        tc_fence_after();
        const uint32_t A0 = smemA + sa * a_stride;
        const uint32_t xlbo16 = a.xblk >> 4;
        if (a.halo) {
          mbar_wait(bfull + 8 * sb, pb);
          tc_fence_after();
          const uint32_t B0 = smemB + sb * b_stride;
          if (a.merge_kw) {
            // one 32-channel input block: the three kw windows are the SAME box shifted by one pixel row (128 bytes), so they
            // are three N blocks of one MMA with a block stride (LBO) of 128 bytes - N = 96 instead of three N = 32 MMAs
            // (a TF32 M = 128 MMA costs ~100 cycles whatever its N: the Cin <= 32 layers at 160 x 160 were MMA-issue bound)
            const uint32_t idesc3 = (idesc & ~(0x3fu << 17)) | ((uint32_t)(96 >> 3) << 17);
#pragma unroll
            for (int ks = 0; ks < WG_PH; ks++)
              umma_tf32(tmem_base, tf_desc(A0 + ks * 1024, lbo16, sbo16, 1),
                        tf_desc(B0 + (uint32_t)(ks * (WG_PW + 2)) * 128, 128 >> 4, sbo16, 1), idesc3, (i > 0 || ks > 0) ? 1u : 0u);
          } else {
            for (int kw = 0; kw < 3; kw++) {
              const uint32_t d_tmem = tmem_base + kw * ncols;
#pragma unroll
              for (int ks = 0; ks < WG_PH; ks++)  // output row ks: its 8 input pixels start at row ks * (PW + 2) + kw of the box
                umma_tf32(d_tmem, tf_desc(A0 + ks * 1024, lbo16, sbo16, 1),
                          tf_desc(B0 + (uint32_t)(ks * (WG_PW + 2) + kw) * 128, xlbo16, sbo16, 1), idesc, (i > 0 || ks > 0) ? 1u : 0u);
            }
          }
          umma_commit(bempty + 8 * sb);
          if (++sb == a.b_stages) { sb = 0; pb ^= 1; }
        } else {
          for (int tt = 0; tt < a.tpc; tt++) {
            mbar_wait(bfull + 8 * sb, pb);
            tc_fence_after();
            const uint32_t B0 = smemB + sb * b_stride;
            const uint32_t d_tmem = tmem_base + tt * ncols;
#pragma unroll
            for (int ks = 0; ks < WG_PW * WG_PH / 8; ks++)  // 8 pixels (K = 8) = 1 KiB = two 4-row swizzle atoms per block
              umma_tf32(d_tmem, tf_desc(A0 + ks * 1024, lbo16, sbo16, 1), tf_desc(B0 + ks * 1024, xlbo16, sbo16, 1), idesc,
                        (i > 0 || ks > 0) ? 1u : 0u);
            umma_commit(bempty + 8 * sb);
            if (++sb == a.b_stages) { sb = 0; pb ^= 1; }
          }
        }
        umma_commit(aempty + 8 * sa);
        if (++sa == WG_A_STAGES) { sa = 0; pa ^= 1; }
      }
      umma_commit(done);
    }
  } else {
    const int q = warp & 3;
    const int co = co_t * 128 + q * 32 + lane;
    if (my_tiles > 0) {
      mbar_wait(done, 0);
      tc_fence_after();
    }
    float* prow = a.part + (((size_t)split * a.co_pad + co) * a.taps) * a.ci_pad + (size_t)ci_t * ncols;
    for (int tt = 0; tt < a.tpc; tt++)
      for (int c0 = 0; c0 < ncols; c0 += 16) {
        const int t = tap0 + tt;
        uint32_t v[16];
        if (my_tiles > 0) {
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + tt * ncols + c0, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; j++) v[j] = 0u;
        }
        if (co < a.Cout) {
#pragma unroll
          for (int g = 0; g < 4; g++)
            *reinterpret_cast<float4*>(prow + (size_t)t * a.ci_pad + c0 + g * 4) =
                make_float4(__uint_as_float(v[g * 4]), __uint_as_float(v[g * 4 + 1]), __uint_as_float(v[g * 4 + 2]),
                            __uint_as_float(v[g * 4 + 3]));
        }
      }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
}

// dw[co][ci][tap] = sum over splits (fixed order) of part[split][co][tap][ci]
// One block per output channel: the partials are read along ci (coalesced), summed over the splits in split order, transposed
// through shared memory ([ci][tap], stride `taps` is odd or 1: no bank conflicts) and written as the contiguous run
// dw[co][:][:].  (Threads over the checkpoint layout read 4-byte words ci_pad apart: 0.86 ms over the 79 folds of a YOLOv11s
// step; reads coalesced over ci with scattered writes: 1.13 ms.)
__global__ void __launch_bounds__(256) tf_wgrad_fold_kernel(const float* __restrict__ part, float* __restrict__ dw, int Cout, int Cin,
                                                            int taps, int splits, int co_pad, int ci_pad) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float fold_sm[];  // [Cin][taps]
  const int co = blockIdx.x;
  const size_t ss = (size_t)co_pad * taps * ci_pad;
  const float* base = part + (size_t)co * taps * ci_pad;
  const int n = Cin * taps;
  for (int e = threadIdx.x; e < n; e += 256) {
    const int t = e / Cin, ci = e - t * Cin;
    const float* p = base + (size_t)t * ci_pad + ci;
    float acc = 0.f;
    int sp = 0;
    for (; sp + 4 <= splits; sp += 4) {
      const float a0 = p[(size_t)sp * ss], a1 = p[(size_t)(sp + 1) * ss], a2 = p[(size_t)(sp + 2) * ss], a3 = p[(size_t)(sp + 3) * ss];
      acc += a0; acc += a1; acc += a2; acc += a3;
    }
    for (; sp < splits; sp++) acc += p[(size_t)sp * ss];
    fold_sm[ci * taps + t] = acc;
  }
  __syncthreads();
  float* out = dw + (size_t)co * n;
  for (int e = threadIdx.x; e < n; e += 256) out[e] = fold_sm[e];
}

struct WgPlan { int nb, co_tiles, ci_tiles, splits, co_pad, ci_pad, pix_tiles, tiles_w, tiles_h, tpc, tap_groups; };
static WgPlan wg_plan(int N, int H, int W, int Cin, int Cout, int k, int stride, int pad) {
  WgPlan p;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const int taps = k * k;
  const int ci_blocks = (Cin + 31) / 32;
  // a CTA owns 128 output channels x (tpc taps x nb * 32 input channels) in TMEM: tpc = 3 (one kh row of a 3x3) x up to
  // 128 channels = 384 columns.  (The first version kept all nine taps of 32 input channels: N = 32 per MMA, and a
  // TF32 M=128 MMA costs ~110 cycles whatever its N - 8.7 % tensor activity, profiles/r2_ncu_tf_wgrad.txt.)
  p.nb = std::min(ci_blocks, 4);
  p.ci_tiles = (ci_blocks + p.nb - 1) / p.nb;
  p.co_tiles = (Cout + 127) / 128;
  p.co_pad = p.co_tiles * 128;
  p.ci_pad = p.ci_tiles * p.nb * 32;
  p.tiles_w = (Wo + WG_PW - 1) / WG_PW;
  p.tiles_h = (Ho + WG_PH - 1) / WG_PH;
  p.pix_tiles = N * p.tiles_w * p.tiles_h;
  p.tpc = taps == 9 ? 3 : 1;
  p.tap_groups = taps / p.tpc;
  const int pairs = p.co_tiles * p.ci_tiles * p.tap_groups;
  // pixel splits: one CTA per SM (189 KiB of shared memory), so the grid should fill ONE wave of SMs, or two when that fills
  // them noticeably better - never a few CTAs over (ceil(2 * 148 / pairs) capped at 64 gave grids of 300 - 320 = a third
  // wave of 4 - 24 CTAs on a third of the layers, and 64-CTA grids on the 1x1 layers with <= 128 channels)
  const int sms = tf_num_sms();
  const int s1 = std::max(1, sms / pairs), s2 = std::max(1, 2 * sms / pairs);
  const double u1 = (double)std::min(pairs * s1, sms) / sms, u2 = (double)std::min(pairs * s2, 2 * sms) / (2.0 * sms);
  p.splits = (pairs <= sms && u2 > u1 + 0.08) ? s2 : s1;
  p.splits = std::max(1, std::min(p.splits, p.pix_tiles));
  return p;
}

size_t tf_conv_workspace_bytes(int N, int H, int W, int Cin, int Cout, int k, int stride) {
  const WgPlan p = wg_plan(N, H, W, Cin, Cout, k, stride, k / 2);
  const size_t part = (size_t)p.splits * p.co_pad * k * k * p.ci_pad * 4;
  const size_t wpk = (size_t)Cout * Cin * k * k * 4;
  return std::max(std::max(part, wpk), (size_t)148 * 4 * 27 * 128 * sizeof(float)) + 256;
}

int tf_conv_backward_weight(const float* x, const float* dz, int N, int H, int W, int Cin, int Cout, int k, int stride, int pad,
                            float* dw, float* ws, size_t ws_bytes, cudaStream_t s, int x_pitch) {
  if (x_pitch && (x_pitch < Cin || x_pitch % 4 || ((uintptr_t)x & 15))) { set_error("tf32 wgrad: input view must be 16-byte aligned with a pitch multiple of 4"); return YB_ERR_SHAPE; }
  const int xp = x_pitch ? x_pitch : Cin;
  if (!tf_shape_ok(Cin, Cout, k, stride, pad)) { set_error("tf32 wgrad: channels must be multiples of 8, k in {1,3}, stride in {1,2}, pad = k/2"); return YB_ERR_SHAPE; }
  TfEncodeFn encode = tf_encode_fn();
  if (!encode) { set_error("cuTensorMapEncodeTiled entry point not found"); return YB_ERR_CUDA; }
  const WgPlan p = wg_plan(N, H, W, Cin, Cout, k, stride, pad);
  const size_t part_bytes = (size_t)p.splits * p.co_pad * k * k * p.ci_pad * 4;
  if (ws_bytes < part_bytes) { set_error("tf32 wgrad: workspace too small"); return YB_ERR_INVALID_ARG; }
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  WgArgs a;
  memset(&a, 0, sizeof(a));
  a.part = ws;
  a.Cout = Cout; a.Cin = Cin; a.taps = k * k; a.ksz = k; a.stride = stride; a.pad = pad;
  a.nb = p.nb; a.co_tiles = p.co_tiles; a.ci_tiles = p.ci_tiles; a.splits = p.splits;
  a.tpc = p.tpc; a.tap_groups = p.tap_groups;
  a.co_blocks = std::min(4, (Cout + 31) / 32);
  a.imgs = N; a.tiles_w = p.tiles_w; a.tiles_h = p.tiles_h; a.pix_tiles = p.pix_tiles;
  a.co_pad = p.co_pad; a.ci_pad = p.ci_pad;
  // 3x3 stride 1: the input tile with its halo is loaded once per pixel tile and the nine taps are row-shifted MMA windows
  // into it (the per-tap form moved 9 x 8 KiB of x per 64 pixels and was bound by L2 -> SM delivery).  The shifted start
  // relies on tcgen05 applying the 32-byte-atom swizzle on absolute shared-memory address bits, as the K-major halo tile
  // of conv_tc.cu does for the 16-byte-atom modes; tests/test_gpu_conv_tc.py's bit-exact case covers it.
  static const bool no_halo = getenv("YB_WGRAD_NO_HALO") != nullptr;
  a.halo = (k == 3 && stride == 1 && !no_halo) ? 1 : 0;
  static const bool no_merge = getenv("YB_WGRAD_NO_MERGE") != nullptr;
  a.merge_kw = (a.halo && p.nb == 1 && !no_merge) ? 1 : 0;
  a.xblk = a.halo ? (uint32_t)(((WG_PW + 2) * WG_PH * 128 + 1023) / 1024 * 1024) : (uint32_t)WG_BLK;
  const size_t b_stride = (size_t)p.nb * a.xblk;
  a.b_stages = (int)std::min<size_t>(16, ((size_t)190 * 1024 - (size_t)WG_A_STAGES * 4 * WG_BLK) / b_stride);
  if (a.b_stages < 2) { set_error("tf32 wgrad: tile does not fit in shared memory"); return YB_ERR_SHAPE; }
  uint32_t cols = 32;
  while (cols < (uint32_t)(a.tpc * p.nb * 32)) cols <<= 1;
  a.tmem_cols = cols;
  {
    cuuint64_t gd[4] = {(cuuint64_t)Cout, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)N};
    cuuint64_t gs[3] = {(cuuint64_t)Cout * 4, (cuuint64_t)Cout * 4 * Wo, (cuuint64_t)Cout * 4 * Wo * Ho};
    cuuint32_t bx[4] = {32, WG_PW, WG_PH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult cr = encode(&a.tmDz, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(dz), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("tf32 wgrad: cuTensorMapEncodeTiled(dz) failed with code " + std::to_string((int)cr)); return YB_ERR_CUDA; }
  }
  {
    cuuint64_t gd[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gs[3] = {(cuuint64_t)xp * 4, (cuuint64_t)xp * 4 * W, (cuuint64_t)xp * 4 * W * H};
    cuuint32_t bx[4] = {32, (cuuint32_t)(a.halo ? WG_PW + 2 : WG_PW * stride), (cuuint32_t)(a.halo ? WG_PH : WG_PH * stride), 1};
    cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult cr = encode(&a.tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("tf32 wgrad: cuTensorMapEncodeTiled(x) failed with code " + std::to_string((int)cr)); return YB_ERR_CUDA; }
  }
  static bool attr_set = false;
  if (!attr_set) {
    YB_CUDA_CHECK(cudaFuncSetAttribute(tf_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const size_t smem = (size_t)WG_A_STAGES * 4 * WG_BLK + (size_t)a.b_stages * b_stride + 1024;
  const int grid = p.co_tiles * p.ci_tiles * p.tap_groups * p.splits;
  YB_CUDA_CHECK(launch_pdl(tf_wgrad_kernel, dim3(grid), dim3(TF_THREADS), smem, s, a));
  const size_t n = (size_t)Cout * Cin * k * k;
  static bool fold_attr = false;
  if (!fold_attr) {
    YB_CUDA_CHECK(cudaFuncSetAttribute(tf_wgrad_fold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    fold_attr = true;
  }
  if ((size_t)Cin * k * k * sizeof(float) > (size_t)160 * 1024) { set_error("tf32 wgrad: Cin * k * k too large for the fold"); return YB_ERR_SHAPE; }
  YB_CUDA_CHECK(launch_pdl(tf_wgrad_fold_kernel, dim3(Cout), dim3(256), (size_t)Cin * k * k * sizeof(float), s, (const float*)ws, dw, Cout, Cin, k * k,
                           p.splits, p.co_pad, p.ci_pad));
  (void)n;
  YB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// The 3-channel stem (model.0: Conv(3, C, k = 3, s = 2), Yolo.cs:53) on CUDA cores, fp32.
// On the tensor-core path the stem ran with its input zero-padded to 8 channels: K = 8 per tap means one tiny MMA per TMA
// box of 128 strided 32-byte rows, and both passes were bound by the TMA row rate - forward 239 us and wgrad 1 288 us of a
// 24.6 ms step (profiles/r2_ncu_tf_conv_big.txt, r2_ncu_tf_wgrad_big.txt), for 0.3 % of the step's FLOPs.  Here:
//   forward   one thread per output pixel and 32-channel group: its 27 inputs in registers, weights [27][C] in shared memory
//   wgrad     dW[co][ci][kh][kw] = sum over pixels of dz[p][co] * x[window(p)][ci][kh][kw]: a block stages 128 pixels (their
//             27-value windows and dz rows) in shared memory, thread (co lane, tap group) accumulates its share of the
//             27 x C products; per-block partials are folded in block order (deterministic)
// x is NHWC with `xc` channels per pixel of which the first 3 are used (the native step keeps an 8-channel input).
// ------------------------------------------------------------------------------------------
constexpr int ST_PIX = 128;

__global__ void __launch_bounds__(256) stem3_forward_kernel(const float* __restrict__ x, int xc, const float* __restrict__ w,
                                                           float* __restrict__ z, int N, int H, int W, int Ho, int Wo, int C) {
  extern __shared__ float sw[];  // [27][C]: k = (kh * 3 + kw) * 3 + ci
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) {
    const int co = i % C, k = i / C;
    const int ci = k % 3, t = k / 3;
    sw[i] = w[(co * 3 + ci) * 9 + t];
  }
  __syncthreads();
  const int groups = (C + 31) / 32;
  const long long total = (long long)N * Ho * Wo * groups;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % groups);
    long long p = idx / groups;
    const int wo = (int)(p % Wo);
    p /= Wo;
    const int ho = (int)(p % Ho);
    const int n = (int)(p / Ho);
    float in[27];
#pragma unroll
    for (int kh = 0; kh < 3; kh++) {
      const int hi = 2 * ho + kh - 1;
#pragma unroll
      for (int kw = 0; kw < 3; kw++) {
        const int wi = 2 * wo + kw - 1;
        const bool ok = hi >= 0 && hi < H && wi >= 0 && wi < W;
        const float* px = x + (((size_t)n * H + (ok ? hi : 0)) * W + (ok ? wi : 0)) * xc;
#pragma unroll
        for (int ci = 0; ci < 3; ci++) in[(kh * 3 + kw) * 3 + ci] = ok ? px[ci] : 0.f;
      }
    }
    const int c0 = g * 32, cn = min(32, C - c0);
    float* out = z + ((size_t)(n * Ho + ho) * Wo + wo) * C + c0;
    for (int c = 0; c < cn; c += 4) {  // C is a multiple of 8
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int k = 0; k < 27; k++) {
        const float4 wv = *reinterpret_cast<const float4*>(sw + k * C + c0 + c);
        a0 = fmaf(in[k], wv.x, a0); a1 = fmaf(in[k], wv.y, a1); a2 = fmaf(in[k], wv.z, a2); a3 = fmaf(in[k], wv.w, a3);
      }
      *reinterpret_cast<float4*>(out + c) = make_float4(a0, a1, a2, a3);
    }
  }
}

// Row-segment version: a tile is up to 128 adjacent output pixels of one output row.  The three input rows it needs are staged
// as [kh][column][4] (one 16-byte load per input pixel, coalesced), so the 3 x 3 window of output pixel p is 3 x three
// adjacent float4 of shared memory; dz of the tile is staged as [pixel][C].  Thread = (channel lane, warp g): warp g takes
// pixels g, g + 8, ... and keeps all 27 taps of its channel(s) in registers: per pixel 9 broadcast LDS.128 + 1 LDS feed
// 27 FMAs (the first version staged im2col windows element by element - ~40 integer instructions per staged value - and
// spent 5 LDS per 4 FMAs: 690 us for 1.4 GFMA).  The 8 warps are folded in warp order through shared memory, the blocks'
// partials by stem3_wgrad_fold_kernel in block order: deterministic.
// grid.x blocks walk the tiles with stride gridDim.x; partial[block][27][C], k = (kh * 3 + kw) * 3 + ci
template <int CG>  // 32-channel groups per thread: ceil(C / 32)
__global__ void __launch_bounds__(256) stem3_wgrad_partial_kernel(const float* __restrict__ x, int xc, const float* __restrict__ dz,
                                                                 float* __restrict__ partial, int N, int H, int W, int Ho, int Wo, int C) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                              // [3][2 * ST_PIX + 1][4]
  float* ds = sm + 3 * (2 * ST_PIX + 1) * 4;   // [ST_PIX][C]; later the fold buffer [8][27][32]
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int segs = (Wo + ST_PIX - 1) / ST_PIX;
  const long long tiles = (long long)N * Ho * segs;
  float acc[CG][27];
#pragma unroll
  for (int b2 = 0; b2 < CG; b2++)
#pragma unroll
    for (int k = 0; k < 27; k++) acc[b2][k] = 0.f;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int seg = (int)(tile % segs);
    const long long q = tile / segs;
    const int ho = (int)(q % Ho), n = (int)(q / Ho);
    const int wo0 = seg * ST_PIX, npx = min(ST_PIX, Wo - wo0);
    const int ncol = 2 * npx + 1, wi0 = 2 * wo0 - 1;
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * ncol; i += 256) {
      const int kh = i / ncol, j = i - kh * ncol;
      const int hi = 2 * ho + kh - 1, wi = wi0 + j;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (hi >= 0 && hi < H && wi >= 0 && wi < W) {
        const float* px = x + (((size_t)n * H + hi) * W + wi) * xc;
        if ((xc & 3) == 0) v = __ldg(reinterpret_cast<const float4*>(px));
        else { v.x = px[0]; v.y = px[1]; v.z = px[2]; }
      }
      *reinterpret_cast<float4*>(xs + ((size_t)kh * (2 * ST_PIX + 1) + j) * 4) = v;
    }
    const float* dzt = dz + (((size_t)n * Ho + ho) * Wo + wo0) * C;  // npx * C contiguous floats (C % 8 == 0: 16-byte aligned)
    for (int i = threadIdx.x; i < npx * C / 4; i += 256)
      *reinterpret_cast<float4*>(ds + (size_t)i * 4) = __ldg(reinterpret_cast<const float4*>(dzt) + i);
    __syncthreads();
    for (int pl = g; pl < npx; pl += 8) {
      float4 w9[9];  // [kh][kw] -> (ci 0..2, pad)
#pragma unroll
      for (int kh = 0; kh < 3; kh++)
#pragma unroll
        for (int kw = 0; kw < 3; kw++)
          w9[kh * 3 + kw] = *reinterpret_cast<const float4*>(xs + ((size_t)kh * (2 * ST_PIX + 1) + 2 * pl + kw) * 4);
#pragma unroll
      for (int b2 = 0; b2 < CG; b2++) {
        {
          const int c = b2 * 32 + lane;
          const float gv = c < C ? ds[pl * C + c] : 0.f;
#pragma unroll
          for (int t = 0; t < 9; t++) {
            acc[b2][t * 3 + 0] = fmaf(gv, w9[t].x, acc[b2][t * 3 + 0]);
            acc[b2][t * 3 + 1] = fmaf(gv, w9[t].y, acc[b2][t * 3 + 1]);
            acc[b2][t * 3 + 2] = fmaf(gv, w9[t].z, acc[b2][t * 3 + 2]);
          }
        }
      }
    }
  }
  // fold the 8 warps in warp order, one 32-channel group at a time
  float* red = ds;  // [8][27][32]
#pragma unroll
  for (int b2 = 0; b2 < CG; b2++) {
    {
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 27; k++) red[(g * 27 + k) * 32 + lane] = acc[b2][k];
      __syncthreads();
      for (int i = threadIdx.x; i < 27 * 32; i += 256) {
        const int k = i >> 5, cl = i & 31;
        float sum = 0.f;
        for (int w8 = 0; w8 < 8; w8++) sum += red[(w8 * 27 + k) * 32 + cl];
        const int c = b2 * 32 + cl;
        if (c < C) partial[((size_t)blockIdx.x * 27 + k) * C + c] = sum;
      }
    }
  }
}
__global__ void stem3_wgrad_fold_kernel(const float* __restrict__ partial, int blocks, int C, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (k, c)
  if (i >= 27 * C) return;
  const int k = i / C, c = i - k * C;
  float s = 0.f;
  for (int b = 0; b < blocks; b++) s += partial[((size_t)b * 27 + k) * C + c];
  const int ci = k % 3, t = k / 3;
  dw[(c * 3 + ci) * 9 + t] = s;
}

int stem3_forward(const float* x, int xc, const float* w, int N, int H, int W, int C, float* z, cudaStream_t s) {
  if (C % 8 || C > 128 || (H & 1) || (W & 1) || xc < 3) { set_error("stem conv: C % 8 == 0, C <= 128, even input size"); return YB_ERR_SHAPE; }
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * Ho * Wo * ((C + 31) / 32);
  const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  stem3_forward_kernel<<<grid, 256, (size_t)27 * C * sizeof(float), s>>>(x, xc, w, z, N, H, W, Ho, Wo, C);
  YB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
size_t stem3_wgrad_workspace_bytes(int C) { return (size_t)148 * 4 * 27 * C * sizeof(float); }
int stem3_backward_weight(const float* x, int xc, const float* dz, int N, int H, int W, int C, float* dw, float* ws, size_t ws_bytes,
                          cudaStream_t s) {
  if (C % 8 || C > 128 || (H & 1) || (W & 1) || xc < 3) { set_error("stem wgrad: C % 8 == 0, C <= 128, even input size"); return YB_ERR_SHAPE; }
  const int blocks = 148 * 4;
  if (ws_bytes < stem3_wgrad_workspace_bytes(C)) { set_error("stem wgrad: workspace too small"); return YB_ERR_INVALID_ARG; }
  const int Ho = H / 2, Wo = W / 2;
  const size_t smem = (size_t)(3 * (2 * ST_PIX + 1) * 4 + std::max(ST_PIX * C, 8 * 27 * 32)) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    YB_CUDA_CHECK(cudaFuncSetAttribute(stem3_wgrad_partial_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    YB_CUDA_CHECK(cudaFuncSetAttribute(stem3_wgrad_partial_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    YB_CUDA_CHECK(cudaFuncSetAttribute(stem3_wgrad_partial_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    YB_CUDA_CHECK(cudaFuncSetAttribute(stem3_wgrad_partial_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr = true;
  }
  switch ((C + 31) / 32) {
    case 1: stem3_wgrad_partial_kernel<1><<<blocks, 256, smem, s>>>(x, xc, dz, ws, N, H, W, Ho, Wo, C); break;
    case 2: stem3_wgrad_partial_kernel<2><<<blocks, 256, smem, s>>>(x, xc, dz, ws, N, H, W, Ho, Wo, C); break;
    case 3: stem3_wgrad_partial_kernel<3><<<blocks, 256, smem, s>>>(x, xc, dz, ws, N, H, W, Ho, Wo, C); break;
    default: stem3_wgrad_partial_kernel<4><<<blocks, 256, smem, s>>>(x, xc, dz, ws, N, H, W, Ho, Wo, C); break;
  }
  stem3_wgrad_fold_kernel<<<(27 * C + 127) / 128, 128, 0, s>>>(ws, blocks, C, dw);
  YB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

static bool tf_have_dev(const char* who) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_error(std::string(who) + ": no CUDA device");
    return false;
  }
  return true;
}

}  // namespace yb

using namespace yb;

extern "C" {

int64_t yb_conv_tc_workspace_bytes(int32_t n, int32_t height, int32_t width, int32_t cin, int32_t cout, int32_t k, int32_t stride) {
  if (n <= 0 || height <= 0 || width <= 0 || cin <= 0 || cout <= 0 || k <= 0 || stride <= 0) return 0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return 0; }
  return (int64_t)tf_conv_workspace_bytes(n, height, width, cin, cout, k, stride);
}

int32_t yb_conv_forward_tc(const float* x, const float* w, const float* bias, int32_t n, int32_t height, int32_t width, int32_t cin,
                           int32_t cout, int32_t k, int32_t stride, int32_t pad, float* z, void* workspace, int64_t workspace_bytes,
                           void* stream) {
  if (!x || !w || !z || !workspace) { set_error("yb_conv_forward_tc: null argument"); return YB_ERR_INVALID_ARG; }
  if (n <= 0 || height <= 0 || width <= 0 || cin <= 0 || cout <= 0) { set_error("yb_conv_forward_tc: bad shape"); return YB_ERR_SHAPE; }
  if (!tf_have_dev("yb_conv_forward_tc")) return YB_ERR_NO_DEVICE;
  return tf_conv_forward(x, w, bias, n, height, width, cin, cout, k, stride, pad, z, (float*)workspace, (size_t)workspace_bytes,
                         (cudaStream_t)stream, 0, nullptr);
}

int32_t yb_conv_backward_data_tc(const float* dz, const float* w, int32_t n, int32_t height, int32_t width, int32_t cin, int32_t cout,
                                 int32_t k, int32_t stride, int32_t pad, float* dx, void* workspace, int64_t workspace_bytes,
                                 void* stream) {
  if (!dz || !w || !dx || !workspace) { set_error("yb_conv_backward_data_tc: null argument"); return YB_ERR_INVALID_ARG; }
  if (n <= 0 || height <= 0 || width <= 0 || cin <= 0 || cout <= 0) { set_error("yb_conv_backward_data_tc: bad shape"); return YB_ERR_SHAPE; }
  if (!tf_have_dev("yb_conv_backward_data_tc")) return YB_ERR_NO_DEVICE;
  return tf_conv_backward_data(dz, w, n, height, width, cin, cout, k, stride, pad, dx, (float*)workspace, (size_t)workspace_bytes,
                               (cudaStream_t)stream, nullptr);
}

int32_t yb_conv_backward_weight_tc(const float* x, const float* dz, int32_t n, int32_t height, int32_t width, int32_t cin,
                                   int32_t cout, int32_t k, int32_t stride, int32_t pad, float* dw, void* workspace,
                                   int64_t workspace_bytes, void* stream) {
  if (!x || !dz || !dw || !workspace) { set_error("yb_conv_backward_weight_tc: null argument"); return YB_ERR_INVALID_ARG; }
  if (n <= 0 || height <= 0 || width <= 0 || cin <= 0 || cout <= 0) { set_error("yb_conv_backward_weight_tc: bad shape"); return YB_ERR_SHAPE; }
  if (!tf_have_dev("yb_conv_backward_weight_tc")) return YB_ERR_NO_DEVICE;
  return tf_conv_backward_weight(x, dz, n, height, width, cin, cout, k, stride, pad, dw, (float*)workspace, (size_t)workspace_bytes,
                                 (cudaStream_t)stream, 0);
}

int32_t yb_stem_conv_forward_f32(const float* x, int32_t x_channels, const float* w, int32_t n, int32_t height, int32_t width,
                                 int32_t cout, float* z, void* stream) {
  if (!x || !w || !z || n <= 0 || height <= 0 || width <= 0 || cout <= 0) { set_error("yb_stem_conv_forward_f32: bad argument"); return YB_ERR_INVALID_ARG; }
  if (!tf_have_dev("yb_stem_conv_forward_f32")) return YB_ERR_NO_DEVICE;
  return stem3_forward(x, x_channels, w, n, height, width, cout, z, (cudaStream_t)stream);
}

int32_t yb_stem_conv_backward_weight_f32(const float* x, int32_t x_channels, const float* dz, int32_t n, int32_t height, int32_t width,
                                         int32_t cout, float* dw, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!x || !dz || !dw || !workspace || n <= 0 || height <= 0 || width <= 0 || cout <= 0) { set_error("yb_stem_conv_backward_weight_f32: bad argument"); return YB_ERR_INVALID_ARG; }
  if (!tf_have_dev("yb_stem_conv_backward_weight_f32")) return YB_ERR_NO_DEVICE;
  return stem3_backward_weight(x, x_channels, dz, n, height, width, cout, dw, (float*)workspace, (size_t)workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
