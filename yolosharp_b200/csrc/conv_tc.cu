// tcgen05 implicit-GEMM convolution for sm_100a (fp16 NHWC activations, fp32 accumulate in TMEM).
//
// Computes the reference's Conv block (Modules/Convs.cs:36-56: Conv2d -> BatchNorm2d -> SiLU, BN
// folded into weights/bias at load) and the Bottleneck shortcut (Block.cs:606) as ONE kernel:
//
//   GEMM view   M = output pixels (tile = 128 rows = a BW x BH rectangle of one image, or 128
//               consecutive pixels of the flattened batch for 1x1 convs)
//               N = Cout (tile = n_tile <= 256), K = taps * Cin walked tap-major in slabs of BK
//   A operand   per (tap, channel slab): one 4-D TMA box {BK, BW, BH, 1} of the NHWC input view at
//               (w0*s + kw - pad, h0*s + kh - pad); out-of-bounds rows/cols are zero-filled by TMA
//               (= conv zero padding), stride-2 convs use the tensor map's traversal stride
//   B operand   weights [Cout][tap][Cin] fp16, 2-D TMA box {BK, n_tile}
//   swizzle     BK = 64/32/16 channels -> SWIZZLE_128B/64B/32B rows, identical in the TMA map and
//               the UMMA shared-memory descriptors
//   MMA         tcgen05.mma.cta_group::1.kind::f16, M=128, N=n_tile, K=16 per instruction, issued
//               by one thread; accumulators double-buffered in TMEM (2 x n_tile columns)
//   epilogue    4 warps: tcgen05.ld (32 lanes x 16 columns) -> +bias -> SiLU -> +residual -> fp16 ->
//               16-byte stores into the channel slice of the (concat) output buffer
//   schedule    persistent CTAs (one per SM), warp-specialised: warp0 = TMA producer, warp1 = MMA
//               issuer (+TMEM alloc), warps2-5 = epilogue; smem ring of `stages` slabs.
#include <cuda.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace yb {

struct TcArgs {
  CUtensorMap tmA;
  CUtensorMap tmB;
  __half* out;
  const __half* res;
  const float* bias;
  int out_pitch, out_coff, res_pitch, res_coff;
  int Ho, Wo;            // output extent the tiles cover (flattened for 1x1: Ho = 1, Wo = B*H*W)
  int imgs;              // images the tiles iterate over (1 for flattened 1x1)
  int tiles_w, tiles_h;  // tiles per image
  int BW, BH;
  int n_tile, n_tiles;
  int ksz, stride, pad;
  int Cin, BK, chunks;   // chunks = Cin / BK
  int act;
  int stages;
  uint32_t a_bytes, b_bytes;      // TMA transaction bytes per slab
  uint32_t a_stride, b_stride;    // smem bytes reserved per slab (1 KiB aligned)
  uint32_t sbo;                   // UMMA stride-byte-offset (8 rows), >> 4
  uint32_t layout_type;           // UMMA LayoutType: 2 = SW128, 4 = SW64, 6 = SW32
  uint32_t tmem_cols;
  int total_tiles;
};

struct TcConvPlan {
  TcArgs args;
  ConvParams p;
  bool flat;   // 1x1 stride-1 conv on the flattened pixel dimension
  size_t smem;
  int grid;
};

// ------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must trap (launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (clock64() - t0 > 4000000000ll) __trap();  // ~2 s
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major operand, swizzled rows (cute::UMMA::SmemDescriptor):
//  [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 |
//  [46,48) version=1 | [61,64) layout type
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t sbo, uint32_t layout_type) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)sbo << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)layout_type << 61);
}

__device__ __forceinline__ float silu_fast(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

constexpr int TC_THREADS = 192;
constexpr int TC_MAX_STAGES = 8;

__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  __shared__ __align__(8) uint64_t bars[2 * TC_MAX_STAGES + 4];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // dynamic smem base rounded up to 1 KiB (SWIZZLE_128B atoms need it)
  const uint32_t smem0 = (smem_u32(tc_smem) + 1023u) & ~1023u;
  const uint32_t smemA = smem0;
  const uint32_t smemB = smem0 + a.stages * a.a_stride;
  const uint32_t full0 = smem_u32(&bars[0]);
  const uint32_t empty0 = smem_u32(&bars[TC_MAX_STAGES]);
  const uint32_t tfull0 = smem_u32(&bars[2 * TC_MAX_STAGES]);
  const uint32_t tempty0 = smem_u32(&bars[2 * TC_MAX_STAGES + 2]);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.stages; s++) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; s++) {
      mbar_init(tfull0 + 8 * s, 1);
      mbar_init(tempty0 + 8 * s, 4);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int ksteps = a.ksz * a.ksz * a.chunks;
  const int tiles_per_img = a.tiles_w * a.tiles_h;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&a.tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&a.tmB) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        const int nt = tile % a.n_tiles;
        const int mt = tile / a.n_tiles;
        const int img = mt / tiles_per_img;
        const int r = mt - img * tiles_per_img;
        const int th = r / a.tiles_w, tw = r - th * a.tiles_w;
        const int wbase = tw * a.BW * a.stride - a.pad;
        const int hbase = th * a.BH * a.stride - a.pad;
        for (int t = 0; t < a.ksz * a.ksz; t++) {
          const int kh = t / a.ksz, kw = t - kh * a.ksz;
          for (int ch = 0; ch < a.chunks; ch++) {
            mbar_wait(empty0 + 8 * stage, phase ^ 1);
            mbar_arrive_expect_tx(full0 + 8 * stage, a.a_bytes + a.b_bytes);
            tma_load_4d(smemA + stage * a.a_stride, &a.tmA, full0 + 8 * stage, ch * a.BK, wbase + kw, hbase + kh, img);
            tma_load_2d(smemB + stage * a.b_stride, &a.tmB, full0 + 8 * stage, t * a.Cin + ch * a.BK, nt * a.n_tile);
            if (++stage == a.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (bit 4), a/b F16, K-major both,
      // n_dim = N>>3 at bit 17, m_dim = 128>>4 at bit 24
      const uint32_t idesc = (1u << 4) | ((uint32_t)(a.n_tile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        mbar_wait(tempty0 + 8 * acc, aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * a.n_tile;
        for (int ks = 0; ks < ksteps; ks++) {
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          const uint64_t ad = umma_desc(smemA + stage * a.a_stride, a.sbo, a.layout_type);
          const uint64_t bd = umma_desc(smemB + stage * a.b_stride, a.sbo, a.layout_type);
          const int kk = a.BK >> 4;
          for (int k = 0; k < kk; k++)  // +32 B per K=16 step inside the swizzled row
            umma_f16(d_tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (ks | k) != 0);
          umma_commit(empty0 + 8 * stage);  // slab free once these MMAs retire
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull0 + 8 * acc);  // accumulator complete
        acc ^= 1;
        if (acc == 0) aphase ^= 1;
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const int nt = tile % a.n_tiles;
      const int mt = tile / a.n_tiles;
      const int img = mt / tiles_per_img;
      const int r = mt - img * tiles_per_img;
      const int th = r / a.tiles_w, tw = r - th * a.tiles_w;
      const int hl = row / a.BW, wl = row - hl * a.BW;
      const int ho = th * a.BH + hl, wo = tw * a.BW + wl;
      const bool valid = hl < a.BH && ho < a.Ho && wo < a.Wo;
      const size_t pix = ((size_t)img * a.Ho + ho) * a.Wo + wo;
      const int n0 = nt * a.n_tile;
      __half* orow = a.out + pix * a.out_pitch + a.out_coff + n0;
      const __half* rrow = a.res ? a.res + pix * a.res_pitch + a.res_coff + n0 : nullptr;

      mbar_wait(tfull0 + 8 * acc, aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * a.n_tile;
      for (int c0 = 0; c0 < a.n_tile; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
        if (valid) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; j++) f[j] = __uint_as_float(v[j]) + __ldg(a.bias + n0 + c0 + j);
          if (a.act == ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 16; j++) f[j] = silu_fast(f[j]);
          }
          if (rrow) {
            const int4 r0 = *reinterpret_cast<const int4*>(rrow + c0);
            const int4 r1 = *reinterpret_cast<const int4*>(rrow + c0 + 8);
            const __half2* h0 = reinterpret_cast<const __half2*>(&r0);
            const __half2* h1 = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const float2 x = __half22float2(h0[j]), y = __half22float2(h1[j]);
              f[2 * j] += x.x; f[2 * j + 1] += x.y;
              f[8 + 2 * j] += y.x; f[8 + 2 * j + 1] += y.y;
            }
          }
          int4 o0, o1;
          __half2* p0 = reinterpret_cast<__half2*>(&o0);
          __half2* p1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
          for (int j = 0; j < 4; j++) {
            p0[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
            p1[j] = __floats2half2_rn(f[8 + 2 * j], f[8 + 2 * j + 1]);
          }
          *reinterpret_cast<int4*>(orow + c0) = o0;
          *reinterpret_cast<int4*>(orow + c0 + 8) = o1;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
      acc ^= 1;
      if (acc == 0) aphase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// Host side: tiling choice + tensor maps
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn(std::string* err) {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
    if (err) *err = "cuTensorMapEncodeTiled entry point not found";
    cudaGetLastError();
    return nullptr;
  }
  fn = (EncodeTiledFn)p;
  return fn;
}

bool tc_conv_supported(const ConvParams& p) {
  if (p.Cin % 16 || p.Cout % 16) return false;
  if (!((p.k == 1 && p.stride == 1) || (p.k == 3 && (p.stride == 1 || p.stride == 2)))) return false;
  if (p.in.coff % 8 || p.in.pitch % 8 || p.out.coff % 8 || p.out.pitch % 8) return false;
  if (p.res.base && (p.res.coff % 8 || p.res.pitch % 8)) return false;
  return true;
}

static int pick_n_tile(int cout) {
  if (cout <= 256) return cout;
  int best = 16;
  for (int n = 16; n <= 256; n += 16)
    if (cout % n == 0) best = n;
  return best;
}

TcConvPlan* tc_conv_plan_create(const ConvParams& p, std::string* err) {
  EncodeTiledFn encode = get_encode_fn(err);
  if (!encode) return nullptr;
  TcConvPlan* plan = new TcConvPlan();
  plan->p = p;
  TcArgs& a = plan->args;
  memset(&a, 0, sizeof(a));
  a.out = reinterpret_cast<__half*>(p.out.base);
  a.res = reinterpret_cast<const __half*>(p.res.base);
  a.bias = p.bias;
  a.out_pitch = p.out.pitch; a.out_coff = p.out.coff;
  a.res_pitch = p.res.pitch; a.res_coff = p.res.coff;
  a.ksz = p.k; a.stride = p.stride; a.pad = p.pad;
  a.Cin = p.Cin;
  a.BK = (p.Cin % 64 == 0) ? 64 : (p.Cin % 32 == 0 ? 32 : 16);
  a.chunks = p.Cin / a.BK;
  a.act = p.act;
  a.n_tile = pick_n_tile(p.Cout);
  a.n_tiles = p.Cout / a.n_tile;
  const CUtensorMapSwizzle swz = a.BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                            : (a.BK == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  a.layout_type = a.BK == 64 ? 2 : (a.BK == 32 ? 4 : 6);
  a.sbo = (8 * a.BK * 2) >> 4;

  plan->flat = (p.k == 1 && p.stride == 1);
  const size_t esz = 2;
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t box[4], estr[4];
  char* base = reinterpret_cast<char*>(p.in.base) + (size_t)p.in.coff * esz;
  if (plan->flat) {
    // all pixels of the batch form one dimension: tiles of 128 consecutive pixels
    const cuuint64_t npix = (cuuint64_t)p.B * p.in.H * p.in.W;
    gdim[0] = p.Cin; gdim[1] = npix; gdim[2] = 1; gdim[3] = 1;
    gstr[0] = (cuuint64_t)p.in.pitch * esz; gstr[1] = gstr[0] * npix; gstr[2] = gstr[1];
    a.BW = 128; a.BH = 1;
    box[0] = a.BK; box[1] = 128; box[2] = 1; box[3] = 1;
    estr[0] = estr[1] = estr[2] = estr[3] = 1;
  } else {
    gdim[0] = p.Cin; gdim[1] = p.in.W; gdim[2] = p.in.H; gdim[3] = p.B;
    gstr[0] = (cuuint64_t)p.in.pitch * esz;
    gstr[1] = gstr[0] * p.in.W;
    gstr[2] = gstr[1] * p.in.H;
    // choose the output rectangle BW x BH (<= 128 rows) with the least padding waste
    double best = -1;
    for (int bw = 1; bw <= std::min(p.Wo, 128); bw++) {
      const int bh = std::min(p.Ho, 128 / bw);
      if (bw * p.stride > 256 || bh * p.stride > 256) continue;
      const double tiles = (double)((p.Wo + bw - 1) / bw) * ((p.Ho + bh - 1) / bh);
      const double eff = (double)p.Wo * p.Ho / (tiles * 128.0);
      if (eff > best + 1e-9 || (eff > best - 1e-9 && bw > a.BW)) { best = eff; a.BW = bw; a.BH = bh; }
    }
    box[0] = a.BK; box[1] = a.BW * p.stride; box[2] = a.BH * p.stride; box[3] = 1;
    estr[0] = 1; estr[1] = p.stride; estr[2] = p.stride; estr[3] = 1;
  }
  CUresult cr = encode(&a.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled(A) failed with code " + std::to_string((int)cr);
    delete plan;
    return nullptr;
  }
  {
    const int K = p.k * p.k * p.Cin;
    cuuint64_t bdim[2] = {(cuuint64_t)K, (cuuint64_t)p.Cout};
    cuuint64_t bstr[1] = {(cuuint64_t)K * esz};
    cuuint32_t bbox[2] = {(cuuint32_t)a.BK, (cuuint32_t)a.n_tile};
    cuuint32_t bes[2] = {1, 1};
    cr = encode(&a.tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(p.w), bdim, bstr, bbox, bes,
                CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      if (err) *err = "cuTensorMapEncodeTiled(B) failed with code " + std::to_string((int)cr);
      delete plan;
      return nullptr;
    }
  }
  a.a_bytes = (uint32_t)(a.BW * a.BH * a.BK * 2);
  a.b_bytes = (uint32_t)(a.n_tile * a.BK * 2);
  a.a_stride = (uint32_t)((128 * a.BK * 2 + 1023) / 1024 * 1024);
  a.b_stride = (uint32_t)((a.n_tile * a.BK * 2 + 1023) / 1024 * 1024);
  const size_t budget = 200 * 1024;
  a.stages = (int)std::min<size_t>(TC_MAX_STAGES, budget / (a.a_stride + a.b_stride));
  if (a.stages < 2) {
    if (err) *err = "tile does not fit in shared memory";
    delete plan;
    return nullptr;
  }
  plan->smem = (size_t)a.stages * (a.a_stride + a.b_stride) + 1024;
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * a.n_tile)) cols <<= 1;
  a.tmem_cols = cols;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    // dynamic limit = ring budget + alignment slack (static smem of the kernel counts against the 227 KiB cap)
    cudaError_t ce = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
    if (ce != cudaSuccess) {
      if (err) *err = std::string("cudaFuncSetAttribute(conv_tc_kernel) failed: ") + cudaGetErrorString(ce);
      num_sms = 0;
      delete plan;
      return nullptr;
    }
  }
  plan->grid = num_sms;
  return plan;
}

void tc_conv_plan_destroy(TcConvPlan* plan) { delete plan; }

int tc_conv_launch(const TcConvPlan* plan, int B, cudaStream_t s) {
  TcArgs a = plan->args;
  const ConvParams& p = plan->p;
  if (plan->flat) {
    a.imgs = 1;
    a.Ho = 1;
    a.Wo = B * p.Ho * p.Wo;
    a.tiles_h = 1;
    a.tiles_w = (a.Wo + 127) / 128;
  } else {
    a.imgs = B;
    a.Ho = p.Ho; a.Wo = p.Wo;
    a.tiles_w = (p.Wo + a.BW - 1) / a.BW;
    a.tiles_h = (p.Ho + a.BH - 1) / a.BH;
  }
  a.total_tiles = a.imgs * a.tiles_w * a.tiles_h * a.n_tiles;
  const int grid = std::min(plan->grid, a.total_tiles);
  conv_tc_kernel<<<grid, TC_THREADS, plan->smem, s>>>(a);
  YB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// Stem: model.0 = Conv(3, C, k3, s2) straight from the NCHW network input (u8 / f16 / f32).
// K = 27 is too thin for the tensor cores and the layer is HBM-bound (reads the image once, writes
// C x H/2 x W/2 fp16); one thread per output pixel, weights broadcast from shared memory.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float stem_load(const void* in, int dtype, size_t i) {
  if (dtype == YB_U8) return (float)reinterpret_cast<const uint8_t*>(in)[i] * (1.0f / 255.0f);
  if (dtype == YB_F16) return __half2float(reinterpret_cast<const __half*>(in)[i]);
  return reinterpret_cast<const float*>(in)[i];
}

__global__ void __launch_bounds__(256) stem_kernel(const void* __restrict__ in, int dtype, int B, int H, int W, int Cout,
                                                   const float* __restrict__ w, const float* __restrict__ bias,
                                                   View out) {
  extern __shared__ float st_smem[];  // [27][Cout] weights + [Cout] bias
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) st_smem[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) st_smem[27 * Cout + i] = bias[i];
  __syncthreads();
  const int Ho = H / 2, Wo = W / 2;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * Ho * Wo) return;
  const int wo = idx % Wo;
  const int ho = (idx / Wo) % Ho;
  const int n = idx / ((size_t)Wo * Ho);
  float x[27];
#pragma unroll
  for (int kh = 0; kh < 3; kh++) {
    const int hi = ho * 2 + kh - 1;
#pragma unroll
    for (int kw = 0; kw < 3; kw++) {
      const int wi = wo * 2 + kw - 1;
      const bool ok = hi >= 0 && hi < H && wi >= 0 && wi < W;
#pragma unroll
      for (int c = 0; c < 3; c++)
        x[(kh * 3 + kw) * 3 + c] = ok ? stem_load(in, dtype, ((size_t)(n * 3 + c) * H + hi) * W + wi) : 0.f;
    }
  }
  __half* o = reinterpret_cast<__half*>(out.base) + idx * out.pitch + out.coff;
  for (int c0 = 0; c0 < Cout; c0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = st_smem[27 * Cout + c0 + j];
#pragma unroll
    for (int k = 0; k < 27; k++) {
      const float* wr = st_smem + k * Cout + c0;
#pragma unroll
      for (int j = 0; j < 8; j++) acc[j] = fmaf(x[k], wr[j], acc[j]);
    }
    int4 ov;
    __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
    for (int j = 0; j < 4; j++) oh[j] = __floats2half2_rn(silu_fast(acc[2 * j]), silu_fast(acc[2 * j + 1]));
    *reinterpret_cast<int4*>(o + c0) = ov;
  }
}

int launch_stem_f16(const void* in, int in_dtype, int B, int H, int W, const float* w, const float* bias,
                    const View& out, cudaStream_t s) {
  if (out.C % 8 || out.coff % 8 || out.pitch % 8) {
    set_error("stem: output channels must be a multiple of 8");
    return YB_ERR_SHAPE;
  }
  const size_t total = (size_t)B * (H / 2) * (W / 2);
  const size_t smem = (size_t)28 * out.C * sizeof(float);
  stem_kernel<<<(unsigned)((total + 255) / 256), 256, smem, s>>>(in, in_dtype, B, H, W, out.C, w, bias, out);
  YB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace yb
