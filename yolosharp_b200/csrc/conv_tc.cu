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
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace yb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn(std::string* err);

enum TcMode { TC_TAP = 0, TC_HALO = 1, TC_S2P = 2 };

struct TcArgs {
  CUtensorMap tmA;
  const uint8_t* wpk;  // weights pre-packed as shared-memory slab images [n tile][tap][chunk][b_stride bytes]
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
  int mode;                       // TcMode
  int stages_a, stages_b;
  uint32_t a_bytes, b_bytes;      // TMA transaction bytes per A slab / B slab
  uint32_t a_stride, b_stride;    // smem bytes reserved per slab (1 KiB aligned)
  uint32_t sbo_a, sbo_b;          // UMMA stride-byte-offset between 8-row groups, >> 4
  uint32_t row_bytes;             // BK * 2
  uint32_t layout_a, layout_b;    // UMMA LayoutType: 2 = SW128, 4 = SW64, 6 = SW32
  uint32_t tmem_cols;
  int n_issuers;                  // 1 or 2 MMA-issuing warps (each with its own half of the rings)
  int n_groups;                   // epilogue groups of 4 warps (2 or 4)
  int n_acc;                      // TMEM accumulator buffers (2 or 4)
  int total_tiles;
  int b_resident;                 // all weight slabs stay in smem for the CTA's lifetime
  int dual;                       // streamed weights: tiles are processed in pairs that share every weight slab
  int ksteps;
  // fused head decode (EpiDecode)
  int epi_mode, dA, dCtot, da0, dch0, dWl, dHW;
  float dstride;
  float* pred;
  // exact x / d for x*d < 2^40 as (x * ceil(2^40/d)) >> 40 (runtime integer division costs ~100+ cycles)
  uint64_t m_ntiles, m_tpi, m_tw, m_bw;
  // layer chaining (TcChain): per-image completion counters instead of a grid-wide dependency
  int* done_ctr;
  const int* dep_ctr;
  int dep_expect;
  uint64_t m_ohw;  // magic number for / (Ho * Wo) of this conv's output (image of a flattened pixel index)
  int* tile_ctr;   // dynamic tile scheduler: global counter of this launch (nullptr = static round-robin)
  int tile_batch;  // consecutive tiles drawn per atomicAdd (one counter address serves the whole grid)
  long long* dbg;  // optional timeline buffer (tools/exp_timeline.py); nullptr in production
};

struct TcConvPlan {
  TcArgs args;
  ConvParams p;
  bool flat;   // 1x1 stride-1 conv on the flattened pixel dimension
  int occ;     // CTAs per SM this plan is sized for
  int threads; // 3 role warps + 4 warps per epilogue group
  bool small;  // <= 2 tiles per CTA
  size_t smem;
  int grid;
  uint8_t* wpk = nullptr;  // device, owned: packed weight slabs
};

__device__ __forceinline__ int fdiv(int x, uint64_t magic) { return (int)(((uint64_t)(uint32_t)x * magic) >> 40); }

__device__ __forceinline__ float silu_fast(float v) { return __fdividef(v, 1.0f + __expf(-v)); }
// x*sigmoid(x) = h + h*tanh(h), h = x/2: one MUFU op instead of two (ex2 + rcp)
__device__ __forceinline__ float silu_tanh(float v) {
  const float h = 0.5f * v;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

constexpr int TC_MAX_THREADS = 608;  // warp 0 TMA producer, warps 1-2 MMA issuers, warps 3.. epilogue groups of 4 warps
constexpr int TC_MAX_GROUPS = 4;
constexpr int TC_ISSUERS = 2;
constexpr int TC_MAX_ACC = 4;
constexpr int TC_MAX_COUT = 1024;
constexpr int TC_MAX_STAGES = 12;
constexpr int HALO_BW = 8, HALO_BH = 16;  // output rectangle of a halo tile (128 rows)

// ------------------------------------------------------------------------------------------
// MMA issue loop.  One thread issues every tcgen05.mma of the CTA, so its per-instruction overhead is
// the kernel's pace for small N (a 128x64x16 MMA occupies the tensor pipe for ~34 cycles; the measured
// single-thread issue floor is ~55 cycles, tools/exp_mma_issue.cu).  Everything that can be hoisted is
// hoisted: descriptor high words are loop constants, low words advance by compile-time amounts
// (KK = BK/16 MMAs per slab, tap shifts of the halo tile), kernel parameters live in registers.
// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// Dynamic tile scheduler.  The producer thread draws tile indices from a global counter (first tile =
// blockIdx.x, then gridDim.x + atomicAdd) and publishes them in a small shared-memory queue that the MMA
// issuers and epilogue groups follow.  CTAs that start late - SMs held by a concurrent kernel (NMS of the
// previous batch, a sibling head branch) - simply take fewer tiles instead of delaying the whole grid with a
// fixed share.  The queue cannot wrap onto unread entries: the producer is at most stages_a + n_acc + groups
// (< 32) tiles ahead of the slowest reader.
// ------------------------------------------------------------------------------------------
constexpr int TQ = 32;
__device__ __forceinline__ void tq_publish(int* s_tile, volatile int* s_head, int li, int tile) {
  s_tile[li & (TQ - 1)] = tile;
  __threadfence_block();
  *s_head = li + 1;
}
__device__ __forceinline__ int tq_get(const int* s_tile, const volatile int* s_head, int li) {
  const long long t0 = clock64();
  while (*s_head <= li)
    if (clock64() - t0 > 4000000000ll) __trap();
  __threadfence_block();
  return reinterpret_cast<const volatile int*>(s_tile)[li & (TQ - 1)];
}

// Producer-side dependency wait of the layer chain: poll the producer launch's counter of image `img` until all of
// its rows are stored (acquire), bounded like every other wait of this kernel.
__device__ __forceinline__ void dep_wait_image(const int* ctr, int img, int expect) {
  const long long t0 = clock64();
  while (true) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ctr + img) : "memory");
    if (v >= expect) break;
    __nanosleep(100);
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}

__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

// Residual (Bottleneck shortcut) values of this lane's own row for one 32/16-column block, issued BEFORE
// the accumulator is awaited so the global-load latency overlaps the MMAs.
__device__ __forceinline__ void res_prefetch(const TcArgs& a, int n0, int cb0, size_t pix, bool valid, int4 (&rv)[4]) {
  const int wb = min(32, a.n_tile - cb0);
  const __half* rrow = a.res + pix * a.res_pitch + a.res_coff + n0 + cb0;
#pragma unroll
  for (int g = 0; g < 4; g++) {
    rv[g] = make_int4(0, 0, 0, 0);
    // L2-only load: with layer chaining a neighbouring row of the same 128-byte line may still be unwritten when
    // this one is read, and a line cached in L1 now would be stale when that row's own tile reads it later
    if (valid && g * 8 < wb) rv[g] = __ldcg(reinterpret_cast<const int4*>(rrow + g * 8));
  }
}

// every lane of the issuing warp waits; the warp is converged again before the next elect.sync
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
  mbar_wait(bar, parity);
  __syncwarp();
}

template <int KK>
__device__ __forceinline__ void mma_role(const TcArgs& a, uint32_t smemA, uint32_t smemB, uint32_t tmem_base,
                                         uint32_t fullA, uint32_t emptyA, uint32_t fullB, uint32_t emptyB,
                                         uint32_t tfull0, uint32_t tempty0, uint32_t bfull, int issuer,
                                         const int* s_tile, const volatile int* s_head) {
  const uint32_t n_tile = a.n_tile;
  const uint32_t idesc = (1u << 4) | ((n_tile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  // descriptor = lo | hi << 32 :  lo = start>>4 | LBO(1)<<16 ;  hi = SBO | version(1)<<14 | layout<<29
  const uint32_t a_hi = a.sbo_a | (1u << 14) | (a.layout_a << 29);
  const uint32_t b_hi = a.sbo_b | (1u << 14) | (a.layout_b << 29);
  const uint32_t lo_flags = 1u << 16;
  const uint32_t a_stride16 = a.a_stride >> 4, b_stride16 = a.b_stride >> 4;
  const uint32_t a_lo0 = ((smemA & 0x3FFFF) >> 4) | lo_flags, b_lo0 = ((smemB & 0x3FFFF) >> 4) | lo_flags;
  const int stages_a = a.stages_a, stages_b = a.stages_b, chunks = a.chunks, ksteps = a.ksteps;
  const bool resident = a.b_resident != 0, halo = a.mode != TC_TAP, s2p = a.mode == TC_S2P;
  const int total_tiles = a.total_tiles, gstride = gridDim.x;
  constexpr uint32_t ROW16 = KK * 2;  // bytes per operand row / 16
  // Two issuer warps take alternate tiles (local tile index li = issuer, issuer+2, ...): one thread tops
  // out at ~55-80 cycles per tcgen05.mma, two keep the tensor pipe fed for N <= 64.  Each issuer owns its
  // own half of the A (and B) ring - the producer fills ring (li % n_issuers) for tile li - so every
  // mbarrier is waited on strictly in phase order by exactly one thread (a parity wait cannot tell
  // "two phases behind" from "done").
  const int nb = a.n_acc;                      // TMEM accumulator buffers
  const int ni = a.n_issuers;
  if (issuer >= ni) return;
  const int ra = stages_a / ni, rb = resident ? 0 : stages_b / ni;  // ring depth per issuer
  const int a_base = issuer * ra, b_base = issuer * rb;
  int sa = 0, sb = 0;
  uint32_t pa = 0, pb = 0;
  if (resident) mbar_wait_warp(bfull, 0);
  const bool dyn = a.tile_ctr != nullptr;
  for (int li = issuer;; li += ni) {
    const int tile = dyn ? tq_get(s_tile, s_head, li) : (int)blockIdx.x + li * gstride;
    __syncwarp();
    if (tile < 0 || tile >= total_tiles) break;
    const int dbg_i = li;
    const int acc = li & (nb - 1);  // nb is 2 or 4
    const uint32_t aphase = (uint32_t)(li >> (nb == 4 ? 2 : 1)) & 1u;
    if (a.dbg && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && dbg_i < 16) a.dbg[dbg_i * 8 + 0] = clock64();
    mbar_wait_warp(tempty0 + 8 * acc, aphase ^ 1);
    tc_fence_after();
    if (a.dbg && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && dbg_i < 16) a.dbg[dbg_i * 8 + 1] = clock64();
    const uint32_t d_tmem = tmem_base + acc * n_tile;
    uint32_t accf = 0;
    if (halo) {
      for (int ch = 0; ch < chunks; ch++) {
        mbar_wait_warp(fullA + 8 * (a_base + sa), pa);
        tc_fence_after();
        if (a.dbg && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && dbg_i < 16 && ch == 0) a.dbg[dbg_i * 8 + 2] = clock64();
        const uint32_t a_lo = a_lo0 + (a_base + sa) * a_stride16;
#pragma unroll
        for (int t = 0; t < 9; t++) {
          // row shift of tap t inside the staged input tile (compile-time constants after unrolling):
          //   halo : (kh * (BW+2) + kw) pixel rows
          //   s2p  : pair rows (input pixels 2q, 2q+1), the tile starts at pair w0-1: kw = 0 is the second half
          //          of pair j, kw = 1 / 2 the two halves of pair j+1; kh advances one input row = (BW+1) pairs
          const uint32_t TAP_HALO = (uint32_t)((t / 3) * (HALO_BW + 2) + (t % 3)) * ROW16;
          const uint32_t TAP_S2P = (uint32_t)((t / 3) * (HALO_BW + 1) + (t % 3 != 0 ? 1 : 0)) * 2 * ROW16 +
                                       (t % 3 != 1 ? ROW16 : 0);
          const uint32_t tap16 = s2p ? TAP_S2P : TAP_HALO;
          uint32_t b_lo;
          if (resident) {
            b_lo = b_lo0 + (t * chunks + ch) * b_stride16;
          } else {
            mbar_wait_warp(fullB + 8 * (b_base + sb), pb);
            tc_fence_after();
            b_lo = b_lo0 + (b_base + sb) * b_stride16;
          }
#pragma unroll
          for (int k = 0; k < KK; k++) {  // +32 B per K=16 step inside the swizzled row
            umma_f16_elect(d_tmem, desc64(a_lo + tap16 + 2 * k, a_hi), desc64(b_lo + 2 * k, b_hi), idesc, accf);
            accf = 1;
          }
          if (!resident) {
            umma_commit_elect(emptyB + 8 * (b_base + sb));
            if (++sb == rb) { sb = 0; pb ^= 1; }
          }
        }
        umma_commit_elect(emptyA + 8 * (a_base + sa));  // halo tile free once its 9 taps retired
        if (++sa == ra) { sa = 0; pa ^= 1; }
      }
    } else {
      for (int ks = 0; ks < ksteps; ks++) {
        mbar_wait_warp(fullA + 8 * (a_base + sa), pa);
        uint32_t b_lo;
        if (resident) {
          b_lo = b_lo0 + ks * b_stride16;
        } else {
          mbar_wait_warp(fullB + 8 * (b_base + sb), pb);
          b_lo = b_lo0 + (b_base + sb) * b_stride16;
        }
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + (a_base + sa) * a_stride16;
#pragma unroll
        for (int k = 0; k < KK; k++) {
          umma_f16_elect(d_tmem, desc64(a_lo + 2 * k, a_hi), desc64(b_lo + 2 * k, b_hi), idesc, accf);
          accf = 1;
        }
        umma_commit_elect(emptyA + 8 * (a_base + sa));  // slab free once these MMAs retire
        if (++sa == ra) { sa = 0; pa ^= 1; }
        if (!resident) {
          umma_commit_elect(emptyB + 8 * (b_base + sb));
          if (++sb == rb) { sb = 0; pb ^= 1; }
        }
      }
    }
    umma_commit_elect(tfull0 + 8 * acc);  // accumulator complete
    if (a.dbg && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && dbg_i < 16) a.dbg[dbg_i * 8 + 3] = clock64();
  }
}

// Streamed-weight layers (weights too large to stay resident): two tiles are accumulated side by side so that
// every weight slab fetched from L2 feeds the MMAs of BOTH.  The timeline of a 3x3 160->160 layer (v8x) showed the
// issuer waiting for weight slabs ~250 cycles per MMA: 148 SMs re-streaming the same 460 KB per 128-pixel tile run
// into the L2 -> SM delivery limit, not into the tensor pipe.  One issuer (wide N: the tensor pipe is the pace).
template <int KK>
__device__ __forceinline__ void mma_role_dual(const TcArgs& a, uint32_t smemA, uint32_t smemB, uint32_t tmem_base,
                                              uint32_t fullA, uint32_t emptyA, uint32_t fullB, uint32_t emptyB,
                                              uint32_t tfull0, uint32_t tempty0) {
  const uint32_t n_tile = a.n_tile;
  const uint32_t idesc = (1u << 4) | ((n_tile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t a_hi = a.sbo_a | (1u << 14) | (a.layout_a << 29);
  const uint32_t b_hi = a.sbo_b | (1u << 14) | (a.layout_b << 29);
  const uint32_t a_stride16 = a.a_stride >> 4, b_stride16 = a.b_stride >> 4;
  const uint32_t a_lo0 = ((smemA & 0x3FFFF) >> 4) | (1u << 16), b_lo0 = ((smemB & 0x3FFFF) >> 4) | (1u << 16);
  const int ra = a.stages_a, rb = a.stages_b, chunks = a.chunks, ksteps = a.ksteps, nb = a.n_acc;
  const bool halo = a.mode != TC_TAP, s2p = a.mode == TC_S2P;
  const int total_tiles = a.total_tiles, gstride = gridDim.x;
  constexpr uint32_t ROW16 = KK * 2;
  int sa = 0, sb = 0;
  uint32_t pa = 0, pb = 0;
  int li = 0;
  for (int tile0 = blockIdx.x; tile0 < total_tiles; tile0 += 2 * gstride, li += 2) {
    const int nT = (tile0 + gstride < total_tiles) ? 2 : 1;
    uint32_t d_tmem[2] = {0, 0};
    int acc[2] = {0, 0};
    for (int q = 0; q < nT; q++) {
      acc[q] = (li + q) & (nb - 1);
      const uint32_t aphase = (uint32_t)((li + q) >> (nb == 4 ? 2 : 1)) & 1u;
      mbar_wait_warp(tempty0 + 8 * acc[q], aphase ^ 1);
      d_tmem[q] = tmem_base + acc[q] * n_tile;
    }
    tc_fence_after();
    uint32_t alo[2] = {0, 0};
    int stq[2] = {0, 0};
    if (halo) {
      for (int ch = 0; ch < chunks; ch++) {
        for (int q = 0; q < nT; q++) {
          mbar_wait_warp(fullA + 8 * sa, pa);
          alo[q] = a_lo0 + sa * a_stride16;
          stq[q] = sa;
          if (++sa == ra) { sa = 0; pa ^= 1; }
        }
        tc_fence_after();
#pragma unroll
        for (int t = 0; t < 9; t++) {
          const uint32_t TAP_HALO = (uint32_t)((t / 3) * (HALO_BW + 2) + (t % 3)) * ROW16;
          const uint32_t TAP_S2P = (uint32_t)((t / 3) * (HALO_BW + 1) + (t % 3 != 0 ? 1 : 0)) * 2 * ROW16 +
                                   (t % 3 != 1 ? ROW16 : 0);
          const uint32_t tap16 = s2p ? TAP_S2P : TAP_HALO;
          mbar_wait_warp(fullB + 8 * sb, pb);
          tc_fence_after();
          const uint32_t b_lo = b_lo0 + sb * b_stride16;
          const uint32_t first = (ch == 0 && t == 0) ? 1u : 0u;
#pragma unroll
          for (int q = 0; q < 2; q++) {
            if (q < nT) {
#pragma unroll
              for (int k = 0; k < KK; k++)
                umma_f16_elect(d_tmem[q], desc64(alo[q] + tap16 + 2 * k, a_hi), desc64(b_lo + 2 * k, b_hi), idesc,
                         (first && k == 0) ? 0u : 1u);
            }
          }
          umma_commit_elect(emptyB + 8 * sb);
          if (++sb == rb) { sb = 0; pb ^= 1; }
        }
        for (int q = 0; q < nT; q++) umma_commit_elect(emptyA + 8 * stq[q]);
      }
    } else {
      for (int ks = 0; ks < ksteps; ks++) {
        for (int q = 0; q < nT; q++) {
          mbar_wait_warp(fullA + 8 * sa, pa);
          alo[q] = a_lo0 + sa * a_stride16;
          stq[q] = sa;
          if (++sa == ra) { sa = 0; pa ^= 1; }
        }
        mbar_wait_warp(fullB + 8 * sb, pb);
        tc_fence_after();
        const uint32_t b_lo = b_lo0 + sb * b_stride16;
#pragma unroll
        for (int q = 0; q < 2; q++) {
          if (q < nT) {
#pragma unroll
            for (int k = 0; k < KK; k++)
              umma_f16_elect(d_tmem[q], desc64(alo[q] + 2 * k, a_hi), desc64(b_lo + 2 * k, b_hi), idesc,
                       (ks == 0 && k == 0) ? 0u : 1u);
          }
        }
        for (int q = 0; q < nT; q++) umma_commit_elect(emptyA + 8 * stq[q]);
        umma_commit_elect(emptyB + 8 * sb);
        if (++sb == rb) { sb = 0; pb ^= 1; }
      }
    }
    for (int q = 0; q < nT; q++) umma_commit_elect(tfull0 + 8 * acc[q]);  // both accumulators complete
  }
}

// ------------------------------------------------------------------------------------------
// Two operand rings feed the single MMA-issuing thread:
//   A ring  : TC_TAP  - one 128-row slab per (tap, channel slab)
//             TC_HALO - one (BH+2)x(BW+2)-pixel halo tile per channel slab; the 9 taps of a 3x3
//                       stride-1 conv are issued from the SAME bytes with the descriptor start
//                       shifted by (kh*(BW+2)+kw) rows and SBO = (BW+2) rows.  tcgen05 applies the
//                       128B/64B/32B swizzle on absolute smem address bits, so a row-shifted start
//                       needs no base-offset (verified on B200 by tools/exp_umma_shift.cu).
//                       -> 180 TMA rows per slab instead of 9 x 128: the kernel is bound by L2
//                       request rate (~30 requests/clk chip-wide), not by bytes.
//   B ring  : one [n_tile x BK] weight slab per (tap, channel slab), or all slabs resident.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(704, 1) conv_tc_kernel(  // 704 = 2 x 352: caps registers at 93 so two 352-thread CTAs fit an SM
const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  __shared__ __align__(8) uint64_t bars[4 * TC_MAX_STAGES + 2 * TC_MAX_ACC + 1];
  __shared__ uint32_t tmem_base_slot;
  __shared__ int s_tile[TQ];
  __shared__ volatile int s_head;
  __shared__ __align__(16) float s_bias[TC_MAX_COUT];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < a.n_tile * a.n_tiles; i += blockDim.x) s_bias[i] = a.bias[i];  // constant data
  // dynamic smem base rounded up to 1 KiB (SWIZZLE_128B atoms need it)
  const uint32_t smem0 = (smem_u32(tc_smem) + 1023u) & ~1023u;
  const uint32_t smemA = smem0;
  const uint32_t smemB = smem0 + a.stages_a * a.a_stride;
  const uint32_t fullA = smem_u32(&bars[0]);
  const uint32_t emptyA = smem_u32(&bars[TC_MAX_STAGES]);
  const uint32_t fullB = smem_u32(&bars[2 * TC_MAX_STAGES]);
  const uint32_t emptyB = smem_u32(&bars[3 * TC_MAX_STAGES]);
  const uint32_t tfull0 = smem_u32(&bars[4 * TC_MAX_STAGES]);
  const uint32_t tempty0 = smem_u32(&bars[4 * TC_MAX_STAGES + TC_MAX_ACC]);
  const uint32_t bfull = smem_u32(&bars[4 * TC_MAX_STAGES + 2 * TC_MAX_ACC]);

  // PDL: let the next kernel of the stream/graph start its prologue while this grid runs; it blocks in
  // its own griddepcontrol.wait until this grid has completed and flushed.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0 && lane == 0) {
    s_head = 0;
    for (int s = 0; s < a.stages_a; s++) {
      mbar_init(fullA + 8 * s, 1);
      mbar_init(emptyA + 8 * s, 1);
    }
    for (int s = 0; s < a.stages_b; s++) {
      mbar_init(fullB + 8 * s, 1);
      mbar_init(emptyB + 8 * s, 1);
    }
    for (int s = 0; s < a.n_acc; s++) {
      mbar_init(tfull0 + 8 * s, 1);
      mbar_init(tempty0 + 8 * s, 4);  // one arrive per warp of the draining epilogue group
    }
    mbar_init(bfull, 1);
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

  const int taps = a.ksz * a.ksz;
  const int tiles_per_img = a.tiles_w * a.tiles_h;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&a.tmA) : "memory");
    if (a.b_resident) {
      // weights do not depend on the previous kernel: fetch them before the grid dependency resolves
      mbar_arrive_expect_tx(bfull, a.b_bytes * a.ksteps);
      for (int ks = 0; ks < a.ksteps; ks++)
        bulk_load_1d(smemB + ks * a.b_stride, a.wpk + (size_t)ks * a.b_stride, a.b_bytes, bfull);
    }
  }
  // activations written by the previous kernel are visible only after this point - unless this launch is chained to
  // its producer by per-image counters (dep_ctr): then tiles start as soon as their images are complete
  if (a.dep_ctr == nullptr) asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (a.dual) {
        // pairs of tiles (same N tile: the grid is a multiple of n_tiles) share every weight slab
        const int ra = a.stages_a, rb = a.stages_b;
        int sa = 0, sb = 0;
        uint32_t pa = 0, pb = 0;
        for (int tile0 = blockIdx.x; tile0 < a.total_tiles; tile0 += 2 * gridDim.x) {
          const int nT = (tile0 + (int)gridDim.x < a.total_tiles) ? 2 : 1;
          int img_[2], wc_[2], hb_[2], nt = 0;
          for (int q = 0; q < nT; q++) {
            const int tile = tile0 + q * gridDim.x;
            const int mt = fdiv(tile, a.m_ntiles);
            nt = tile - mt * a.n_tiles;
            img_[q] = fdiv(mt, a.m_tpi);
            const int r = mt - img_[q] * tiles_per_img;
            const int th = fdiv(r, a.m_tw), tw = r - th * a.tiles_w;
            wc_[q] = a.mode == TC_S2P ? tw * a.BW - 1 : tw * a.BW * a.stride - a.pad;
            hb_[q] = th * a.BH * a.stride - a.pad;
          }
          auto load_a = [&](int q, int c0, int dw, int dh) {
            mbar_wait(emptyA + 8 * sa, pa ^ 1);
            mbar_arrive_expect_tx(fullA + 8 * sa, a.a_bytes);
            tma_load_4d(smemA + sa * a.a_stride, &a.tmA, fullA + 8 * sa, c0, wc_[q] + dw, hb_[q] + dh, img_[q]);
            if (++sa == ra) { sa = 0; pa ^= 1; }
          };
          auto load_b = [&](int t, int ch) {
            mbar_wait(emptyB + 8 * sb, pb ^ 1);
            mbar_arrive_expect_tx(fullB + 8 * sb, a.b_bytes);
            bulk_load_1d(smemB + sb * a.b_stride, a.wpk + (size_t)((nt * taps + t) * a.chunks + ch) * a.b_stride, a.b_bytes,
                         fullB + 8 * sb);
            if (++sb == rb) { sb = 0; pb ^= 1; }
          };
          if (a.mode != TC_TAP) {
            for (int ch = 0; ch < a.chunks; ch++) {
              for (int q = 0; q < nT; q++) load_a(q, ch * a.BK, 0, 0);
              for (int t = 0; t < taps; t++) load_b(t, ch);
            }
          } else {
            for (int t = 0; t < taps; t++) {
              const int kh = a.ksz == 3 ? (t >= 6 ? 2 : (t >= 3 ? 1 : 0)) : 0, kw = t - kh * a.ksz;
              for (int ch = 0; ch < a.chunks; ch++) {
                for (int q = 0; q < nT; q++) load_a(q, ch * a.BK, kw, kh);
                load_b(t, ch);
              }
            }
          }
        }
      }
      const int ni = a.n_issuers;
      const int ra = a.stages_a / ni, rb = a.b_resident ? 0 : a.stages_b / ni;
      int sa_[2] = {0, 0}, sb_[2] = {0, 0};
      uint32_t pa_[2] = {0, 0}, pb_[2] = {0, 0};
      int li = 0;
      const bool dyn = a.tile_ctr != nullptr;
      int tile = a.dual ? a.total_tiles : (int)blockIdx.x;
      int nxt = a.total_tiles, left = 1;
      const int batch = a.tile_batch;
      // the next batch of tiles is drawn one batch ahead, so the atomic's latency never sits on the load path
      if (dyn && tile < a.total_tiles) nxt = (int)gridDim.x + atomicAdd(a.tile_ctr, batch);
      int dep_ready = -1;  // last image of the producer known to be complete
      for (; tile < a.total_tiles; li++) {
        if (a.dep_ctr != nullptr) {
          // images this tile reads: its own image, or for a flattened 1x1 tile the images of its first / last pixel
          const int mt0 = fdiv(tile, a.m_ntiles);
          int i0, i1;
          if (a.imgs == 1 && a.Ho == 1) {
            const int p0 = mt0 * 128, p1 = min(p0 + 127, a.Wo - 1);
            i0 = fdiv(p0, a.m_ohw); i1 = fdiv(p1, a.m_ohw);
          } else {
            i0 = i1 = fdiv(mt0, a.m_tpi);
          }
          if (i1 != dep_ready || i0 != i1) {
            for (int im = i0; im <= i1; im++)
              if (im != dep_ready) dep_wait_image(a.dep_ctr, im, a.dep_expect);
            dep_ready = i1;
            asm volatile("fence.proxy.async.global;" ::: "memory");  // generic-proxy acquire -> async-proxy (TMA) reads
          }
        }
        if (dyn) tq_publish(s_tile, &s_head, li, tile);
        const int rg = ni == 2 ? (li & 1) : 0;  // ring (= issuer) of this tile
        int& sa = sa_[rg]; int& sb = sb_[rg];
        uint32_t& pa = pa_[rg]; uint32_t& pb = pb_[rg];
        const int a_base = rg * ra, b_base = rg * rb;
        const int mt = fdiv(tile, a.m_ntiles);
        const int nt = tile - mt * a.n_tiles;
        const int img = fdiv(mt, a.m_tpi);
        const int r = mt - img * tiles_per_img;
        const int th = fdiv(r, a.m_tw), tw = r - th * a.tiles_w;
        const int wbase = tw * a.BW * a.stride - a.pad;
        const int hbase = th * a.BH * a.stride - a.pad;
        if (a.mode != TC_TAP) {
          for (int ch = 0; ch < a.chunks; ch++) {
            mbar_wait(emptyA + 8 * (a_base + sa), pa ^ 1);
            mbar_arrive_expect_tx(fullA + 8 * (a_base + sa), a.a_bytes);
            tma_load_4d(smemA + (a_base + sa) * a.a_stride, &a.tmA, fullA + 8 * (a_base + sa), ch * a.BK,
                        a.mode == TC_S2P ? tw * a.BW - 1 : wbase, hbase, img);
            if (++sa == ra) { sa = 0; pa ^= 1; }
            if (!a.b_resident)
              for (int t = 0; t < taps; t++) {
                mbar_wait(emptyB + 8 * (b_base + sb), pb ^ 1);
                mbar_arrive_expect_tx(fullB + 8 * (b_base + sb), a.b_bytes);
                bulk_load_1d(smemB + (b_base + sb) * a.b_stride,
                             a.wpk + (size_t)((nt * taps + t) * a.chunks + ch) * a.b_stride, a.b_bytes,
                             fullB + 8 * (b_base + sb));
                if (++sb == rb) { sb = 0; pb ^= 1; }
              }
          }
        } else {
          for (int t = 0; t < taps; t++) {
            const int kh = a.ksz == 3 ? (t >= 6 ? 2 : (t >= 3 ? 1 : 0)) : 0, kw = t - kh * a.ksz;
            for (int ch = 0; ch < a.chunks; ch++) {
              mbar_wait(emptyA + 8 * (a_base + sa), pa ^ 1);
              mbar_arrive_expect_tx(fullA + 8 * (a_base + sa), a.a_bytes);
              tma_load_4d(smemA + (a_base + sa) * a.a_stride, &a.tmA, fullA + 8 * (a_base + sa), ch * a.BK, wbase + kw,
                          hbase + kh, img);
              if (++sa == ra) { sa = 0; pa ^= 1; }
              if (!a.b_resident) {
                mbar_wait(emptyB + 8 * (b_base + sb), pb ^ 1);
                mbar_arrive_expect_tx(fullB + 8 * (b_base + sb), a.b_bytes);
                bulk_load_1d(smemB + (b_base + sb) * a.b_stride,
                             a.wpk + (size_t)((nt * taps + t) * a.chunks + ch) * a.b_stride, a.b_bytes,
                             fullB + 8 * (b_base + sb));
                if (++sb == rb) { sb = 0; pb ^= 1; }
              }
            }
          }
        }
        if (dyn) {
          if (--left > 0) {
            tile++;
          } else {
            tile = nxt;
            left = batch;
            if (tile < a.total_tiles) nxt = (int)gridDim.x + atomicAdd(a.tile_ctr, batch);
          }
        } else {
          tile += gridDim.x;
        }
      }
      if (dyn && !a.dual) {  // end marks for every reader (2 issuers, up to 4 epilogue groups)
        for (int k = 0; k < 3; k++) s_tile[(li + k) & (TQ - 1)] = -1;
        tq_publish(s_tile, &s_head, li + 3, -1);
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ===================== MMA issuers (whole warp, elect.sync issues) =====================
    {
      const int issuer = warp - 1;
      if (a.dual) {
        if (issuer == 0) {
          switch (a.BK) {
            case 64: mma_role_dual<4>(a, smemA, smemB, tmem_base, fullA, emptyA, fullB, emptyB, tfull0, tempty0); break;
            case 32: mma_role_dual<2>(a, smemA, smemB, tmem_base, fullA, emptyA, fullB, emptyB, tfull0, tempty0); break;
            default: mma_role_dual<1>(a, smemA, smemB, tmem_base, fullA, emptyA, fullB, emptyB, tfull0, tempty0); break;
          }
        }
      } else
      switch (a.BK) {
        case 64: mma_role<4>(a, smemA, smemB, tmem_base, fullA, emptyA, fullB, emptyB, tfull0, tempty0, bfull, issuer, s_tile, &s_head); break;
        case 32: mma_role<2>(a, smemA, smemB, tmem_base, fullA, emptyA, fullB, emptyB, tfull0, tempty0, bfull, issuer, s_tile, &s_head); break;
        default: mma_role<1>(a, smemA, smemB, tmem_base, fullA, emptyA, fullB, emptyB, tfull0, tempty0, bfull, issuer, s_tile, &s_head); break;
      }
    }
  } else {
    // ===================== epilogue (warps 3 ..) =====================
    // G groups of four warps (one warp per TMEM lane quarter: a warp may only read lanes
    // 32*(warp%4)..+31; any four consecutive warps cover all quarters) drain different tiles: group g
    // takes local tiles g, g+G, ...  The per-tile epilogue is a latency-bound ~2000-4000 cycle chain
    // (tcgen05.ld -> bias -> SiLU -> pack -> store), so its THROUGHPUT comes from having several tiles in
    // their epilogue at once (timeline: one tile at a time paced the whole kernel for N <= 64).
    const int ew = warp - 3;
    const int q = warp & 3;
    const int grp = ew >> 2;
    const int G = a.n_groups;
    const int row = q * 32 + lane;
    const bool dyn = a.tile_ctr != nullptr;
    int pend_img = -1, pend_cnt = 0;  // layer chaining: rows stored but not yet published
    for (int li = grp;; li += G) {
      const int tile = dyn ? tq_get(s_tile, &s_head, li) : (int)blockIdx.x + li * (int)gridDim.x;
      if (tile < 0 || tile >= a.total_tiles) break;
      const int acc = li & (a.n_acc - 1);  // n_acc is 2 or 4
      const uint32_t aphase = (uint32_t)(li >> (a.n_acc == 4 ? 2 : 1)) & 1u;
      const int mt = fdiv(tile, a.m_ntiles);
      const int nt = tile - mt * a.n_tiles;
      const int img = fdiv(mt, a.m_tpi);
      const int r = mt - img * tiles_per_img;
      const int th = fdiv(r, a.m_tw), tw = r - th * a.tiles_w;
      const int hl = fdiv(row, a.m_bw), wl = row - hl * a.BW;
      const int ho = th * a.BH + hl, wo = tw * a.BW + wl;
      const bool valid = hl < a.BH && ho < a.Ho && wo < a.Wo;
      const size_t pix = ((size_t)img * a.Ho + ho) * a.Wo + wo;
      const int n0 = nt * a.n_tile;
      const float* bias = s_bias + n0;
      int4 rv[4];
      const bool use_res = a.epi_mode == EPI_STORE && a.res != nullptr;
      if (use_res) res_prefetch(a, n0, 0, pix, valid, rv);

      const int dbg_t = li;
      const bool dbg_on = a.dbg && blockIdx.x == 0 && (warp & 3) == 3 && lane == 0 && dbg_t < 16;
      if (dbg_on) a.dbg[dbg_t * 8 + 4] = clock64();
      mbar_wait(tfull0 + 8 * acc, aphase);
      tc_fence_after();
      if (dbg_on) a.dbg[dbg_t * 8 + 5] = clock64();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * a.n_tile;
      if (a.epi_mode != EPI_STORE) {
        // fused Detect tail (flattened 1x1 conv): this thread's row is pixel `wo` of the whole batch
        const int n = valid ? wo / a.dHW : 0;
        const int i = wo - n * a.dHW;
        float* po = a.pred + (size_t)n * a.dCtot * a.dA + a.da0 + i;
        if (a.epi_mode == EPI_DFL_BOX) {
          // DFL (Block.cs:44): softmax over the 16 bins of each side, expectation with weights 0..15;
          // then dist2bbox(xywh) * stride (Tal.cs:338-356, Head.cs:221)
          float d[4];
#pragma unroll
          for (int sd = 0; sd < 4; sd++) {
            uint32_t v[16];
            tmem_ld16(taddr + sd * 16, v);
            tmem_ld_wait();
            float f[16], mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 16; j++) { f[j] = __uint_as_float(v[j]) + bias[sd * 16 + j]; mx = fmaxf(mx, f[j]); }
            float sum = 0.f, ex = 0.f;
#pragma unroll
            for (int j = 0; j < 16; j++) { const float e = __expf(f[j] - mx); sum += e; ex = fmaf(e, (float)j, ex); }
            d[sd] = __fdividef(ex, sum);
          }
          if (valid) {
            const int y = i / a.dWl, x = i - y * a.dWl;
            const float ax = (float)x + 0.5f, ay = (float)y + 0.5f;
            const float x1 = ax - d[0], y1 = ay - d[1], x2 = ax + d[2], y2 = ay + d[3];
            po[0] = (x1 + x2) * 0.5f * a.dstride;
            po[(size_t)a.dA] = (y1 + y2) * 0.5f * a.dstride;
            po[(size_t)2 * a.dA] = (x2 - x1) * a.dstride;
            po[(size_t)3 * a.dA] = (y2 - y1) * a.dstride;
          }
        } else {
          for (int c0 = 0; c0 < a.n_tile; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + c0, v);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
              for (int j = 0; j < 16; j++) {
                float f = __uint_as_float(v[j]) + bias[c0 + j];
                if (a.epi_mode == EPI_SIGMOID) f = __fdividef(1.0f, 1.0f + __expf(-f));
                po[(size_t)(a.dch0 + c0 + j) * a.dA] = f;  // lanes = consecutive anchors: coalesced
              }
            }
          }
        }
      } else {
        // Store path, 32 columns at a time.  A lane owns one output pixel (TMEM lane) and writes its own
        // 64 contiguous bytes per block with four 16-byte stores.
        // (A smem-transposed variant that made every warp store cover whole rows cut L2 requests 8x but
        // cost ~1000 cycles of shuffles / smem round trips per tile; ncu shows L2 far from saturated,
        // so the short instruction path wins.)
        __half* orow = a.out + pix * a.out_pitch + a.out_coff + n0;
        for (int cb0 = 0; cb0 < a.n_tile; cb0 += 32) {
          const int wb = min(32, a.n_tile - cb0);  // 32 or 16 channels
          uint32_t v[32];
          if (wb == 32) tmem_ld32(taddr + cb0, v); else tmem_ld16(taddr + cb0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; g++) {  // 8 channels -> one 16-byte store
            if (g * 8 < wb) {
              const float4 b0 = *reinterpret_cast<const float4*>(bias + cb0 + g * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(bias + cb0 + g * 8 + 4);
              float f[8];
              f[0] = __uint_as_float(v[g * 8 + 0]) + b0.x; f[1] = __uint_as_float(v[g * 8 + 1]) + b0.y;
              f[2] = __uint_as_float(v[g * 8 + 2]) + b0.z; f[3] = __uint_as_float(v[g * 8 + 3]) + b0.w;
              f[4] = __uint_as_float(v[g * 8 + 4]) + b1.x; f[5] = __uint_as_float(v[g * 8 + 5]) + b1.y;
              f[6] = __uint_as_float(v[g * 8 + 6]) + b1.z; f[7] = __uint_as_float(v[g * 8 + 7]) + b1.w;
              if (a.act == ACT_SILU) {
#pragma unroll
                for (int j = 0; j < 8; j++) f[j] = silu_tanh(f[j]);
              }
              if (use_res) {
                const __half2* h = reinterpret_cast<const __half2*>(&rv[g]);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                  const float2 x = __half22float2(h[j]);
                  f[2 * j] += x.x; f[2 * j + 1] += x.y;
                }
              }
              int4 o;
              __half2* ph = reinterpret_cast<__half2*>(&o);
#pragma unroll
              for (int j = 0; j < 4; j++) ph[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
              if (valid) *reinterpret_cast<int4*>(orow + cb0 + g * 8) = o;
            }
          }
          if (use_res && cb0 + 32 < a.n_tile) res_prefetch(a, n0, cb0 + 32, pix, valid, rv);  // next block
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
      if (a.done_ctr != nullptr && a.epi_mode == EPI_STORE) {
        // Publish this warp's rows per image.  A publish is fence (every lane's stores before the count) + warp barrier
        // + one atomic; the fence costs several hundred cycles on the epilogue's latency chain, so rows are counted in
        // registers and published only when the warp moves on to another image (tiles arrive in image order) or runs
        // out of tiles.  A flattened tile may span two images.
        const int my_img = (a.imgs == 1 && a.Ho == 1) ? fdiv(min(wo, a.Wo - 1), a.m_ohw) : img;
        const unsigned vm = __ballot_sync(0xffffffffu, valid);
        if (vm) {
          const int first = __shfl_sync(0xffffffffu, my_img, __ffs(vm) - 1);
          const unsigned same = __ballot_sync(0xffffffffu, valid && my_img == first);
          if (pend_img >= 0 && pend_img != first) {
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicAdd(a.done_ctr + pend_img, pend_cnt);
            pend_cnt = 0;
          }
          pend_img = first;
          pend_cnt += __popc(same);
          if (vm != same) {
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicAdd(a.done_ctr + first, pend_cnt);
            pend_img = first + 1;
            pend_cnt = __popc(vm ^ same);
          }
        }
      }
      if (dbg_on) a.dbg[dbg_t * 8 + 6] = clock64();
    }
    if (pend_img >= 0) {
      __threadfence();
      __syncwarp();
      if (lane == 0) atomicAdd(a.done_ctr + pend_img, pend_cnt);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
}

// One-time weight packing: w [Cout][tap][Cin] fp16 -> slab images [n tile][tap][chunk][b_stride bytes]; inside a
// slab row r (output channel) holds BK channels, its 16-byte piece c sits at the K-major swizzled position the
// UMMA descriptor expects: c ^ (r & 7) for 128-byte rows, c ^ ((r >> 1) & 3) for 64-byte, c ^ ((r >> 2) & 1) for 32-byte.
__global__ void pack_weights_kernel(const __half* __restrict__ w, uint8_t* __restrict__ out, int n_tile, int n_tiles,
                                    int taps, int chunks, int BK, int Cin, uint32_t b_stride, long long pieces) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pieces) return;
  const int c16 = BK / 8;
  const int c = (int)(i % c16);
  long long q = i / c16;
  const int r = (int)(q % n_tile); q /= n_tile;
  const int ch = (int)(q % chunks); q /= chunks;
  const int t = (int)(q % taps);
  const int nt = (int)(q / taps);
  const size_t K = (size_t)taps * Cin;
  if (ch * BK + c * 8 >= Cin) return;  // ragged last slab: stays zero
  const int4 v = *reinterpret_cast<const int4*>(w + (size_t)(nt * n_tile + r) * K + (size_t)t * Cin + ch * BK + c * 8);
  const int sw = BK == 64 ? (r & 7) : (BK == 32 ? ((r >> 1) & 3) : ((r >> 2) & 1));
  uint8_t* slab = out + (size_t)((nt * taps + t) * chunks + ch) * b_stride;
  *reinterpret_cast<int4*>(slab + (size_t)r * BK * 2 + ((c ^ sw) << 4)) = v;
}

// ------------------------------------------------------------------------------------------
// Host side: tiling choice + tensor maps
// ------------------------------------------------------------------------------------------
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
  if (p.Cin % 16 || p.Cout % 16 || p.Cout > TC_MAX_COUT) return false;
  if (!((p.k == 1 && p.stride == 1) || (p.k == 3 && (p.stride == 1 || p.stride == 2)))) return false;
  if (p.in.coff % 8 || p.in.pitch % 8 || p.out.coff % 8 || p.out.pitch % 8) return false;
  if (p.res.base && (p.res.coff % 8 || p.res.pitch % 8)) return false;
  return true;
}

// Experiment knob (tools/exp_dual.py): YB_PLAN_SMALL=1 sizes every plan for half an SM (<= 104 KiB of rings, <= 256
// TMEM columns, 352 threads) and launches one CTA per SM, so that kernels of TWO concurrent streams (two half-batch
// engines) are co-resident on every SM and each fills the other's pipeline bubbles.
static int plan_small() {
  const char* v = getenv("YB_PLAN_SMALL");
  return v ? atoi(v) : 0;
}

static int pick_n_tile(int cout) {
  const int cap = plan_small() ? 128 : 256;
  if (cout <= cap) return cout;
  int best = 16;
  for (int n = 16; n <= cap; n += 16)
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
  // stride-2 3x3 convs over a whole-buffer view with <= 32 channels use pair rows (below)
  const bool s2p_ok = p.k == 3 && p.stride == 2 && p.in.coff == 0 && p.in.pitch == p.Cin && p.Cin <= 32 && p.in.W % 2 == 0;
  a.mode = (p.k == 3 && p.stride == 1) ? TC_HALO : (s2p_ok ? TC_S2P : TC_TAP);
  // channel slab: the widest of 64 / 32 / 16 channels (128 / 64 / 32-byte operand rows) that pads K by at most
  // 35 %; a ragged last slab
  // is zero-filled by TMA (activations, dim 0 bound = Cin) and by the weight packing.  (80 / 160 / 400-channel
  // layers of v8x ran with 16 / 32-channel slabs = 32 / 64-byte rows and stayed at ~30 % of the tensor peak while
  // the 320 / 640-channel layers reached 45-75 %.)
  auto padded = [&](int bk) { return (p.Cin + bk - 1) / bk * bk; };
  a.BK = padded(64) * 100 <= p.Cin * 135 ? 64 : (padded(32) * 100 <= p.Cin * 135 ? 32 : 16);  // <= 35 % zero K
  a.chunks = (p.Cin + a.BK - 1) / a.BK;
  a.act = p.act;
  a.n_tile = pick_n_tile(p.Cout);
  a.n_tiles = p.Cout / a.n_tile;
  const CUtensorMapSwizzle swz = a.BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                            : (a.BK == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  a.layout_a = a.layout_b = a.BK == 64 ? 2 : (a.BK == 32 ? 4 : 6);
  a.row_bytes = a.BK * 2;
  a.sbo_a = a.sbo_b = (8 * a.row_bytes) >> 4;
  // stride-2 3x3, pair rows: two horizontally adjacent pixels (input columns 2q, 2q+1) are contiguous in a
  // whole-buffer NHWC view, so the tensor map declares them as ONE row of 2*Cin channels with the swizzle of
  // that width.  One dense box of (2BH+1) input rows x (BW+1) pairs then serves all 9 taps as row / K-slice
  // shifts (see mma_role); the left / top zero padding is TMA out-of-bounds fill of pair -1 / row -1.
  // (A strided box per tap moved 9 x 128 rows of Cin*2 bytes per tile and was bound by the TMA row rate:
  //  93 G sectors/s on model.1, profiles/r1_ncu_s2_tap_mode.txt.)
  CUtensorMapSwizzle swz_a = swz;
  if (a.mode == TC_S2P) {
    swz_a = a.BK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    a.layout_a = a.BK == 32 ? 2 : 4;
  }

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
    if (a.mode == TC_HALO) {
      // 8 x 16 output pixels; one TMA box carries the 10 x 18 input halo of a channel slab
      a.BW = HALO_BW; a.BH = HALO_BH;
      a.sbo_a = ((a.BW + 2) * a.row_bytes) >> 4;
      box[0] = a.BK; box[1] = a.BW + 2; box[2] = a.BH + 2; box[3] = 1;
      estr[0] = estr[1] = estr[2] = estr[3] = 1;
    } else if (a.mode == TC_S2P) {
      // 8 x 16 output pixels from 9 pairs x 33 input rows; consecutive output rows are two input rows =
      // 18 pair rows apart
      a.BW = HALO_BW; a.BH = HALO_BH;
      a.sbo_a = (2 * (a.BW + 1) * 2 * a.row_bytes) >> 4;
      gdim[0] = 2 * p.Cin; gdim[1] = p.in.W / 2;
      gstr[0] = (cuuint64_t)2 * p.in.pitch * esz;
      box[0] = 2 * a.BK; box[1] = a.BW + 1; box[2] = 2 * a.BH + 1; box[3] = 1;
      estr[0] = estr[1] = estr[2] = estr[3] = 1;
    } else {
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
  }
  CUresult cr = encode(&a.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, swz_a, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled(A) failed with code " + std::to_string((int)cr);
    delete plan;
    return nullptr;
  }
  const int a_rows = a.mode == TC_HALO ? (a.BW + 2) * (a.BH + 2)
                                       : (a.mode == TC_S2P ? 2 * (a.BW + 1) * (2 * a.BH + 1) : a.BW * a.BH);
  a.a_bytes = (uint32_t)(a_rows * a.row_bytes);
  a.b_bytes = (uint32_t)(a.n_tile * a.row_bytes);
  a.a_stride = (uint32_t)((std::max(a_rows, 128) * a.row_bytes + 1023) / 1024 * 1024);
  a.b_stride = (uint32_t)((a.n_tile * a.row_bytes + 1023) / 1024 * 1024);
  a.ksteps = p.k * p.k * a.chunks;
  {
    // Weights are constant: store every [n_tile x BK] slab in global memory exactly as it must look in shared
    // memory (swizzled rows, padded to b_stride), so the kernel fetches a slab with ONE contiguous bulk copy.
    // (As 2-D TMA boxes the slabs cost ~3 cycles per row in the TMA unit - 7200 rows per tile for a 3x3 160->160
    //  layer, more than its MMAs - and bounded the wide layers of v8s / v8x at ~30 % of the tensor peak.)
    const size_t total = (size_t)a.n_tiles * a.ksteps * a.b_stride;
    if (cudaMalloc(&plan->wpk, total) != cudaSuccess) {
      if (err) *err = "cudaMalloc(packed weights) failed";
      cudaGetLastError();
      delete plan;
      return nullptr;
    }
    cudaMemset(plan->wpk, 0, total);
    const int chunks16 = a.BK / 8;  // 16-byte pieces per row
    const long long pieces = (long long)a.n_tiles * a.ksteps * a.n_tile * chunks16;
    pack_weights_kernel<<<(unsigned)((pieces + 255) / 256), 256>>>(reinterpret_cast<const __half*>(p.w), plan->wpk, a.n_tile,
                                                                    a.n_tiles, p.k * p.k, a.chunks, a.BK, p.Cin,
                                                                    a.b_stride, pieces);
    if (cudaDeviceSynchronize() != cudaSuccess) {
      if (err) *err = std::string("pack_weights_kernel failed: ") + cudaGetErrorString(cudaGetLastError());
      cudaFree(plan->wpk);
      delete plan;
      return nullptr;
    }
    a.wpk = plan->wpk;
  }
  // Two CTAs per SM (each <= ~100 KiB smem, <= 256 TMEM columns) double the tiles in flight per SM and
  // hide the producer -> MMA -> epilogue hand-off latencies of the HBM-bound high-resolution layers;
  // layers whose resident weights or wide N tiles do not fit run one CTA per SM with the full budget.
  // 4 accumulator buffers when they fit in 256 TMEM columns (keeps two CTAs per SM possible), else 2
  a.n_acc = 4 * a.n_tile <= 256 ? 4 : 2;
  uint32_t cols = 32;
  while (cols < (uint32_t)(a.n_acc * a.n_tile)) cols <<= 1;
  a.tmem_cols = cols;
  const size_t b_all = (size_t)a.ksteps * a.b_stride;
  const int m_tiles = plan->flat ? (p.B * p.Ho * p.Wo + 127) / 128
                                 : p.B * ((p.Wo + a.BW - 1) / a.BW) * ((p.Ho + a.BH - 1) / a.BH);
  plan->occ = 1;
  for (int occ = 2; occ >= 1; occ--) {
    const size_t budget = occ == 2 ? 104 * 1024 : 200 * 1024;  // operand rings
    // layers with <= 2 tiles per CTA gain nothing from deep rings or resident weights; a small footprint
    // lets the NEXT kernel's CTAs (PDL / sibling branches) become resident while this one drains
    const bool small = m_tiles * a.n_tiles <= 2 * 148;
    if (occ == 2 && cols > 256) continue;
    // keep the whole weight matrix in smem when it leaves room for >= 3 activation slabs: removes the
    // weight re-fetch per tile
    a.b_resident = (a.n_tiles == 1 && b_all + 3 * (size_t)a.a_stride <= budget) ? 1 : 0;
    // resident weights at one CTA/SM beat re-fetched weights at two CTAs/SM
    if (occ == 2 && !plan_small() && !small && !a.b_resident && a.n_tiles == 1 && b_all + 3 * (size_t)a.a_stride <= 200 * 1024) continue;
    if (a.b_resident) {
      a.stages_a = (int)std::min<size_t>(a.mode != TC_TAP ? (a.chunks > 1 ? 8 : 6) : TC_MAX_STAGES, (budget - b_all) / a.a_stride);
      a.stages_b = 0;
      plan->smem = (size_t)a.stages_a * a.a_stride + b_all + 1024;
    } else if (a.mode != TC_TAP) {
      // A slab serves 9 B slabs: a short A ring and as many weight slabs as fit
      a.stages_a = (int)std::min<size_t>(3, std::max<size_t>(2, (budget / 3) / a.a_stride));
      const size_t rest = budget > (size_t)a.stages_a * a.a_stride ? budget - (size_t)a.stages_a * a.a_stride : 0;
      a.stages_b = (int)std::min<size_t>(TC_MAX_STAGES, rest / a.b_stride);
      plan->smem = (size_t)a.stages_a * a.a_stride + (size_t)a.stages_b * a.b_stride + 1024;
    } else {
      a.stages_a = a.stages_b = (int)std::min<size_t>(8, budget / (a.a_stride + a.b_stride));
      plan->smem = (size_t)a.stages_a * (a.a_stride + a.b_stride) + 1024;
    }
    const bool fits = a.stages_a >= 2 && (a.b_resident || a.stages_b >= ((occ == 2 && a.mode != TC_TAP) ? 3 : 2)) &&
                      (size_t)a.stages_a * a.a_stride + (a.b_resident ? b_all : (size_t)a.stages_b * a.b_stride) <= budget;
    if (fits && (occ == 1 || small || a.stages_a >= 3 || plan_small())) { plan->occ = occ; break; }
    if (occ == 1) a.stages_a = 0;  // reported below
  }
  // epilogue groups: one per accumulator buffer at one CTA/SM (608 threads); two when two CTAs share the SM
  a.n_groups = (plan->occ == 1 && a.n_acc == 4) ? 4 : 2;
  plan->threads = 96 + 128 * a.n_groups;
  // two issuers need >= 2 slabs per half ring
  a.n_issuers = (a.stages_a >= 4 && (a.b_resident || a.stages_b >= 4)) ? 2 : 1;
  {
    // Each issuer owns half of the ring and the single producer fills tiles in order: when a half ring cannot hold
    // one tile's activation slabs plus one of the next tile, the producer blocks inside a tile and the other
    // issuer starves (timeline of the 3-slab 80->80 layers of v8x: the two issuers alternated instead of
    // overlapping).  Such layers run one issuer on the whole ring.
    const int slabs_per_tile = a.mode != TC_TAP ? a.chunks : a.ksteps;
    if (a.n_issuers == 2 && a.stages_a / 2 < slabs_per_tile + 1) a.n_issuers = 1;
  }
  if (a.n_issuers == 2) {
    a.stages_a &= ~1;  // even split
    if (!a.b_resident) a.stages_b &= ~1;
  }
  if (!a.b_resident) {
    // streamed weights: tile pairs share the weight slabs (mma_role_dual); two activation slabs per step
    const size_t budget = plan->occ == 2 ? 104 * 1024 : 200 * 1024;
    int sa2, sb2;
    if (a.mode != TC_TAP) {
      sa2 = (size_t)4 * a.a_stride + 3 * (size_t)a.b_stride <= budget ? 4 : 2;
      sb2 = budget > (size_t)sa2 * a.a_stride ? (int)std::min<size_t>(TC_MAX_STAGES, (budget - (size_t)sa2 * a.a_stride) / a.b_stride) : 0;
    } else {
      const int S = (int)std::min<size_t>(6, budget / (2 * (size_t)a.a_stride + a.b_stride));
      sa2 = 2 * S;
      sb2 = S;
    }
    if (sa2 >= 2 && sb2 >= 2) {
      a.dual = 1;
      a.n_issuers = 1;
      a.stages_a = sa2;
      a.stages_b = sb2;
      plan->smem = (size_t)a.stages_a * a.a_stride + (size_t)a.stages_b * a.b_stride + 1024;
    }
  }
  if (a.stages_a < 2 || (!a.b_resident && a.stages_b < 2)) {
    if (err) *err = "tile does not fit in shared memory";
    delete plan;
    return nullptr;
  }
  a.epi_mode = p.dec.mode;
  a.dA = p.dec.A; a.dCtot = p.dec.Ctot; a.da0 = p.dec.a0; a.dch0 = p.dec.ch0; a.dWl = p.dec.Wl; a.dHW = p.dec.HW;
  a.dstride = p.dec.stride;
  if (a.epi_mode != EPI_STORE && !(plan->flat && a.n_tiles == 1 && (a.epi_mode != EPI_DFL_BOX || a.n_tile == 64))) {
    if (err) *err = "fused decode epilogue needs a flattened 1x1 conv with a single N tile";
    delete plan;
    return nullptr;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    // dynamic limit = ring budget + alignment slack (static smem of the kernel counts against the 227 KiB cap)
    cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaError_t ce = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
    if (ce != cudaSuccess) {
      if (err) *err = std::string("cudaFuncSetAttribute(conv_tc_kernel) failed: ") + cudaGetErrorString(ce);
      num_sms = 0;
      delete plan;
      return nullptr;
    }
  }
  plan->grid = num_sms * plan->occ;
  if (plan_small() && plan->occ == 2) plan->grid = num_sms;  // the SM's other half belongs to the other stream's kernel
  plan->small = m_tiles * a.n_tiles <= 2 * 148;
  return plan;
}

std::string tc_conv_plan_describe(const TcConvPlan* plan) {
  const TcArgs& a = plan->args;
  char buf[256];
  snprintf(buf, sizeof(buf), "mode %d BK %d chunks %d n_tile %d x%d stages a/b %d/%d resident %d dual %d issuers %d groups %d acc %d occ %d "
           "threads %d smem %zu KiB grid %d", a.mode, a.BK, a.chunks, a.n_tile, a.n_tiles, a.stages_a, a.stages_b, a.b_resident, a.dual,
           a.n_issuers, a.n_groups, a.n_acc, plan->occ, plan->threads, plan->smem / 1024, plan->grid);
  return buf;
}

void tc_conv_plan_destroy(TcConvPlan* plan) {
  if (plan && plan->wpk) cudaFree(plan->wpk);
  delete plan;
}

long long* g_tc_dbg = nullptr;  // set by yb_debug_timeline: next tcgen05 conv launches write their timeline here
int g_tc_dbg_countdown = -1;

int tc_conv_rows_per_image(const TcConvPlan* plan) { return plan->p.Ho * plan->p.Wo * plan->args.n_tiles; }

int tc_conv_launch(const TcConvPlan* plan, int B, float* pred, int* tile_ctr, cudaStream_t s, const TcChain* chain) {
  TcArgs a = plan->args;
  a.pred = pred;
  if (chain) {
    a.done_ctr = chain->done_ctr;
    if (!plan->args.dual) {  // tile pairs (streamed weights, static order) keep the grid-wide dependency
      a.dep_ctr = chain->dep_ctr;
      a.dep_expect = chain->dep_expect;
    }
  }
  // Tile pairs keep the static order (both tiles need the same N tile).  (Restricting the queue to long-tile layers
  // made the isolated short-tile layers of v8n ~10 % faster but the overlapped step no faster, and cost v8s 4 %.)
  a.tile_ctr = a.dual ? nullptr : tile_ctr;
  a.dbg = nullptr;
  if (g_tc_dbg && g_tc_dbg_countdown >= 0 && g_tc_dbg_countdown-- == 0) a.dbg = g_tc_dbg;
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
  auto magic = [](int d) { return (uint64_t)(((((unsigned __int128)1) << 40) + d - 1) / (unsigned)d); };
  a.m_ntiles = magic(a.n_tiles); a.m_tpi = magic(a.tiles_w * a.tiles_h); a.m_tw = magic(a.tiles_w); a.m_bw = magic(a.BW);
  a.m_ohw = magic(p.Ho * p.Wo);
  int grid = std::min(plan->grid, a.total_tiles);
  // chained layers: one CTA per SM for the two-CTA plans, so that the other slot of every SM is free for the NEXT
  // layer's CTA - consecutive layers then run side by side, the later one on the images the earlier one has finished
  static const int chain_grid1 = getenv("YB_CHAIN_GRID1") ? atoi(getenv("YB_CHAIN_GRID1")) : 1;
  if (chain && chain_grid1 && plan->occ == 2 && !a.dual) grid = std::min(grid, std::max(1, plan->grid / 2));
  // concurrent head branches: a latency-bound layer with ~1 tile per CTA gives up half of its CTAs (each
  // then pipelines 2-3 tiles) so that a sibling branch can occupy the other SMs at the same time
  if (plan->p.share_sms && a.total_tiles <= 4 * plan->grid && a.ksteps * (a.BK >> 4) <= 40)
    grid = std::max(1, std::min(grid, (a.total_tiles + 2) / 3));
  if (a.dual && grid >= a.n_tiles) grid -= grid % a.n_tiles;  // tile t and t + grid must share their N tile
  // one atomic per ~quarter of a CTA's share (every atomic of the grid hits the same L2 address: per-tile draws
  // cost the 6400-tile layers 6-8 us)
  a.tile_batch = std::max(1, std::min(8, a.total_tiles / (4 * grid)));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(plan->threads);
  cfg.dynamicSmemBytes = plan->smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // PDL (see griddepcontrol in the kernel)
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static const bool no_pdl = getenv("YB_DEBUG_NO_PDL") != nullptr;  // experiments only (tools/exp_fixed_cost.py)
  cfg.numAttrs = no_pdl ? 0 : 1;
  YB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_tc_kernel, a));
  return 0;
}

// ------------------------------------------------------------------------------------------
// Stem: model.0 = Conv(3, C, k3, s2) straight from the NCHW network input (u8 / f16 / f32) on the tensor
// cores.  K ordering is chosen so that im2col needs no element shuffling: for output pixel (ho, wo) and each
// of the 9 (kh, c) input rows, the FOUR contiguous input columns 2wo-2 .. 2wo+1 form one 8-byte K group
//     k = (kh*3 + c)*4 + slot,   slot 0 -> column 2wo-2 (weight 0), slots 1..3 -> kw = 0..2
// so a thread builds its A row from 9 aligned 8-byte loads (K = 36, padded to 48 = 3 MMAs) instead of 27
// scalar loads + repacking (the first version spent 650 instructions per warp and tile on that and was
// issue-bound at 147 us; profiles/r1_ncu_stem_*.txt).
//   A tile   128 rows x 128 B, SWIZZLE_128B layout written by hand (16-byte piece p of row r at p ^ (r & 7))
//   weights  [Cout][64] fp16 resident in smem, same layout
//   MMA      3 x tcgen05.mma (M=128, N=Cout, K=16) by warp 4
//   epilogue same 128 threads: tcgen05.ld -> +bias -> SiLU -> fp16 NHWC store (Cout*2 contiguous bytes)
// Four CTAs per SM overlap each other's load / MMA / store latencies; the layer is HBM-bound
// (reads the image once, writes Cout x H/2 x W/2 fp16).
// (A TMA box load of the NCHW patch faulted with "illegal instruction" on B200 for rank-3 f32 maps;
//  plain loads are used instead.)
// ------------------------------------------------------------------------------------------
constexpr int ST_TW = 16, ST_TH = 8;
constexpr int ST_THREADS = 160;  // warps 0-3: im2col + epilogue, warp 4: MMA issue

struct StemArgs {
  const void* in;
  const __half* w16;   // [Cout][64] fp16, k = (kh*3 + c)*4 + kw + 1, other k zero
  const float* bias;
  __half* out;
  int out_pitch, out_coff;
  int B, H, W, Ho, Wo, Cout;
  // Detector.ImagePredict pads the image right / bottom to a multiple of 32 with the value 114 before the /255
  // (Models/Detector.cs:35-41): the caller's tensor is (B, 3, src_H, src_W) with src <= H, W and the padding is
  // produced here instead of by a separate pad kernel.  pad_h2 = the padded pixel as two fp16 (already scaled).
  int src_H, src_W;
  uint32_t pad_h2;
  int tiles_w, tiles_h, total_tiles;
  uint32_t tmem_cols;
  uint64_t m_tpi, m_tw;  // magic numbers for / tiles_per_img and / tiles_w (fdiv)
};

__device__ __forceinline__ uint32_t pack_h2(float x, float y) {
  const __half2 h = __floats2half2_rn(x, y);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// four input columns col .. col+3 (col even, may be -2) of one row as 4 halves; `lo_ok` = col >= 0
template <int DT>
__device__ __forceinline__ uint2 stem_load4(const void* in, size_t rowbase, int col, bool lo_ok) {
  uint2 r = make_uint2(0u, 0u);
  if (DT == YB_F16) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(reinterpret_cast<const __half*>(in) + rowbase + col);
    if (lo_ok) r.x = __ldg(p);
    r.y = __ldg(p + 1);
  } else if (DT == YB_U8) {
    // two bytes -> half2 without integer->float conversions: 0x6400 | b is the fp16 number 1024 + b, and
    // fma(1024 + b, k, -1024 k) = b * k rounded once (k = fp16(1/255); Detector.cs:41 divides by 255)
    const uint16_t* p = reinterpret_cast<const uint16_t*>(reinterpret_cast<const uint8_t*>(in) + rowbase + col);
    const __half2 k2 = __float2half2_rn(1.0f / 255.0f);
    const __half2 c2 = __hmul2(k2, __float2half2_rn(-1024.0f));
    auto cvt = [&](uint32_t v) {
      const uint32_t x = __byte_perm(v, 0x64006400u, 0x7150);
      const __half2 h = __hfma2(*reinterpret_cast<const __half2*>(&x), k2, c2);
      return *reinterpret_cast<const uint32_t*>(&h);
    };
    if (lo_ok) r.x = cvt(__ldg(p));
    r.y = cvt(__ldg(p + 1));
  } else {
    const float2* p = reinterpret_cast<const float2*>(reinterpret_cast<const float*>(in) + rowbase + col);
    if (lo_ok) { const float2 v = __ldg(p); r.x = pack_h2(v.x, v.y); }
    const float2 v = __ldg(p + 1);
    r.y = pack_h2(v.x, v.y);
  }
  return r;
}

// one input element as fp16 bits (same arithmetic as stem_load4)
template <int DT>
__device__ __forceinline__ uint32_t stem_load1(const void* in, size_t i) {
  if (DT == YB_F16) return (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(in) + i);
  if (DT == YB_U8) {
    const __half k = __float2half_rn(1.0f / 255.0f);
    const uint32_t x = 0x6400u | (uint32_t)__ldg(reinterpret_cast<const uint8_t*>(in) + i);
    const unsigned short xs = (unsigned short)x;
    const __half h = __hfma(*reinterpret_cast<const __half*>(&xs), k, __hmul(k, __float2half_rn(-1024.0f)));
    return (uint32_t)*reinterpret_cast<const unsigned short*>(&h);
  }
  const __half h = __float2half_rn(__ldg(reinterpret_cast<const float*>(in) + i));
  return (uint32_t)*reinterpret_cast<const unsigned short*>(&h);
}

// four columns col .. col+3 of a SOURCE row that may end before them (ragged width / odd row alignment): columns
// < 0 are the conv's zero padding, columns >= src_W the 114-padding of the image
template <int DT>
__device__ __forceinline__ uint2 stem_load4_ragged(const void* in, size_t rowbase, int col, int src_W, uint32_t pad1) {
  uint32_t e[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int c = col + j;
    e[j] = c < 0 ? 0u : (c >= src_W ? pad1 : stem_load1<DT>(in, rowbase + c));
  }
  return make_uint2(e[0] | (e[1] << 16), e[2] | (e[3] << 16));
}

template <int DT>
__global__ void __launch_bounds__(ST_THREADS, 4) stem_tc_kernel(const __grid_constant__ StemArgs a) {
  extern __shared__ __align__(1024) uint8_t st_smem[];
  __shared__ __align__(8) uint64_t bars[2];  // a_ready, mma_done
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bias[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const uint32_t base = (smem_u32(st_smem) + 1023u) & ~1023u;
  const uint32_t smA = base;              // 128 rows x 128 B
  const uint32_t smB = base + 16 * 1024;  // Cout rows x 128 B
  const uint32_t a_ready = smem_u32(&bars[0]), mma_done = smem_u32(&bars[1]);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (tid == 0) {
    mbar_init(a_ready, 128);
    mbar_init(mma_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(a.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // weights -> smem with the SWIZZLE_128B pattern; the K padding of the A rows is zeroed once
  for (int i = tid; i < a.Cout * 8; i += ST_THREADS) {
    const int n = i >> 3, pc = i & 7;
    const int4 v = *reinterpret_cast<const int4*>(a.w16 + n * 64 + pc * 8);
    st_shared_v4(smB + n * 128 + ((pc ^ (n & 7)) << 4), v);
  }
  for (int i = tid; i < 128 * 8; i += ST_THREADS) st_shared_v4(smA + i * 16, make_int4(0, 0, 0, 0));
  for (int i = tid; i < a.Cout; i += ST_THREADS) s_bias[i] = a.bias[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const int tiles_per_img = a.tiles_w * a.tiles_h;
  if (warp == 4) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(a.Cout >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t hi = 64u | (1u << 14) | (2u << 29);  // SBO = 8 rows x 128 B, version 1, SWIZZLE_128B
      const uint32_t a_lo = ((smA & 0x3FFFF) >> 4) | (1u << 16), b_lo = ((smB & 0x3FFFF) >> 4) | (1u << 16);
      int it = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, it++) {
        mbar_wait(a_ready, it & 1);
        tc_fence_after();
        umma_f16(tmem, desc64(a_lo, hi), desc64(b_lo, hi), idesc, 0);
        umma_f16(tmem, desc64(a_lo + 2, hi), desc64(b_lo + 2, hi), idesc, 1);
        umma_f16(tmem, desc64(a_lo + 4, hi), desc64(b_lo + 4, hi), idesc, 1);
        umma_commit(mma_done);
      }
    }
  } else {
    const int tx = tid & (ST_TW - 1), ty = tid / ST_TW;  // output pixel inside the tile
    const size_t plane = (size_t)a.src_H * a.src_W;
    const bool ragged = a.src_W != a.W || a.src_H != a.H;  // padded source: rows may be misaligned and end early
    const uint32_t pad1 = a.pad_h2 & 0xffffu;
    auto gather = [&](int tile, uint2 (&v)[9]) {
      const int n = fdiv(tile, a.m_tpi), r = tile - n * tiles_per_img;
      const int th = fdiv(r, a.m_tw);
      const int ho = th * ST_TH + ty, wo = (r - th * a.tiles_w) * ST_TW + tx;
      const bool pix_ok = ho < a.Ho && wo < a.Wo;
      const int col = 2 * wo - 2;
#pragma unroll
      for (int kh = 0; kh < 3; kh++) {
        const int hi_ = ho * 2 + kh - 1;
        const bool row_ok = pix_ok && hi_ >= 0 && hi_ < a.H;
        const bool in_src = hi_ < a.src_H;
        const size_t rowbase = (size_t)n * 3 * plane + (size_t)((row_ok && in_src) ? hi_ : 0) * a.src_W;
#pragma unroll
        for (int c = 0; c < 3; c++) {
          uint2 g = make_uint2(0u, 0u);
          if (row_ok) {
            if (!ragged) g = stem_load4<DT>(a.in, rowbase + c * plane, col, wo > 0);
            else if (in_src) g = stem_load4_ragged<DT>(a.in, rowbase + c * plane, col, a.src_W, pad1);
            else g = make_uint2(wo > 0 ? a.pad_h2 : 0u, a.pad_h2);  // a row of the bottom padding
          }
          v[kh * 3 + c] = g;
        }
      }
    };
    uint2 v[9];
    if (blockIdx.x < a.total_tiles) gather(blockIdx.x, v);
    const uint32_t a_row = smA + tid * 128;
    const uint32_t sw = (uint32_t)(tid & 7);
    int it = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, it++) {
      const int n = fdiv(tile, a.m_tpi), r = tile - n * tiles_per_img;
      const int th = fdiv(r, a.m_tw);
      const int ho = th * ST_TH + ty, wo = (r - th * a.tiles_w) * ST_TW + tx;
#pragma unroll
      for (int pc = 0; pc < 4; pc++)
        st_shared_v4(a_row + ((pc ^ sw) << 4), make_int4((int)v[2 * pc].x, (int)v[2 * pc].y, (int)v[2 * pc + 1].x, (int)v[2 * pc + 1].y));
      st_shared_v4(a_row + ((4u ^ sw) << 4), make_int4((int)v[8].x, (int)v[8].y, 0, 0));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
      mbar_arrive(a_ready);
      if (tile + (int)gridDim.x < a.total_tiles) gather(tile + gridDim.x, v);  // next tile's loads fly during the MMA
      mbar_wait(mma_done, it & 1);
      tc_fence_after();
      const bool valid = ho < a.Ho && wo < a.Wo;
      __half* o = a.out + ((size_t)(n * a.Ho + ho) * a.Wo + wo) * a.out_pitch + a.out_coff;
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
      for (int c0 = 0; c0 < a.Cout; c0 += 16) {
        uint32_t acc[16];
        tmem_ld16(taddr + c0, acc);
        tmem_ld_wait();
        if (valid) {
          int4 o0, o1;
          __half2* p0 = reinterpret_cast<__half2*>(&o0);
          __half2* p1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
          for (int j = 0; j < 4; j++) {
            p0[j] = __floats2half2_rn(silu_tanh(__uint_as_float(acc[2 * j]) + s_bias[c0 + 2 * j]),
                                      silu_tanh(__uint_as_float(acc[2 * j + 1]) + s_bias[c0 + 2 * j + 1]));
            p1[j] = __floats2half2_rn(silu_tanh(__uint_as_float(acc[8 + 2 * j]) + s_bias[c0 + 8 + 2 * j]),
                                      silu_tanh(__uint_as_float(acc[8 + 2 * j + 1]) + s_bias[c0 + 8 + 2 * j + 1]));
          }
          *reinterpret_cast<int4*>(o + c0) = o0;
          *reinterpret_cast<int4*>(o + c0 + 8) = o1;
        }
      }
      tc_fence_before();  // TMEM reads done before the next tile's MMA overwrites the accumulator
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(a.tmem_cols) : "memory");
  }
}

int launch_stem_f16(const void* in, int in_dtype, int B, int H, int W, const __half* w16, const float* bias,
                    const View& out, cudaStream_t s, int src_H, int src_W) {
  if (out.C % 16 || out.C > 256 || out.coff % 8 || out.pitch % 8 || (W & 1) || (H & 1)) {
    set_error("stem: output channels must be a multiple of 16 and <= 256, input height/width even");
    return YB_ERR_SHAPE;
  }
  StemArgs a;
  memset(&a, 0, sizeof(a));
  a.in = in; a.w16 = w16; a.bias = bias;
  a.out = reinterpret_cast<__half*>(out.base);
  a.out_pitch = out.pitch; a.out_coff = out.coff;
  a.B = B; a.H = H; a.W = W; a.Ho = H / 2; a.Wo = W / 2; a.Cout = out.C;
  a.src_H = src_H > 0 ? src_H : H; a.src_W = src_W > 0 ? src_W : W;
  if (a.src_H > H || a.src_W > W) { set_error("stem: source image larger than the planned input"); return YB_ERR_SHAPE; }
  {
    // the padded pixel: u8 input -> half(114 * half(1/255)) exactly as a loaded 114 would come out; float inputs are
    // already scaled by the caller (pad(x, 114) / 255, Detector.cs:41) -> half(114 / 255)
    const __half k = __float2half_rn(1.0f / 255.0f);
    const __half pv = in_dtype == YB_U8 ? __float2half_rn(114.0f * __half2float(k)) : __float2half_rn(114.0f / 255.0f);
    const unsigned short bits = *reinterpret_cast<const unsigned short*>(&pv);
    a.pad_h2 = (uint32_t)bits | ((uint32_t)bits << 16);
  }
  a.tiles_w = (a.Wo + ST_TW - 1) / ST_TW;
  a.tiles_h = (a.Ho + ST_TH - 1) / ST_TH;
  a.total_tiles = B * a.tiles_w * a.tiles_h;
  auto magic = [](int d) { return (uint64_t)(((((unsigned __int128)1) << 40) + d - 1) / (unsigned)d); };
  a.m_tpi = magic(a.tiles_w * a.tiles_h); a.m_tw = magic(a.tiles_w);
  uint32_t cols = 32;
  while (cols < (uint32_t)a.Cout) cols <<= 1;
  a.tmem_cols = cols;
  static int num_sms = 0;
  const size_t smem = 1024 + 16 * 1024 + (size_t)a.Cout * 128;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    YB_CUDA_CHECK(cudaFuncSetAttribute(stem_tc_kernel<YB_U8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    YB_CUDA_CHECK(cudaFuncSetAttribute(stem_tc_kernel<YB_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    YB_CUDA_CHECK(cudaFuncSetAttribute(stem_tc_kernel<YB_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  }
  const int per_sm = std::max(1, std::min(4, (int)(512 / cols)));
  const int grid = std::min(a.total_tiles, num_sms * per_sm);
  if (in_dtype == YB_U8) stem_tc_kernel<YB_U8><<<grid, ST_THREADS, smem, s>>>(a);
  else if (in_dtype == YB_F16) stem_tc_kernel<YB_F16><<<grid, ST_THREADS, smem, s>>>(a);
  else stem_tc_kernel<YB_F32><<<grid, ST_THREADS, smem, s>>>(a);
  YB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace yb
