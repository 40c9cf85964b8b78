// Train-mode BatchNorm2d + SiLU of the reference's Conv block (Modules/Convs.cs:36-56 with `yolo.train()`,
// YoloBaseTaskModel.cs:299,325): forward with batch statistics + running-statistics update, and backward.
//
//   forward   z (M = B*H*W rows, C channels, NHWC fp32) -> mean_c, var_c over the M rows (biased, two-pass),
//             y = SiLU(gamma * (z - mean) * invstd + beta), invstd = 1/sqrt(var + eps);
//             running_mean = (1-m) running_mean + m mean, running_var = (1-m) running_var + m var * M/(M-1)
//             (PyTorch BatchNorm2d semantics; the reference builds BatchNorm2d(eps 1e-3, momentum 0.03))
//   backward  u = gamma * xhat + beta, g = dy * SiLU'(u);  dbeta = sum g, dgamma = sum g*xhat,
//             dz = gamma * invstd * (g - dbeta/M - xhat * dgamma/M)
//
// All three passes are column reductions over a tall NHWC matrix: every block owns 32 channels x a slab of rows
// (32 x 8 threads, 128-byte coalesced rows), writes one partial per (slab, channel), and a second tiny kernel folds
// the partials in a fixed order - deterministic, no atomics.  HBM-bound: forward reads z twice (+1 in the
// elementwise pass) and writes y once; backward reads z, dy twice and writes dz.
#include <algorithm>

#include "common.cuh"

namespace yb {

namespace {

constexpr int BN_TX = 32, BN_TY = 8, BN_ROWS_PER_BLOCK = 256, BN_MAX_SLABS = 512;
// rows per slab: 256, or more for tall matrices so that the serial fold of the per-slab partials (bn_finish_kernel, one
// thread per channel) stays <= 512 steps (the 1.6 M-row first layers had 6 400 slabs: 16 us per fold, 243 folds a step)
static int bn_rows_per_block(long long M) {
  const long long r = (M + BN_MAX_SLABS - 1) / BN_MAX_SLABS;
  return (int)std::max<long long>(BN_ROWS_PER_BLOCK, (r + BN_TY - 1) / BN_TY * BN_TY);
}

__device__ __forceinline__ float silu_f(float u) { return u / (1.f + expf(-u)); }
__device__ __forceinline__ float silu_grad(float u) {
  const float s = 1.f / (1.f + expf(-u));
  return s * (1.f + u * (1.f - s));
}

// mode 0: sum z                      -> p0
// mode 1: sum (z - mean)^2           -> p0
// mode 2: sum g, sum g * xhat        -> p0, p1   (g = dy * SiLU'(gamma*xhat+beta))
// mode 3: sum (z - K), sum (z - K)^2 -> p0, p1   one pass over z for mean AND variance; K = z[row 0][c] is a shift close to
//         the mean (a sample of the channel), so var = (S2 - S1^2 / M) / M cancels at most a few bits
//         (relative error ~ eps * (1 + (K - mean)^2 / var)); the two-pass modes 0 / 1 read every activation twice
__global__ void __launch_bounds__(BN_TX* BN_TY) bn_partial_kernel(int mode, int act, const float* __restrict__ z, const float* __restrict__ dy,
                                                                long long M, int C, int pitch, int dpitch,
                                                                const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                float* __restrict__ p0, float* __restrict__ p1, int rpb) {
  const int c = blockIdx.x * BN_TX + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * rpb;
  float a0 = 0.f, a1 = 0.f;
  if (c < C) {
    const float mu = mode == 3 ? z[c] : (mode ? mean[c] : 0.f);
    const float is = mode == 2 ? invstd[c] : 0.f, ga = mode == 2 ? gamma[c] : 0.f, be = mode == 2 ? beta[c] : 0.f;
    for (long long r = r0 + threadIdx.y; r < min(M, r0 + rpb); r += BN_TY) {
      const float v = z[r * pitch + c];
      if (mode == 0) {
        a0 += v;
      } else if (mode == 1) {
        const float d = v - mu;
        a0 += d * d;
      } else if (mode == 3) {
        const float d = v - mu;
        a0 += d;
        a1 = fmaf(d, d, a1);
      } else {
        const float xh = (v - mu) * is;
        const float g = dy[r * dpitch + c] * (act ? silu_grad(ga * xh + be) : 1.f);
        a0 += g;
        a1 += g * xh;
      }
    }
  }
  __shared__ float s0[BN_TY][BN_TX], s1[BN_TY][BN_TX];
  s0[threadIdx.y][threadIdx.x] = a0;
  s1[threadIdx.y][threadIdx.x] = a1;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t0 = 0.f, t1 = 0.f;
    for (int k = 0; k < BN_TY; k++) { t0 += s0[k][threadIdx.x]; t1 += s1[k][threadIdx.x]; }
    p0[(size_t)blockIdx.y * C + c] = t0;
    if (mode >= 2) p1[(size_t)blockIdx.y * C + c] = t1;
  }
}

// (A 4-channels-per-thread variant with 16-byte loads was no faster: 3.5 vs 3.1 ms over the 162 launches of a YOLOv11s step -
// the kernel is bound by rows in flight, not by load width.)
// fold the per-slab partials in a FIXED order; step 0 -> mean, step 1 -> var / invstd / running stats, step 2 -> dgamma,
// dbeta, step 3 -> mean / var from shifted sums.  Block = 32 channels x 8 lanes: lane y sums slabs y, y+8, ... in order,
// the 8 partial sums are then added in lane order (one thread per channel walking 512 slabs serially took 12 us per fold,
// 243 folds a step).
__global__ void __launch_bounds__(256) bn_finish_kernel(int step, const float* __restrict__ p0, const float* __restrict__ p1, int slabs, int C,
                                 long long M, float eps, float momentum, float* __restrict__ mean, float* __restrict__ invstd,
                                 float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ dgamma,
                                 float* __restrict__ dbeta) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  __shared__ float f0[8][32], f1[8][32];
  float a0 = 0.f, a1 = 0.f;
  if (c < C)
    for (int s = threadIdx.y; s < slabs; s += 8) {
      a0 += p0[(size_t)s * C + c];
      if (step >= 2) a1 += p1[(size_t)s * C + c];
    }
  f0[threadIdx.y][threadIdx.x] = a0;
  f1[threadIdx.y][threadIdx.x] = a1;
  __syncthreads();
  if (threadIdx.y != 0 || c >= C) return;
  float t0 = 0.f, t1 = 0.f;
  for (int k = 0; k < 8; k++) { t0 += f0[k][threadIdx.x]; t1 += f1[k][threadIdx.x]; }
  if (step == 3) {  // shifted sums (mode 3): running_mean carries the shift pointer's value z[0][c] via `dgamma`
    const float K = dgamma[c];
    const float d = t0 / (float)M;
    const float mu = K + d;
    const float var = fmaxf(t1 / (float)M - d * d, 0.f);  // biased: used for normalisation
    mean[c] = mu;
    invstd[c] = 1.f / sqrtf(var + eps);
    if (running_mean) {
      const float unbiased = M > 1 ? var * ((float)M / (float)(M - 1)) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
  } else if (step == 0) {
    mean[c] = t0 / (float)M;
  } else if (step == 1) {
    const float var = t0 / (float)M;  // biased: used for normalisation
    invstd[c] = 1.f / sqrtf(var + eps);
    if (running_mean) {
      const float unbiased = M > 1 ? t0 / (float)(M - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean[c];
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
  } else {
    dbeta[c] = t0;
    dgamma[c] = t1;
  }
}

__global__ void bn_silu_apply_kernel(const float* __restrict__ z, long long M, int C, int pitch, int opitch,
                                     const float* __restrict__ mean, const float* __restrict__ invstd,
                                     const float* __restrict__ gamma, const float* __restrict__ beta, int act, float* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * C) return;
  const long long r = i / C;
  const int c = (int)(i - r * C);
  const float u = gamma[c] * (z[r * pitch + c] - mean[c]) * invstd[c] + beta[c];
  y[r * opitch + c] = act ? silu_f(u) : u;
}

__global__ void bn_silu_dz_kernel(const float* __restrict__ z, const float* __restrict__ dy, long long M, int C, int pitch, int dpitch,
                                  int opitch, const float* __restrict__ mean, const float* __restrict__ invstd,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                                  const float* __restrict__ dgamma, const float* __restrict__ dbeta, float* __restrict__ dz) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * C) return;
  const long long r = i / C;
  const int c = (int)(i - r * C);
  const float xh = (z[r * pitch + c] - mean[c]) * invstd[c];
  const float g = dy[r * dpitch + c] * (act ? silu_grad(gamma[c] * xh + beta[c]) : 1.f);
  const float inv_m = 1.f / (float)M;
  dz[r * opitch + c] = gamma[c] * invstd[c] * (g - dbeta[c] * inv_m - xh * dgamma[c] * inv_m);
}

// 4 channels per thread (C, pitches multiples of 4; 32-bit index math): the scalar kernels above spend a 64-bit division
// per element
__global__ void bn_silu_apply4_kernel(const float* __restrict__ z, int total4, int C4, int pitch, int opitch,
                                      const float* __restrict__ mean, const float* __restrict__ invstd,
                                      const float* __restrict__ gamma, const float* __restrict__ beta, int act, float* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int r = i / C4, c = (i - r * C4) * 4;
  const float4 v = *reinterpret_cast<const float4*>(z + (size_t)r * pitch + c);
  const float4 mu = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
  const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
  float4 u;
  u.x = ga.x * (v.x - mu.x) * is.x + be.x; u.y = ga.y * (v.y - mu.y) * is.y + be.y;
  u.z = ga.z * (v.z - mu.z) * is.z + be.z; u.w = ga.w * (v.w - mu.w) * is.w + be.w;
  if (act) { u.x = silu_f(u.x); u.y = silu_f(u.y); u.z = silu_f(u.z); u.w = silu_f(u.w); }
  *reinterpret_cast<float4*>(y + (size_t)r * opitch + c) = u;
}

__global__ void bn_silu_dz4_kernel(const float* __restrict__ z, const float* __restrict__ dy, int total4, int C4, int pitch, int dpitch,
                                   int opitch, float inv_m, const float* __restrict__ mean, const float* __restrict__ invstd,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                                   const float* __restrict__ dgamma, const float* __restrict__ dbeta, float* __restrict__ dz) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int r = i / C4, c = (i - r * C4) * 4;
  const float4 v = *reinterpret_cast<const float4*>(z + (size_t)r * pitch + c);
  const float4 d = *reinterpret_cast<const float4*>(dy + (size_t)r * dpitch + c);
  const float vv[4] = {v.x, v.y, v.z, v.w}, dd[4] = {d.x, d.y, d.z, d.w};
  float o[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const float is = invstd[c + j], ga = gamma[c + j];
    const float xh = (vv[j] - mean[c + j]) * is;
    const float g = dd[j] * (act ? silu_grad(ga * xh + beta[c + j]) : 1.f);
    o[j] = ga * is * (g - dbeta[c + j] * inv_m - xh * dgamma[c + j] * inv_m);
  }
  *reinterpret_cast<float4*>(dz + (size_t)r * opitch + c) = make_float4(o[0], o[1], o[2], o[3]);
}

}  // namespace

int bn_silu_train_forward(const float* z, long long M, int C, int pitch, const float* gamma, const float* beta, float eps,
                          float momentum, int act, float* running_mean, float* running_var, float* y, int ypitch,
                          float* save_mean, float* save_invstd, cudaStream_t s) {
  if (M <= 0 || C <= 0 || pitch < C || ypitch < C) {
    set_error("yb_bn_silu_train_forward: bad shape");
    return YB_ERR_SHAPE;
  }
  const int rpb = bn_rows_per_block(M);
  const int slabs = (int)((M + rpb - 1) / rpb);
  float* part = nullptr;
  YB_CUDA_CHECK(cudaMallocAsync((void**)&part, (size_t)2 * slabs * C * sizeof(float), s));
  const dim3 grid((C + BN_TX - 1) / BN_TX, slabs), block(BN_TX, BN_TY);
  const dim3 fb(32, 8);
  const int fg = (C + 31) / 32;
  // one pass over z: per-slab sums of (z - K) and (z - K)^2 about the shift K = z[row 0], folded in slab order
  bn_partial_kernel<<<grid, block, 0, s>>>(3, act, z, nullptr, M, C, pitch, 0, nullptr, nullptr, nullptr, nullptr, part,
                                          part + (size_t)slabs * C, rpb);
  bn_finish_kernel<<<fg, fb, 0, s>>>(3, part, part + (size_t)slabs * C, slabs, C, M, eps, momentum, save_mean, save_invstd, running_mean,
                                     running_var, const_cast<float*>(z) /* the shift row, read only */, nullptr);
  const long long total = M * C;
  if (C % 4 == 0 && pitch % 4 == 0 && ypitch % 4 == 0 && total / 4 < (1ll << 31) && ((uintptr_t)z % 16 == 0) && ((uintptr_t)y % 16 == 0)) {
    const int total4 = (int)(total / 4);
    bn_silu_apply4_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, s>>>(z, total4, C / 4, pitch, ypitch, save_mean, save_invstd, gamma,
                                                                          beta, act, y);
  } else {
    bn_silu_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(z, M, C, pitch, ypitch, save_mean, save_invstd, gamma, beta, act, y);
  }
  YB_CUDA_CHECK(cudaGetLastError());
  YB_CUDA_CHECK(cudaFreeAsync(part, s));
  return YB_OK;
}

int bn_silu_backward(const float* z, const float* dy, long long M, int C, int pitch, int dpitch, const float* gamma,
                     const float* beta, const float* save_mean, const float* save_invstd, int act, float* dz, int zpitch,
                     float* dgamma, float* dbeta, cudaStream_t s) {
  if (M <= 0 || C <= 0 || pitch < C || dpitch < C || zpitch < C) {
    set_error("yb_bn_silu_backward: bad shape");
    return YB_ERR_SHAPE;
  }
  const int rpb = bn_rows_per_block(M);
  const int slabs = (int)((M + rpb - 1) / rpb);
  float* part = nullptr;
  YB_CUDA_CHECK(cudaMallocAsync((void**)&part, (size_t)2 * slabs * C * sizeof(float), s));
  const dim3 grid((C + BN_TX - 1) / BN_TX, slabs), block(BN_TX, BN_TY);
  bn_partial_kernel<<<grid, block, 0, s>>>(2, act, z, dy, M, C, pitch, dpitch, save_mean, save_invstd, gamma, beta, part,
                                          part + (size_t)slabs * C, rpb);
  bn_finish_kernel<<<(C + 31) / 32, dim3(32, 8), 0, s>>>(2, part, part + (size_t)slabs * C, slabs, C, M, 0.f, 0.f, nullptr, nullptr, nullptr,
                                                   nullptr, dgamma, dbeta);
  const long long total = M * C;
  if (C % 4 == 0 && pitch % 4 == 0 && dpitch % 4 == 0 && zpitch % 4 == 0 && total / 4 < (1ll << 31) && ((uintptr_t)z % 16 == 0) &&
      ((uintptr_t)dy % 16 == 0) && ((uintptr_t)dz % 16 == 0)) {
    const int total4 = (int)(total / 4);
    bn_silu_dz4_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, s>>>(z, dy, total4, C / 4, pitch, dpitch, zpitch, 1.f / (float)M, save_mean,
                                                                       save_invstd, gamma, beta, act, dgamma, dbeta, dz);
  } else {
    bn_silu_dz_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(z, dy, M, C, pitch, dpitch, zpitch, save_mean, save_invstd, gamma, beta,
                                                                  act, dgamma, dbeta, dz);
  }
  YB_CUDA_CHECK(cudaGetLastError());
  YB_CUDA_CHECK(cudaFreeAsync(part, s));
  return YB_OK;
}

// AdamW step (torch.optim.AdamW semantics, the optimizer the reference builds in YoloBaseTaskModel.cs:142-160):
//   p *= 1 - lr*wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  float pi = p[i] * (1.f - lr * wd);
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = pi - (lr / bc1) * (mi / denom);
}

int adamw_step(float* p, const float* g, float* m, float* v, long long n, int step, float lr, float b1, float b2, float eps,
               float wd, cudaStream_t s) {
  if (n <= 0 || step < 1) {
    set_error("yb_adamw_step: need n > 0 and step >= 1");
    return YB_ERR_INVALID_ARG;
  }
  const float bc1 = 1.f - powf(b1, (float)step), bc2_sqrt = sqrtf(1.f - powf(b2, (float)step));
  adamw_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, g, m, v, n, lr, b1, b2, eps, wd, bc1, bc2_sqrt);
  YB_CUDA_CHECK(cudaGetLastError());
  return YB_OK;
}

}  // namespace yb

// ------------------------------------------------------------------------------------------
// Convolution backward, fp32 CUDA-core parity path (the training-side twin of conv_generic_kernel: correct and
// deterministic first; the tcgen05 dgrad / wgrad kernels of the throughput path will be checked against it).
//   dz (N, Ho, Wo, Cout) NHWC, weights in the reference's checkpoint layout (Cout, Cin, k, k)
//   dgrad  dx[n,h,w,ci] = sum_{kh,kw,co} dz[n,(h+pad-kh)/s,(w+pad-kw)/s,co] * W[co,ci,kh,kw]   (only exact divisions)
//   wgrad  dW[co,ci,kh,kw] = sum_{n,ho,wo} dz[n,ho,wo,co] * x[n,ho*s+kh-pad,wo*s+kw-pad,ci]
// ------------------------------------------------------------------------------------------
namespace yb {

namespace {

__global__ void conv_dgrad_kernel(const float* __restrict__ dz, const float* __restrict__ w, int N, int H, int W, int Cin, int Ho,
                                  int Wo, int Cout, int k, int stride, int pad, float* __restrict__ dx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * H * W * Cin;
  if (i >= total) return;
  const int ci = (int)(i % Cin);
  long long p = i / Cin;
  const int x = (int)(p % W);
  p /= W;
  const int y = (int)(p % H);
  const int n = (int)(p / H);
  float acc = 0.f;
  for (int kh = 0; kh < k; kh++) {
    const int hn = y + pad - kh;
    if (hn < 0 || hn % stride) continue;
    const int ho = hn / stride;
    if (ho >= Ho) continue;
    for (int kw = 0; kw < k; kw++) {
      const int wn = x + pad - kw;
      if (wn < 0 || wn % stride) continue;
      const int wo = wn / stride;
      if (wo >= Wo) continue;
      const float* dzp = dz + (((size_t)n * Ho + ho) * Wo + wo) * Cout;
      const float* wp = w + ((size_t)ci * k + kh) * k + kw;  // + co * Cin*k*k
      for (int co = 0; co < Cout; co++) acc = fmaf(dzp[co], wp[(size_t)co * Cin * k * k], acc);
    }
  }
  dx[i] = acc;
}

// one block per (co, tap, slab of output rows): threads over ci, serial over the slab's pixels; the per-slab
// partials are folded in slab order by conv_wgrad_reduce_kernel (deterministic sums, no atomics)
__global__ void __launch_bounds__(128) conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dz, int N, int H, int W,
                                                        int Cin, int Ho, int Wo, int Cout, int k, int stride, int pad,
                                                        int rows_per_slab, float* __restrict__ part) {
  const int tap = blockIdx.x, co = blockIdx.y, slab = blockIdx.z;
  const int kh = tap / k, kw = tap - kh * k;
  const int r0 = slab * rows_per_slab, r1 = min(N * Ho, r0 + rows_per_slab);  // rows of the flattened (n, ho) index
  const size_t dw_size = (size_t)Cout * Cin * k * k;
  for (int ci = threadIdx.x; ci < Cin; ci += blockDim.x) {
    float acc = 0.f;
    for (int r = r0; r < r1; r++) {
      const int n = r / Ho, ho = r - n * Ho;
      const int hi = ho * stride + kh - pad;
      if (hi < 0 || hi >= H) continue;
      for (int wo = 0; wo < Wo; wo++) {
        const int wi = wo * stride + kw - pad;
        if (wi < 0 || wi >= W) continue;
        acc = fmaf(dz[(((size_t)n * Ho + ho) * Wo + wo) * Cout + co], x[(((size_t)n * H + hi) * W + wi) * Cin + ci], acc);
      }
    }
    part[(size_t)slab * dw_size + (((size_t)co * Cin + ci) * k + kh) * k + kw] = acc;
  }
}

__global__ void conv_wgrad_reduce_kernel(const float* __restrict__ part, int slabs, size_t dw_size, float* __restrict__ dw) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dw_size) return;
  float t = 0.f;
  for (int s = 0; s < slabs; s++) t += part[(size_t)s * dw_size + i];
  dw[i] = t;
}

}  // namespace

int conv_backward_data(const float* dz, const float* w, int N, int H, int W, int Cin, int Cout, int k, int stride, int pad,
                       float* dx, cudaStream_t s) {
  if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || k <= 0 || stride <= 0 || pad < 0) {
    set_error("yb_conv_backward_data: bad shape");
    return YB_ERR_SHAPE;
  }
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const long long total = (long long)N * H * W * Cin;
  conv_dgrad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(dz, w, N, H, W, Cin, Ho, Wo, Cout, k, stride, pad, dx);
  YB_CUDA_CHECK(cudaGetLastError());
  return YB_OK;
}

int conv_backward_weight(const float* x, const float* dz, int N, int H, int W, int Cin, int Cout, int k, int stride, int pad,
                         float* dw, cudaStream_t s) {
  if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || k <= 0 || stride <= 0 || pad < 0) {
    set_error("yb_conv_backward_weight: bad shape");
    return YB_ERR_SHAPE;
  }
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const size_t dw_size = (size_t)Cout * Cin * k * k;
  // enough slabs to fill the machine on the high-resolution layers, bounded partial buffer (<= 64 MB)
  int slabs = std::max(1, std::min(64, (N * Ho * Wo) / 2048));
  slabs = (int)std::max<size_t>(1, std::min<size_t>((size_t)slabs, ((size_t)16 << 20) / dw_size));
  slabs = std::min(slabs, N * Ho);
  const int rows_per_slab = (N * Ho + slabs - 1) / slabs;
  slabs = (N * Ho + rows_per_slab - 1) / rows_per_slab;
  float* part = nullptr;
  YB_CUDA_CHECK(cudaMallocAsync((void**)&part, (size_t)slabs * dw_size * sizeof(float), s));
  conv_wgrad_kernel<<<dim3(k * k, Cout, slabs), 128, 0, s>>>(x, dz, N, H, W, Cin, Ho, Wo, Cout, k, stride, pad, rows_per_slab, part);
  conv_wgrad_reduce_kernel<<<(unsigned)((dw_size + 255) / 256), 256, 0, s>>>(part, slabs, dw_size, dw);
  YB_CUDA_CHECK(cudaGetLastError());
  YB_CUDA_CHECK(cudaFreeAsync(part, s));
  return YB_OK;
}

}  // namespace yb
