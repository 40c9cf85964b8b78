// Validation-side post-processing on the GPU, batched over the images of a step:
//   * Metrics.box_iou            Utils/Metrics.cs:16-34   IoU = inter / (area1 + area2 - inter + eps), fp32
//   * match_predictions          Models/YoloBaseTaskModel.cs:377-446   the (detections x 10 IoU thresholds) true-positive
//                                 matrix that Detector.Val feeds to ap_per_class (Models/Detector.cs:103-120)
// The reference runs match_predictions per image on the host: nonzero -> argsort by IoU -> "unique by detection, then
// unique by label" with first occurrences found in a scalar loop of .item() calls.  What that procedure keeps is
//   best(d)  = the label with the highest IoU among the labels of d's class with IoU >= thr   (first unique: rows come
//              out ordered by detection index)
//   correct[d] = best(d) exists and d is the LOWEST detection index among { d' : best(d') == best(d) }   (second unique)
// which is what the kernel computes directly, one CTA per (image, threshold).
#include "common.cuh"

namespace yb {

__device__ __forceinline__ float box_iou_rn(float ax1, float ay1, float ax2, float ay2, float bx1, float by1, float bx2,
                                            float by2, float eps) {
  const float w = fmaxf(__fsub_rn(fminf(ax2, bx2), fmaxf(ax1, bx1)), 0.f);
  const float h = fmaxf(__fsub_rn(fminf(ay2, by2), fmaxf(ay1, by1)), 0.f);
  const float inter = __fmul_rn(w, h);
  const float a1 = __fmul_rn(__fsub_rn(ax2, ax1), __fsub_rn(ay2, ay1));
  const float a2 = __fmul_rn(__fsub_rn(bx2, bx1), __fsub_rn(by2, by1));
  return __fdiv_rn(inter, __fadd_rn(__fsub_rn(__fadd_rn(a1, a2), inter), eps));
}

__global__ void box_iou_kernel(const float* __restrict__ b1, int n, const float* __restrict__ b2, int m, float eps,
                               float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * m) return;
  const int i = idx / m, j = idx - i * m;
  out[idx] = box_iou_rn(b1[4 * i], b1[4 * i + 1], b1[4 * i + 2], b1[4 * i + 3], b2[4 * j], b2[4 * j + 1], b2[4 * j + 2],
                        b2[4 * j + 3], eps);
}

constexpr int MP_MAX_DET = 1024, MP_MAX_LABELS = 2048, MP_MAX_THR = 16;
struct MatchThr { float v[MP_MAX_THR]; };

__global__ void __launch_bounds__(256) match_predictions_kernel(const float* __restrict__ dets, const int* __restrict__ counts,
                                                                int max_det, int row_w, const float* __restrict__ labels,
                                                                int n_labels, MatchThr thr, int n_thr,
                                                                unsigned char* __restrict__ correct) {
  __shared__ int best[MP_MAX_DET];
  __shared__ int first_det[MP_MAX_LABELS];
  const int b = blockIdx.x, ti = blockIdx.y;
  const int n = min(counts[b], max_det);
  const float t = thr.v[ti];
  for (int l = threadIdx.x; l < n_labels; l += blockDim.x) first_det[l] = 0x7fffffff;
  __syncthreads();
  for (int d = threadIdx.x; d < n; d += blockDim.x) {
    const float* r = dets + ((size_t)b * max_det + d) * row_w;
    const float x1 = r[0], y1 = r[1], x2 = r[2], y2 = r[3], cls = r[5];
    int bl = -1;
    float bi = 0.f;
    for (int l = 0; l < n_labels; l++) {
      const float* g = labels + (size_t)l * 6;
      if ((int)g[0] != b || g[1] != cls) continue;  // iou * correct_class: other classes contribute 0 (< every threshold)
      const float iou = box_iou_rn(g[2], g[3], g[4], g[5], x1, y1, x2, y2, 1e-7f);
      if (iou >= t && (bl < 0 || iou > bi)) { bl = l; bi = iou; }
    }
    best[d] = bl;
    if (bl >= 0) atomicMin(&first_det[bl], d);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < max_det; d += blockDim.x) {
    unsigned char c = 0;
    if (d < n && best[d] >= 0 && first_det[best[d]] == d) c = 1;
    correct[((size_t)b * max_det + d) * n_thr + ti] = c;
  }
}


// Metrics.mask_iou (Utils/Metrics.cs:120-125): iou[i][j] = inter / (sum(mask1[i]) + sum(mask2[j]) - inter + eps) with
// inter = max(mask1[i] . mask2[j], 0) over the n = H * W pixels of the flattened masks.  One block per (i, j): the three sums
// in one pass (warp shuffles, then the 8 warp partials in order).  Seg validation compares a few dozen masks per image
// (Models/Segmenter.cs:142): a latency-sized problem, not a GEMM worth tensor cores.
__global__ void __launch_bounds__(256) mask_iou_kernel(const float* __restrict__ m1, const float* __restrict__ m2, int n, float eps,
                                                       float* __restrict__ out, int M) {
  const int i = blockIdx.y, j = blockIdx.x;
  const float* a = m1 + (size_t)i * n;
  const float* b = m2 + (size_t)j * n;
  float inter = 0.f, sa = 0.f, sb = 0.f;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const float x = a[k], y = b[k];
    inter = fmaf(x, y, inter);
    sa += x;
    sb += y;
  }
  __shared__ float red[3][8];
  for (int o = 16; o; o >>= 1) {
    inter += __shfl_down_sync(0xffffffffu, inter, o);
    sa += __shfl_down_sync(0xffffffffu, sa, o);
    sb += __shfl_down_sync(0xffffffffu, sb, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = inter; red[1][warp] = sa; red[2][warp] = sb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
    for (int w = 0; w < 8; w++) { t0 += red[0][w]; t1 += red[1][w]; t2 += red[2][w]; }
    t0 = fmaxf(t0, 0.f);
    out[(size_t)i * M + j] = __fdiv_rn(t0, __fadd_rn(__fsub_rn(__fadd_rn(t1, t2), t0), eps));
  }
}

}  // namespace yb

using namespace yb;

extern "C" {

int32_t yb_box_iou(const float* box1, int32_t n, const float* box2, int32_t m, float eps, float* out, void* stream) {
  if (n < 0 || m < 0 || ((n > 0 && m > 0) && (!box1 || !box2 || !out))) { set_error("yb_box_iou: bad argument"); return YB_ERR_INVALID_ARG; }
  if (n == 0 || m == 0) return YB_OK;
  box_iou_kernel<<<(n * m + 255) / 256, 256, 0, (cudaStream_t)stream>>>(box1, n, box2, m, eps, out);
  YB_CUDA_CHECK(cudaGetLastError());
  return YB_OK;
}

int32_t yb_mask_iou(const float* mask1, int32_t n1, const float* mask2, int32_t n2, int32_t pixels, float eps, float* out, void* stream) {
  if (n1 < 0 || n2 < 0 || pixels <= 0 || ((n1 > 0 && n2 > 0) && (!mask1 || !mask2 || !out))) { set_error("yb_mask_iou: bad argument"); return YB_ERR_INVALID_ARG; }
  if (n1 == 0 || n2 == 0) return YB_OK;
  if (n1 > 65535) { set_error("yb_mask_iou: at most 65535 rows in mask1"); return YB_ERR_INVALID_ARG; }
  mask_iou_kernel<<<dim3(n2, n1), 256, 0, (cudaStream_t)stream>>>(mask1, mask2, pixels, eps, out, n2);
  YB_CUDA_CHECK(cudaGetLastError());
  return YB_OK;
}

int32_t yb_match_predictions(const float* dets, const int32_t* counts, int32_t batch, int32_t max_det, int32_t row_width,
                             const float* labels, int32_t n_labels, const float* iou_thresholds_host, int32_t n_thresholds,
                             uint8_t* correct, void* stream) {
  if (!dets || !counts || !correct || !iou_thresholds_host || (n_labels > 0 && !labels)) { set_error("yb_match_predictions: null argument"); return YB_ERR_INVALID_ARG; }
  if (batch <= 0 || max_det <= 0 || max_det > MP_MAX_DET || row_width < 6 || n_labels < 0 || n_labels > MP_MAX_LABELS ||
      n_thresholds <= 0 || n_thresholds > MP_MAX_THR) {
    set_error("yb_match_predictions: need 0 < max_det <= 1024, row_width >= 6, n_labels <= 2048, 0 < n_thresholds <= 16");
    return YB_ERR_INVALID_ARG;
  }
  MatchThr t;
  for (int i = 0; i < MP_MAX_THR; i++) t.v[i] = i < n_thresholds ? iou_thresholds_host[i] : 2.f;
  match_predictions_kernel<<<dim3(batch, n_thresholds), 256, 0, (cudaStream_t)stream>>>(dets, counts, max_det, row_width, labels,
                                                                                      n_labels, t, n_thresholds, correct);
  YB_CUDA_CHECK(cudaGetLastError());
  return YB_OK;
}

}  // extern "C"
