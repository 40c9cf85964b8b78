// Detection training loss on the GPU: TaskAlignedAssigner + CIoU / DFL / BCE with gradients.
//
// Restates Utils/Loss.cs:328-485 (v8DetectionLoss), :122-167 (BboxLoss), :94-120 (DFLoss),
// Utils/Tal.cs:13-311 (TaskAlignedAssigner) and Utils/Metrics.cs:36-111 (bbox_iou, CIoU) of the reference for the
// raw head outputs of the train-mode forward ("boxes" (B, 4*reg_max, A) distribution logits, "scores" (B, nc, A)
// class logits).  Outputs the three loss items and, optionally, d(sum of items * batch)/d(boxes, scores) - the
// tensors the backward pass of the network starts from.  The assignment runs on detached values exactly as in the
// reference (no gradient through the targets).
//
// Kernels (all fp32, HBM/latency-bound on tiny data next to the network itself):
//   loss_decode_kernel   per (b, a): softmax-expectation of the 4 x reg_max bins -> xyxy box in grid units
//   tal_metric_kernel    per (b, g, a): anchor inside (widened) gt?  overlap = clamp(CIoU, 0), align = s^0.5 ov^6
//   tal_topk_kernel      per (b, g): top-k anchors by align metric (ties: lowest anchor index)
//   tal_resolve_kernel   per (b, a): anchors claimed by several gts keep the gt of largest overlap
//   tal_norm_kernel      per (b, g): max align / max overlap over its positives
//   loss_grad_kernel     per (b, a): BCE over classes, CIoU and DFL of foreground anchors, gradients; block sums
// CIoU derivatives come from forward-mode dual numbers (the reference keeps `alpha` in the autograd graph,
// Metrics.cs:101, so the closed form of the usual CIoU gradient does not apply).
#include <cmath>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace yb {

namespace {

// ---------------------------------------------------------------- dual numbers over the 4 box coordinates
struct D4 {
  float v, d[4];
};
__device__ __forceinline__ D4 dconst(float c) { return D4{c, {0.f, 0.f, 0.f, 0.f}}; }
__device__ __forceinline__ D4 dvar(float c, int i) {
  D4 r = dconst(c);
  r.d[i] = 1.f;
  return r;
}
__device__ __forceinline__ D4 operator+(const D4& a, const D4& b) {
  D4 r;
  r.v = a.v + b.v;
  for (int i = 0; i < 4; i++) r.d[i] = a.d[i] + b.d[i];
  return r;
}
__device__ __forceinline__ D4 operator-(const D4& a, const D4& b) {
  D4 r;
  r.v = a.v - b.v;
  for (int i = 0; i < 4; i++) r.d[i] = a.d[i] - b.d[i];
  return r;
}
__device__ __forceinline__ D4 operator*(const D4& a, const D4& b) {
  D4 r;
  r.v = a.v * b.v;
  for (int i = 0; i < 4; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
__device__ __forceinline__ D4 operator/(const D4& a, const D4& b) {
  D4 r;
  const float inv = 1.f / b.v;
  r.v = a.v * inv;
  for (int i = 0; i < 4; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
__device__ __forceinline__ D4 dmin(const D4& a, const D4& b) { return a.v <= b.v ? a : b; }   // torch.minimum: grad to the
__device__ __forceinline__ D4 dmax(const D4& a, const D4& b) { return a.v >= b.v ? a : b; }   // selected operand
__device__ __forceinline__ D4 dclamp_min(const D4& a, float lo) { return a.v < lo ? dconst(lo) : a; }
__device__ __forceinline__ D4 datan(const D4& a) {
  D4 r;
  r.v = atanf(a.v);
  const float g = 1.f / (1.f + a.v * a.v);
  for (int i = 0; i < 4; i++) r.d[i] = a.d[i] * g;
  return r;
}

// Metrics.cs:36-111 (xywh = false, CIoU = true); box1 carries the derivatives, box2 is constant.
__device__ __forceinline__ D4 ciou_dual(const D4 b1[4], const float b2[4]) {
  const float eps = 1e-7f;
  const D4 w1 = b1[2] - b1[0], h1 = dclamp_min(b1[3] - b1[1], eps);
  const float w2 = b2[2] - b2[0], h2 = fmaxf(b2[3] - b2[1], eps);
  const D4 iw = dclamp_min(dmin(b1[2], dconst(b2[2])) - dmax(b1[0], dconst(b2[0])), 0.f);
  const D4 ih = dclamp_min(dmin(b1[3], dconst(b2[3])) - dmax(b1[1], dconst(b2[1])), 0.f);
  const D4 inter = iw * ih;
  const D4 uni = w1 * h1 + dconst(w2 * h2) - inter + dconst(eps);
  const D4 iou = inter / uni;
  const D4 cw = dmax(b1[2], dconst(b2[2])) - dmin(b1[0], dconst(b2[0]));
  const D4 ch = dmax(b1[3], dconst(b2[3])) - dmin(b1[1], dconst(b2[1]));
  const D4 c2 = cw * cw + ch * ch + dconst(eps);
  const D4 dx = dconst(b2[0] + b2[2]) - b1[0] - b1[2], dy = dconst(b2[1] + b2[3]) - b1[1] - b1[3];
  const D4 rho2 = (dx * dx + dy * dy) * dconst(0.25f);
  const D4 da = dconst(atanf(w2 / h2)) - datan(w1 / h1);
  const D4 v = da * da * dconst(4.0f / (3.14159265358979323846f * 3.14159265358979323846f));
  const D4 alpha = v / (v - iou + dconst(1.0f + eps));  // in the graph, as in the reference
  return iou - (rho2 / c2 + v * alpha);
}

__device__ __forceinline__ float ciou_value(const float b1[4], const float b2[4]) {
  D4 d[4];
  for (int i = 0; i < 4; i++) d[i] = dconst(b1[i]);
  return ciou_dual(d, b2).v;
}

struct LossGeom {
  int B, A, nc, reg_max, n_max;  // n_max = padded ground truths per image
  int lvl_w[3], lvl_h[3], lvl_a0[3];
  float lvl_stride[3];
  int topk;
};

__device__ __forceinline__ void anchor_of(const LossGeom& g, int a, float& ax, float& ay, float& stride) {
  int l = a >= g.lvl_a0[2] ? 2 : (a >= g.lvl_a0[1] ? 1 : 0);
  const int i = a - g.lvl_a0[l];
  const int y = i / g.lvl_w[l];
  ax = (float)(i - y * g.lvl_w[l]) + 0.5f;  // Tal.cs:313-335, grid_cell_offset 0.5
  ay = (float)y + 0.5f;
  stride = g.lvl_stride[l];
}

// bbox_decode (Loss.cs:397-408): softmax over reg_max bins, expectation, dist2bbox(xyxy) in grid units
__global__ void loss_decode_kernel(LossGeom g, const float* __restrict__ boxes, float* __restrict__ pbox) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (a >= g.A) return;
  float ax, ay, st;
  anchor_of(g, a, ax, ay, st);
  float dist[4];
  for (int s = 0; s < 4; s++) {
    const float* p = boxes + ((size_t)b * 4 * g.reg_max + (size_t)s * g.reg_max) * g.A + a;
    float mx = -INFINITY;
    for (int i = 0; i < g.reg_max; i++) mx = fmaxf(mx, p[(size_t)i * g.A]);
    float sum = 0.f, ex = 0.f;
    for (int i = 0; i < g.reg_max; i++) {
      const float e = expf(p[(size_t)i * g.A] - mx);
      sum += e;
      ex += e * (float)i;
    }
    dist[s] = ex / sum;
  }
  float* o = pbox + ((size_t)b * g.A + a) * 4;
  o[0] = ax - dist[0]; o[1] = ay - dist[1]; o[2] = ax + dist[2]; o[3] = ay + dist[3];
}

// get_pos_mask / get_box_metrics / select_candidates_in_gts (Tal.cs:90-142, 213-235)
//   gts: (B, n_max, 5) = [cls, x1, y1, x2, y2] pixels, zero rows = padding (mask_gt = sum(xyxy) > 0, Loss.cs:429)
__global__ void tal_metric_kernel(LossGeom g, const float* __restrict__ scores, const float* __restrict__ pbox,
                                  const float* __restrict__ gts, float* __restrict__ ov, float* __restrict__ am,
                                  unsigned char* __restrict__ in_gts) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x, gi = blockIdx.y, b = blockIdx.z;
  if (a >= g.A) return;
  const float* gt = gts + ((size_t)b * g.n_max + gi) * 5;
  const size_t o = ((size_t)b * g.n_max + gi) * g.A + a;
  const bool valid = (gt[1] + gt[2] + gt[3] + gt[4]) > 0.0f;
  float ax, ay, st;
  anchor_of(g, a, ax, ay, st);
  const float px = ax * st, py = ay * st;
  // boxes narrower / lower than the smallest stride are widened to stride_val = strides[1] (Tal.cs:217-222)
  const float cx = (gt[1] + gt[3]) / 2, cy = (gt[2] + gt[4]) / 2;
  float w = gt[3] - gt[1], h = gt[4] - gt[2];
  if (valid && w < g.lvl_stride[0]) w = g.lvl_stride[1];
  if (valid && h < g.lvl_stride[0]) h = g.lvl_stride[1];
  const float x1 = cx - w / 2, y1 = cy - h / 2, x2 = cx + w / 2, y2 = cy + h / 2;
  const bool inside = fminf(fminf(px - x1, py - y1), fminf(x2 - px, y2 - py)) > 1e-9f;
  float overlap = 0.f, metric = 0.f;
  if (inside && valid) {
    const float* pb = pbox + ((size_t)b * g.A + a) * 4;
    const float pp[4] = {pb[0] * st, pb[1] * st, pb[2] * st, pb[3] * st};  // pred_bboxes * stride_tensor (Loss.cs:436)
    const float gg[4] = {gt[1], gt[2], gt[3], gt[4]};
    D4 d[4];
    for (int i = 0; i < 4; i++) d[i] = dconst(gg[i]);
    overlap = fmaxf(ciou_dual(d, pp).v, 0.f);  // iou_calculation(gt, pd).clamp(0), Tal.cs:139-142
    const int cls = (int)gt[0];
    const float x = scores[((size_t)b * g.nc + cls) * g.A + a];
    const float s = 1.f / (1.f + expf(-x));
    const float o2 = overlap * overlap;
    metric = sqrtf(s) * (o2 * o2 * o2);  // alpha = 0.5, beta = 6 (Loss.cs:358)
  }
  ov[o] = overlap;
  am[o] = metric;
  in_gts[o] = (inside && valid) ? 1 : 0;
}

// select_topk_candidates (Tal.cs:144-167): top-k anchors of every valid gt by align metric; ties -> lowest index.
// (Zero-metric picks carry target score 0 and weight 0: they contribute neither loss nor gradient.)
constexpr int TOPK_MAX = 16;
__global__ void __launch_bounds__(256) tal_topk_kernel(LossGeom g, const float* __restrict__ am,
                                                       const unsigned char* __restrict__ in_gts,
                                                       const float* __restrict__ gts, unsigned char* __restrict__ mask_pos) {
  const int gi = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const float* gt = gts + ((size_t)b * g.n_max + gi) * 5;
  const size_t base = ((size_t)b * g.n_max + gi) * g.A;
  for (int a = tid; a < g.A; a += 256) mask_pos[base + a] = 0;
  if (!((gt[1] + gt[2] + gt[3] + gt[4]) > 0.0f)) return;
  __shared__ float s_v[256 * TOPK_MAX];
  __shared__ int s_i[256 * TOPK_MAX];
  float bv[TOPK_MAX];
  int bi[TOPK_MAX];
  for (int k = 0; k < g.topk; k++) { bv[k] = -1.f; bi[k] = 0x7fffffff; }
  for (int a = tid; a < g.A; a += 256) {
    const float v = am[base + a];
    int k = g.topk;
    while (k > 0 && (v > bv[k - 1] || (v == bv[k - 1] && a < bi[k - 1]))) k--;
    if (k < g.topk) {
      for (int j = g.topk - 1; j > k; j--) { bv[j] = bv[j - 1]; bi[j] = bi[j - 1]; }
      bv[k] = v; bi[k] = a;
    }
  }
  for (int k = 0; k < g.topk; k++) { s_v[tid * TOPK_MAX + k] = bv[k]; s_i[tid * TOPK_MAX + k] = bi[k]; }
  __syncthreads();
  if (tid == 0) {
    int head[256];
    for (int t = 0; t < 256; t++) head[t] = 0;
    for (int k = 0; k < g.topk; k++) {  // k-way merge of the 256 sorted per-thread lists
      float best = -2.f;
      int besti = 0x7fffffff, bt = -1;
      for (int t = 0; t < 256; t++) {
        if (head[t] >= g.topk) continue;
        const float v = s_v[t * TOPK_MAX + head[t]];
        const int i = s_i[t * TOPK_MAX + head[t]];
        if (v > best || (v == best && i < besti)) { best = v; besti = i; bt = t; }
      }
      if (bt < 0 || besti == 0x7fffffff) break;
      head[bt]++;
      mask_pos[base + besti] = in_gts[base + besti];  // mask_topk * mask_in_gts * mask_gt
    }
  }
}

// select_highest_overlaps (Tal.cs:237-266, topk2 == topk) + get_targets: per anchor the assigned gt, fg flag
__global__ void tal_resolve_kernel(LossGeom g, const float* __restrict__ ov, unsigned char* __restrict__ mask_pos,
                                   int* __restrict__ gt_idx, unsigned char* __restrict__ fg) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (a >= g.A) return;
  int cnt = 0, first = 0;
  for (int gi = g.n_max - 1; gi >= 0; gi--)
    if (mask_pos[((size_t)b * g.n_max + gi) * g.A + a]) { cnt++; first = gi; }
  if (cnt > 1) {  // keep the gt with the largest overlap (argmax over ALL gts, first maximum)
    int best = 0;
    float bo = ov[((size_t)b * g.n_max) * g.A + a];
    for (int gi = 1; gi < g.n_max; gi++) {
      const float o = ov[((size_t)b * g.n_max + gi) * g.A + a];
      if (o > bo) { bo = o; best = gi; }
    }
    for (int gi = 0; gi < g.n_max; gi++) mask_pos[((size_t)b * g.n_max + gi) * g.A + a] = gi == best ? 1 : 0;
    first = best;
    cnt = 1;
  }
  gt_idx[(size_t)b * g.A + a] = cnt ? first : 0;
  fg[(size_t)b * g.A + a] = cnt ? 1 : 0;
}

// pos_align_metrics / pos_overlaps (Tal.cs:84-86): per gt the maxima over its positive anchors
__global__ void __launch_bounds__(256) tal_norm_kernel(LossGeom g, const float* __restrict__ ov, const float* __restrict__ am,
                                                       const unsigned char* __restrict__ mask_pos, float* __restrict__ pos_am,
                                                       float* __restrict__ pos_ov) {
  const int gi = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const size_t base = ((size_t)b * g.n_max + gi) * g.A;
  float ma = 0.f, mo = 0.f;
  for (int a = tid; a < g.A; a += 256)
    if (mask_pos[base + a]) { ma = fmaxf(ma, am[base + a]); mo = fmaxf(mo, ov[base + a]); }
  __shared__ float sa[256], so[256];
  sa[tid] = ma; so[tid] = mo;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) { sa[tid] = fmaxf(sa[tid], sa[tid + s]); so[tid] = fmaxf(so[tid], so[tid + s]); }
    __syncthreads();
  }
  if (tid == 0) { pos_am[b * g.n_max + gi] = sa[0]; pos_ov[b * g.n_max + gi] = so[0]; }
}

// target score of every anchor (Tal.cs:86-88) and their sum
__global__ void tal_score_kernel(LossGeom g, const float* __restrict__ ov, const float* __restrict__ am,
                                 const int* __restrict__ gt_idx, const unsigned char* __restrict__ fg,
                                 const float* __restrict__ pos_am, const float* __restrict__ pos_ov,
                                 float* __restrict__ tscore, float* __restrict__ tss) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  float v = 0.f;
  if (a < g.A && fg[(size_t)b * g.A + a]) {
    const int gi = gt_idx[(size_t)b * g.A + a];
    const size_t o = ((size_t)b * g.n_max + gi) * g.A + a;
    v = am[o] * pos_ov[b * g.n_max + gi] / (pos_am[b * g.n_max + gi] + 1e-9f);
  }
  if (a < g.A) tscore[(size_t)b * g.A + a] = v;
  // block sum -> one atomic
  __shared__ float red[256];
  red[threadIdx.x] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0 && red[0] != 0.f) atomicAdd(tss, red[0]);
}

// Loss.cs:443-465 + BboxLoss + DFLoss with gradients of (box + cls + dfl) * gains * batch
__global__ void __launch_bounds__(128) loss_grad_kernel(LossGeom g, const float* __restrict__ boxes, const float* __restrict__ scores,
                                                        const float* __restrict__ pbox, const float* __restrict__ gts,
                                                        const int* __restrict__ gt_idx, const unsigned char* __restrict__ fg,
                                                        const float* __restrict__ tscore, const float* __restrict__ tss_p,
                                                        float hyp_box, float hyp_cls, float hyp_dfl, float* __restrict__ loss,
                                                        float* __restrict__ gboxes, float* __restrict__ gscores) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  const float tss = fmaxf(*tss_p, 1.0f);  // Math.Max(target_scores.sum(), 1), Loss.cs:440
  const float gscale = (float)g.B / tss;   // the returned loss is loss * batch_size (Loss.cs:473)
  float l_box = 0.f, l_cls = 0.f, l_dfl = 0.f;
  if (a < g.A) {
    const bool is_fg = fg[(size_t)b * g.A + a] != 0;
    const float ts = tscore[(size_t)b * g.A + a];
    const float* gt = gts + ((size_t)b * g.n_max + (is_fg ? gt_idx[(size_t)b * g.A + a] : 0)) * 5;
    const int tcls = is_fg ? (int)gt[0] : -1;
    // ---- classification: BCEWithLogits(pred_scores, target_scores).sum() / tss
    for (int c = 0; c < g.nc; c++) {
      const size_t o = ((size_t)b * g.nc + c) * g.A + a;
      const float x = scores[o], t = c == tcls ? ts : 0.f;
      l_cls += fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
      if (gscores) gscores[o] = (1.f / (1.f + expf(-x)) - t) * hyp_cls * gscale;
    }
    float ax, ay, st;
    anchor_of(g, a, ax, ay, st);
    if (is_fg) {
      const float weight = ts;  // target_scores.sum(-1)
      const float* pb = pbox + ((size_t)b * g.A + a) * 4;
      const float tb[4] = {gt[1] / st, gt[2] / st, gt[3] / st, gt[4] / st};  // target_bboxes / stride_tensor
      D4 d[4];
      for (int i = 0; i < 4; i++) d[i] = dvar(pb[i], i);
      const D4 iou = ciou_dual(d, tb);
      l_box = (1.0f - iou.v) * weight;
      // d(box loss)/d(pred xyxy) = -weight * d(iou); x1 = ax - l, y1 = ay - t, x2 = ax + r, y2 = ay + b
      const float gd[4] = {weight * iou.d[0], weight * iou.d[1], -weight * iou.d[2], -weight * iou.d[3]};  // wrt l, t, r, b
      // DFL target: bbox2dist(anchor, target, reg_max - 1) then clamp (Tal.cs:364-378, Loss.cs:106)
      const float tl4[4] = {ax - tb[0], ay - tb[1], tb[2] - ax, tb[3] - ay};
      for (int s = 0; s < 4; s++) {
        const float* p = boxes + ((size_t)b * 4 * g.reg_max + (size_t)s * g.reg_max) * g.A + a;
        float mx = -INFINITY;
        for (int i = 0; i < g.reg_max; i++) mx = fmaxf(mx, p[(size_t)i * g.A]);
        float sum = 0.f, ex = 0.f;
        for (int i = 0; i < g.reg_max; i++) {
          const float e = expf(p[(size_t)i * g.A] - mx);
          sum += e;
          ex += e * (float)i;
        }
        const float dist = ex / sum, lse = mx + logf(sum);
        float t = fminf(fmaxf(tl4[s], 0.f), (float)(g.reg_max - 1) - 0.01f);
        t = fminf(fmaxf(t, 0.f), (float)(g.reg_max - 1) - 0.01f);
        const int il = (int)t;
        const float wl = (float)(il + 1) - t, wr = 1.f - wl;
        l_dfl += ((lse - p[(size_t)il * g.A]) * wl + (lse - p[(size_t)(il + 1) * g.A]) * wr) * 0.25f * weight;
        if (gboxes) {
          for (int i = 0; i < g.reg_max; i++) {
            const float pi = expf(p[(size_t)i * g.A] - mx) / sum;
            const float g_iou = gd[s] * pi * ((float)i - dist);                       // through the softmax expectation
            const float g_dfl = (pi - (i == il ? wl : 0.f) - (i == il + 1 ? wr : 0.f)) * 0.25f * weight;
            gboxes[((size_t)b * 4 * g.reg_max + (size_t)s * g.reg_max + i) * g.A + a] =
                (g_iou * hyp_box + g_dfl * hyp_dfl) * gscale;
          }
        }
      }
    } else if (gboxes) {
      for (int c = 0; c < 4 * g.reg_max; c++) gboxes[((size_t)b * 4 * g.reg_max + c) * g.A + a] = 0.f;
    }
  }
  __shared__ float red[3][128];
  red[0][threadIdx.x] = l_box; red[1][threadIdx.x] = l_cls; red[2][threadIdx.x] = l_dfl;
  __syncthreads();
  for (int s = 64; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int k = 0; k < 3; k++) red[k][threadIdx.x] += red[k][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    atomicAdd(&loss[0], red[0][0] * hyp_box / tss);
    atomicAdd(&loss[1], red[1][0] * hyp_cls / tss);
    atomicAdd(&loss[2], red[2][0] * hyp_dfl / tss);
  }
}

}  // namespace

// preprocess (Loss.cs:363-389) on the host: rows [img, cls, x, y, w, h] (normalised xywh) -> (B, n_max, 5) [cls, xyxy pixels],
// zero rows as padding.  Split from the launch so that a training step can stage its targets BEFORE it queues the forward
// pass (no host synchronisation in the middle of the step).
int detection_loss_prepare(const float* targets_host, int n_targets, int B, int nc, int H, int W, std::vector<float>& gts, int* n_max_out) {
  if (B <= 0 || n_targets < 0 || (n_targets > 0 && !targets_host)) { set_error("yb_detection_loss: bad targets"); return YB_ERR_INVALID_ARG; }
  std::vector<int> count(B, 0);
  for (int i = 0; i < n_targets; i++) {
    const int bi = (int)targets_host[(size_t)i * 6];
    if (bi < 0 || bi >= B) {
      set_error("yb_detection_loss: target image index out of range");
      return YB_ERR_INVALID_ARG;
    }
    const int c = (int)targets_host[(size_t)i * 6 + 1];
    if (c < 0 || c >= nc) {
      set_error("yb_detection_loss: target class out of range");
      return YB_ERR_INVALID_ARG;
    }
    count[bi]++;
  }
  int n_max = 0;
  for (int b = 0; b < B; b++) n_max = std::max(n_max, count[b]);
  n_max = std::max(n_max, 1);  // an all-padding row keeps the kernels uniform when there are no targets
  gts.assign((size_t)B * n_max * 5, 0.f);
  std::fill(count.begin(), count.end(), 0);
  for (int i = 0; i < n_targets; i++) {
    const float* t = targets_host + (size_t)i * 6;
    const int bi = (int)t[0];
    float* o = gts.data() + ((size_t)bi * n_max + count[bi]++) * 5;
    const float x = t[2] * (float)W, y = t[3] * (float)H, w = t[4] * (float)W, h = t[5] * (float)H;  // imgsz[[1,0,1,0]]
    o[0] = t[1];
    o[1] = x - w / 2; o[2] = y - h / 2; o[3] = x + w / 2; o[4] = y + h / 2;  // xywh2xyxy, Ops.cs:68-81
  }
  *n_max_out = n_max;
  return YB_OK;
}

// the device part: d_gts (B, n_max, 5) already on the device (stream-ordered before this call)
int detection_loss_launch_dev(const float* boxes, const float* scores, int B, int nc, int reg_max, int H, int W, const float* d_gts,
                              int n_max, int topk, float hyp_box, float hyp_cls, float hyp_dfl, float* loss_items, float* grad_boxes,
                              float* grad_scores, unsigned char* fg_out, int* gt_idx_out, float* tscore_out, cudaStream_t s) {
  if (B <= 0 || nc <= 0 || nc >= 4096 || reg_max < 2 || reg_max > 32 || H % 32 || W % 32 || H <= 0 || W <= 0 || topk < 1 ||
      topk > TOPK_MAX || n_max < 1) {
    set_error("yb_detection_loss: unsupported shape (need H, W multiples of 32, 2 <= reg_max <= 32, 1 <= topk <= 16)");
    return YB_ERR_SHAPE;
  }
  LossGeom g;
  g.B = B; g.nc = nc; g.reg_max = reg_max; g.topk = topk;
  int a0 = 0;
  for (int l = 0; l < 3; l++) {
    const int st = 8 << l;
    g.lvl_w[l] = W / st; g.lvl_h[l] = H / st; g.lvl_a0[l] = a0; g.lvl_stride[l] = (float)st;
    a0 += g.lvl_w[l] * g.lvl_h[l];
  }
  g.A = a0;
  g.n_max = n_max;
  const size_t nga = (size_t)B * g.n_max * g.A, na = (size_t)B * g.A;
  // scratch: pbox | ov | am | pos_am | pos_ov | tscore | tss | gt_idx | in_gts | mask_pos | fg
  const size_t f_count = na * 4 + nga * 2 + (size_t)B * g.n_max * 2 + na + 1;
  char* scratch = nullptr;
  const size_t bytes = f_count * 4 + na * 4 + nga * 2 + na + 64;
  YB_CUDA_CHECK(cudaMallocAsync((void**)&scratch, bytes, s));
  float* pbox = reinterpret_cast<float*>(scratch);
  float* ov = pbox + na * 4;
  float* am = ov + nga;
  float* pos_am = am + nga;
  float* pos_ov = pos_am + (size_t)B * g.n_max;
  float* tscore = pos_ov + (size_t)B * g.n_max;
  float* tss = tscore + na;
  int* gt_idx = reinterpret_cast<int*>(tss + 1);
  unsigned char* in_gts = reinterpret_cast<unsigned char*>(gt_idx + na);
  unsigned char* mask_pos = in_gts + nga;
  unsigned char* fg = mask_pos + nga;
  // from here on every error path must release `scratch`: keep the first error and fall through to the free
  cudaError_t ce = cudaMemsetAsync(tss, 0, 4, s);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(loss_items, 0, 12, s);
  if (ce == cudaSuccess) {
    const dim3 ga((g.A + 255) / 256, B);
    loss_decode_kernel<<<ga, 256, 0, s>>>(g, boxes, pbox);
    tal_metric_kernel<<<dim3((g.A + 255) / 256, g.n_max, B), 256, 0, s>>>(g, scores, pbox, d_gts, ov, am, in_gts);
    tal_topk_kernel<<<dim3(g.n_max, B), 256, 0, s>>>(g, am, in_gts, d_gts, mask_pos);
    tal_resolve_kernel<<<ga, 256, 0, s>>>(g, ov, mask_pos, gt_idx, fg);
    tal_norm_kernel<<<dim3(g.n_max, B), 256, 0, s>>>(g, ov, am, mask_pos, pos_am, pos_ov);
    tal_score_kernel<<<ga, 256, 0, s>>>(g, ov, am, gt_idx, fg, pos_am, pos_ov, tscore, tss);
    loss_grad_kernel<<<dim3((g.A + 127) / 128, B), 128, 0, s>>>(g, boxes, scores, pbox, d_gts, gt_idx, fg, tscore, tss, hyp_box,
                                                                hyp_cls, hyp_dfl, loss_items, grad_boxes, grad_scores);
    ce = cudaGetLastError();
  }
  if (ce == cudaSuccess && fg_out) ce = cudaMemcpyAsync(fg_out, fg, na, cudaMemcpyDeviceToDevice, s);
  if (ce == cudaSuccess && gt_idx_out) ce = cudaMemcpyAsync(gt_idx_out, gt_idx, na * 4, cudaMemcpyDeviceToDevice, s);
  if (ce == cudaSuccess && tscore_out) ce = cudaMemcpyAsync(tscore_out, tscore, na * 4, cudaMemcpyDeviceToDevice, s);
  cudaFreeAsync(scratch, s);
  if (ce != cudaSuccess) { set_error(std::string("yb_detection_loss: ") + cudaGetErrorString(ce)); return YB_ERR_CUDA; }
  return YB_OK;
}

int detection_loss_launch(const float* boxes, const float* scores, int B, int nc, int reg_max, int H, int W,
                          const float* targets_host, int n_targets, int topk, float hyp_box, float hyp_cls, float hyp_dfl,
                          float* loss_items, float* grad_boxes, float* grad_scores, unsigned char* fg_out, int* gt_idx_out,
                          float* tscore_out, cudaStream_t s) {
  std::vector<float> gts;
  int n_max = 0;
  if (int rc = detection_loss_prepare(targets_host, n_targets, B, nc, H, W, gts, &n_max)) return rc;
  float* d_gts = nullptr;
  YB_CUDA_CHECK(cudaMallocAsync((void**)&d_gts, gts.size() * 4, s));
  cudaError_t ce = cudaMemcpyAsync(d_gts, gts.data(), gts.size() * 4, cudaMemcpyHostToDevice, s);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);  // `gts` is a host temporary
  if (ce != cudaSuccess) {
    cudaFreeAsync(d_gts, s);
    set_error(std::string("yb_detection_loss: ") + cudaGetErrorString(ce));
    return YB_ERR_CUDA;
  }
  const int rc = detection_loss_launch_dev(boxes, scores, B, nc, reg_max, H, W, d_gts, n_max, topk, hyp_box, hyp_cls, hyp_dfl, loss_items,
                                           grad_boxes, grad_scores, fg_out, gt_idx_out, tscore_out, s);
  cudaFreeAsync(d_gts, s);
  return rc;
}

}  // namespace yb
