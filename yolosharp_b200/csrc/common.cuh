// Shared declarations for the yolob200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string>

#include "yolob200.h"

namespace yb {

// A channel-slice view of an NHWC activation buffer: element (n,h,w,c) lives at
// base[((n*H + h)*W + w) * pitch + coff + c].  Concat/chunk of the reference graph
// (Modules/Block.cs:391-396) become views of one wider buffer - no copies.
struct View {
  void* base = nullptr;  // device pointer to the buffer start (element type = engine storage type)
  int H = 0, W = 0;
  int pitch = 0;  // channels of the whole buffer
  int coff = 0;   // first channel of this view
  int C = 0;      // channels of this view
};

enum Act { ACT_NONE = 0, ACT_SILU = 1 };

// Head-tail fusion (tcgen05 path): the final 1x1 convs of the Detect branches write straight into the
// prediction tensor (B, Ctot, A) instead of an NHWC buffer (Modules/Head.cs:204-223 decode).
enum EpiMode { EPI_STORE = 0, EPI_DFL_BOX = 1, EPI_SIGMOID = 2, EPI_RAW = 3 };
struct EpiDecode {
  int mode = EPI_STORE;
  int A = 0, Ctot = 0;  // anchors per image, channels of pred
  int a0 = 0;           // first anchor of this pyramid level
  int ch0 = 0;          // first pred channel written (4 for class probs, 4+nc for mask coeffs)
  int Wl = 0, HW = 0;   // level width, pixels per image
  float stride = 0.f;
};

struct ConvParams {
  View in, out, res;  // res.base == nullptr -> no residual
  View out2;          // optional second destination: nearest-2x upsampled copy (out2.base != nullptr)
  EpiDecode dec;
  int share_sms = 0;  // 1: op runs concurrently with sibling branches - size its grid to half the SMs
  const void* w;      // packed weights, layout depends on kernel
  const float* bias;  // [Cout] fp32 (folded BN beta - mean*scale, or conv bias)
  int B;
  int Cin, Cout;
  int k, stride, pad;
  int Ho, Wo;
  int act;
};

void set_error(const std::string& msg);

#define YB_CUDA_CHECK(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::yb::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " +   \
                      __FILE__ + ":" + std::to_string(__LINE__));                            \
      return YB_ERR_CUDA;                                                                    \
    }                                                                                        \
  } while (0)

// ---- programmatic dependent launch for the kernels of the training step ----
// A kernel launched through launch_pdl may be scheduled while its predecessor in the stream is still running: its CTAs
// become resident as the predecessor's retire, run their prologue (parameter loads, barrier / TMEM setup, descriptor
// prefetch) and block in pdl_wait() until the predecessor has completed and its writes are visible.  Every such kernel
// calls pdl_wait() before its first global-memory access and pdl_trigger() right AFTER it: the successor can be scheduled
// once all CTAs of this kernel have passed their wait, so at most two kernels of the chain are ever in flight (this one
// finishing, the next one in its prologue).  (Triggering before the wait lets a whole chain of small kernels become
// resident at once; with that form a seven-kernel loss chain read a scalar before its producer's atomics - not understood,
// so the conservative order is used everywhere.)  A dependent step of ~800 small kernels otherwise pays ~1.8 us of drain +
// launch + fill per boundary (profiles/r2_train_profile_v11s_native.txt: 16.8 ms of kernels in an 18.2 ms step).
// Both instructions are no-ops in a kernel launched the ordinary way, so a kernel may be launched either way.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();  // false when YB_NO_PDL is set (A/B measurements)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(static_cast<Args&&>(args))...);
}
#endif

// ---- kernels_generic.cu : CUDA-core kernels, templated on storage type (float | __half) ----
template <typename T>
int launch_conv_generic(const ConvParams& p, cudaStream_t s);
template <typename T>
int launch_dwconv3x3(const ConvParams& p, cudaStream_t s);  // depthwise 3x3 s1 p1, w [9][C] fp32
template <typename T>
int launch_sppf_pool(const View& in, const View& out5, const View& out9, const View& out13, int B,
                     cudaStream_t s);
template <typename T>
int launch_upsample2x(const View& in, const View& out, int B, cudaStream_t s);
// network input (B,3,H,W) NCHW of dtype u8/f16/f32 -> NHWC T with pitch/coff from `out`
template <typename T>
int launch_input_to_nhwc(const void* in, int in_dtype, const View& out, int B, cudaStream_t s, int src_H = 0, int src_W = 0);
// read back a view as NCHW fp32 (debug)
template <typename T>
int launch_view_to_nchw_f32(const View& in, float* out, int B, cudaStream_t s);

// Head decode (Modules/Head.cs:204-223 + Block.cs:15-45 + Tal.cs:313-356):
// box logits (64 ch) + class logits (nc ch) [+ mask coeffs] of one level, NHWC ->
// rows [0,4) xywh*stride, [4,4+nc) sigmoid, [4+nc,..) coeffs of pred (B, Ctot, A), anchors
// [a0, a0 + H*W) of this level.
template <typename T>
int launch_decode_level(const View& box, const View& cls, const View* coef, int B, int nc, int nm,
                        int reg_max, float stride, int a0, int A, int Ctot, float* pred,
                        cudaStream_t s);
template <typename T>
int launch_pixel_shuffle2(const View& in, const View& out, int B, cudaStream_t s);
// C2PSA attention core: qkv (B,N,nh*(2kd+hd)) -> out (B,N,nh*hd) and the dense v copy for the pe conv
template <typename T>
int launch_attention(const View& qkv, const View& out, const View& vout, int B, int nh, int kd, int hd, float scale,
                     cudaStream_t s);
// Tiled attention forward for kd = 32, hd = 64 (every YOLOv11 size: num_heads = c / 64, key_dim = 32), shared by the
// inference engine (T = __half / float, q|k|v interleaved per head in the qkv conv output) and the training step (fp32,
// separate q, k, v): element strides describe where token t of head h of image b lives.
struct AttnIO {
  const void *q, *k, *v;          // element type T
  long long in_tok, in_img;       // element strides between tokens / images of q and k
  long long v_tok, v_img;         // the same for v
  long long q_head, k_head, v_head;  // element offset of head h: h * q_head etc. (already includes nothing else)
  void* out;                      // type T, (B, N, nh * 64)-like with out_tok / out_img / head offset h * 64
  long long out_tok, out_img;
  void* vout;                     // optional dense copy of v (same addressing as out), type T
  float *row_max, *row_sum;       // optional (B, nh, N) fp32 softmax statistics for the backward pass
};
template <typename T>
int launch_attention_tiled_32x64(const AttnIO& io, int B, int N, int nh, float scale, cudaStream_t s);
bool attention_tiled_32x64_fits(int N);
// proto (B,h,w,32) NHWC T -> (B,32,h,w) fp32
template <typename T>
int launch_proto_out(const View& in, float* out, int B, cudaStream_t s);

// ---- conv_tc.cu : tcgen05 implicit-GEMM conv (fp16 storage, fp32 accumulate in TMEM) ----
struct TcConvPlan;  // opaque: tensor maps + tiling for one conv layer
TcConvPlan* tc_conv_plan_create(const ConvParams& p, std::string* err);
void tc_conv_plan_destroy(TcConvPlan* plan);
std::string tc_conv_plan_describe(const TcConvPlan* plan);  // tiling summary (YB_DEBUG_PLANS)
// Cross-layer overlap (DESIGN 4.1 "layer chaining"): instead of waiting for the whole previous grid
// (griddepcontrol.wait) a conv may start a tile as soon as the images it reads are complete in its producer.
//   done_ctr   this launch's per-image completion counters (rows x N tiles stored), or nullptr
//   dep_ctr    the producer launch's counters, or nullptr = wait for the previous grid as a whole
//   dep_expect rows x N tiles the producer stores per image
struct TcChain {
  int* done_ctr = nullptr;
  const int* dep_ctr = nullptr;
  int dep_expect = 0;
};
int tc_conv_launch(const TcConvPlan* plan, int B, float* pred, int* tile_ctr, cudaStream_t s, const TcChain* chain = nullptr);
int tc_conv_rows_per_image(const TcConvPlan* plan);  // rows x N tiles one image contributes to done_ctr
bool tc_conv_supported(const ConvParams& p);
// stem: NCHW u8/f16/f32 input -> 3x3 s2 conv (Cin=3) + bias + SiLU -> NHWC fp16
int launch_stem_f16(const void* in, int in_dtype, int B, int H, int W, const __half* w16 /*[Cout][32]*/,
                    const float* bias, const View& out, cudaStream_t s, int src_H = 0, int src_W = 0);  // src_*: unpadded source size

// ---- nms.cu ----
int nms_launch(const float* pred, int B, int C, int A, int nc, float conf, float iou, int max_det,
               int max_nms, int max_wh, float* dets, int* counts, int* keep_idx, cudaStream_t s);
int detection_loss_launch(const float* boxes, const float* scores, int B, int nc, int reg_max, int H, int W,
                          const float* targets_host, int n_targets, int topk, float hyp_box, float hyp_cls, float hyp_dfl,
                          float* loss_items, float* grad_boxes, float* grad_scores, unsigned char* fg_out, int* gt_idx_out,
                          float* tscore_out, cudaStream_t s);
// counters: optional >= 64 zeroed words owned by ONE stream (the statistics kernels leave them zero); nullptr = allocate and
// clear per call
int bn_silu_train_forward(const float* z, long long M, int C, int pitch, const float* gamma, const float* beta, float eps,
                          float momentum, int act, float* running_mean, float* running_var, float* y, int ypitch,
                          float* save_mean, float* save_invstd, cudaStream_t s, unsigned* counters = nullptr);
int bn_silu_backward(const float* z, const float* dy, long long M, int C, int pitch, int dpitch, const float* gamma,
                     const float* beta, const float* save_mean, const float* save_invstd, int act, float* dz, int zpitch,
                     float* dgamma, float* dbeta, cudaStream_t s, unsigned* counters = nullptr);
int adamw_step(float* p, const float* g, float* m, float* v, long long n, int step, float lr, float b1, float b2, float eps,
               float wd, cudaStream_t s);
int conv_backward_data(const float* dz, const float* w, int N, int H, int W, int Cin, int Cout, int k, int stride, int pad,
                       float* dx, cudaStream_t s);
int conv_backward_weight(const float* x, const float* dz, int N, int H, int W, int Cin, int Cout, int k, int stride, int pad,
                         float* dw, cudaStream_t s);
int masks_launch(const float* proto, const float* dets, const int* counts, int B, int max_det, int nm,
                 int mh, int mw, int H, int W, uint8_t* masks, cudaStream_t s, int mask_cap = 0);  // mask_cap: masks per image (0 = max_det)

}  // namespace yb
