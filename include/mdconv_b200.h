/*
 * mdconv_b200.h -- C ABI of libmdconv_b200.so: deformable / modulated-deformable convolution
 * forward + backward (2D and 3D) for NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for the reference's extension module MDCONV_CUDA
 * (CHONSPQX/modulated-deform-conv).  The reference binds eight C++ functions taking at::Tensor
 * handles through pybind11; here the same eight entry points exist with plain device pointers,
 * sizes and a CUDA stream, so any host language can bind them (ctypes, cgo, JNI, N-API ...).
 * The Python module `MDCONV_CUDA.py` shipped in this repo is such a binding and keeps the
 * reference's positional signatures.
 *
 * Reference interface each entry replaces (paths under the reference repo):
 *   deform_conv2d_forward_cuda              src/deformable_conv.cu:117-123
 *   deform_conv2d_backward_cuda             src/deformable_conv.cu:327-333
 *   modulated_deform_conv2d_forward_cuda    src/mdeformable_conv.cu:120-126
 *   modulated_deform_conv2d_backward_cuda   src/mdeformable_conv.cu:361-366
 *   deform_conv3d_forward_cuda              src/deformable_conv3d.cu:160-167
 *   deform_conv3d_backward_cuda             src/deformable_conv3d.cu:434-442
 *   modulated_deform_conv3d_forward_cuda    src/mdeformable_conv3d.cu:170-177
 *   modulated_deform_conv3d_backward_cuda   src/mdeformable_conv3d.cu:443-451
 *   (pybind registration: src/mdeformable_conv.cu:460-465, src/warp.cpp:1-19)
 *
 * Conventions
 *   - All tensor pointers are DEVICE pointers to dense, contiguous tensors in the reference's
 *     layouts:  input [B,C,H,W(,L)], weight [Co,C/g,kh,kw(,kl)], bias [Co] (NULL when
 *     with_bias == 0), offset [B, dg*nd*K, Ho,Wo(,Lo)] (per deformable group, per tap: h,w(,l)),
 *     mask [B, dg*K, Ho,Wo(,Lo)], output / grad_output [B,Co,Ho,Wo(,Lo)]; every gradient has
 *     the shape of its primal.  K = kh*kw(*kl);  out = (in + 2*pad - (dil*(k-1)+1))/stride + 1.
 *   - All tensors of one call share one dtype (MDCONV_F32 / F16 / BF16 / F64).
 *   - Nothing is allocated inside: the caller passes a workspace of at least
 *     mdconv_workspace_bytes() bytes (device memory, 256-byte aligned).
 *   - Work is enqueued on `stream` (a cudaStream_t; NULL = default stream) of the CURRENT device
 *     and the call returns without synchronising.
 *   - Outputs are fully overwritten (gradients are zero-initialised by the callee).
 *   - `in_step` is accepted for signature compatibility and ignored: no column buffer exists,
 *     so results never depend on it (the reference chunks the batch by gcd(B, in_step),
 *     src/config.h:43-60).
 *   - Return value: 0 on success, otherwise an MDCONV_ERR_* code (mdconv_status_string()).
 *     Unlike the reference (which printf()s launch errors and returns 0), launch failures are
 *     reported.
 */
#ifndef MDCONV_B200_H_
#define MDCONV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDCONV_B200_ABI_VERSION 3

enum mdconv_dtype { MDCONV_F32 = 0, MDCONV_F16 = 1, MDCONV_BF16 = 2, MDCONV_F64 = 3 };

enum mdconv_status {
  MDCONV_OK = 0,
  MDCONV_ERR_BAD_GEOMETRY = 1,   /* nd, sizes, divisibility, empty output */
  MDCONV_ERR_NULL_POINTER = 2,
  MDCONV_ERR_WORKSPACE = 3,      /* workspace too small or misaligned */
  MDCONV_ERR_DTYPE = 4,
  MDCONV_ERR_CUDA = 5,           /* a CUDA runtime / launch error (see mdconv_last_cuda_error) */
  MDCONV_ERR_TOO_LARGE = 6       /* a tensor exceeds 2^31-1 elements */
};

/* Gradient rule at exactly-integer sampling coordinates (the reference's two behaviours):
 * A = modulated 2D (src/mdeformable_conv.cu:259-315), B = plain 2D and both 3D ops
 * (src/deformable_conv.cu:254-261, src/mdeformable_conv3d.cu:340-342). */
enum mdconv_grad_rule { MDCONV_RULE_A = 0, MDCONV_RULE_B = 1 };

/* Kernel family selection (mostly for tests and benchmarks). */
enum mdconv_path {
  MDCONV_PATH_AUTO = 0,     /* tcgen05 path when the shape/dtype qualifies, else generic */
  MDCONV_PATH_GENERIC = 1,  /* fused SIMT implicit-GEMM kernels, every dtype/shape */
  MDCONV_PATH_TENSOR = 2    /* tcgen05/TMEM path; MDCONV_ERR_BAD_GEOMETRY if it does not qualify */
};

/* Memory layout of the ACTIVATION tensors of a call (input, output, grad_output, grad_input).  Weight, bias, offset
 * and mask always keep the reference's layouts.  CHANNELS_LAST = [B, spatial..., C] (torch.channels_last /
 * channels_last_3d): the layout the kernels gather from, so no internal re-layout pass runs (ABI version 3). */
enum mdconv_layout { MDCONV_LAYOUT_NCHW = 0, MDCONV_LAYOUT_CHANNELS_LAST = 1 };

typedef struct mdconv_geom {
  int32_t nd;                 /* 2 or 3 */
  int32_t batch, c_in, c_out;
  int32_t in_size[3];         /* H, W, L  (L ignored for nd == 2) */
  int32_t kernel[3];
  int32_t stride[3];
  int32_t pad[3];
  int32_t dilation[3];
  int32_t groups;
  int32_t deformable_groups;
  int32_t dtype;              /* enum mdconv_dtype */
  int32_t modulated;          /* 1: mask tensor present */
  int32_t grad_rule;          /* enum mdconv_grad_rule */
  int32_t with_bias;
  int32_t path;               /* enum mdconv_path */
  int32_t io_layout;          /* enum mdconv_layout (ABI version 3) */
} mdconv_geom_t;

int mdconv_abi_version(void);
const char* mdconv_status_string(int status);
/* Text of the last CUDA error seen by this thread ("" if none). */
const char* mdconv_last_cuda_error(void);

/* out[0..nd) = output spatial size.  Returns MDCONV_ERR_BAD_GEOMETRY if any is < 1. */
int mdconv_out_size(const mdconv_geom_t* g, int32_t out[3]);

/* Bytes of device workspace mdconv_forward / mdconv_backward need for this geometry. */
size_t mdconv_workspace_bytes(const mdconv_geom_t* g, int backward);

/* Which path AUTO would take for this geometry: MDCONV_PATH_GENERIC or MDCONV_PATH_TENSOR. */
int mdconv_selected_path(const mdconv_geom_t* g, int backward);

/* Kernel launches enqueued by the most recent forward/backward call, and by all calls of this
 * process so far (process-wide counters: frameworks run backward on their own threads). */
int mdconv_last_launch_count(void);
long long mdconv_total_launch_count(void);

/* Per-kernel timing for benchmarks: while enabled, every kernel this library launches is bracketed
 * by CUDA events on the launching stream.  mdconv_profile_read() waits for them, writes one
 * "kernel_name=milliseconds\n" line per launch (launch order) into buf, clears the list and returns
 * the number of entries.  Process-wide; off by default. */
int mdconv_profile_enable(int on);
int mdconv_profile_read(char* buf, size_t cap);

/* The backward builds its inverse index on an internal side stream (forked from / joined to the
 * caller's stream with events) so it overlaps the layout passes.  on = 0 keeps every kernel on the
 * caller's stream (use it with mdconv_profile_enable: a forked kernel's events also count the time it
 * waits for free SMs).  Default on; MDCONV_B200_NO_SIDE_STREAM=1 in the environment also disables it. */
int mdconv_side_stream_enable(int on);

int mdconv_forward(const mdconv_geom_t* g, const void* input, const void* weight, const void* bias,
                   const void* offset, const void* mask, void* output, void* workspace,
                   size_t workspace_bytes, void* stream);

/* Optional extras of a forward call (ABI version 3); every field may be 0 / NULL.
 *   input_channels_last: a buffer mdconv_to_channels_last() filled from THIS input (mdconv_channels_last_bytes()
 *     bytes).  The forward then skips its own re-layout pass, so several operators over one input -- a Pack module's
 *     offset/mask producer convolution and its deformable convolution (reference modulated_deform_conv.py:781-785)
 *     -- share one.  If it is the head of `workspace` itself, the saved-layout hand-off to the backward still works;
 *     if it lives elsewhere the workspace only has to hold the re-ordered weight (c_out * (c_in/groups) * K * 2
 *     bytes rounded up to 256).
 *   regular_grid = 1: there is no offset tensor; every tap samples its integer grid position, i.e. a plain
 *     convolution with this geometry's kernel / stride / pad / dilation (modulated must be 0, deformable_groups 1,
 *     offset and mask NULL).  This is the producer of the Pack modules, run on the same kernel as the deformable
 *     forward (fp16 / bf16, c_in a multiple of 64).  c_out must be a multiple of 32: pad the filter bank with zero
 *     rows and say how many channels are real in c_out_valid; channels [split, c_out_valid) go to output2
 *     [B, c_out_valid - split, P] and channels [0, split) to output [B, split, P] (output2 NULL: one tensor
 *     [B, c_out_valid, P]). */
typedef struct mdconv_fwd_extra {
  const void* input_channels_last;
  int32_t regular_grid;
  int32_t c_out_valid;
  void* output2;
  int32_t split;
} mdconv_fwd_extra_t;

int mdconv_forward_ex(const mdconv_geom_t* g, const void* input, const void* weight, const void* bias,
                      const void* offset, const void* mask, void* output, void* workspace,
                      size_t workspace_bytes, const mdconv_fwd_extra_t* extra, void* stream);
/* Bytes of the channels-last copy (0 when the tcgen05 path does not take this geometry), and the pass itself. */
size_t mdconv_channels_last_bytes(const mdconv_geom_t* g);
int mdconv_to_channels_last(const mdconv_geom_t* g, const void* input, void* channels_last, void* stream);

/* Any of grad_bias / grad_mask may be NULL (with_bias == 0 / modulated == 0). */
int mdconv_backward(const mdconv_geom_t* g, const void* input, const void* weight, const void* offset,
                    const void* mask, const void* grad_output, void* grad_input, void* grad_weight,
                    void* grad_bias, void* grad_offset, void* grad_mask, void* workspace,
                    size_t workspace_bytes, void* stream);

/* Forward -> backward hand-off of the internal layouts (ABI version 2).  The reference recomputes everything
 * in its backward (it re-runs the im2col gather, src/mdeformable_conv.cu:421-433); here the tcgen05 path's
 * forward leaves the channels-last input and the re-ordered weight in the first mdconv_saved_bytes() bytes of
 * ITS workspace.  A caller that keeps that workspace untouched until the matching backward (same input, weight
 * and geometry) passes it as `saved` and the backward skips both layout passes.  mdconv_saved_bytes() is 0
 * when either direction would not run on the tensor path; saved == NULL is mdconv_backward().
 * weight_ready_event: NULL, or a cudaEvent_t the library records on `stream` as soon as grad_weight and grad_bias
 * are final -- the tensor path computes them FIRST, so a data-parallel caller can run its all-reduce of the two
 * parameter gradients (the only exchange of this path) on another stream while grad_offset / grad_mask /
 * grad_input are still being computed. */
size_t mdconv_saved_bytes(const mdconv_geom_t* g);
int mdconv_backward_saved(const mdconv_geom_t* g, const void* input, const void* weight, const void* offset,
                          const void* mask, const void* grad_output, void* grad_input, void* grad_weight,
                          void* grad_bias, void* grad_offset, void* grad_mask, void* workspace,
                          size_t workspace_bytes, const void* saved, size_t saved_bytes, void* weight_ready_event,
                          void* stream);

/* Optional extras of a backward call (ABI version 3); every field may be 0 / NULL.
 *   saved, saved_bytes, weight_ready_event: as in mdconv_backward_saved.
 *   grad_weight_f32 / grad_bias_f32: device buffers of c_out*(c_in/groups)*K resp. c_out floats that receive the
 *     parameter gradients BEFORE they are rounded to the storage dtype (fp16 / bf16 calls; fp32 calls get a copy).
 *     The parameter gradients are the only quantities of this path that sum over samples (reference:
 *     src/mdeformable_conv.cu:436-444), i.e. the only ones a batch-sharded caller exchanges: all-reducing the fp32
 *     values and rounding once afterwards gives the single-GPU result instead of a sum of rounded partial sums. */
typedef struct mdconv_bwd_extra {
  const void* saved;
  size_t saved_bytes;
  void* weight_ready_event;
  float* grad_weight_f32;
  float* grad_bias_f32;
} mdconv_bwd_extra_t;

int mdconv_backward_ex(const mdconv_geom_t* g, const void* input, const void* weight, const void* offset,
                       const void* mask, const void* grad_output, void* grad_input, void* grad_weight,
                       void* grad_bias, void* grad_offset, void* grad_mask, void* workspace, size_t workspace_bytes,
                       const mdconv_bwd_extra_t* extra, void* stream);

/* Streaming multiprocessors the persistent data-gradient kernels leave free once weight_ready_event has been
 * recorded, so the caller's collective kernel has somewhere to run (default 8, or MDCONV_B200_COMM_SMS; 0..64).
 * Match it to the CTA count of the communicator (e.g. NCCL max_ctas).  Returns the previous value. */
int mdconv_set_comm_sms(int sms);

/* fp32 calls whose channel counts qualify run their contractions on the tensor cores with every operand split into
 * two bf16 halves (hi*hi + hi*lo + lo*hi, fp32 accumulation: ~2^-17 relative per product, results within the 1e-4
 * bar of the exact-fp32 kernels).  on = 0 sends fp32 back to the exact SIMT kernels; MDCONV_B200_FP32_TC=0 in the
 * environment does the same.  Returns the previous setting. */
int mdconv_set_fp32_tensor_cores(int on);

/* grad_input of 16-bit calls is a col2im GATHER over grad_columns ([samples][K][P][C] in the storage dtype -- the
 * reference's own buffer, src/mdeformable_conv.cu:396-419) that the grad_offset / grad_mask kernel streams out on its
 * way; the buffer is part of the backward workspace mdconv_workspace_bytes() reports.  `megabytes` (default 1024, or
 * MDCONV_B200_GCOL_MAX_MB) caps it: a batch whose grad_columns are larger is walked in chunks of whole samples that
 * fit (slower: more, smaller launches).  0 -- and fp32 calls -- run grad_input as a third contraction over the inverse
 * index instead, which needs no such buffer.  Returns the previous cap (-1: default not resolved yet). */
long long mdconv_set_gcol_budget_mb(long long megabytes);

/* ---- The reference's eight entry points, C ABI form --------------------------------------
 * Argument order follows the reference signatures; tensor handles become device pointers and the
 * sizes the reference read from the tensors are passed explicitly (batch, channels, spatial
 * sizes, channels_out).  Trailing (dtype, workspace, workspace_bytes, stream) are new. */

int deform_conv2d_forward_cuda(
    const void* input, const void* weight, const void* bias, const void* offset, void* output,
    int batch, int channels, int height, int width, int channels_out,
    int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
    int dilation_h, int dilation_w, int group, int deformable_group, int in_step, int with_bias,
    int dtype, void* workspace, size_t workspace_bytes, void* stream);

int deform_conv2d_backward_cuda(
    const void* input, const void* weight, const void* bias, const void* offset,
    void* grad_input, void* grad_weight, void* grad_bias, void* grad_offset, const void* grad_output,
    int batch, int channels, int height, int width, int channels_out,
    int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
    int dilation_h, int dilation_w, int group, int deformable_group, int in_step, int with_bias,
    int dtype, void* workspace, size_t workspace_bytes, void* stream);

int modulated_deform_conv2d_forward_cuda(
    const void* input, const void* weight, const void* bias, const void* offset, const void* mask,
    void* output,
    int batch, int channels, int height, int width, int channels_out,
    int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
    int dilation_h, int dilation_w, int group, int deformable_group, int in_step, int with_bias,
    int dtype, void* workspace, size_t workspace_bytes, void* stream);

int modulated_deform_conv2d_backward_cuda(
    const void* input, const void* weight, const void* bias, const void* offset, const void* mask,
    const void* grad_output,
    void* grad_input, void* grad_offset, void* grad_mask, void* grad_weight, void* grad_bias,
    int batch, int channels, int height, int width, int channels_out,
    int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
    int dilation_h, int dilation_w, int group, int deformable_group, int in_step, int with_bias,
    int dtype, void* workspace, size_t workspace_bytes, void* stream);

int deform_conv3d_forward_cuda(
    const void* input, const void* weight, const void* bias, const void* offset, void* output,
    int batch, int channels, int height, int width, int length, int channels_out,
    int kernel_h, int kernel_w, int kernel_l, int stride_h, int stride_w, int stride_l,
    int pad_h, int pad_w, int pad_l, int dilation_h, int dilation_w, int dilation_l,
    int group, int deformable_group, int in_step, int with_bias,
    int dtype, void* workspace, size_t workspace_bytes, void* stream);

int deform_conv3d_backward_cuda(
    const void* input, const void* weight, const void* bias, const void* offset,
    void* grad_input, void* grad_weight, void* grad_bias, void* grad_offset, const void* grad_output,
    int batch, int channels, int height, int width, int length, int channels_out,
    int kernel_h, int kernel_w, int kernel_l, int stride_h, int stride_w, int stride_l,
    int pad_h, int pad_w, int pad_l, int dilation_h, int dilation_w, int dilation_l,
    int group, int deformable_group, int in_step, int with_bias,
    int dtype, void* workspace, size_t workspace_bytes, void* stream);

int modulated_deform_conv3d_forward_cuda(
    const void* input, const void* weight, const void* bias, const void* offset, const void* mask,
    void* output,
    int batch, int channels, int height, int width, int length, int channels_out,
    int kernel_h, int kernel_w, int kernel_l, int stride_h, int stride_w, int stride_l,
    int pad_h, int pad_w, int pad_l, int dilation_h, int dilation_w, int dilation_l,
    int group, int deformable_group, int in_step, int with_bias,
    int dtype, void* workspace, size_t workspace_bytes, void* stream);

int modulated_deform_conv3d_backward_cuda(
    const void* input, const void* weight, const void* bias, const void* offset, const void* mask,
    void* grad_input, void* grad_weight, void* grad_bias, void* grad_offset, void* grad_mask,
    const void* grad_output,
    int batch, int channels, int height, int width, int length, int channels_out,
    int kernel_h, int kernel_w, int kernel_l, int stride_h, int stride_w, int stride_l,
    int pad_h, int pad_w, int pad_l, int dilation_h, int dilation_w, int dilation_l,
    int group, int deformable_group, int in_step, int with_bias,
    int dtype, void* workspace, size_t workspace_bytes, void* stream);

/* Workspace size for the legacy-named entry points (same meaning as mdconv_workspace_bytes). */
size_t mdconv_legacy_workspace_bytes(int nd, int modulated, int backward, int batch, int channels,
                                     int height, int width, int length, int channels_out,
                                     int kernel_h, int kernel_w, int kernel_l,
                                     int stride_h, int stride_w, int stride_l,
                                     int pad_h, int pad_w, int pad_l,
                                     int dilation_h, int dilation_w, int dilation_l,
                                     int group, int deformable_group, int with_bias, int dtype);

#ifdef __cplusplus
}
#endif
#endif /* MDCONV_B200_H_ */
