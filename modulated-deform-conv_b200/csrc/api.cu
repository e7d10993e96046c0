// api.cu -- C ABI of libmdconv_b200.so (see include/mdconv_b200.h): validation, geometry
// resolution, path selection and the reference's eight entry-point names.
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "tensor_path.cuh"
#include "tc_common.cuh"

using namespace mdconv;

namespace {

thread_local char g_cuda_err[256] = "";
std::atomic<int> g_last_launches{0};   // process-wide: autograd runs backward on its own thread
std::atomic<long long> g_total_launches{0};

struct ProfEntry {
  const char* name;
  cudaEvent_t a, b;
};
std::atomic<bool> g_prof_on{false};
std::mutex g_prof_mu;
std::vector<ProfEntry>* g_prof = nullptr;

int resolve(const mdconv_geom_t* in, DevGeom& g) {
  if (in == nullptr) return MDCONV_ERR_NULL_POINTER;
  if (in->nd != 2 && in->nd != 3) return MDCONV_ERR_BAD_GEOMETRY;
  if (in->dtype < MDCONV_F32 || in->dtype > MDCONV_F64) return MDCONV_ERR_DTYPE;
  memset(&g, 0, sizeof(g));
  g.nd = in->nd; g.B = in->batch; g.C = in->c_in; g.Co = in->c_out;
  g.g = in->groups; g.dg = in->deformable_groups;
  if (g.B < 1 || g.C < 1 || g.Co < 1 || g.g < 1 || g.dg < 1) return MDCONV_ERR_BAD_GEOMETRY;
  if (g.C % g.g || g.Co % g.g || g.C % g.dg) return MDCONV_ERR_BAD_GEOMETRY;
  g.K = 1; g.P = 1; g.IN = 1;
  long long K = 1, P = 1, IN = 1;
  for (int a = 0; a < 3; ++a) {
    if (a < g.nd) {
      g.in[a] = in->in_size[a]; g.k[a] = in->kernel[a]; g.s[a] = in->stride[a];
      g.p[a] = in->pad[a]; g.d[a] = in->dilation[a];
      if (g.in[a] < 1 || g.k[a] < 1 || g.s[a] < 1 || g.d[a] < 1 || g.p[a] < 0) return MDCONV_ERR_BAD_GEOMETRY;
    } else {
      g.in[a] = 1; g.k[a] = 1; g.s[a] = 1; g.p[a] = 0; g.d[a] = 1;
    }
    // reference: src/mdeformable_conv.cu:150-153
    const int span = g.in[a] + 2 * g.p[a] - (g.d[a] * (g.k[a] - 1) + 1);
    if (span < 0) return MDCONV_ERR_BAD_GEOMETRY;
    g.out[a] = span / g.s[a] + 1;
    K *= g.k[a]; P *= g.out[a]; IN *= g.in[a];
  }
  g.Cg = g.C / g.g; g.Cog = g.Co / g.g; g.Cdg = g.C / g.dg;
  const long long lim = 2147483647LL;
  const long long Kg = (long long)g.Cg * K;
  if (K > 4096 || P > lim || IN > lim || Kg > lim) return MDCONV_ERR_TOO_LARGE;
  if ((long long)g.B * g.C * IN > lim || (long long)g.B * g.Co * P > lim ||
      (long long)g.B * g.dg * g.nd * K * P > lim || (long long)g.Co * Kg > lim)
    return MDCONV_ERR_TOO_LARGE;
  g.K = (int)K; g.P = (int)P; g.IN = (int)IN; g.Kg = (int)Kg;
  g.modulated = in->modulated ? 1 : 0;
  g.rule = in->grad_rule ? 1 : 0;
  g.with_bias = in->with_bias ? 1 : 0;
  if (in->io_layout != MDCONV_LAYOUT_NCHW && in->io_layout != MDCONV_LAYOUT_CHANNELS_LAST)
    return MDCONV_ERR_BAD_GEOMETRY;
  g.cl = in->io_layout == MDCONV_LAYOUT_CHANNELS_LAST ? 1 : 0;
  g.esize = in->dtype == MDCONV_F32 ? 4 : (in->dtype == MDCONV_F64 ? 8 : 2);
  for (int a = 0; a < 3; ++a) {
    g.fd_out[a] = make_fastdiv(g.out[a]);
    g.fd_k[a] = make_fastdiv(g.k[a]);
  }
  g.fd_K = make_fastdiv(g.K);
  g.fd_Cdg = make_fastdiv(g.Cdg);
  return MDCONV_OK;
}

int pick_path(const mdconv_geom_t* in, const DevGeom& g, int backward) {
  const bool can = tensor_path_supported(g, in->dtype, backward);
  if (g.cl) return (can && in->path != MDCONV_PATH_GENERIC) ? MDCONV_PATH_TENSOR : -1;   // tensor kernels only
  if (in->path == MDCONV_PATH_GENERIC) return MDCONV_PATH_GENERIC;
  if (in->path == MDCONV_PATH_TENSOR) return can ? MDCONV_PATH_TENSOR : -1;
  return can ? MDCONV_PATH_TENSOR : MDCONV_PATH_GENERIC;
}

int finish(LaunchCtx& ctx) {
  g_last_launches = ctx.launches;
  g_total_launches += ctx.launches;
  if (ctx.err != cudaSuccess) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", cudaGetErrorName(ctx.err), cudaGetErrorString(ctx.err));
    return MDCONV_ERR_CUDA;
  }
  return MDCONV_OK;
}

}  // namespace

namespace mdconv {
void prof_begin(LaunchCtx& ctx, const char* name) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof) g_prof = new std::vector<ProfEntry>();
  ProfEntry e{name, nullptr, nullptr};
  cudaEventCreate(&e.a);
  cudaEventCreate(&e.b);
  cudaEventRecord(e.a, ctx.stream);
  g_prof->push_back(e);
}
void prof_end(LaunchCtx& ctx) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof || g_prof->empty()) return;
  cudaEventRecord(g_prof->back().b, ctx.stream);
}
}  // namespace mdconv

extern "C" {

int mdconv_profile_enable(int on) {
  g_prof_on = on != 0;
  return MDCONV_OK;
}

// Waits for the recorded events, writes "name=milliseconds\n" per timed kernel (in launch order)
// into buf and clears the list.  Returns the number of entries.
int mdconv_profile_read(char* buf, size_t cap) {
  int n = 0;
  std::string out;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_prof) {
    cudaEvent_t first = g_prof->empty() ? nullptr : g_prof->front().a;
    for (ProfEntry& e : *g_prof) {
      float ms = -1.f, at = 0.f;
      if (cudaEventSynchronize(e.b) == cudaSuccess) cudaEventElapsedTime(&ms, e.a, e.b);
      if (first != nullptr && e.a != first) cudaEventElapsedTime(&at, first, e.a);   // start, relative to the first launch
      char line[160];
      snprintf(line, sizeof(line), "%s=%.6f@%.6f\n", e.name, ms, at);
      out += line;
      ++n;
    }
    for (ProfEntry& e : *g_prof) {   // only now: `first` is the first entry's start event
      cudaEventDestroy(e.a);
      cudaEventDestroy(e.b);
    }
    g_prof->clear();
  }
  if (buf && cap) {
    snprintf(buf, cap, "%s", out.c_str());
  }
  return n;
}

int mdconv_side_stream_enable(int on) {
  mdconv::tc::set_side_stream_enabled(on != 0);
  return MDCONV_OK;
}

int mdconv_abi_version(void) { return MDCONV_B200_ABI_VERSION; }

const char* mdconv_status_string(int status) {
  switch (status) {
    case MDCONV_OK: return "ok";
    case MDCONV_ERR_BAD_GEOMETRY: return "bad geometry (nd/sizes/divisibility/empty output, or path not applicable)";
    case MDCONV_ERR_NULL_POINTER: return "null pointer";
    case MDCONV_ERR_WORKSPACE: return "workspace too small or not 256-byte aligned";
    case MDCONV_ERR_DTYPE: return "unsupported dtype";
    case MDCONV_ERR_CUDA: return "CUDA error";
    case MDCONV_ERR_TOO_LARGE: return "tensor exceeds 2^31-1 elements";
    default: return "unknown status";
  }
}

const char* mdconv_last_cuda_error(void) { return g_cuda_err; }
int mdconv_last_launch_count(void) { return g_last_launches; }
long long mdconv_total_launch_count(void) { return g_total_launches; }

int mdconv_out_size(const mdconv_geom_t* gi, int32_t out[3]) {
  DevGeom g;
  const int rc = resolve(gi, g);
  if (rc != MDCONV_OK) return rc;
  if (out == nullptr) return MDCONV_ERR_NULL_POINTER;
  out[0] = g.out[0]; out[1] = g.out[1]; out[2] = g.out[2];
  return MDCONV_OK;
}

int mdconv_selected_path(const mdconv_geom_t* gi, int backward) {
  DevGeom g;
  if (resolve(gi, g) != MDCONV_OK) return -1;
  return pick_path(gi, g, backward);
}

size_t mdconv_workspace_bytes(const mdconv_geom_t* gi, int backward) {
  DevGeom g;
  if (resolve(gi, g) != MDCONV_OK) return 0;
  const int path = pick_path(gi, g, backward);
  if (path == MDCONV_PATH_TENSOR) return tensor_workspace_bytes(g, gi->dtype, backward);
  return generic_workspace_bytes(g, gi->dtype, backward);
}

int mdconv_forward(const mdconv_geom_t* gi, const void* input, const void* weight, const void* bias,
                   const void* offset, const void* mask, void* output, void* workspace, size_t workspace_bytes,
                   void* stream) {
  return mdconv_forward_ex(gi, input, weight, bias, offset, mask, output, workspace, workspace_bytes, nullptr, stream);
}

int mdconv_forward_ex(const mdconv_geom_t* gi, const void* input, const void* weight, const void* bias,
                      const void* offset, const void* mask, void* output, void* workspace, size_t workspace_bytes,
                      const mdconv_fwd_extra_t* extra, void* stream) {
  DevGeom g;
  int rc = resolve(gi, g);
  if (rc != MDCONV_OK) return rc;
  FwdExtra ex;
  if (extra != nullptr) {
    ex.xt = extra->input_channels_last;
    ex.grid = extra->regular_grid ? 1 : 0;
    ex.co_valid = extra->c_out_valid;
    ex.y2 = extra->output2;
    ex.split = extra->split;
  }
  if (ex.grid) {
    if (g.modulated || g.dg != 1 || g.cl || offset != nullptr || mask != nullptr) return MDCONV_ERR_BAD_GEOMETRY;
    if (ex.co_valid < 0 || ex.co_valid > g.Co || (ex.y2 != nullptr && (ex.split < 1 || ex.split >= (ex.co_valid ? ex.co_valid : g.Co))))
      return MDCONV_ERR_BAD_GEOMETRY;
    if (g.g != 1 && ex.co_valid != 0 && ex.co_valid != g.Co) return MDCONV_ERR_BAD_GEOMETRY;
  } else if (ex.y2 != nullptr || ex.co_valid != 0) {
    return MDCONV_ERR_BAD_GEOMETRY;
  }
  if (!input || !weight || !output || (!offset && !ex.grid)) return MDCONV_ERR_NULL_POINTER;
  if (g.modulated && !mask) return MDCONV_ERR_NULL_POINTER;
  if (g.with_bias && !bias) return MDCONV_ERR_NULL_POINTER;
  const int path = pick_path(gi, g, 0);
  if (path < 0) return MDCONV_ERR_BAD_GEOMETRY;
  if ((ex.grid || ex.xt != nullptr) && path != MDCONV_PATH_TENSOR) return MDCONV_ERR_BAD_GEOMETRY;   // tcgen05 kernels only
  if (ex.xt != nullptr && g.cl) return MDCONV_ERR_BAD_GEOMETRY;
  size_t need = path == MDCONV_PATH_TENSOR ? tensor_workspace_bytes(g, gi->dtype, 0)
                                           : generic_workspace_bytes(g, gi->dtype, 0);
  if (ex.xt != nullptr && ex.xt != workspace) need = tensor_weight_layout_bytes(g);   // only the re-ordered weight
  if (need > 0 && (workspace == nullptr || workspace_bytes < need || ((uintptr_t)workspace & 255)))
    return MDCONV_ERR_WORKSPACE;
  LaunchCtx ctx{(cudaStream_t)stream, 0, cudaSuccess};
  (void)cudaGetLastError();
  if (path == MDCONV_PATH_TENSOR)
    tensor_forward(g, gi->dtype, input, weight, bias, offset, mask, output, workspace, ex, ctx);
  else
    generic_forward(g, gi->dtype, input, weight, bias, offset, mask, output, ctx);
  return finish(ctx);
}

size_t mdconv_channels_last_bytes(const mdconv_geom_t* gi) {
  DevGeom g;
  if (resolve(gi, g) != MDCONV_OK || g.cl) return 0;
  if (pick_path(gi, g, 0) != MDCONV_PATH_TENSOR) return 0;
  return tensor_channels_last_bytes(g);
}

int mdconv_to_channels_last(const mdconv_geom_t* gi, const void* input, void* channels_last, void* stream) {
  DevGeom g;
  int rc = resolve(gi, g);
  if (rc != MDCONV_OK) return rc;
  if (!input || !channels_last) return MDCONV_ERR_NULL_POINTER;
  if (g.cl || pick_path(gi, g, 0) != MDCONV_PATH_TENSOR || ((uintptr_t)channels_last & 255)) return MDCONV_ERR_BAD_GEOMETRY;
  LaunchCtx ctx{(cudaStream_t)stream, 0, cudaSuccess};
  (void)cudaGetLastError();
  tensor_to_channels_last(g, gi->dtype, input, channels_last, ctx);
  return finish(ctx);
}

size_t mdconv_saved_bytes(const mdconv_geom_t* gi) {
  DevGeom g;
  if (resolve(gi, g) != MDCONV_OK) return 0;
  if (pick_path(gi, g, 0) != MDCONV_PATH_TENSOR || pick_path(gi, g, 1) != MDCONV_PATH_TENSOR) return 0;
  return tensor_saved_bytes(g, gi->dtype);
}

int mdconv_backward(const mdconv_geom_t* gi, const void* input, const void* weight, const void* offset,
                    const void* mask, const void* grad_output, void* grad_input, void* grad_weight, void* grad_bias,
                    void* grad_offset, void* grad_mask, void* workspace, size_t workspace_bytes, void* stream) {
  return mdconv_backward_saved(gi, input, weight, offset, mask, grad_output, grad_input, grad_weight, grad_bias,
                               grad_offset, grad_mask, workspace, workspace_bytes, nullptr, 0, nullptr, stream);
}

int mdconv_backward_saved(const mdconv_geom_t* gi, const void* input, const void* weight, const void* offset,
                          const void* mask, const void* grad_output, void* grad_input, void* grad_weight,
                          void* grad_bias, void* grad_offset, void* grad_mask, void* workspace, size_t workspace_bytes,
                          const void* saved, size_t saved_bytes, void* weight_ready_event, void* stream) {
  mdconv_bwd_extra_t ex;
  memset(&ex, 0, sizeof(ex));
  ex.saved = saved; ex.saved_bytes = saved_bytes; ex.weight_ready_event = weight_ready_event;
  return mdconv_backward_ex(gi, input, weight, offset, mask, grad_output, grad_input, grad_weight, grad_bias,
                            grad_offset, grad_mask, workspace, workspace_bytes, &ex, stream);
}

int mdconv_backward_ex(const mdconv_geom_t* gi, const void* input, const void* weight, const void* offset,
                       const void* mask, const void* grad_output, void* grad_input, void* grad_weight,
                       void* grad_bias, void* grad_offset, void* grad_mask, void* workspace, size_t workspace_bytes,
                       const mdconv_bwd_extra_t* extra, void* stream) {
  DevGeom g;
  int rc = resolve(gi, g);
  if (rc != MDCONV_OK) return rc;
  if (!input || !weight || !offset || !grad_output || !grad_input || !grad_weight || !grad_offset)
    return MDCONV_ERR_NULL_POINTER;
  if (g.modulated && (!mask || !grad_mask)) return MDCONV_ERR_NULL_POINTER;
  if (g.with_bias && !grad_bias) return MDCONV_ERR_NULL_POINTER;
  const int path = pick_path(gi, g, 1);
  if (path < 0) return MDCONV_ERR_BAD_GEOMETRY;
  const size_t need = path == MDCONV_PATH_TENSOR ? tensor_workspace_bytes(g, gi->dtype, 1)
                                                 : generic_workspace_bytes(g, gi->dtype, 1);
  if (need > 0 && (workspace == nullptr || workspace_bytes < need || ((uintptr_t)workspace & 255)))
    return MDCONV_ERR_WORKSPACE;
  LaunchCtx ctx{(cudaStream_t)stream, 0, cudaSuccess};
  const void* saved = nullptr;
  if (extra != nullptr) {
    ctx.weight_ready = (cudaEvent_t)extra->weight_ready_event;
    if (gi->dtype != MDCONV_F64) {   // fp64 gradients are already wider than the fp32 copy
      ctx.gw32 = extra->grad_weight_f32;
      ctx.gb32 = g.with_bias ? extra->grad_bias_f32 : nullptr;
    }
    saved = extra->saved;
    if (saved != nullptr) {
      // only meaningful when both directions run on the tensor path; anything else is a caller error
      const size_t want = mdconv_saved_bytes(gi);
      if (want == 0 || extra->saved_bytes < want || ((uintptr_t)saved & 255)) return MDCONV_ERR_WORKSPACE;
    }
  }
  (void)cudaGetLastError();
  if (path == MDCONV_PATH_TENSOR)
    tensor_backward(g, gi->dtype, input, weight, offset, mask, grad_output, grad_input, grad_weight, grad_bias,
                    grad_offset, grad_mask, workspace, saved, ctx);
  else
    generic_backward(g, gi->dtype, input, weight, offset, mask, grad_output, grad_input, grad_weight, grad_bias,
                     grad_offset, grad_mask, workspace, ctx);
  signal_weight_ready(ctx);   // paths that did not signal earlier: everything is enqueued by now
  return finish(ctx);
}

int mdconv_set_comm_sms(int sms) { return mdconv::set_comm_spare_sms(sms); }
int mdconv_set_fp32_tensor_cores(int on) { return mdconv::set_fp32_on_tensor_cores(on); }
long long mdconv_set_gcol_budget_mb(long long megabytes) { return mdconv::tc::set_gcol_budget_mb(megabytes); }

// ---- legacy-named entry points -----------------------------------------------------------------
static mdconv_geom_t legacy_geom(int nd, int modulated, int batch, int channels, int height, int width, int length,
                                 int channels_out, int kh, int kw, int kl, int sh, int sw, int sl, int ph, int pw,
                                 int pl, int dh, int dw, int dl, int group, int dgroup, int with_bias, int dtype) {
  mdconv_geom_t g;
  memset(&g, 0, sizeof(g));
  g.nd = nd; g.batch = batch; g.c_in = channels; g.c_out = channels_out;
  g.in_size[0] = height; g.in_size[1] = width; g.in_size[2] = length;
  g.kernel[0] = kh; g.kernel[1] = kw; g.kernel[2] = kl;
  g.stride[0] = sh; g.stride[1] = sw; g.stride[2] = sl;
  g.pad[0] = ph; g.pad[1] = pw; g.pad[2] = pl;
  g.dilation[0] = dh; g.dilation[1] = dw; g.dilation[2] = dl;
  g.groups = group; g.deformable_groups = dgroup;
  g.dtype = dtype; g.modulated = modulated;
  // rule A only for the modulated 2D op (reference src/mdeformable_conv.cu:259-315)
  g.grad_rule = (nd == 2 && modulated) ? MDCONV_RULE_A : MDCONV_RULE_B;
  g.with_bias = with_bias ? 1 : 0;
  g.path = MDCONV_PATH_AUTO;
  return g;
}

size_t mdconv_legacy_workspace_bytes(int nd, int modulated, int backward, int batch, int channels, int height,
                                     int width, int length, int channels_out, int kh, int kw, int kl, int sh,
                                     int sw, int sl, int ph, int pw, int pl, int dh, int dw, int dl, int group,
                                     int dgroup, int with_bias, int dtype) {
  mdconv_geom_t g = legacy_geom(nd, modulated, batch, channels, height, width, length, channels_out, kh, kw, kl, sh,
                                sw, sl, ph, pw, pl, dh, dw, dl, group, dgroup, with_bias, dtype);
  return mdconv_workspace_bytes(&g, backward);
}

int deform_conv2d_forward_cuda(const void* input, const void* weight, const void* bias, const void* offset,
                               void* output, int batch, int channels, int height, int width, int channels_out,
                               int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                               int dilation_h, int dilation_w, int group, int deformable_group, int in_step,
                               int with_bias, int dtype, void* workspace, size_t workspace_bytes, void* stream) {
  (void)in_step;
  mdconv_geom_t g = legacy_geom(2, 0, batch, channels, height, width, 1, channels_out, kernel_h, kernel_w, 1,
                                stride_h, stride_w, 1, pad_h, pad_w, 0, dilation_h, dilation_w, 1, group,
                                deformable_group, with_bias, dtype);
  return mdconv_forward(&g, input, weight, with_bias ? bias : nullptr, offset, nullptr, output, workspace,
                        workspace_bytes, stream);
}

int deform_conv2d_backward_cuda(const void* input, const void* weight, const void* bias, const void* offset,
                                void* grad_input, void* grad_weight, void* grad_bias, void* grad_offset,
                                const void* grad_output, int batch, int channels, int height, int width,
                                int channels_out, int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h,
                                int pad_w, int dilation_h, int dilation_w, int group, int deformable_group,
                                int in_step, int with_bias, int dtype, void* workspace, size_t workspace_bytes,
                                void* stream) {
  (void)in_step; (void)bias;
  mdconv_geom_t g = legacy_geom(2, 0, batch, channels, height, width, 1, channels_out, kernel_h, kernel_w, 1,
                                stride_h, stride_w, 1, pad_h, pad_w, 0, dilation_h, dilation_w, 1, group,
                                deformable_group, with_bias, dtype);
  return mdconv_backward(&g, input, weight, offset, nullptr, grad_output, grad_input, grad_weight,
                         with_bias ? grad_bias : nullptr, grad_offset, nullptr, workspace, workspace_bytes, stream);
}

int modulated_deform_conv2d_forward_cuda(const void* input, const void* weight, const void* bias,
                                         const void* offset, const void* mask, void* output, int batch,
                                         int channels, int height, int width, int channels_out, int kernel_h,
                                         int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                                         int dilation_h, int dilation_w, int group, int deformable_group,
                                         int in_step, int with_bias, int dtype, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  (void)in_step;
  mdconv_geom_t g = legacy_geom(2, 1, batch, channels, height, width, 1, channels_out, kernel_h, kernel_w, 1,
                                stride_h, stride_w, 1, pad_h, pad_w, 0, dilation_h, dilation_w, 1, group,
                                deformable_group, with_bias, dtype);
  return mdconv_forward(&g, input, weight, with_bias ? bias : nullptr, offset, mask, output, workspace,
                        workspace_bytes, stream);
}

int modulated_deform_conv2d_backward_cuda(const void* input, const void* weight, const void* bias,
                                          const void* offset, const void* mask, const void* grad_output,
                                          void* grad_input, void* grad_offset, void* grad_mask, void* grad_weight,
                                          void* grad_bias, int batch, int channels, int height, int width,
                                          int channels_out, int kernel_h, int kernel_w, int stride_h, int stride_w,
                                          int pad_h, int pad_w, int dilation_h, int dilation_w, int group,
                                          int deformable_group, int in_step, int with_bias, int dtype,
                                          void* workspace, size_t workspace_bytes, void* stream) {
  (void)in_step; (void)bias;
  mdconv_geom_t g = legacy_geom(2, 1, batch, channels, height, width, 1, channels_out, kernel_h, kernel_w, 1,
                                stride_h, stride_w, 1, pad_h, pad_w, 0, dilation_h, dilation_w, 1, group,
                                deformable_group, with_bias, dtype);
  return mdconv_backward(&g, input, weight, offset, mask, grad_output, grad_input, grad_weight,
                         with_bias ? grad_bias : nullptr, grad_offset, grad_mask, workspace, workspace_bytes, stream);
}

int deform_conv3d_forward_cuda(const void* input, const void* weight, const void* bias, const void* offset,
                               void* output, int batch, int channels, int height, int width, int length,
                               int channels_out, int kernel_h, int kernel_w, int kernel_l, int stride_h,
                               int stride_w, int stride_l, int pad_h, int pad_w, int pad_l, int dilation_h,
                               int dilation_w, int dilation_l, int group, int deformable_group, int in_step,
                               int with_bias, int dtype, void* workspace, size_t workspace_bytes, void* stream) {
  (void)in_step;
  mdconv_geom_t g = legacy_geom(3, 0, batch, channels, height, width, length, channels_out, kernel_h, kernel_w,
                                kernel_l, stride_h, stride_w, stride_l, pad_h, pad_w, pad_l, dilation_h, dilation_w,
                                dilation_l, group, deformable_group, with_bias, dtype);
  return mdconv_forward(&g, input, weight, with_bias ? bias : nullptr, offset, nullptr, output, workspace,
                        workspace_bytes, stream);
}

int deform_conv3d_backward_cuda(const void* input, const void* weight, const void* bias, const void* offset,
                                void* grad_input, void* grad_weight, void* grad_bias, void* grad_offset,
                                const void* grad_output, int batch, int channels, int height, int width, int length,
                                int channels_out, int kernel_h, int kernel_w, int kernel_l, int stride_h,
                                int stride_w, int stride_l, int pad_h, int pad_w, int pad_l, int dilation_h,
                                int dilation_w, int dilation_l, int group, int deformable_group, int in_step,
                                int with_bias, int dtype, void* workspace, size_t workspace_bytes, void* stream) {
  (void)in_step; (void)bias;
  mdconv_geom_t g = legacy_geom(3, 0, batch, channels, height, width, length, channels_out, kernel_h, kernel_w,
                                kernel_l, stride_h, stride_w, stride_l, pad_h, pad_w, pad_l, dilation_h, dilation_w,
                                dilation_l, group, deformable_group, with_bias, dtype);
  return mdconv_backward(&g, input, weight, offset, nullptr, grad_output, grad_input, grad_weight,
                         with_bias ? grad_bias : nullptr, grad_offset, nullptr, workspace, workspace_bytes, stream);
}

int modulated_deform_conv3d_forward_cuda(const void* input, const void* weight, const void* bias,
                                         const void* offset, const void* mask, void* output, int batch,
                                         int channels, int height, int width, int length, int channels_out,
                                         int kernel_h, int kernel_w, int kernel_l, int stride_h, int stride_w,
                                         int stride_l, int pad_h, int pad_w, int pad_l, int dilation_h,
                                         int dilation_w, int dilation_l, int group, int deformable_group,
                                         int in_step, int with_bias, int dtype, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  (void)in_step;
  mdconv_geom_t g = legacy_geom(3, 1, batch, channels, height, width, length, channels_out, kernel_h, kernel_w,
                                kernel_l, stride_h, stride_w, stride_l, pad_h, pad_w, pad_l, dilation_h, dilation_w,
                                dilation_l, group, deformable_group, with_bias, dtype);
  return mdconv_forward(&g, input, weight, with_bias ? bias : nullptr, offset, mask, output, workspace,
                        workspace_bytes, stream);
}

int modulated_deform_conv3d_backward_cuda(const void* input, const void* weight, const void* bias,
                                          const void* offset, const void* mask, void* grad_input, void* grad_weight,
                                          void* grad_bias, void* grad_offset, void* grad_mask,
                                          const void* grad_output, int batch, int channels, int height, int width,
                                          int length, int channels_out, int kernel_h, int kernel_w, int kernel_l,
                                          int stride_h, int stride_w, int stride_l, int pad_h, int pad_w, int pad_l,
                                          int dilation_h, int dilation_w, int dilation_l, int group,
                                          int deformable_group, int in_step, int with_bias, int dtype,
                                          void* workspace, size_t workspace_bytes, void* stream) {
  (void)in_step; (void)bias;
  mdconv_geom_t g = legacy_geom(3, 1, batch, channels, height, width, length, channels_out, kernel_h, kernel_w,
                                kernel_l, stride_h, stride_w, stride_l, pad_h, pad_w, pad_l, dilation_h, dilation_w,
                                dilation_l, group, deformable_group, with_bias, dtype);
  return mdconv_backward(&g, input, weight, offset, mask, grad_output, grad_input, grad_weight,
                         with_bias ? grad_bias : nullptr, grad_offset, grad_mask, workspace, workspace_bytes, stream);
}

}  // extern "C"
