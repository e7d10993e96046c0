// common.cuh -- shared device-side definitions for libmdconv_b200 (sm_100a only).
//
// Semantics implemented here restate SURVEY.md appendix A (derived from the reference's
// src/mdeformable_conv.cu:5-35,49-86,222-317 and twins); the code itself is written for this
// library's fused kernels, not transcribed.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mdconv_b200.h"

namespace mdconv {

constexpr float kEps = 1.1920929e-7f;  // FLT_EPSILON, reference src/config.h:18

// n / d for 0 <= n < 2^31 without the ~30-instruction integer-division sequence (the sampling-plan
// producers decode a position and a tap per plan).  q = umulhi(n, mul) >> shr, mul = ceil(2^(32+shr) / d).
struct FastDiv {
  uint32_t mul, shr;
  int d;
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = d;
  f.mul = 0; f.shr = 0;
  if (d > 1) {
    int lg = 0;
    while ((1ll << lg) < d) ++lg;  // ceil(log2 d)
    const unsigned long long m = ((1ull << (31 + lg)) + (unsigned long long)d - 1) / (unsigned long long)d;
    f.mul = (uint32_t)m;
    f.shr = (uint32_t)(lg - 1);
  }
  return f;
}
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) {
  return f.d == 1 ? n : (int)(__umulhi((uint32_t)n, f.mul) >> f.shr);
}

// Fully resolved geometry, passed to kernels by value.
struct DevGeom {
  int nd, B, C, Co;
  int in[3], k[3], s[3], p[3], d[3], out[3];
  int g, dg;
  int K;    // taps
  int P;    // output positions per sample
  int IN;   // input positions per sample and channel
  int Cg, Cog, Cdg;
  int Kg;   // Cg * K, reduction length of the per-group contraction
  int modulated, rule, with_bias;
  int esize;  // bytes per element of the call's dtype
  int cl;   // 1: activations (input, output, grad_output, grad_input) are channels-last [B, spatial..., C]
  FastDiv fd_out[3], fd_k[3];  // divisors out[a], k[a]
  FastDiv fd_K, fd_Cdg;        // divisors K (taps), Cdg (channels per deformable group)
};

// ---- storage <-> compute conversions --------------------------------------------------------
template <typename F, typename T> __device__ __forceinline__ F ld_as(const T* p);
template <> __device__ __forceinline__ float ld_as<float, float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_as<float, __half>(const __half* p) { return __half2float(__ldg(p)); }
template <> __device__ __forceinline__ float ld_as<float, __nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(__ldg(p));
}
template <> __device__ __forceinline__ double ld_as<double, double>(const double* p) { return __ldg(p); }

template <typename T, typename F> __device__ __forceinline__ T cvt_to(F v);
template <> __device__ __forceinline__ float cvt_to<float, float>(float v) { return v; }
template <> __device__ __forceinline__ __half cvt_to<__half, float>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_to<__nv_bfloat16, float>(float v) {
  return __float2bfloat16_rn(v);
}
template <> __device__ __forceinline__ double cvt_to<double, double>(double v) { return v; }

template <typename F> __device__ __forceinline__ F ffloor(F v);
template <> __device__ __forceinline__ float ffloor<float>(float v) { return floorf(v); }
template <> __device__ __forceinline__ double ffloor<double>(double v) { return floor(v); }

// ---- one sampling point ----------------------------------------------------------------------
// Corner numbering: bit a of the corner id set <=> "high" on axis a (axis 0 = h, 1 = w, 2 = l).
template <int ND, typename F>
struct SamplePoint {
  int lo[ND];
  F fr[ND];
  bool in_range;  // every coordinate strictly inside (-1, size)
};

// Sampling coordinate p_a = out_a*stride_a - pad_a + tap_a*dil_a + offset_a, formed in the
// compute type F exactly as the reference forms it in scalar_t (int + scalar_t -> scalar_t).
template <int ND, typename T, typename F>
__device__ __forceinline__ void make_sample_point(const DevGeom& g, const T* __restrict__ off_b, int dgi, int tap,
                                                  int p, SamplePoint<ND, F>& s) {
  int o[3], t[3];
  if (ND == 2) {
    o[0] = fdiv(p, g.fd_out[1]); o[1] = p - o[0] * g.out[1];
    t[0] = fdiv(tap, g.fd_k[1]); t[1] = tap - t[0] * g.k[1];
  } else {
    const int r = fdiv(p, g.fd_out[2]); o[2] = p - r * g.out[2];
    o[0] = fdiv(r, g.fd_out[1]); o[1] = r - o[0] * g.out[1];
    const int q = fdiv(tap, g.fd_k[2]); t[2] = tap - q * g.k[2];
    t[0] = fdiv(q, g.fd_k[1]); t[1] = q - t[0] * g.k[1];
  }
  const T* op = off_b + ((size_t)(dgi * g.K + tap) * ND) * g.P + p;
  bool ok = true;
#pragma unroll
  for (int a = 0; a < ND; ++a) {
    const int base = o[a] * g.s[a] - g.p[a] + t[a] * g.d[a];
    const F pos = (F)base + ld_as<F>(op + (size_t)a * g.P);
    const F fl = ffloor(pos);
    s.lo[a] = (int)fl;
    s.fr[a] = pos - fl;
    ok = ok && (pos > (F)-1) && (pos < (F)g.in[a]);
  }
  s.in_range = ok;
}

// Forward interpolation of one channel plane at a sample point (value is 0 when out of range;
// a corner contributes only when it lies inside the plane).
template <int ND, typename T, typename F>
__device__ __forceinline__ F interp_forward(const DevGeom& g, const T* __restrict__ plane, const SamplePoint<ND, F>& s) {
  if (!s.in_range) return (F)0;
  F val = (F)0;
#pragma unroll
  for (int c = 0; c < (1 << ND); ++c) {
    F w = (F)1;
    bool ok = true;
    int idx = 0;
#pragma unroll
    for (int a = 0; a < ND; ++a) {
      const int hi = (c >> a) & 1;
      const int ia = s.lo[a] + hi;
      ok = ok && (hi ? (ia <= g.in[a] - 1) : (ia >= 0));
      w *= hi ? s.fr[a] : ((F)1 - s.fr[a]);
      idx = idx * g.in[a] + ia;
    }
    if (ok) val += w * ld_as<F>(plane + idx);
  }
  return val;
}

// Corner plan for the backward rule: per corner the linear index, the interpolation weight,
// whether its value is loaded and whether its grad_input add happens.
template <int ND, typename F>
struct CornerPlan {
  int idx[1 << ND];
  F wgt[1 << ND];
  unsigned load_mask;  // bit c: corner value participates (else treated as 0)
  unsigned add_mask;   // bit c: grad_input add happens
  bool offset_grad;    // grad_offset is produced for this point
};

template <int ND, typename F>
__device__ __forceinline__ void make_corner_plan(const DevGeom& g, const SamplePoint<ND, F>& s, bool rule_b,
                                                 CornerPlan<ND, F>& cp) {
  bool lo_ok[ND], hi_load[ND], hi_add[ND];
#pragma unroll
  for (int a = 0; a < ND; ++a) {
    const int lo = s.lo[a], hi = lo + 1;
    lo_ok[a] = (lo >= 0) && (lo <= g.in[a] - 1);
    const bool hi_in = (hi >= 0) && (hi <= g.in[a] - 1);
    const bool frac = fabs((double)s.fr[a]) > (double)kEps;
    // rule A: values always loaded, only zero-weight adds skipped; rule B: the corner is absent.
    hi_load[a] = rule_b ? (hi_in && frac) : hi_in;
    hi_add[a] = hi_in && frac;
  }
  cp.load_mask = 0;
  cp.add_mask = 0;
#pragma unroll
  for (int c = 0; c < (1 << ND); ++c) {
    bool okl = true, oka = true;
    int idx = 0;
    F w = (F)1;
#pragma unroll
    for (int a = 0; a < ND; ++a) {
      const int hi = (c >> a) & 1;
      const int ia = s.lo[a] + hi;
      okl = okl && (hi ? hi_load[a] : lo_ok[a]);
      oka = oka && (hi ? hi_add[a] : lo_ok[a]);
      w *= hi ? s.fr[a] : ((F)1 - s.fr[a]);
      idx = idx * g.in[a] + ia;
    }
    cp.idx[c] = idx;
    cp.wgt[c] = w;
    if (okl) cp.load_mask |= 1u << c;
    if (oka) cp.add_mask |= 1u << c;
  }
  cp.offset_grad = rule_b ? true : s.in_range;
}

// d(value)/d(coordinate a) given the corner values v[] (already zeroed where not loaded).
template <int ND, typename F>
__device__ __forceinline__ F coord_derivative(const SamplePoint<ND, F>& s, const F* v, int a) {
  F acc = (F)0;
#pragma unroll
  for (int c = 0; c < (1 << ND); ++c) {
    F w = ((c >> a) & 1) ? (F)1 : (F)-1;
#pragma unroll
    for (int b = 0; b < ND; ++b) {
      if (b == a) continue;
      w *= ((c >> b) & 1) ? s.fr[b] : ((F)1 - s.fr[b]);
    }
    acc += w * v[c];
  }
  return acc;
}

// ---- host-side launch bookkeeping -------------------------------------------------------------
struct LaunchCtx {
  cudaStream_t stream;
  int launches;
  cudaError_t err;
  // backward only: recorded on `stream` as soon as grad_weight and grad_bias are final (nullptr: not wanted), so a
  // data-parallel caller can start their all-reduce while the data gradients are still being computed
  cudaEvent_t weight_ready = nullptr;
  // backward only: the join event of a side-stream fork made earlier in this call (nullptr: nothing forked)
  cudaEvent_t side_join = nullptr;
  // SMs the persistent kernels launched from now on leave free (set once the weight-ready event has been
  // recorded: the caller's all-reduce kernel needs somewhere to run while the data gradients are computed)
  int spare_sms = 0;
  // backward only: optional fp32 copies of the parameter gradients before rounding (mdconv_bwd_extra_t)
  float* gw32 = nullptr;
  float* gb32 = nullptr;
  // backward only: row stride (elements) of the grad_output rows handed to the bias-gradient kernel (0: P)
  int go_ld = 0;
};
// SMs left to a concurrent collective after the weight-ready event (MDCONV_B200_COMM_SMS, default 8; api.cu)
int comm_spare_sms();
int set_comm_spare_sms(int sms);   // returns the previous value
int set_fp32_on_tensor_cores(int on);   // returns the previous value (tensor_path.cu)
namespace tc {
long long set_gcol_budget_mb(long long mb);   // grad_columns budget of the col2im-gather grad_input (tensor_bwd_input.cu)
}
inline void signal_weight_ready(LaunchCtx& ctx) {
  if (ctx.weight_ready != nullptr) {
    cudaError_t e = cudaEventRecord(ctx.weight_ready, ctx.stream);
    if (e != cudaSuccess && ctx.err == cudaSuccess) ctx.err = e;
    ctx.weight_ready = nullptr;
    ctx.spare_sms = comm_spare_sms();
  }
}

#define MDCONV_CHECK_LAUNCH(ctx)                       \
  do {                                                 \
    (ctx).launches++;                                  \
    cudaError_t e__ = cudaGetLastError();              \
    if (e__ != cudaSuccess && (ctx).err == cudaSuccess) (ctx).err = e__; \
  } while (0)

// Optional per-kernel timing (mdconv_profile_enable): CUDA events recorded on the launching
// stream around one kernel.  No-ops unless profiling is on.  Implemented in api.cu.
void prof_begin(LaunchCtx& ctx, const char* name);
void prof_end(LaunchCtx& ctx);
struct ProfScope {
  LaunchCtx& ctx;
  ProfScope(LaunchCtx& c, const char* name) : ctx(c) { prof_begin(ctx, name); }
  ~ProfScope() { prof_end(ctx); }
};

// generic (SIMT) path, generic_kernels.cu
size_t generic_workspace_bytes(const DevGeom& g, int dtype, int backward);
void generic_forward(const DevGeom& g, int dtype, const void* x, const void* w, const void* bias, const void* off,
                     const void* mask, void* y, LaunchCtx& ctx);
// `parts` selects which gradients the generic kernels produce (the tensor path delegates the ones
// it does not compute itself).
enum { GENERIC_BWD_DATA = 1, GENERIC_BWD_WEIGHT = 2, GENERIC_BWD_BIAS = 4, GENERIC_BWD_ALL = 7 };
void generic_backward(const DevGeom& g, int dtype, const void* x, const void* w, const void* off, const void* mask,
                      const void* go, void* gx, void* gw, void* gb, void* goff, void* gmask, void* ws, LaunchCtx& ctx,
                      int parts = GENERIC_BWD_ALL);

}  // namespace mdconv
