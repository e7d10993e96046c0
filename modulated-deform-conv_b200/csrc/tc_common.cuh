// tc_common.cuh -- pieces shared by the tcgen05 kernels (forward, grad-weight, grad-data):
// the per-(pixel, tap) sampling plan, the 8-channel corner gather, and the pre-pass kernels.
#pragma once
#include <atomic>
#include <type_traits>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace mdconv {
namespace tc {

constexpr int BK = 64;  // K elements per block = 128 bytes = one 128B-swizzle row

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// Channels-last copy of the input: B*IN positions x C channels, plus ONE all-zero position at the end.  The sampling
// plans point every absent corner (sample out of range, corner outside the input) at that position: the reference
// returns an exact 0 there WITHOUT reading the input (src/mdeformable_conv.cu:19-29,80), so an Inf/NaN elsewhere in x
// must not turn into 0 * Inf = NaN through a zero-weight load of some unrelated pixel.
inline size_t xt_bytes(const DevGeom& g) { return g.cl ? 0 : align256(((size_t)g.B * g.IN + 1) * g.C * g.esize); }
// re-ordered filter bank(s): one [Co][Kg] 16-bit bank, or two (bf16 hi, bf16 lo) for fp32 callers
inline size_t wr_bytes(const DevGeom& g) { return align256((size_t)g.Co * g.Kg * 2) * (g.esize == 4 ? 2 : 1); }

// Shared by the forward and grad_input epilogues when the caller's activations are channels-last: one 64-channel
// chunk of a 128-row accumulator tile goes TMEM -> registers -> (+bias, convert) -> a 128B-swizzled staging tile in
// shared memory -> ONE TMA store (cp.async.bulk.tensor) of [128 rows x 64 channels]; rows past the end of the
// tensor are clipped by the tensor map.  Two staging tiles alternate so a store overlaps the next chunk.
constexpr int CL_STAGE_BYTES = 128 * 128;

// ---- explicit shared-state-space accesses (generic LD/ST through a computed pointer is slower) ----
// `loaded & zero` (zero: a run-time 0) as ONE instruction the compiler can neither fold, duplicate nor sink into a branch:
// the first read of a register some earlier load wrote, i.e. the place where the thread waits for that load's
// scoreboard.  The result (0) is added to the addresses of whatever must not start before that wait.
__device__ __forceinline__ uint32_t drain_point(uint32_t loaded, int zero) {
  uint32_t r;
  asm volatile("and.b32 %0, %1, %2;" : "=r"(r) : "r"(loaded), "r"(zero));
  return r;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// byte offset of 16-byte chunk `chunk` of row `row` inside a [rows][128 B] tile with 128B swizzle
__device__ __forceinline__ uint32_t sw128(int row, int chunk) {
  return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}

// ---- sampling plan --------------------------------------------------------------------------------
// One (output position, tap): per corner the address of the 8-channel vector in the channels-last
// input (uint4 units, channel offset excluded) and the interpolation weight (x mask), zero for
// corners outside the input or points outside the sampling range (reference rule:
// src/mdeformable_conv.cu:19-29,80).  Computed once per (position, tap), shared by all channels
// of the deformable group.
template <int ND>
struct __align__(16) PlanEntry {
  int idx[1 << ND];
  uint32_t w[1 << ND];  // the weight x mask twice, in the storage dtype: half2 (w, w) / bfloat162 (w, w)
};

template <int ND>
struct RowCoord {
  int b, p;
  int base[ND];  // out*stride - pad per axis
  bool valid;
};

template <int ND>
__device__ __forceinline__ void decode_row(const DevGeom& g, int b, int p, bool valid, RowCoord<ND>& rc) {
  rc.b = valid ? b : 0;
  rc.p = valid ? p : 0;
  rc.valid = valid;
  int o[3];
  if (ND == 2) {
    o[0] = fdiv(rc.p, g.fd_out[1]); o[1] = rc.p - o[0] * g.out[1];
  } else {
    const int r2 = fdiv(rc.p, g.fd_out[2]); o[2] = rc.p - r2 * g.out[2];
    o[0] = fdiv(r2, g.fd_out[1]); o[1] = r2 - o[0] * g.out[1];
  }
#pragma unroll
  for (int a = 0; a < ND; ++a) rc.base[a] = o[a] * g.s[a] - g.p[a];
}

__device__ __forceinline__ uint32_t pack_w(float w, __half) {
  __half2 h = __float2half2_rn(w);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_w(float w, float) { return __float_as_uint(w); }   // fp32 callers: as is
__device__ __forceinline__ uint32_t pack_w(float w, __nv_bfloat16) {
  __nv_bfloat162 h = __float2bfloat162_rn(w);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Sampling position of (row, tap): floor, fraction, range flag, mask value.
template <int ND>
struct SamplePos {
  int lo[ND];
  float fr[ND];
  bool in_range;
  float m;
};

// Raw offset / mask values of one (row, deformable group, tap): the only global loads of a plan.
// Split from the arithmetic so a producer can fetch the next plan's values while it builds this one.
template <int ND>
struct RawSample {
  float o[ND];
  float m;
};

template <typename T, int ND, bool MOD>
__device__ __forceinline__ RawSample<ND> fetch_sample(const DevGeom& g, const T* __restrict__ off,
                                                      const T* __restrict__ mask, const RowCoord<ND>& rc, int dgi,
                                                      int tap) {
  RawSample<ND> r;
  const size_t ch = (size_t)rc.b * g.dg * g.K + (size_t)dgi * g.K + tap;
  const T* op = off + ch * ND * g.P + rc.p;
#pragma unroll
  for (int a = 0; a < ND; ++a) r.o[a] = ld_as<float>(op + (size_t)a * g.P);
  r.m = 1.f;
  if (MOD) r.m = ld_as<float>(mask + ch * g.P + rc.p);
  return r;
}

template <int ND>
__device__ __forceinline__ void build_pos(const DevGeom& g, const RowCoord<ND>& rc, int tap, const RawSample<ND>& r,
                                          SamplePos<ND>& sp) {
  int tt[3];
  if (ND == 2) {
    tt[0] = fdiv(tap, g.fd_k[1]); tt[1] = tap - tt[0] * g.k[1];
  } else {
    const int q = fdiv(tap, g.fd_k[2]); tt[2] = tap - q * g.k[2];
    tt[0] = fdiv(q, g.fd_k[1]); tt[1] = q - tt[0] * g.k[1];
  }
  sp.in_range = true;
#pragma unroll
  for (int a = 0; a < ND; ++a) {
    const float pos = (float)(rc.base[a] + tt[a] * g.d[a]) + r.o[a];
    const float fl = floorf(pos);
    sp.lo[a] = (int)fl;
    sp.fr[a] = pos - fl;
    sp.in_range = sp.in_range && (pos > -1.f) && (pos < (float)g.in[a]);
  }
  sp.m = r.m;
}

template <typename T, int ND, bool MOD>
__device__ __forceinline__ void sample_pos(const DevGeom& g, const T* __restrict__ off, const T* __restrict__ mask,
                                           const RowCoord<ND>& rc, int dgi, int tap, SamplePos<ND>& sp) {
  const RawSample<ND> r = fetch_sample<T, ND, MOD>(g, off, mask, rc, dgi, tap);
  build_pos<ND>(g, rc, tap, r, sp);
}

// corner indices + weights of an already-located sample
template <typename T, int ND>
__device__ __forceinline__ PlanEntry<ND> plan_from_pos(const DevGeom& g, const RowCoord<ND>& rc,
                                                       const SamplePos<ND>& sp, int C8) {
  PlanEntry<ND> e;
  // NCHW callers: the all-zero position behind the last sample of the workspace copy (xt_bytes).  Channels-last
  // callers are gathered from in place (nothing can be appended to their tensor): a zero-weight load of the
  // sample's own first pixel stands in, the one case where a non-finite input can differ from the reference.
  const int zero_pos = (g.cl ? rc.b : g.B) * g.IN * C8;
#pragma unroll
  for (int c = 0; c < (1 << ND); ++c) { e.idx[c] = zero_pos; e.w[c] = 0u; }
  if (!rc.valid || !sp.in_range) return e;
#pragma unroll
  for (int c = 0; c < (1 << ND); ++c) {
    float w = sp.m;
    bool ok = true;
    int pix = 0;
#pragma unroll
    for (int a = 0; a < ND; ++a) {
      const int hi = (c >> a) & 1;
      const int ia = sp.lo[a] + hi;
      ok = ok && (hi ? (ia <= g.in[a] - 1) : (ia >= 0));
      w *= hi ? sp.fr[a] : (1.f - sp.fr[a]);
      pix = pix * g.in[a] + ia;
    }
    if (ok) {
      e.idx[c] = (rc.b * g.IN + pix) * C8;
      e.w[c] = pack_w(w, T());
    }
  }
  return e;
}

template <typename T, int ND, bool MOD>
__device__ __forceinline__ PlanEntry<ND> make_plan_entry(const DevGeom& g, const T* __restrict__ off,
                                                         const T* __restrict__ mask, const RowCoord<ND>& rc, int c0,
                                                         int tap, int C8) {
  SamplePos<ND> sp;
  sp.in_range = false;
  if (rc.valid) sample_pos<T, ND, MOD>(g, off, mask, rc, c0 / g.Cdg, tap, sp);
  return plan_from_pos<T, ND>(g, rc, sp, C8);
}

// Regular-grid plan (plain convolution): tap (row, tap) reads ONE input position, base + tap * dilation, or the
// all-zero position when that lies in the padding.  Only idx[0] is used by the consumer.
template <typename T, int ND>
__device__ __forceinline__ PlanEntry<ND> make_grid_entry(const DevGeom& g, const RowCoord<ND>& rc, int tap, int C8) {
  PlanEntry<ND> e;
  const int zero_pos = (g.cl ? rc.b : g.B) * g.IN * C8;
#pragma unroll
  for (int c = 0; c < (1 << ND); ++c) { e.idx[c] = zero_pos; e.w[c] = 0u; }
  if (!rc.valid) return e;
  int tt[3];
  if (ND == 2) {
    tt[0] = fdiv(tap, g.fd_k[1]); tt[1] = tap - tt[0] * g.k[1];
  } else {
    const int q = fdiv(tap, g.fd_k[2]); tt[2] = tap - q * g.k[2];
    tt[0] = fdiv(q, g.fd_k[1]); tt[1] = q - tt[0] * g.k[1];
  }
  bool ok = true;
  int pix = 0;
#pragma unroll
  for (int a = 0; a < ND; ++a) {
    const int ia = rc.base[a] + tt[a] * g.d[a];
    ok = ok && ia >= 0 && ia < g.in[a];
    pix = pix * g.in[a] + ia;
  }
  if (ok) e.idx[0] = (rc.b * g.IN + pix) * C8;
  return e;
}

template <int ND>
__device__ __forceinline__ void store_plan(uint32_t saddr, const PlanEntry<ND>& e) {
  const uint4* q = reinterpret_cast<const uint4*>(&e);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(PlanEntry<ND>) / 16); ++i) sts128(saddr + 16 * i, q[i]);
}
template <int ND>
__device__ __forceinline__ PlanEntry<ND> load_plan(uint32_t saddr) {
  PlanEntry<ND> e;
  uint4* q = reinterpret_cast<uint4*>(&e);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(PlanEntry<ND>) / 16); ++i) q[i] = lds128(saddr + 16 * i);
  return e;
}

// ---- 8-channel interpolation ------------------------------------------------------------------------
// Packed HFMA2 in the storage dtype (HFMA2 for fp16, HFMA2.BF16 for bf16) with the weight x mask broadcast to both
// halves: 4 instructions per corner for 8 channels.  The result feeds a 16-bit tensor-core operand anyway; the extra
// roundings of the 4- / 8-term sum stay well inside the parity bars (profiles/r2_parity_margins.json).
template <typename H2>
__device__ __forceinline__ uint4 interp8_packed(const uint4* v, const uint32_t* w, int nc) {
  uint4 acc = make_uint4(0u, 0u, 0u, 0u);   // +0.0 in both formats
  H2* a = reinterpret_cast<H2*>(&acc);
  for (int c = 0; c < nc; ++c) {
    const H2 wh = *reinterpret_cast<const H2*>(&w[c]);
    const H2* x = reinterpret_cast<const H2*>(&v[c]);
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __hfma2(wh, x[i], a[i]);
  }
  return acc;
}
__device__ __forceinline__ uint4 interp8(const uint4* v, const uint32_t* w, int nc, __half) {
  return interp8_packed<__half2>(v, w, nc);
}
__device__ __forceinline__ uint4 interp8(const uint4* v, const uint32_t* w, int nc, __nv_bfloat16) {
  return interp8_packed<__nv_bfloat162>(v, w, nc);
}

// column value (8 channels) of one plan entry; xc already points at the channel chunk
template <typename T, int ND>
__device__ __forceinline__ uint4 gather8(const uint4* __restrict__ xc, const PlanEntry<ND>& e) {
  uint4 v[1 << ND];
#pragma unroll
  for (int c = 0; c < (1 << ND); ++c) v[c] = __ldg(xc + e.idx[c]);
  return interp8(v, e.w, 1 << ND, T());
}

// ---- channels-last epilogue (forward output, grad_input) ---------------------------------------------------
__device__ __forceinline__ uint32_t pack2_storage(float a, float b, __half) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack2_storage(float a, float b, __nv_bfloat16) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// Called by all 128 epilogue threads (4 warps; thread `row` owns TMEM lane `row` = tile row) once the accumulator
// of a tile is complete.  taddr: TMEM address of this warp's 32 lanes at the tile's first column; ncols (multiple of
// 64) columns go out as ncols/64 TMA stores at (column col0 + 64*i, row row0) of the [rows][channels] tensor behind
// `tmap`.  bias: NULL or the bias of column col0 (added in fp32 before the conversion).  `buf` is the caller's
// staging-buffer toggle (kept across tiles); `leader` = the one thread that issues / tracks the bulk stores.
// bar_id: a named barrier reserved for these 128 threads.
template <typename T>
__device__ __forceinline__ void epilogue_store_channels_last(uint32_t taddr, int ncols, const T* __restrict__ bias,
                                                             uint8_t* sStage, int& buf, const CUtensorMap* tmap,
                                                             int col0, int row0, int row, bool leader, int bar_id) {
  for (int ch = 0; ch < ncols; ch += 64) {
    if (leader) ptx::bulk_wait_group_read<1>();   // the store that last read this staging tile has drained it
    ptx::named_bar_sync(bar_id, 128);
    const uint32_t sb = ptx::smem_u32(sStage + buf * CL_STAGE_BYTES);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t v[16];
      ptx::tmem_ld_32x16(taddr + ch + q * 16, v);
      ptx::tmem_ld_wait();
      float f[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        f[j] = __uint_as_float(v[j]);
        if (bias != nullptr) f[j] += ld_as<float>(bias + ch + q * 16 + j);
      }
      uint4 o0, o1;
      o0.x = pack2_storage(f[0], f[1], T());   o0.y = pack2_storage(f[2], f[3], T());
      o0.z = pack2_storage(f[4], f[5], T());   o0.w = pack2_storage(f[6], f[7], T());
      o1.x = pack2_storage(f[8], f[9], T());   o1.y = pack2_storage(f[10], f[11], T());
      o1.z = pack2_storage(f[12], f[13], T()); o1.w = pack2_storage(f[14], f[15], T());
      sts128(sb + sw128(row, 2 * q), o0);
      sts128(sb + sw128(row, 2 * q + 1), o1);
    }
    ptx::fence_proxy_async_smem();                // generic-proxy writes -> visible to the TMA (async proxy) read
    ptx::named_bar_sync(bar_id, 128);
    if (leader) {
      ptx::tma_store_2d(tmap, sStage + buf * CL_STAGE_BYTES, col0 + ch, row0);
      ptx::bulk_commit_group();
    }
    buf ^= 1;
  }
}

// ---- pre-pass kernels (tensor_path.cu) ----------------------------------------------------------------
// zero_row: also clear the extra all-zero position behind the last sample (xt_bytes)
// out_ld > 0: row stride of the output (the inverse direction, see k_to_channels_last)
template <typename T> void launch_to_channels_last(const T* x, T* xt, int B, int C, int IN, LaunchCtx& ctx,
                                                   bool zero_row = false, int out_ld = 0);
// W [Co][Cg][K] -> Wr [Co][(cc*K + tap)*64 + ci]   (c = cc*64 + ci)
template <typename T> void launch_reorder_weight(const T* w, void* wr, int Co, int Cg, int K, LaunchCtx& ctx);

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_fn();
int sm_count();   // of the CURRENT device (cached per device ordinal)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute of a kernel: opt in once per
// (kernel instantiation, device ordinal).  `seen` is a function-local static bitmask of the ordinals already done;
// two threads racing on the first call both set the attribute, which is harmless.
template <typename Kernel>
inline void opt_in_dynamic_smem(Kernel* kernel, size_t smem, std::atomic<unsigned long long>& seen, LaunchCtx& ctx) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  const bool tracked = dev >= 0 && dev < 64;
  const unsigned long long bit = tracked ? (1ull << dev) : 0ull;
  if (tracked && (seen.load(std::memory_order_acquire) & bit)) return;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    if (ctx.err == cudaSuccess) ctx.err = e;
    return;
  }
  if (tracked) seen.fetch_or(bit, std::memory_order_release);
}

template <typename T> inline CUtensorMapDataType tmap_dtype();
// [rows][channels] row-major tensor (a channels-last activation), box = 64 channels x 128 rows, 128B swizzle:
// the map the channels-last epilogues store through
template <typename T>
inline bool encode_rows_by_channels(CUtensorMap* m, void* base, size_t rows, int channels) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)channels, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)channels * sizeof(T)};
  const cuuint32_t box[2] = {64, 128};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, tmap_dtype<T>(), 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename T>
inline CUtensorMapDataType tmap_dtype() {
  return std::is_same<T, __half>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
}
bool fp32_on_tensor_cores();
template <typename T>
constexpr int umma_fmt() {
  return std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
}

// Tail split of a persistent tile loop (see TcFwdParams): `tiles` 128-row tiles dealt round-robin to
// min(tiles, sms) CTAs.  If the last, incomplete round holds at most half a grid of tiles, each of them is cut into
// up to 4 row ranges (multiples of 8 rows: one gather warp owns 8 rows) so the ranges fill the grid.
inline void plan_tail_split(int tiles, int sms, bool allowed, int& full_tiles, int& tail_tiles, int& split, int& rows) {
  const int grid = tiles < sms ? tiles : sms;
  full_tiles = grid > 0 ? (tiles / grid) * grid : 0;
  tail_tiles = tiles - full_tiles;
  split = 1;
  if (allowed && tail_tiles > 0) {
    split = grid / tail_tiles;
    if (split > 4) split = 4;
    if (split < 1) split = 1;
  }
  rows = ((128 + split - 1) / split + 7) & ~7;
  if (split > 1 && rows * (split - 1) >= 128) {   // the last range would be empty: fewer, fuller ranges
    split = (128 + rows - 1) / rows;
  }
}

// Side stream of the current device (created on first use) + fork/join helpers: small independent
// passes (the inverse-index build) run beside the layout passes of the main stream.
struct SideStream {
  cudaStream_t stream;
  cudaEvent_t fork, join;
};
SideStream* side_stream();   // nullptr if disabled or it cannot be created (callers then stay on the main stream)
void set_side_stream_enabled(bool on);

// backward kernels (tensor_bwd.cu)
template <typename T>
void tensor_backward_t(const DevGeom& g, const void* x, const void* w, const void* off, const void* mask,
                       const void* go, void* gx, void* gw, void* gb, void* goff, void* gmask, void* ws, const void* saved,
                       LaunchCtx& ctx);
size_t tensor_backward_workspace(const DevGeom& g);
bool tensor_backward_supported(const DevGeom& g, int dtype);

// grad_offset / grad_mask (tensor_bwd_sample.cu)
bool grad_sample_supported(const DevGeom& g);
template <typename T>
void tensor_grad_sample(const DevGeom& g, const T* xt, const void* wr, const void* off, const void* mask, const void* go,
                        const void* go_lo, int go_ld, void* goff, void* gmask, void* gcol, int sample0, int samples,
                        LaunchCtx& ctx);
// grad_input (tensor_bwd_input.cu)
bool grad_input_supported(const DevGeom& g);
size_t grad_input_workspace(const DevGeom& g);
template <typename T>
void tensor_grad_input_prepare(const DevGeom& g, const void* off, const void* mask, void* ws, LaunchCtx& ctx);
template <typename T>
void tensor_grad_input(const DevGeom& g, const void* w, const void* go, void* gx, void* ws, LaunchCtx& ctx);
// col2im-gather path of grad_input (16-bit callers): the batch is walked in chunks of gcol_chunk_samples() samples;
// per chunk k_tc_grad_sample streams its grad_columns into ONE buffer of the workspace -- small enough to stay in L2,
// re-used by every chunk -- and a gather over the inverse index turns it into that chunk's grad_input
bool grad_input_uses_gcol(const DevGeom& g);
int gcol_chunk_samples(const DevGeom& g);
void* grad_input_gcol(const DevGeom& g, void* ws);            // the grad_columns buffer (or nullptr: contraction path)
long long set_gcol_budget_mb(long long mb);                   // returns the previous budget (-1: not set yet)
template <typename T>
void tensor_grad_input_from_gcol(const DevGeom& g, void* gx, void* ws, int sample0, int samples, LaunchCtx& ctx);

}  // namespace tc
}  // namespace mdconv
