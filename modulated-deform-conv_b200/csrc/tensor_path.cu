// tensor_path.cu -- tcgen05 / TMEM / TMA path (fp16, bf16) for sm_100a.
//
// Forward = implicit GEMM  D[pixels x Co] = Col[pixels x (Ci/g)*K] * W^T  per group, with the
// offset-driven gather fused into the A-operand staging: the column matrix only ever exists as
// 128x64 swizzled shared-memory tiles that tcgen05.mma consumes; nothing column-shaped touches HBM.
//
//   warp 0       TMA producer: weight K-slices [BN x 64] (128B swizzle) -> smem ring
//   warp 1       MMA issuer: one lane issues tcgen05.mma (M=128, N=BN, K=16) x4 per K-block,
//                accumulators in TMEM (2 stages x 256 columns), tcgen05.commit -> mbarriers
//   warp 2       TMEM allocator
//   warps 4-7    epilogue: tcgen05.ld -> +bias -> convert -> NCHW stores (overlaps the next tile)
//   warps 8-11   plan producers: one sampling plan per (pixel row, tap) -- corner indices + weight*mask,
//                once per (pixel, tap), not per channel as the reference does
//                (src/mdeformable_conv.cu:59-83) -- into a 4-slot ring that runs ahead of the gather
//   warps 12-27  gather producers: per K-block (64 channels x 1 tap) 128-bit channels-last loads of the
//                corners (all loads of a thread in flight together), HFMA2 interpolation, 16-byte
//                stores into the swizzled A tile, fence.proxy.async, mbarrier arrive
//
// Data layout: the kernel reads x as channels-last [B][positions][C] (one 128-byte line = 64
// channels of one position, so 8 lanes x 16 B cover a line) and the weight re-ordered to
// [Co][(channel-block, tap, 64 channels)] so a K-block is 64 contiguous elements; both are produced
// by two small pre-pass kernels into the caller's workspace.
//
// Replaces reference K1-K4 + G1 (src/mdeformable_conv.cu:38-87,172-184 and twins).
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "tc_common.cuh"
#include "tensor_path.cuh"

namespace mdconv {

using tc::PlanEntry;
using tc::RowCoord;

namespace {

constexpr int TC_BM = 128;        // pixels per tile (UMMA M)
constexpr int TC_BK = tc::BK;     // K elements per block = 128 bytes = one swizzle row
#ifndef MDCONV_TC_STAGES
#define MDCONV_TC_STAGES 3
#endif
constexpr int TC_STAGES = MDCONV_TC_STAGES;      // A/B operand ring
constexpr int TC_PLAN_SLOTS = 4;  // sampling-plan ring
constexpr int TC_MAX_BN = 256;
constexpr int TC_PLAN_WARP0 = 8;      // warps 8-11: plan producers (one tile row per thread)
constexpr int TC_GATHER_WARP0 = 12;   // warps 12-27: gather producers
constexpr int TC_GATHER_WARPS = 16;
constexpr int TC_GATHER_THREADS = TC_GATHER_WARPS * 32;
constexpr int TC_THREADS = (TC_GATHER_WARP0 + TC_GATHER_WARPS) * 32;  // 896
constexpr int TC_A_BYTES = TC_BM * 128;
constexpr int TC_B_BYTES = TC_MAX_BN * 128;

using tc::align256;

struct TcFwdParams {
  DevGeom g;
  const void* xt;    // channels-last input
  const void* off;
  const void* mask;
  const void* bias;
  void* y;
  void* y2;          // regular-grid producer: channels [split, co_valid) go here ([B][co_valid - split][P]), else NULL
  int split;         // first channel of y2 (co_valid when there is no second output)
  int co_valid;      // output channels that are really stored (the rest are zero-filter padding up to a multiple of 32)
  int BN;            // output channels per tile
  int m_tiles, n_tiles, total_tiles;
  // Tail split: the tiles are dealt round-robin to the persistent CTAs; when the last round would keep only a
  // fraction of them busy (784 tiles on 148 SMs: 44 of 148), each of its `tail_tiles` tiles is cut into `tail_split`
  // row ranges of `tail_rows` rows and every range goes to its own CTA.  The kernels are paced by the gather, whose
  // cost follows the number of live rows, so the last round then takes ~1/tail_split of a full one.
  int full_tiles;    // tiles dealt in complete rounds (a multiple of the grid size)
  int tail_tiles, tail_split, tail_rows;
  int ccs;           // channel blocks per group (Cg / 64)
  int bpd;           // channel blocks per deformable group (Cdg / 64)
  int kblocks;       // ccs * K
  int zero;          // always 0 at run time, opaque to the compiler: see the gather loop (load ordering)
};

// K-block order inside one (tile, group): channel blocks are walked in runs that share a deformable
// group; inside a run the tap is the outer loop, so ONE sampling plan (per position and tap) serves
// every channel block of the run.  All warp roles iterate this same sequence.
//   F(cc, tap, first_of_phase, blocks_in_phase)
template <typename F>
__device__ __forceinline__ void for_each_kblock(const TcFwdParams& prm, int grp, F&& f) {
  const int gb0 = grp * prm.ccs;
  int cc = 0;
  while (cc < prm.ccs) {
    const int room = prm.bpd - (gb0 + cc) % prm.bpd;
    const int run_end = cc + room < prm.ccs ? cc + room : prm.ccs;
    for (int tap = 0; tap < prm.g.K; ++tap)
      for (int c = cc; c < run_end; ++c) f(c, tap, c == cc, run_end - cc);
    cc = run_end;
  }
}

// Work items of a CTA: tiles blockIdx.x, +gridDim.x, ... through the complete rounds, then at most ONE row range
// of a tail tile (tile index >= full_tiles marks it; its rows are a per-CTA constant, see tail_rows_of_cta).
__device__ __forceinline__ int next_tile(const TcFwdParams& prm, int t) {   // -1: no more work
  if (t >= prm.full_tiles) return -1;                       // the tail item is always the last one
  const int n = t + (int)gridDim.x;
  if (n < prm.full_tiles) return n;
  const int j = (int)blockIdx.x;                            // == n - full_tiles
  return j < prm.tail_tiles * prm.tail_split ? prm.full_tiles + j / prm.tail_split : -1;
}
__device__ __forceinline__ void tail_rows_of_cta(const TcFwdParams& prm, int& lo, int& hi) {
  lo = ((int)blockIdx.x % prm.tail_split) * prm.tail_rows;
  hi = min(TC_BM, lo + prm.tail_rows);
}

// The same sequence as for_each_kblock, but as an explicit iterator over (tile, run, tap, channel
// block) so the gather warps can look one K-block ahead (loads of K-block k+1 are issued before
// K-block k is interpolated).
struct KbIter {
  int t, grp, run0, run_end, tap, cc;
  bool valid;
  __device__ __forceinline__ void set_run(const TcFwdParams& prm, int c0) {
    run0 = c0;
    const int room = prm.bpd - (grp * prm.ccs + c0) % prm.bpd;
    run_end = c0 + room < prm.ccs ? c0 + room : prm.ccs;
  }
  __device__ __forceinline__ void set_tile(const TcFwdParams& prm, int tile) {
    t = tile;
    valid = tile >= 0;
    grp = valid ? (tile / prm.m_tiles) / prm.n_tiles : 0;
    tap = 0;
    set_run(prm, 0);
    cc = 0;
  }
  __device__ __forceinline__ void advance(const TcFwdParams& prm) {
    if (++cc < run_end) return;
    if (++tap < prm.g.K) { cc = run0; return; }
    tap = 0;
    if (run_end < prm.ccs) { set_run(prm, run_end); cc = run0; return; }
    set_tile(prm, next_tile(prm, t));
  }
  __device__ __forceinline__ bool first() const { return cc == run0; }
  __device__ __forceinline__ bool last() const { return cc == run_end - 1; }
};

// fp32 callers (T = float) run the SAME implicit GEMM on the bf16 tensor pipe with every operand split in two:
// v = hi + lo, hi = bf16(v), lo = bf16(v - hi) (16 mantissa bits together), and three products per K step,
// hi*hi + hi*lo + lo*hi, accumulated in the fp32 TMEM tile; the dropped lo*lo term and the split residue are
// ~2^-17 relative per product, far inside the 1e-4 bar (what the reference gets from an SGEMM,
// src/deformable_conv.cu:176-178).  The gathered column tile is interpolated in fp32 and split on its way into
// shared memory (two A tiles per stage), the filters are split once by the re-order pass (two B tiles per stage).
template <typename T> struct FwdCfg {
  static constexpr bool F32 = false;
  static constexpr int A_STAGE = TC_A_BYTES, PLAN_SLOTS = TC_PLAN_SLOTS, MAX_BN = TC_MAX_BN;
};
template <> struct FwdCfg<float> {
  static constexpr bool F32 = true;
  static constexpr int A_STAGE = 2 * TC_A_BYTES, PLAN_SLOTS = 2, MAX_BN = 128;   // B stage: hi rows, then lo rows
};
__device__ __forceinline__ uint32_t bf16x2_bits(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// 8 fp32 values -> their bf16 "hi" parts and the bf16 of what is left
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  float r[8];
  uint32_t* h = reinterpret_cast<uint32_t*>(&hi);
  uint32_t* l = reinterpret_cast<uint32_t*>(&lo);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    const float2 back = __bfloat1622float2(hh);
    r[2 * i] = v[2 * i] - back.x;
    r[2 * i + 1] = v[2 * i + 1] - back.y;
    l[i] = bf16x2_bits(r[2 * i], r[2 * i + 1]);
  }
}

// Register budget after setmaxnreg: the CTA owns TC_THREADS x 72 registers (what it is launched with);
// the control / epilogue / plan warpgroups give some back, the gather warpgroups take them.
constexpr uint32_t TC_REGS_LAUNCH = 72;
constexpr uint32_t TC_REGS_CTL = 32, TC_REGS_EPI = 48, TC_REGS_PLAN = 48, TC_REGS_GATHER = 88;
static_assert(128 * (TC_REGS_CTL + TC_REGS_EPI + TC_REGS_PLAN) + TC_GATHER_THREADS * TC_REGS_GATHER <=
                  TC_THREADS * TC_REGS_LAUNCH,
              "setmaxnreg.inc would wait forever: the CTA does not own that many registers");

// GRID: no offset tensor -- every tap samples its integer grid position (a plain convolution: the offset / mask
// producer of the Pack modules, reference modulated_deform_conv.py:734-738,759-768).  One corner per (position, tap),
// weight 1, no interpolation arithmetic; everything else (roles, rings, MMA, epilogue) is the deformable kernel.
// CL: the caller's activations are channels-last (the output leaves through TMA stores); compile-time so the NCHW
// instantiation carries none of that epilogue's registers or code.
template <typename T, int ND, bool MOD, bool GRID, bool CL>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_forward(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_y,
             const __grid_constant__ CUtensorMap tmap_w2, const TcFwdParams prm) {
  constexpr bool F32 = FwdCfg<T>::F32;
  constexpr int A_STAGE = FwdCfg<T>::A_STAGE;
  constexpr int PLAN_SLOTS = FwdCfg<T>::PLAN_SLOTS;
  static_assert(!(F32 && (GRID || CL)), "the fp32 split path has no producer / channels-last variant");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve-up (all stage buffers 1024-byte aligned for the 128B swizzle); the two output staging tiles exist only
  // for channels-last callers (the launch asks for the extra shared memory only then)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = sA + TC_STAGES * A_STAGE;
  uint8_t* sStage = sB + TC_STAGES * TC_B_BYTES;
  PlanEntry<ND>* sPlan = reinterpret_cast<PlanEntry<ND>*>(sStage + (CL ? 2 * tc::CL_STAGE_BYTES : 0));
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sPlan) +
                                               PLAN_SLOTS * TC_BM * sizeof(PlanEntry<ND>));
  uint64_t* full_a = bars;
  uint64_t* full_b = full_a + TC_STAGES;
  uint64_t* empty = full_b + TC_STAGES;
  uint64_t* plan_full = empty + TC_STAGES;
  uint64_t* plan_empty = plan_full + PLAN_SLOTS;
  uint64_t* tmem_full = plan_empty + PLAN_SLOTS;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const DevGeom& g = prm.g;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_w);
    if (CL) ptx::prefetch_tensormap(&tmap_y);
    if (F32) ptx::prefetch_tensormap(&tmap_w2);
    for (int s = 0; s < TC_STAGES; ++s) {
      ptx::mbar_init(&full_a[s], ptx::arrivals(TC_GATHER_THREADS));
      ptx::mbar_init(&full_b[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < PLAN_SLOTS; ++s) {
      ptx::mbar_init(&plan_full[s], ptx::arrivals(TC_BM));
      ptx::mbar_init(&plan_empty[s], ptx::arrivals(TC_GATHER_THREADS));
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], ptx::arrivals(128));
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<512>(tmem_ptr);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  const int BN = prm.BN;
  const long long total_pix = (long long)g.B * g.P;

  if (warp < 4) {
    ptx::reg_dealloc<TC_REGS_CTL>();
  if (warp == 0) {
    // ===================== TMA producer: weight K-slices =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t >= 0; t = next_tile(prm, t)) {
        const int rest = t / prm.m_tiles;
        const int nt = rest % prm.n_tiles, grp = rest / prm.n_tiles;
        const int row0 = grp * g.Cog + nt * BN;
        for_each_kblock(prm, grp, [&](int cc, int tap, bool, int) {
          ptx::mbar_wait_relaxed(&empty[s], ph ^ 1, 64);
          ptx::mbar_arrive_expect_tx(&full_b[s], (uint32_t)BN * 128u * (F32 ? 2u : 1u));
          ptx::tma_load_2d(sB + s * TC_B_BYTES, &tmap_w, &full_b[s], (cc * g.K + tap) * TC_BK, row0);
          if (F32)   // the "lo" halves of the same filter slice
            ptx::tma_load_2d(sB + s * TC_B_BYTES + TC_B_BYTES / 2, &tmap_w2, &full_b[s], (cc * g.K + tap) * TC_BK, row0);
          if (++s == TC_STAGES) { s = 0; ph ^= 1; }
        });
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr int fmt = F32 ? 1 : tc::umma_fmt<T>();   // fp32 callers: bf16 halves
#ifdef MDCONV_EXP_HALFN   // timing experiment only (wrong results): half-width MMAs, everything else unchanged
      const uint32_t idesc = ptx::umma_idesc_f16(fmt, 0, 0, TC_BM, BN / 2);
#else
      const uint32_t idesc = ptx::umma_idesc_f16(fmt, 0, 0, TC_BM, BN);
#endif
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int t = blockIdx.x; t >= 0; t = next_tile(prm, t)) {
        ptx::mbar_wait(&tmem_empty[as], aph ^ 1);
        ptx::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * TC_MAX_BN);
        for (int kb = 0; kb < prm.kblocks; ++kb) {
          ptx::mbar_wait(&full_b[s], ph);
          ptx::mbar_wait(&full_a[s], ph);
          ptx::tc_fence_after_sync();
          const uint32_t a_addr = ptx::smem_u32(sA + s * A_STAGE);
          const uint32_t b_addr = ptx::smem_u32(sB + s * TC_B_BYTES);
#pragma unroll
#ifdef MDCONV_EXP_NOMMA    // timing experiment only: no MMAs at all (the commits still run the pipeline)
          for (int k = 0; k < 0; ++k) {
#else
          for (int k = 0; k < TC_BK / 16; ++k) {
#endif
            const uint64_t ad = ptx::umma_smem_desc_sw128(a_addr + k * 32, 0, 1024);
            const uint64_t bd = ptx::umma_smem_desc_sw128(b_addr + k * 32, 0, 1024);
            ptx::umma_f16(d_tmem, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            if (F32) {   // + hi*lo + lo*hi
              const uint64_t al = ptx::umma_smem_desc_sw128(a_addr + TC_A_BYTES + k * 32, 0, 1024);
              const uint64_t bl = ptx::umma_smem_desc_sw128(b_addr + TC_B_BYTES / 2 + k * 32, 0, 1024);
              ptx::umma_f16(d_tmem, ad, bl, idesc, 1u);
              ptx::umma_f16(d_tmem, al, bd, idesc, 1u);
            }
          }
          ptx::umma_commit(&empty[s]);
          if (++s == TC_STAGES) { s = 0; ph ^= 1; }
        }
        ptx::umma_commit(&tmem_full[as]);
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  }
  } else if (warp < 8) {
    // ===================== epilogue =====================
    ptx::reg_dealloc<TC_REGS_EPI>();
    const int ew = warp & 3;
    T* __restrict__ y = reinterpret_cast<T*>(prm.y);
    const T* __restrict__ bias = reinterpret_cast<const T*>(prm.bias);
    int as = 0;
    uint32_t aph = 0;
    int cl_buf = 0;
    int tail_lo, tail_hi;
    tail_rows_of_cta(prm, tail_lo, tail_hi);
    const int trow = ew * 32 + lane;
    const bool in_tail_rows = trow >= tail_lo && trow < tail_hi;
    for (int t = blockIdx.x; t >= 0; t = next_tile(prm, t)) {
      const int mt = t % prm.m_tiles;
      const int rest = t / prm.m_tiles;
      const int nt = rest % prm.n_tiles, grp = rest / prm.n_tiles;
      const long long gp = (long long)mt * TC_BM + trow;
      const bool valid = gp < total_pix && (t < prm.full_tiles || in_tail_rows);
      const int b = valid ? (int)(gp / g.P) : 0;
      const int p = valid ? (int)(gp % g.P) : 0;
      const int co0 = grp * g.Cog + nt * BN;
      T* yrow = y + ((size_t)b * g.Co + co0) * g.P + p;
      ptx::mbar_wait_relaxed(&tmem_full[as], aph);
      ptx::tc_fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * TC_MAX_BN);
      if constexpr (CL) {
        // channels-last output [B*P][Co]: 64-channel chunks through shared memory and TMA stores
        tc::epilogue_store_channels_last<T>(taddr, BN, g.with_bias ? bias + co0 : nullptr, sStage, cl_buf, &tmap_y, co0,
                                            mt * TC_BM, ew * 32 + lane, ew == 0 && lane == 0, 1);
      } else {
        for (int c0 = 0; c0 < BN; c0 += 16) {
          uint32_t v[16];
          ptx::tmem_ld_32x16(taddr + c0, v);
          ptx::tmem_ld_wait();
          if (valid) {
            if (!GRID) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float f = __uint_as_float(v[j]);
                if (g.with_bias) f += ld_as<float>(bias + co0 + c0 + j);
                yrow[(size_t)(c0 + j) * g.P] = cvt_to<T, float>(f);
              }
            } else {
              // producer: only co_valid channels exist; channels >= split belong to the second tensor
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int co = co0 + c0 + j;
                if (co < prm.co_valid) {
                  float f = __uint_as_float(v[j]);
                  if (g.with_bias) f += ld_as<float>(bias + co);
                  T* dst = co < prm.split
                               ? y + ((size_t)b * prm.split + co) * g.P + p
                               : reinterpret_cast<T*>(prm.y2) + ((size_t)b * (prm.co_valid - prm.split) + (co - prm.split)) * g.P + p;
                  *dst = cvt_to<T, float>(f);
                }
              }
            }
          }
        }
      }
      ptx::tc_fence_before_sync();
      ptx::mbar_arrive_warp(&tmem_empty[as]);
      if (++as == 2) { as = 0; aph ^= 1; }
    }
    if (CL && ew == 0 && lane == 0) ptx::bulk_wait_group<0>();   // the staging tiles must outlive their stores
  } else if (warp < TC_GATHER_WARP0) {
    // ===================== plan producers =====================
    ptx::reg_dealloc<TC_REGS_PLAN>();
    // One thread per tile row: the sampling plan of (row, tap) for the current deformable-group run
    // (corner indices + weight*mask, reference src/mdeformable_conv.cu:59-83 computes this per
    // channel), written into a ring of plan slots that runs ahead of the gather warps.
    const int row = threadIdx.x - TC_PLAN_WARP0 * 32;
    const T* __restrict__ off = reinterpret_cast<const T*>(prm.off);
    const T* __restrict__ mask = reinterpret_cast<const T*>(prm.mask);
    const int C8 = F32 ? g.C >> 2 : g.C >> 3;   // 16-byte vectors per position
    int ps = 0;
    uint32_t pph = 0;
    int tail_lo, tail_hi;
    tail_rows_of_cta(prm, tail_lo, tail_hi);
    const bool in_tail_rows = row >= tail_lo && row < tail_hi;
    for (int t = blockIdx.x; t >= 0; t = next_tile(prm, t)) {
      const int mt = t % prm.m_tiles;
      const int grp = (t / prm.m_tiles) / prm.n_tiles;
      const long long gp = (long long)mt * TC_BM + row;
      const bool v = gp < total_pix && (t < prm.full_tiles || in_tail_rows);
      RowCoord<ND> rc;
      tc::decode_row<ND>(g, v ? (int)(gp / g.P) : 0, v ? (int)(gp % g.P) : 0, v, rc);
      for_each_kblock(prm, grp, [&](int cc, int tap, bool first, int) {
        if (!first) return;
        const PlanEntry<ND> e = GRID ? tc::make_grid_entry<T, ND>(g, rc, tap, C8)
                                     : tc::make_plan_entry<T, ND, MOD>(g, off, mask, rc, grp * g.Cg + cc * TC_BK, tap, C8);
        ptx::mbar_wait_relaxed(&plan_empty[ps], pph ^ 1);
        tc::store_plan<ND>(ptx::smem_u32(sPlan + ps * TC_BM + row), e);
        ptx::mbar_arrive_warp(&plan_full[ps]);
        if (++ps == PLAN_SLOTS) { ps = 0; pph ^= 1; }
      });
    }
  } else {
    // ===================== gather producers =====================
    ptx::reg_alloc<TC_REGS_GATHER>();
    // 16 warps x 2 passes x 4 rows = 128 rows per K-block; 8 lanes x 16 bytes cover the 64 channels
    // of one corner.  Software-pipelined one K-block deep: the corner loads of K-block k+1 are in
    // flight while K-block k is interpolated and stored (two register buffers, ping-pong).
    constexpr int NC = GRID ? 1 : (1 << ND);          // corners per sample
    constexpr int NQ = (1 << ND) / 4;                 // 16-byte words of a plan row's index (and weight) block
    const int gw = warp - TC_GATHER_WARP0;
    const int chunk = lane & 7;
    const int row0 = gw * 8 + (lane >> 3);
    const uint4* __restrict__ xq = reinterpret_cast<const uint4*>(prm.xt);
    const uint32_t plan_row = ptx::smem_u32(sPlan + row0);
    constexpr uint32_t kSlotBytes = TC_BM * (uint32_t)sizeof(PlanEntry<ND>);
    constexpr uint32_t kRow4 = 4 * (uint32_t)sizeof(PlanEntry<ND>);
    int s = 0;                 // operand ring position (consume side)
    uint32_t ph = 0;
    int ips = 0, cps = 0;      // plan ring positions: issue side / consume side
    uint32_t ipph = 0;
    // (rows outside a tail item's range carry an all-zero plan: their loads all hit the one cached all-zero
    // position, a quarter of the L1 wavefronts of a live row, and nothing reads the rows back)

    // issue the corner loads of K-block `it` into (v0, v1)
    // `after` (always 0) chains the loads behind a value read from the OTHER register buffer, see the loop below
    auto issue = [&](const KbIter& it, uint4 (&v0)[NC], uint4 (&v1)[NC], uint32_t after) {
      if constexpr (F32) return;
      if (it.first()) ptx::mbar_wait(&plan_full[ips], ipph);
      const uint32_t pa = plan_row + ips * kSlotBytes + after;
      const uint4* xc = xq + ((it.grp * g.Cg + it.cc * TC_BK) >> 3) + chunk;
      int i0[4 * NQ], i1[4 * NQ];
#pragma unroll
      for (int q = 0; q < (GRID ? 1 : NQ); ++q) {
        const uint4 a = tc::lds128(pa + 16 * q), b = tc::lds128(pa + kRow4 + 16 * q);
        i0[4 * q] = (int)a.x; i0[4 * q + 1] = (int)a.y; i0[4 * q + 2] = (int)a.z; i0[4 * q + 3] = (int)a.w;
        i1[4 * q] = (int)b.x; i1[4 * q + 1] = (int)b.y; i1[4 * q + 2] = (int)b.z; i1[4 * q + 3] = (int)b.w;
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) v0[c] = __ldg(xc + i0[c]);
#ifndef MDCONV_EXP_HALFROWS   // timing experiment only: the second row of every thread is not gathered
#pragma unroll
      for (int c = 0; c < NC; ++c) v1[c] = __ldg(xc + i1[c]);
#endif
      if (it.last()) {
        if (++ips == PLAN_SLOTS) { ips = 0; ipph ^= 1; }
      }
    };
    // interpolate K-block `it` from (v0, v1) into the A stage, publish it, release the plan slot
    auto finish = [&](const KbIter& it, const uint4 (&v0)[NC], const uint4 (&v1)[NC], uint32_t after = 0u) {
      ptx::mbar_wait(&empty[s], ph ^ 1);            // MMA has finished reading A[s]
      const uint32_t a_addr = ptx::smem_u32(sA + s * A_STAGE) + after;
      if (GRID) {
        // the loaded vector IS the column value (absent taps were pointed at the all-zero position)
        tc::sts128(a_addr + tc::sw128(row0, chunk), v0[0]);
        tc::sts128(a_addr + tc::sw128(row0 + 4, chunk), v1[0]);
      } else {
        const uint32_t pa = plan_row + cps * kSlotBytes + 4 * (1 << ND);  // weights follow the indices
        uint32_t w0[4 * NQ], w1[4 * NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const uint4 a = tc::lds128(pa + 16 * q), b = tc::lds128(pa + kRow4 + 16 * q);
          w0[4 * q] = a.x; w0[4 * q + 1] = a.y; w0[4 * q + 2] = a.z; w0[4 * q + 3] = a.w;
          w1[4 * q] = b.x; w1[4 * q + 1] = b.y; w1[4 * q + 2] = b.z; w1[4 * q + 3] = b.w;
        }
        if constexpr (!F32) {
          tc::sts128(a_addr + tc::sw128(row0, chunk), tc::interp8(v0, w0, NC, T()));
#ifndef MDCONV_EXP_HALFROWS
          tc::sts128(a_addr + tc::sw128(row0 + 4, chunk), tc::interp8(v1, w1, NC, T()));
#endif
        }
      }
#ifndef MDCONV_EXP_NOFENCE    // timing experiment only
      ptx::fence_proxy_async_smem();
#endif
      ptx::mbar_arrive_warp(&full_a[s]);
      if (++s == TC_STAGES) { s = 0; ph ^= 1; }
      if (it.last()) {
        ptx::mbar_arrive_warp(&plan_empty[cps]);
        if (++cps == PLAN_SLOTS) cps = 0;
      }
    };
    // fp32: one K-block start to finish (no look-ahead: 8 channels of a corner are two 16-byte loads, a row's
    // corners already fill the registers).  Each of the thread's two rows: load, interpolate in fp32, split into
    // hi / lo bf16, store both tiles.
    auto block_f32 = [&](const KbIter& it) {
      if (it.first()) ptx::mbar_wait(&plan_full[ips], ipph);
      const uint32_t pa = plan_row + ips * kSlotBytes;
      const uint4* xc = xq + ((it.grp * g.Cg + it.cc * TC_BK) >> 2) + 2 * chunk;
      ptx::mbar_wait(&empty[s], ph ^ 1);            // MMA has finished reading A[s]
      const uint32_t a_addr = ptx::smem_u32(sA + s * A_STAGE);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const uint4 iv = tc::lds128(pa + r * kRow4 + 16 * q);
          const uint4 wv = tc::lds128(pa + r * kRow4 + 4 * (1 << ND) + 16 * q);
          const int idx[4] = {(int)iv.x, (int)iv.y, (int)iv.z, (int)iv.w};
          const float wgt[4] = {__uint_as_float(wv.x), __uint_as_float(wv.y), __uint_as_float(wv.z), __uint_as_float(wv.w)};
          uint4 va[4], vb[4];   // channels 0-3 and 4-7 of this lane's 8, for four corners
#pragma unroll
          for (int c = 0; c < 4; ++c) { va[c] = __ldg(xc + idx[c]); vb[c] = __ldg(xc + idx[c] + 1); }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            acc[0] = fmaf(wgt[c], __uint_as_float(va[c].x), acc[0]);
            acc[1] = fmaf(wgt[c], __uint_as_float(va[c].y), acc[1]);
            acc[2] = fmaf(wgt[c], __uint_as_float(va[c].z), acc[2]);
            acc[3] = fmaf(wgt[c], __uint_as_float(va[c].w), acc[3]);
            acc[4] = fmaf(wgt[c], __uint_as_float(vb[c].x), acc[4]);
            acc[5] = fmaf(wgt[c], __uint_as_float(vb[c].y), acc[5]);
            acc[6] = fmaf(wgt[c], __uint_as_float(vb[c].z), acc[6]);
            acc[7] = fmaf(wgt[c], __uint_as_float(vb[c].w), acc[7]);
          }
        }
        uint4 hi, lo;
        split8(acc, hi, lo);
        tc::sts128(a_addr + tc::sw128(row0 + 4 * r, chunk), hi);
        tc::sts128(a_addr + TC_A_BYTES + tc::sw128(row0 + 4 * r, chunk), lo);
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive_warp(&full_a[s]);
      if (++s == TC_STAGES) { s = 0; ph ^= 1; }
      if (it.last()) {
        ptx::mbar_arrive_warp(&plan_empty[ips]);
        if (++ips == PLAN_SLOTS) { ips = 0; ipph ^= 1; }
      }
    };

    KbIter ka, kb2;
    ka.set_tile(prm, blockIdx.x);
    if (F32) {
      while (ka.valid) {
        block_f32(ka);
        ka.advance(prm);
      }
    } else if (ND == 2 || GRID) {
      // All corner loads of a thread retire through ONE hardware scoreboard (ptxas gives every LDG of this loop the
      // same one), and waiting on a scoreboard means waiting for EVERYTHING outstanding on it: if the loads of K-block
      // k+1 were issued before the first use of K-block k's registers, that use would also wait for the loads just
      // issued and the look-ahead would hide nothing.  So the wait comes first: one instruction reads a register of
      // the buffer about to be interpolated (`& prm.zero`: a run-time 0 the compiler cannot fold), the address of the
      // next loads depends on its result, and the loads then fly while the current block is interpolated and stored.
      uint4 a0[NC], a1[NC], b0[NC], b1[NC];
      if (ka.valid) issue(ka, a0, a1, 0u);
      while (ka.valid) {
        kb2 = ka;
        kb2.advance(prm);
        const uint32_t da = tc::drain_point(a1[NC - 1].w, prm.zero);
        if (kb2.valid) issue(kb2, b0, b1, da);
        finish(ka, a0, a1, da);
        if (!kb2.valid) break;
        ka = kb2;
        ka.advance(prm);
        const uint32_t db = tc::drain_point(b1[NC - 1].w, prm.zero);
        if (ka.valid) issue(ka, a0, a1, db);
        finish(kb2, b0, b1, db);
      }
    } else {
      // 3D: 16 corner loads per thread already fill the register budget; no look-ahead
      uint4 a0[NC], a1[NC];
      while (ka.valid) {
        issue(ka, a0, a1, 0u);
        finish(ka, a0, a1);
        ka.advance(prm);
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<512>(tmem_base);
}

// ---- pre-pass kernels ---------------------------------------------------------------------------
// x [B][C][IN] -> xt [B][IN][C].  64 channels x 64 positions per block; 16-byte loads along the
// positions, 16-byte stores along the channels, transposed through shared memory (row stride 66
// elements = 33 words keeps both phases at most 2-way bank conflicted).  VEC == false is the scalar
// form for sizes / pointers that are not 16-byte friendly.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) k_to_channels_last(const T* __restrict__ x, T* __restrict__ xt, int C, int IN,
                                                          int zero_row, int out_ld) {
  // out_ld: row stride of the output in elements (C for the dense channels-last copy; the same kernel also turns a
  // channels-last grad_output [P][Co] back into rows [Co][ld] padded to 16 bytes: C := P, IN := Co, out_ld := ld)
  __shared__ __align__(16) T tile[64][66];  // [position][channel]
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 64, i0 = blockIdx.x * 64;
  const T* xb = x + (size_t)b * C * IN;
  T* xtb = xt + (size_t)b * IN * out_ld;
  if (zero_row && blockIdx.x == 0 && b == 0 && threadIdx.x < 64 && c0 + (int)threadIdx.x < C)
    xt[(size_t)gridDim.z * C * IN + c0 + threadIdx.x] = cvt_to<T, float>(0.f);   // the all-zero position (xt_bytes)
  if (VEC) {
    // load: thread -> (channel, 8-position chunk); 512 (channel, chunk) pairs, two per thread
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int q = threadIdx.x + 256 * h;
      const int c = q >> 3, ch = q & 7;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (c0 + c < C && i0 + ch * 8 < IN) v = __ldg(reinterpret_cast<const uint4*>(xb + (size_t)(c0 + c) * IN + i0 + ch * 8));
      const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
      for (int k = 0; k < 8; ++k) tile[ch * 8 + k][c] = e[k];
    }
    __syncthreads();
    // store: thread -> (position, 8-channel chunk)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int q = threadIdx.x + 256 * h;
      const int i = q >> 3, ch = q & 7;
      if (i0 + i < IN && c0 + ch * 8 < C) {
        uint4 v;
        uint32_t* o = reinterpret_cast<uint32_t*>(&v);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&tile[i][ch * 8]);
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = src[k];
        *reinterpret_cast<uint4*>(xtb + (size_t)(i0 + i) * out_ld + c0 + ch * 8) = v;
      }
    }
  } else {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      const int c = ty + 8 * rr;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = tx + 32 * h;
        if (c0 + c < C && i0 + i < IN) tile[i][c] = xb[(size_t)(c0 + c) * IN + i0 + i];
      }
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      const int i = ty + 8 * rr;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = tx + 32 * h;
        if (c0 + c < C && i0 + i < IN) xtb[(size_t)(i0 + i) * out_ld + c0 + c] = tile[i][c];
      }
    }
  }
}

// W [Co][Cg][K] -> Wr [Co][(cc*K + tap)*64 + ci],  c = cc*64 + ci
template <typename T>
__global__ void k_reorder_weight(const T* __restrict__ w, T* __restrict__ wr, int Co, int Cg, int K) {
  const size_t n = (size_t)Co * Cg * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int Kg = Cg * K;
    const int co = (int)(i / Kg);
    const int r = (int)(i % Kg);
    const int ci = r & 63;
    const int q = r >> 6;
    const int tap = q % K, cc = q / K;
    wr[i] = w[(size_t)co * Kg + (size_t)(cc * 64 + ci) * K + tap];
  }
}

// fp32 filters: the same re-order, split into bf16 hi / lo banks (w = hi + lo up to 2^-17 relative)
__global__ void k_reorder_weight_split(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi,
                                       __nv_bfloat16* __restrict__ lo, int Co, int Cg, int K) {
  const size_t n = (size_t)Co * Cg * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int Kg = Cg * K;
    const int co = (int)(i / Kg);
    const int r = (int)(i % Kg);
    const int ci = r & 63;
    const int q = r >> 6;
    const int tap = q % K, cc = q / K;
    const float v = w[(size_t)co * Kg + (size_t)(cc * 64 + ci) * K + tap];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// ---- host side --------------------------------------------------------------------------------
int pick_bn(const DevGeom& g) {
  const int mx = g.esize == 4 ? FwdCfg<float>::MAX_BN : TC_MAX_BN;
  return g.Cog <= mx ? g.Cog : mx;
}

template <typename T, int ND>
size_t fwd_smem_bytes(bool cl) {
  return 1024 + (size_t)TC_STAGES * (FwdCfg<T>::A_STAGE + TC_B_BYTES) + (cl ? 2 * tc::CL_STAGE_BYTES : 0) +
         (size_t)FwdCfg<T>::PLAN_SLOTS * TC_BM * sizeof(PlanEntry<ND>) +
         (3 * TC_STAGES + 2 * FwdCfg<T>::PLAN_SLOTS + 4) * sizeof(uint64_t) + 16;
}

struct FwdMaps {
  CUtensorMap w, y, w2;   // filters (hi for fp32), channels-last output, filters lo (fp32 only)
};
template <typename T, int ND, bool MOD, bool GRID, bool CL>
void launch_tc_forward_cl(const FwdMaps& m, const TcFwdParams& prm, LaunchCtx& ctx) {
  const size_t smem = fwd_smem_bytes<T, ND>(CL);
  static std::atomic<unsigned long long> configured{0};
  tc::opt_in_dynamic_smem(k_tc_forward<T, ND, MOD, GRID, CL>, smem, configured, ctx);
  const int grid = prm.total_tiles < tc::sm_count() ? prm.total_tiles : tc::sm_count();
  ProfScope prof(ctx, GRID ? "tc_producer_conv" : "tc_forward");
  k_tc_forward<T, ND, MOD, GRID, CL><<<grid, TC_THREADS, smem, ctx.stream>>>(m.w, m.y, m.w2, prm);
  MDCONV_CHECK_LAUNCH(ctx);
}
template <typename T, int ND, bool MOD, bool GRID = false>
void launch_tc_forward(const FwdMaps& m, const TcFwdParams& prm, LaunchCtx& ctx) {
  if constexpr (std::is_same<T, float>::value) {
    launch_tc_forward_cl<T, ND, MOD, false, false>(m, prm, ctx);
  } else {
    if (!GRID && prm.g.cl) launch_tc_forward_cl<T, ND, MOD, false, !GRID>(m, prm, ctx);
    else launch_tc_forward_cl<T, ND, MOD, GRID, false>(m, prm, ctx);
  }
}

// bf16 tensor map over a re-ordered filter bank [Co][Kg], box = 64 K-elements x BN rows
template <typename E>
static bool encode_filter_map(CUtensorMap* m, const void* wr, const DevGeom& g, int BN) {
  tc::PFN_encodeTiled enc = tc::get_encode_fn();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)g.Kg, (cuuint64_t)g.Co};
  const cuuint64_t strides[1] = {(cuuint64_t)g.Kg * sizeof(E)};
  const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)BN};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, tc::tmap_dtype<E>(), 2, const_cast<void*>(wr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename T>
void tensor_forward_t(const DevGeom& g, const void* x, const void* w, const void* bias, const void* off,
                      const void* mask, void* y, void* ws, const FwdExtra& ex, LaunchCtx& ctx) {
  constexpr bool F32 = std::is_same<T, float>::value;
  // channels-last callers are gathered from in place; NCHW callers get a channels-last copy in the workspace,
  // unless an earlier call already made one (FwdExtra::xt)
  const T* xt = g.cl ? reinterpret_cast<const T*>(x) : reinterpret_cast<const T*>(ws);
  // (a channels-last copy that lives OUTSIDE this workspace leaves only the re-ordered weight here, at offset 0)
  const bool xt_elsewhere = !g.cl && ex.xt != nullptr && ex.xt != ws;
  char* wr = reinterpret_cast<char*>(ws) + (xt_elsewhere ? 0 : tc::xt_bytes(g));
  if (!g.cl && ex.xt != nullptr) xt = reinterpret_cast<const T*>(ex.xt);
  else if (!g.cl) tc::launch_to_channels_last<T>((const T*)x, reinterpret_cast<T*>(ws), g.B, g.C, g.IN, ctx, /*zero_row=*/true);
  tc::launch_reorder_weight<T>((const T*)w, wr, g.Co, g.Cg, g.K, ctx);   // fp32: hi bank, then lo bank
  TcFwdParams prm;
  prm.g = g;
  prm.xt = xt; prm.off = off; prm.mask = mask; prm.bias = bias; prm.y = y;
  prm.y2 = ex.y2;
  prm.co_valid = ex.co_valid > 0 ? ex.co_valid : g.Co;
  prm.split = ex.y2 != nullptr ? ex.split : prm.co_valid;
  prm.BN = pick_bn(g);
  const long long total_pix = (long long)g.B * g.P;
  prm.m_tiles = (int)((total_pix + TC_BM - 1) / TC_BM);
  prm.n_tiles = g.Cog / prm.BN;
  prm.total_tiles = prm.m_tiles * prm.n_tiles * g.g;
  prm.ccs = g.Cg / TC_BK;
  prm.bpd = g.Cdg / TC_BK;
  prm.kblocks = prm.ccs * g.K;
  prm.zero = 0;
  tc::plan_tail_split(prm.total_tiles, tc::sm_count(), !g.cl, prm.full_tiles, prm.tail_tiles, prm.tail_split,
                      prm.tail_rows);

  FwdMaps maps;
  memset(&maps, 0, sizeof(maps));
  bool ok;
  if constexpr (F32) {
    ok = encode_filter_map<__nv_bfloat16>(&maps.w, wr, g, prm.BN) &&
         encode_filter_map<__nv_bfloat16>(&maps.w2, wr + tc::wr_bytes(g) / 2, g, prm.BN);
  } else {
    ok = encode_filter_map<T>(&maps.w, wr, g, prm.BN);
    if (ok && g.cl) ok = tc::encode_rows_by_channels<T>(&maps.y, y, (size_t)g.B * g.P, g.Co);
  }
  if (!ok) {
    if (ctx.err == cudaSuccess) ctx.err = cudaErrorInvalidValue;
    return;
  }
  if constexpr (!F32) {
    if (ex.grid) {
      if (g.nd == 2) launch_tc_forward<T, 2, false, true>(maps, prm, ctx);
      else launch_tc_forward<T, 3, false, true>(maps, prm, ctx);
      return;
    }
  }
  if (g.nd == 2) {
    if (g.modulated) launch_tc_forward<T, 2, true>(maps, prm, ctx);
    else launch_tc_forward<T, 2, false>(maps, prm, ctx);
  } else {
    if (g.modulated) launch_tc_forward<T, 3, true>(maps, prm, ctx);
    else launch_tc_forward<T, 3, false>(maps, prm, ctx);
  }
}

}  // namespace

namespace tc {

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

static std::atomic<bool> g_side_stream_on{true};
void set_side_stream_enabled(bool on) { g_side_stream_on = on; }

SideStream* side_stream() {
  if (!g_side_stream_on) return nullptr;
  static SideStream per_dev[64];
  static bool made[64] = {false};
  static bool failed[64] = {false};
  // MDCONV_B200_NO_SIDE_STREAM=1 keeps everything on the caller's stream (per-kernel event timings are
  // only meaningful that way: a forked kernel's events also count the time it waits for free SMs)
  static const bool disabled = [] { const char* e = getenv("MDCONV_B200_NO_SIDE_STREAM"); return e && e[0] == '1'; }();
  if (disabled) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (failed[dev]) return nullptr;
  if (!made[dev]) {
    SideStream s;
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) {
      failed[dev] = true;
      (void)cudaGetLastError();
      return nullptr;
    }
    per_dev[dev] = s;
    made[dev] = true;
  }
  return &per_dev[dev];
}

int sm_count() {
  static std::atomic<int> per_dev[64];   // zero-initialised; 0 = not queried yet
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  const bool tracked = dev >= 0 && dev < 64;
  int n = tracked ? per_dev[dev].load(std::memory_order_relaxed) : 0;
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (tracked) per_dev[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

template <typename T>
void launch_to_channels_last(const T* x, T* xt, int B, int C, int IN, LaunchCtx& ctx, bool zero_row, int out_ld) {
  ProfScope prof(ctx, out_ld > 0 ? "from_channels_last" : "to_channels_last");
  if (out_ld <= 0) out_ld = C;
  dim3 grid((unsigned)((IN + 63) / 64), (unsigned)((C + 63) / 64), (unsigned)B);
  const bool vec = sizeof(T) == 2 && IN % 8 == 0 && C % 8 == 0 && out_ld % 8 == 0 && ((uintptr_t)x & 15) == 0 &&
                   ((uintptr_t)xt & 15) == 0;   // the 16-byte form moves 8 two-byte elements per access
  if (vec) k_to_channels_last<T, true><<<grid, 256, 0, ctx.stream>>>(x, xt, C, IN, zero_row ? 1 : 0, out_ld);
  else k_to_channels_last<T, false><<<grid, 256, 0, ctx.stream>>>(x, xt, C, IN, zero_row ? 1 : 0, out_ld);
  MDCONV_CHECK_LAUNCH(ctx);
}
template <typename T>
void launch_reorder_weight(const T* w, void* wr, int Co, int Cg, int K, LaunchCtx& ctx) {
  ProfScope prof(ctx, "reorder_weight");
  const size_t n = (size_t)Co * Cg * K;
  const int blocks = (int)((n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256);
  if constexpr (std::is_same<T, float>::value) {
    // two bf16 banks, hi then lo, each align256(Co * Kg * 2) bytes (wr_bytes)
    __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(wr);
    __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(wr) + align256(n * 2));
    k_reorder_weight_split<<<blocks, 256, 0, ctx.stream>>>(w, hi, lo, Co, Cg, K);
  } else {
    k_reorder_weight<T><<<blocks, 256, 0, ctx.stream>>>(w, reinterpret_cast<T*>(wr), Co, Cg, K);
  }
  MDCONV_CHECK_LAUNCH(ctx);
}
template void launch_to_channels_last<__half>(const __half*, __half*, int, int, int, LaunchCtx&, bool, int);
template void launch_to_channels_last<__nv_bfloat16>(const __nv_bfloat16*, __nv_bfloat16*, int, int, int, LaunchCtx&,
                                                     bool, int);
template void launch_to_channels_last<float>(const float*, float*, int, int, int, LaunchCtx&, bool, int);
template void launch_reorder_weight<__half>(const __half*, void*, int, int, int, LaunchCtx&);
template void launch_reorder_weight<__nv_bfloat16>(const __nv_bfloat16*, void*, int, int, int, LaunchCtx&);
template void launch_reorder_weight<float>(const float*, void*, int, int, int, LaunchCtx&);

}  // namespace tc

static std::atomic<int> g_comm_sms{-1};   // -1: not set yet -> MDCONV_B200_COMM_SMS or 8
// fp32 calls take the split-bf16 tensor-core path by default; MDCONV_B200_FP32_TC=0 (or mdconv_set_fp32_tensor_cores)
// sends them back to the exact-fp32 SIMT kernels.
static std::atomic<int> g_fp32_tc{-1};
bool fp32_on_tensor_cores() {
  int v = g_fp32_tc.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("MDCONV_B200_FP32_TC");
    v = (e && e[0] == '0') ? 0 : 1;
    g_fp32_tc.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}
int set_fp32_on_tensor_cores(int on) {
  const int prev = fp32_on_tensor_cores() ? 1 : 0;
  g_fp32_tc.store(on ? 1 : 0, std::memory_order_relaxed);
  return prev;
}

int comm_spare_sms() {
  int v = g_comm_sms.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("MDCONV_B200_COMM_SMS");
    v = e ? atoi(e) : 8;
    v = v < 0 ? 0 : (v > 64 ? 64 : v);
    g_comm_sms.store(v, std::memory_order_relaxed);
  }
  return v;
}
int set_comm_spare_sms(int sms) {
  const int prev = comm_spare_sms();
  g_comm_sms.store(sms < 0 ? 0 : (sms > 64 ? 64 : sms), std::memory_order_relaxed);
  return prev;
}

bool tensor_path_supported(const DevGeom& g, int dtype, int backward) {
  if (dtype != MDCONV_F16 && dtype != MDCONV_BF16 && dtype != MDCONV_F32) return false;
  if (g.Cg % TC_BK != 0 || g.Cdg % TC_BK != 0) return false;
  const int vec = dtype == MDCONV_F32 ? 4 : 8;   // elements per 16-byte vector
  if (((long long)g.B * g.IN + 1) * (g.C / vec) > 2147483647LL) return false;
  if (g.cl && (dtype == MDCONV_F32 || g.Cog % 64 != 0 || g.Co % 8 != 0 || g.C % 8 != 0)) return false;   // 64-channel TMA store chunks
  if (dtype == MDCONV_F32 && !fp32_on_tensor_cores()) return false;
  if (backward) return tc::tensor_backward_supported(g, dtype);
  if (g.Cog % 32 != 0) return false;
  const int mx = dtype == MDCONV_F32 ? FwdCfg<float>::MAX_BN : TC_MAX_BN;
  if (g.Cog > mx && g.Cog % mx != 0) return false;
  return true;
}

size_t tensor_workspace_bytes(const DevGeom& g, int dtype, int backward) {
  (void)dtype;
  if (backward) return tc::tensor_backward_workspace(g);
  return tc::xt_bytes(g) + tc::wr_bytes(g);
}

void tensor_forward(const DevGeom& g, int dtype, const void* x, const void* w, const void* bias, const void* off,
                    const void* mask, void* y, void* ws, const FwdExtra& ex, LaunchCtx& ctx) {
  if (dtype == MDCONV_F16) tensor_forward_t<__half>(g, x, w, bias, off, mask, y, ws, ex, ctx);
  else if (dtype == MDCONV_BF16) tensor_forward_t<__nv_bfloat16>(g, x, w, bias, off, mask, y, ws, ex, ctx);
  else tensor_forward_t<float>(g, x, w, bias, off, mask, y, ws, ex, ctx);
}

size_t tensor_channels_last_bytes(const DevGeom& g) { return tc::xt_bytes(g); }
size_t tensor_weight_layout_bytes(const DevGeom& g) { return tc::wr_bytes(g); }
void tensor_to_channels_last(const DevGeom& g, int dtype, const void* x, void* xt, LaunchCtx& ctx) {
  if (dtype == MDCONV_F16)
    tc::launch_to_channels_last<__half>((const __half*)x, (__half*)xt, g.B, g.C, g.IN, ctx, /*zero_row=*/true);
  else if (dtype == MDCONV_BF16)
    tc::launch_to_channels_last<__nv_bfloat16>((const __nv_bfloat16*)x, (__nv_bfloat16*)xt, g.B, g.C, g.IN, ctx,
                                               /*zero_row=*/true);
  else
    tc::launch_to_channels_last<float>((const float*)x, (float*)xt, g.B, g.C, g.IN, ctx, /*zero_row=*/true);
}

void tensor_backward(const DevGeom& g, int dtype, const void* x, const void* w, const void* off, const void* mask,
                     const void* go, void* gx, void* gw, void* gb, void* goff, void* gmask, void* ws, const void* saved,
                     LaunchCtx& ctx) {
  if (dtype == MDCONV_F16) tc::tensor_backward_t<__half>(g, x, w, off, mask, go, gx, gw, gb, goff, gmask, ws, saved, ctx);
  else if (dtype == MDCONV_BF16)
    tc::tensor_backward_t<__nv_bfloat16>(g, x, w, off, mask, go, gx, gw, gb, goff, gmask, ws, saved, ctx);
  else tc::tensor_backward_t<float>(g, x, w, off, mask, go, gx, gw, gb, goff, gmask, ws, saved, ctx);
}

// forward workspace prefix the backward can reuse: [channels-last input | re-ordered weight]
size_t tensor_saved_bytes(const DevGeom& g, int dtype) {
  (void)dtype;
  return tc::xt_bytes(g) + tc::wr_bytes(g);
}

}  // namespace mdconv
