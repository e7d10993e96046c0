// tensor_bwd.cu -- tcgen05 backward kernels (fp16 / bf16).
//
// grad_weight:  gW[co, kk] = sum over positions of go[co, pos] * col[kk, pos]   (reference G3,
//   src/mdeformable_conv.cu:436-439, which reads a materialised 462 MB column buffer at cfg2).
//   Here it is one implicit GEMM with the reduction over positions split across CTAs:
//     A = go tile [128 co x 64 pos], K-major, straight from the NCHW grad_output by TMA (3D map);
//     B = re-gathered column tile [64 pos x 64*NB kk], MN-major (the same swizzled 128-byte rows
//         the forward kernel produces, read by the tensor core "transposed");
//     D = fp32 [co x kk] in TMEM, written once per CTA as a partial; a second kernel folds the
//         partials in a fixed order (deterministic) and un-permutes kk -> (channel, tap).
//
// This file also orchestrates the whole tensor-path backward (tensor_backward_t): grad_weight + grad_bias first
// (the weight-ready event is recorded after them), then grad_offset / grad_mask (tensor_bwd_sample.cu) and
// grad_input (tensor_bwd_input.cu); shapes those two refuse fall back to the generic data-gradient kernel.
#include <string.h>

#include "tc_common.cuh"

namespace mdconv {
namespace tc {

namespace {

constexpr int GW_PLAN_WARP0 = 8;     // warps 8-11: plan producers
constexpr int GW_GATHER_WARP0 = 12;  // warps 12-27: gather producers
constexpr int GW_GATHER_WARPS = 16;
constexpr int GW_GATHER_THREADS = GW_GATHER_WARPS * 32;
constexpr int GW_THREADS = (GW_GATHER_WARP0 + GW_GATHER_WARPS) * 32;  // 896
#ifndef MDCONV_GW_STAGES
#define MDCONV_GW_STAGES 3
#endif
constexpr int GW_STAGES = MDCONV_GW_STAGES;
constexpr int GW_PLAN_SLOTS = 4;
constexpr int GW_NB = 2;                       // 64-wide kk blocks per CTA tile (N = 128)
constexpr int GW_GO_BYTES = 2 * 128 * 128;     // two M tiles of [128 co x 64 pos]
constexpr int GW_COL_BYTES = GW_NB * 64 * 128; // NB blocks of [64 pos x 64 kk]
// registers after setmaxnreg (the CTA owns GW_THREADS x 72)
constexpr uint32_t GW_REGS_CTL = 32, GW_REGS_EPI = 56, GW_REGS_PLAN = 56, GW_REGS_GATHER = 88;
static_assert(128 * (GW_REGS_CTL + GW_REGS_EPI + GW_REGS_PLAN) + GW_GATHER_THREADS * GW_REGS_GATHER <= GW_THREADS * 72,
              "setmaxnreg.inc would wait forever");

// fp32 callers: both operands as bf16 hi / lo pairs, three products per K step (see tensor_path.cu, FwdCfg).
// The grad_output tile comes from two pre-split bf16 copies (one 128-row M tile per CTA instead of two), the
// column tile is interpolated in fp32 and split on its way into shared memory.
template <typename T> struct GwCfg {
  static constexpr bool F32 = false;
  static constexpr int COL_STAGE = GW_COL_BYTES, M_TILES = 2;
};
template <> struct GwCfg<float> {
  static constexpr bool F32 = true;
  static constexpr int COL_STAGE = 2 * GW_COL_BYTES, M_TILES = 1;
};
__device__ __forceinline__ void gw_split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t* h = reinterpret_cast<uint32_t*>(&hi);
  uint32_t* l = reinterpret_cast<uint32_t*>(&lo);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    const float2 back = __bfloat1622float2(hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * i] - back.x, v[2 * i + 1] - back.y);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
}

struct TcGwParams {
  DevGeom g;
  const void* xt;
  const void* off;
  const void* mask;
  float* part;   // [S][Co][Kg] fp32, kk in (channel-block, tap, 64 ch) order
  int nblocks;   // 64-wide kk blocks per group = (Cg/64) * K
  int n_tiles;   // ceil(nblocks / GW_NB)
  int m_pairs;   // ceil(Cog / 256)
  int S;         // splits of the position reduction
  int pbi;       // 64-position blocks per image
  int total_pb;  // B * pbi
  int zero;      // always 0 at run time (tc::drain_point)
};

template <typename T, int ND, bool MOD>
__global__ void __launch_bounds__(GW_THREADS, 1)
k_tc_grad_weight(const __grid_constant__ CUtensorMap tmap_go, const __grid_constant__ CUtensorMap tmap_go2,
                 const TcGwParams prm) {
  constexpr bool F32 = GwCfg<T>::F32;
  constexpr int COL_STAGE = GwCfg<T>::COL_STAGE;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sGo = smem;   // fp32: [hi M tile | lo M tile], else two M tiles
  uint8_t* sCol = sGo + GW_STAGES * GW_GO_BYTES;
  PlanEntry<ND>* sPlan = reinterpret_cast<PlanEntry<ND>*>(sCol + GW_STAGES * COL_STAGE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sPlan) +
                                               GW_PLAN_SLOTS * 128 * sizeof(PlanEntry<ND>));
  uint64_t* full_go = bars;
  uint64_t* full_col = full_go + GW_STAGES;
  uint64_t* empty = full_col + GW_STAGES;
  uint64_t* plan_full = empty + GW_STAGES;
  uint64_t* plan_empty = plan_full + GW_PLAN_SLOTS;
  uint64_t* done = plan_empty + GW_PLAN_SLOTS;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);

  const DevGeom& g = prm.g;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work item
  int w = blockIdx.x;
  const int split = w % prm.S; w /= prm.S;
  const int nt = w % prm.n_tiles; w /= prm.n_tiles;
  const int mp = w % prm.m_pairs;
  const int grp = w / prm.m_pairs;
  const int nb = min(GW_NB, prm.nblocks - nt * GW_NB);     // kk blocks in this tile
  const int co_per_cta = 128 * GwCfg<T>::M_TILES;
  const int co_left = g.Cog - mp * co_per_cta;
  const int MT = (!F32 && co_left > 128) ? 2 : 1;            // 128-row M tiles in this CTA
  const int pb0 = (int)((long long)prm.total_pb * split / prm.S);
  const int pb1 = (int)((long long)prm.total_pb * (split + 1) / prm.S);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_go);
    if (F32) ptx::prefetch_tensormap(&tmap_go2);
    for (int s = 0; s < GW_STAGES; ++s) {
      ptx::mbar_init(&full_go[s], 1);
      ptx::mbar_init(&full_col[s], ptx::arrivals(GW_GATHER_THREADS));
      ptx::mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < GW_PLAN_SLOTS; ++s) {
      ptx::mbar_init(&plan_full[s], ptx::arrivals(128));
      ptx::mbar_init(&plan_empty[s], ptx::arrivals(GW_GATHER_THREADS));
    }
    ptx::mbar_init(done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<256>(tmem_ptr);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < 4) {
    ptx::reg_dealloc<GW_REGS_CTL>();
  if (warp == 0) {
    // ===================== TMA producer: grad_output tiles =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const int co_base = grp * g.Cog + mp * co_per_cta;
      for (int pb = pb0; pb < pb1; ++pb) {
        const int b = pb / prm.pbi, p0 = (pb - b * prm.pbi) * 64;
        ptx::mbar_wait_relaxed(&empty[s], ph ^ 1, 64);
        ptx::mbar_arrive_expect_tx(&full_go[s], (uint32_t)(F32 ? 2 : MT) * 16384u);
        for (int mt = 0; mt < MT; ++mt)
          ptx::tma_load_3d(sGo + s * GW_GO_BYTES + mt * 16384, &tmap_go, &full_go[s], p0, co_base + mt * 128, b);
        if (F32) ptx::tma_load_3d(sGo + s * GW_GO_BYTES + 16384, &tmap_go2, &full_go[s], p0, co_base, b);   // lo halves
        if (++s == GW_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_f16(F32 ? 1 : umma_fmt<T>(), /*A K-major*/ 0, /*B MN-major*/ 1, 128, nb * 64);
      int s = 0;
      uint32_t ph = 0;
      for (int pb = pb0; pb < pb1; ++pb) {
        ptx::mbar_wait(&full_go[s], ph);
        ptx::mbar_wait(&full_col[s], ph);
        ptx::tc_fence_after_sync();
        const uint32_t go_addr = ptx::smem_u32(sGo + s * GW_GO_BYTES);
        const uint32_t col_addr = ptx::smem_u32(sCol + s * COL_STAGE);
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 16 positions per MMA
            const uint64_t ad = ptx::umma_smem_desc_sw128(go_addr + mt * 16384 + k * 32, 0, 1024);
            const uint64_t bd = ptx::umma_smem_desc_sw128(col_addr + k * 2048, 8192, 1024);
            ptx::umma_f16(tmem_base + (uint32_t)(mt * 128), ad, bd, idesc, (pb > pb0 || k > 0) ? 1u : 0u);
            if (F32) {   // + hi*lo + lo*hi
              const uint64_t al = ptx::umma_smem_desc_sw128(go_addr + 16384 + k * 32, 0, 1024);
              const uint64_t bl = ptx::umma_smem_desc_sw128(col_addr + GW_COL_BYTES + k * 2048, 8192, 1024);
              ptx::umma_f16(tmem_base, ad, bl, idesc, 1u);
              ptx::umma_f16(tmem_base, al, bd, idesc, 1u);
            }
          }
        }
        ptx::umma_commit(&empty[s]);
        if (++s == GW_STAGES) { s = 0; ph ^= 1; }
      }
      ptx::umma_commit(done);
    }
  }
  } else if (warp < GW_GATHER_WARP0) {
    // ===================== plan producers (warps 4-11), then epilogue (warps 4-7) =====================
    // Two teams of 128 threads take alternate position blocks (team 0 = warps 8-11: even blocks, team 1 =
    // warps 4-7: odd blocks), so two plans are always being built; a single in-order team is the
    // bottleneck otherwise (a plan is ~150 dependent instructions and is consumed once).
    ptx::reg_dealloc<GW_REGS_PLAN>();
    {
      const int team = warp < 8 ? 1 : 0;
      const int pt = threadIdx.x & 127;
      const T* __restrict__ off = reinterpret_cast<const T*>(prm.off);
      const T* __restrict__ mask = reinterpret_cast<const T*>(prm.mask);
      const int C8 = F32 ? g.C >> 2 : g.C >> 3;   // 16-byte vectors per position
      const int prow = pt & 63, pblk = pt >> 6;
      const int kkb = nt * GW_NB + pblk;
      const int pcc = kkb / g.K, ptap = kkb - pcc * g.K;
      const int pdg = (grp * g.Cg + pcc * BK) / g.Cdg;
      const bool active = pblk < nb;
      // thread (blk, row): the sampling plan of position p0 + row for this CTA's kk block `blk`; the
      // offset / mask values of the team's NEXT position block are fetched while this one is built.
      auto coord = [&](int pb, RowCoord<ND>& rc) {
        const int b = pb / prm.pbi, p0 = (pb - b * prm.pbi) * 64;
        decode_row<ND>(g, b, p0 + prow, active && (p0 + prow < g.P), rc);
      };
      RowCoord<ND> rc_next;
      RawSample<ND> raw_next;
      const int first = pb0 + team;
      if (first < pb1) {
        coord(first, rc_next);
        if (rc_next.valid) raw_next = fetch_sample<T, ND, MOD>(g, off, mask, rc_next, pdg, ptap);
      }
      for (int pb = first; pb < pb1; pb += 2) {
        const int it = pb - pb0;                      // iteration index -> ring slot and parity
        const int ps = it % GW_PLAN_SLOTS;
        const uint32_t pph = (uint32_t)(it / GW_PLAN_SLOTS) & 1u;
        const RowCoord<ND> rc = rc_next;
        const RawSample<ND> raw = raw_next;
        if (pb + 2 < pb1) {
          coord(pb + 2, rc_next);
          if (rc_next.valid) raw_next = fetch_sample<T, ND, MOD>(g, off, mask, rc_next, pdg, ptap);
        }
        SamplePos<ND> sp;
        sp.in_range = false;
        if (rc.valid) build_pos<ND>(g, rc, ptap, raw, sp);
        const PlanEntry<ND> e = plan_from_pos<T, ND>(g, rc, sp, C8);
        ptx::mbar_wait_relaxed(&plan_empty[ps], pph ^ 1, 128);
        store_plan<ND>(ptx::smem_u32(sPlan + ps * 128 + pt), e);
        ptx::mbar_arrive_warp(&plan_full[ps]);
      }
    }
    if (warp < 8) {
      // ---- epilogue: TMEM -> fp32 partial ----
      const int ew = warp & 3;
      ptx::mbar_wait_relaxed(done, 0, 1024);
      ptx::tc_fence_after_sync();
      const int N = nb * 64;
      for (int mt = 0; mt < MT; ++mt) {
        const int co = mp * co_per_cta + mt * 128 + ew * 32 + lane;  // within the group
        const bool valid = co < g.Cog;
        float* dst = prm.part + ((size_t)split * g.Co + (size_t)grp * g.Cog + (valid ? co : 0)) * g.Kg +
                     (size_t)nt * GW_NB * 64;
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(mt * 128);
        for (int c0 = 0; c0 < N; c0 += 16) {
          uint32_t v[16];
          ptx::tmem_ld_32x16(taddr + c0, v);
          ptx::tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(dst + c0 + j) =
                  make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                              __uint_as_float(v[j + 3]));
          }
        }
      }
    }
  } else {
    // ===================== gather producers: column tiles =====================
    ptx::reg_alloc<GW_REGS_GATHER>();
    // 128 plan rows = NB blocks x 64 positions; warp gw owns rows gw*8 .. gw*8+7 (two passes of 4).
    // One position block of look-ahead: the corner loads of block pb+1 are in flight while block pb is
    // interpolated and stored (ping-pong register buffers; 3D has no room for the second buffer).
    constexpr int NC = 1 << ND;
    const int gw = warp - GW_GATHER_WARP0;
    const int chunk = lane & 7;
    const int q0 = gw * 8 + (lane >> 3);     // plan row of pass 0 (pass 1: +4)
    const int jb = q0 >> 6;                  // kk block of this warp
    const int r0 = q0 & 63;                  // position row inside the block
    const uint4* __restrict__ xq = reinterpret_cast<const uint4*>(prm.xt);
    const int c0_blk = grp * g.Cg + ((nt * GW_NB + jb) / g.K) * BK;
    const uint4* xc = F32 ? xq + (c0_blk >> 2) + 2 * chunk : xq + (c0_blk >> 3) + chunk;
    const bool active = jb < nb;
    const uint32_t plan_row = ptx::smem_u32(sPlan + q0);
    constexpr uint32_t kSlotBytes = 128 * (uint32_t)sizeof(PlanEntry<ND>);
    constexpr uint32_t kRow4 = 4 * (uint32_t)sizeof(PlanEntry<ND>);
    int s = 0, ips = 0, cps = 0;
    uint32_t ph = 0, ipph = 0;

    auto issue = [&](uint4 (&v0)[NC], uint4 (&v1)[NC], uint32_t after) {
      if constexpr (F32) return;
      ptx::mbar_wait(&plan_full[ips], ipph);
      if (active) {
        const uint32_t pa = plan_row + ips * kSlotBytes + after;
        int i0[NC], i1[NC];
#pragma unroll
        for (int q = 0; q < NC / 4; ++q) {
          const uint4 a = lds128(pa + 16 * q), b = lds128(pa + kRow4 + 16 * q);
          i0[4 * q] = (int)a.x; i0[4 * q + 1] = (int)a.y; i0[4 * q + 2] = (int)a.z; i0[4 * q + 3] = (int)a.w;
          i1[4 * q] = (int)b.x; i1[4 * q + 1] = (int)b.y; i1[4 * q + 2] = (int)b.z; i1[4 * q + 3] = (int)b.w;
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) v0[c] = __ldg(xc + i0[c]);
#pragma unroll
        for (int c = 0; c < NC; ++c) v1[c] = __ldg(xc + i1[c]);
      }
      if (++ips == GW_PLAN_SLOTS) { ips = 0; ipph ^= 1; }
    };
    auto finish = [&](const uint4 (&v0)[NC], const uint4 (&v1)[NC], uint32_t after = 0u) {
      ptx::mbar_wait(&empty[s], ph ^ 1);
      if (active) {
        const uint32_t pa = plan_row + cps * kSlotBytes + 4 * NC + after;
        uint32_t w0[NC], w1[NC];
#pragma unroll
        for (int q = 0; q < NC / 4; ++q) {
          const uint4 a = lds128(pa + 16 * q), b = lds128(pa + kRow4 + 16 * q);
          w0[4 * q] = a.x; w0[4 * q + 1] = a.y; w0[4 * q + 2] = a.z; w0[4 * q + 3] = a.w;
          w1[4 * q] = b.x; w1[4 * q + 1] = b.y; w1[4 * q + 2] = b.z; w1[4 * q + 3] = b.w;
        }
        if constexpr (!F32) {
          const uint32_t col_addr = ptx::smem_u32(sCol + s * COL_STAGE) + jb * 8192;
          sts128(col_addr + sw128(r0, chunk), interp8(v0, w0, NC, T()));
          sts128(col_addr + sw128(r0 + 4, chunk), interp8(v1, w1, NC, T()));
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive_warp(&full_col[s]);
      ptx::mbar_arrive_warp(&plan_empty[cps]);
      if (++s == GW_STAGES) { s = 0; ph ^= 1; }
      if (++cps == GW_PLAN_SLOTS) cps = 0;
    };

    // fp32: one position block start to finish (no look-ahead); per row: load, interpolate in fp32, split, store
    auto block_f32 = [&]() {
      ptx::mbar_wait(&plan_full[ips], ipph);
      ptx::mbar_wait(&empty[s], ph ^ 1);
      if (active) {
        const uint32_t pa = plan_row + ips * kSlotBytes;
        const uint32_t col_addr = ptx::smem_u32(sCol + s * COL_STAGE) + jb * 8192;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          float acc[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
          for (int q = 0; q < NC / 4; ++q) {
            const uint4 iv = lds128(pa + r * kRow4 + 16 * q);
            const uint4 wv = lds128(pa + r * kRow4 + 4 * NC + 16 * q);
            const int idx[4] = {(int)iv.x, (int)iv.y, (int)iv.z, (int)iv.w};
            const float wgt[4] = {__uint_as_float(wv.x), __uint_as_float(wv.y), __uint_as_float(wv.z),
                                  __uint_as_float(wv.w)};
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
          gw_split8(acc, hi, lo);
          sts128(col_addr + sw128(r0 + 4 * r, chunk), hi);
          sts128(col_addr + GW_COL_BYTES + sw128(r0 + 4 * r, chunk), lo);
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive_warp(&full_col[s]);
      ptx::mbar_arrive_warp(&plan_empty[ips]);
      if (++s == GW_STAGES) { s = 0; ph ^= 1; }
      if (++ips == GW_PLAN_SLOTS) { ips = 0; ipph ^= 1; }
    };

    if (F32) {
      for (int pb = pb0; pb < pb1; ++pb) block_f32();
    } else if (ND == 2) {
      uint4 a0[NC], a1[NC], b0[NC], b1[NC];
      int pb = pb0;
      // wait for the block about to be interpolated BEFORE the next block's loads are issued: every LDG of the loop
      // retires through one scoreboard (see the forward kernel's gather loop, tensor_path.cu)
      if (pb < pb1) issue(a0, a1, 0u);
      while (pb < pb1) {
        const uint32_t da = tc::drain_point(a1[NC - 1].w, prm.zero);
        if (pb + 1 < pb1) issue(b0, b1, da);
        finish(a0, a1, da);
        if (++pb >= pb1) break;
        const uint32_t db = tc::drain_point(b1[NC - 1].w, prm.zero);
        if (pb + 1 < pb1) issue(a0, a1, db);
        finish(b0, b1, db);
        ++pb;
      }
    } else {
      uint4 a0[NC], a1[NC];
      for (int pb = pb0; pb < pb1; ++pb) {
        issue(a0, a1, 0u);
        finish(a0, a1);
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<256>(tmem_base);
}

// gW[co][c][tap] = sum_s part[s][co][((c/64)*K + tap)*64 + c%64]   (fixed order => deterministic)
template <typename T>
__global__ void k_reduce_grad_weight(const float* __restrict__ part, T* __restrict__ gw, float* __restrict__ gw32, int S,
                                     int Co, int Cg, int K) {
  const size_t n = (size_t)Co * Cg * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int Kg = Cg * K;
    const int co = (int)(i / Kg);
    const int r = (int)(i % Kg);
    const int c = r / K, tap = r - c * K;
    const size_t src = (size_t)co * Kg + (size_t)((c >> 6) * K + tap) * 64 + (c & 63);
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc += part[(size_t)s * n + src];
    gw[i] = cvt_to<T, float>(acc);
    if (gw32 != nullptr) gw32[i] = acc;   // un-rounded copy for a data-parallel caller's fp32 all-reduce
  }
}

// grad_output rows [B*Co][P] -> [B*Co][ld] with a zero tail (ld = P rounded up to 8 elements): TMA needs every
// global stride to be a multiple of 16 bytes, which NCHW rows of 14x14 / 7x7 / odd maps are not.
template <typename T>
__global__ void __launch_bounds__(256) k_pad_rows(const T* __restrict__ src, T* __restrict__ dst, size_t rows, int P,
                                                  int ld) {
  const size_t n = rows * (size_t)ld;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / ld;
    const int p = (int)(i - r * ld);
    dst[i] = p < P ? src[r * P + p] : cvt_to<T, float>(0.f);
  }
}

// fp32 grad_output [rows][P] -> two bf16 copies [rows][ld] (hi, lo), zero tail
__global__ void __launch_bounds__(256) k_split_rows(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi,
                                                    __nv_bfloat16* __restrict__ lo, size_t rows, int P, int ld) {
  const size_t n = rows * (size_t)ld;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / ld;
    const int p = (int)(i - r * ld);
    const float v = p < P ? src[r * P + p] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

static int gw_co_per_cta(const DevGeom& g) { return g.esize == 4 ? 128 : 256; }

int gw_splits(const DevGeom& g) {
  const int nblocks = (g.Cg / BK) * g.K;
  const int n_tiles = (nblocks + GW_NB - 1) / GW_NB;
  const int m_pairs = (g.Cog + gw_co_per_cta(g) - 1) / gw_co_per_cta(g);
  const int items = g.g * m_pairs * n_tiles;
  const int pbi = (g.P + 63) / 64;
  const long long total_pb = (long long)g.B * pbi;
  int S = sm_count() / items;  // one wave: S * items <= number of SMs
  if (S < 1) S = 1;
  if (S > 32) S = 32;
  if ((long long)S > total_pb) S = (int)total_pb;
  return S;
}

template <typename T, int ND>
size_t gw_smem_bytes() {
  return 1024 + (size_t)GW_STAGES * (GW_GO_BYTES + GwCfg<T>::COL_STAGE) +
         (size_t)GW_PLAN_SLOTS * 128 * sizeof(PlanEntry<ND>) + (3 * GW_STAGES + 2 * GW_PLAN_SLOTS + 1) * sizeof(uint64_t) + 16;
}

template <typename T, int ND, bool MOD>
void launch_grad_weight(const CUtensorMap& tmap, const CUtensorMap& tmap2, const TcGwParams& prm, int grid,
                        LaunchCtx& ctx) {
  const size_t smem = gw_smem_bytes<T, ND>();
  static std::atomic<unsigned long long> configured{0};
  opt_in_dynamic_smem(k_tc_grad_weight<T, ND, MOD>, smem, configured, ctx);
  ProfScope prof(ctx, "tc_grad_weight");
  k_tc_grad_weight<T, ND, MOD><<<grid, GW_THREADS, smem, ctx.stream>>>(tmap, tmap2, prm);
  MDCONV_CHECK_LAUNCH(ctx);
}

}  // namespace

static bool data_on_tensor_cores(const DevGeom& g);
bool tensor_backward_supported(const DevGeom& g, int dtype) {
  (void)dtype;
  // maps with P % 8 != 0 go through a row-padded copy of grad_output (go_ld); channels-last callers need every
  // gradient on the tensor kernels (the SIMT fallback of the data gradients reads and writes NCHW)
  return g.cl ? data_on_tensor_cores(g) : true;
}

// leading dimension (elements) of the grad_output rows the TMA maps see, and the bytes of the padded copy
int go_ld(const DevGeom& g) { return (g.P + 7) & ~7; }
static size_t go_pad_bytes(const DevGeom& g) {
  if (g.esize == 4) return 2 * align256((size_t)g.B * g.Co * go_ld(g) * 2);   // fp32: bf16 hi copy + bf16 lo copy
  return (g.P % 8 == 0 && !g.cl) ? 0 : align256((size_t)g.B * g.Co * go_ld(g) * 2);
}

// grad_offset / grad_mask / grad_input on the tensor kernels: callers (16-bit, or fp32 as split bf16) whose shapes
// those two kernels take; everything else keeps the SIMT data-gradient kernel
static bool data_on_tensor_cores(const DevGeom& g) {
  return (g.esize == 2 || g.esize == 4) && grad_sample_supported(g) && grad_input_supported(g);
}

size_t tensor_backward_workspace(const DevGeom& g) {
  const size_t part = align256((size_t)gw_splits(g) * g.Co * g.Kg * sizeof(float));
  const size_t data = data_on_tensor_cores(g) ? grad_input_workspace(g)
                                              : generic_workspace_bytes(g, g.esize == 4 ? MDCONV_F32 : MDCONV_F16, 1);
  return xt_bytes(g) + wr_bytes(g) + part + go_pad_bytes(g) + data;
}

template <typename T>
void tensor_backward_t(const DevGeom& g, const void* x, const void* w, const void* off, const void* mask,
                       const void* go, void* gx, void* gw, void* gb, void* goff, void* gmask, void* ws, const void* saved,
                       LaunchCtx& ctx) {
  constexpr bool F32 = std::is_same<T, float>::value;
  char* base = reinterpret_cast<char*>(ws);
  const T* xt = reinterpret_cast<const T*>(base);
  base += xt_bytes(g);
  const void* wr = base;
  base += wr_bytes(g);
  if (saved != nullptr) {
    // the forward of this very call pair left both layouts in its workspace (same offsets: tensor_forward_t)
    xt = reinterpret_cast<const T*>(saved);
    wr = reinterpret_cast<const char*>(saved) + xt_bytes(g);
  }
  if (g.cl) xt = reinterpret_cast<const T*>(x);   // channels-last callers: the input itself is the gather source
  const int S = gw_splits(g);
  float* part = reinterpret_cast<float*>(base);
  base += align256((size_t)S * g.Co * g.Kg * sizeof(float));
  // grad_output as the TMA maps read it: the caller's tensor, or a copy with rows padded to 16 bytes
  const int ld = go_ld(g);
  const void* go_tma = go;
  const void* go_tma_lo = nullptr;
  if constexpr (F32) {
    // fp32: two bf16 copies (hi, lo) with padded rows; the bias gradient still reads the caller's fp32 tensor
    __nv_bfloat16* gh = reinterpret_cast<__nv_bfloat16*>(base);
    __nv_bfloat16* gl = reinterpret_cast<__nv_bfloat16*>(base + go_pad_bytes(g) / 2);
    base += go_pad_bytes(g);
    ProfScope prof(ctx, "split_grad_output");
    const size_t rows = (size_t)g.B * g.Co, n = rows * ld;
    const int blocks = (int)((n + 255) / 256 > 148 * 8 ? 148 * 8 : (n + 255) / 256);
    k_split_rows<<<blocks, 256, 0, ctx.stream>>>((const float*)go, gh, gl, rows, g.P, ld);
    MDCONV_CHECK_LAUNCH(ctx);
    go_tma = gh;
    go_tma_lo = gl;
  } else if (g.cl) {
    // channels-last grad_output [B][P][Co] -> rows [B][Co][ld] for the two kernels that read it by TMA as a
    // [channels x positions] operand (grad_weight, grad_sample) and for the bias gradient
    T* gp = reinterpret_cast<T*>(base);
    base += go_pad_bytes(g);
    launch_to_channels_last<T>((const T*)go, gp, g.B, g.P, g.Co, ctx, false, ld);
    go_tma = gp;
  } else if (g.P % 8 != 0) {
    T* gp = reinterpret_cast<T*>(base);
    base += go_pad_bytes(g);
    ProfScope prof(ctx, "pad_grad_output");
    const size_t rows = (size_t)g.B * g.Co, n = rows * ld;
    const int blocks = (int)((n + 255) / 256 > 148 * 8 ? 148 * 8 : (n + 255) / 256);
    k_pad_rows<T><<<blocks, 256, 0, ctx.stream>>>((const T*)go, gp, rows, g.P, ld);
    MDCONV_CHECK_LAUNCH(ctx);
    go_tma = gp;
  }
  void* data_ws = base;
  const int dtype = F32 ? MDCONV_F32 : (std::is_same<T, __half>::value ? MDCONV_F16 : MDCONV_BF16);
  const bool data_tc = data_on_tensor_cores(g);

  // the inverse index of grad_input only needs offset/mask: build it on the side stream, beside the
  // layout passes and ahead of the big kernels
  if (data_tc) tensor_grad_input_prepare<T>(g, off, mask, data_ws, ctx);
  if (saved == nullptr) {
    if (!g.cl) launch_to_channels_last<T>((const T*)x, const_cast<T*>(xt), g.B, g.C, g.IN, ctx, /*zero_row=*/true);
    if (data_tc) launch_reorder_weight<T>((const T*)w, const_cast<void*>(wr), g.Co, g.Cg, g.K, ctx);
  }

  // ---- grad_weight first (it is what a data-parallel caller all-reduces) ----
  TcGwParams prm;
  prm.g = g;
  prm.xt = xt; prm.off = off; prm.mask = mask; prm.part = part;
  prm.nblocks = (g.Cg / BK) * g.K;
  prm.n_tiles = (prm.nblocks + GW_NB - 1) / GW_NB;
  prm.m_pairs = (g.Cog + gw_co_per_cta(g) - 1) / gw_co_per_cta(g);
  prm.S = S;
  prm.pbi = (g.P + 63) / 64;
  prm.total_pb = g.B * prm.pbi;
  prm.zero = 0;
  const int grid = g.g * prm.m_pairs * prm.n_tiles * S;

  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    if (ctx.err == cudaSuccess) ctx.err = cudaErrorNotSupported;
    return;
  }
  using E = typename std::conditional<F32, __nv_bfloat16, T>::type;   // element type of the TMA-read copies
  CUtensorMap tmap, tmap2;
  memset(&tmap2, 0, sizeof(tmap2));
  const cuuint64_t dims[3] = {(cuuint64_t)g.P, (cuuint64_t)g.Co, (cuuint64_t)g.B};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(E), (cuuint64_t)g.Co * ld * sizeof(E)};
  const cuuint32_t box[3] = {64, 128, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&tmap, tmap_dtype<E>(), 3, const_cast<void*>(go_tma), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS && F32)
    r = enc(&tmap2, tmap_dtype<E>(), 3, const_cast<void*>(go_tma_lo), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (ctx.err == cudaSuccess) ctx.err = cudaErrorInvalidValue;
    return;
  }
  if (g.nd == 2) {
    if (g.modulated) launch_grad_weight<T, 2, true>(tmap, tmap2, prm, grid, ctx);
    else launch_grad_weight<T, 2, false>(tmap, tmap2, prm, grid, ctx);
  } else {
    if (g.modulated) launch_grad_weight<T, 3, true>(tmap, tmap2, prm, grid, ctx);
    else launch_grad_weight<T, 3, false>(tmap, tmap2, prm, grid, ctx);
  }
  {
    ProfScope prof(ctx, "reduce_grad_weight");
    const size_t n = (size_t)g.Co * g.Kg;
    const int blocks = (int)((n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256);
    k_reduce_grad_weight<T><<<blocks, 256, 0, ctx.stream>>>(part, (T*)gw, ctx.gw32, S, g.Co, g.Cg, g.K);
    MDCONV_CHECK_LAUNCH(ctx);
  }

  // grad_bias next: with it both parameter gradients are final and the caller's all-reduce may start
  ctx.go_ld = F32 ? 0 : ld;
  generic_backward(g, dtype, x, w, off, mask, F32 ? go : go_tma, gx, gw, gb, goff, gmask, nullptr, ctx, GENERIC_BWD_BIAS);
  ctx.go_ld = 0;
  signal_weight_ready(ctx);

  // ---- grad_offset / grad_mask / grad_input ----
  if (data_tc) {
    // 16-bit callers: chunk by chunk, grad_sample also streams grad_columns into an L2-sized buffer and grad_input is
    // a gather over it; otherwise grad_input runs its own contraction (tensor_bwd_input.cu)
    void* gcol = grad_input_gcol(g, data_ws);
    if (gcol != nullptr) {
      const int step = gcol_chunk_samples(g);
      for (int b0 = 0; b0 < g.B; b0 += step) {
        const int nb = g.B - b0 < step ? g.B - b0 : step;
        tensor_grad_sample<T>(g, xt, wr, off, mask, go_tma, go_tma_lo, ld, goff, gmask, gcol, b0, nb, ctx);
        tensor_grad_input_from_gcol<T>(g, gx, data_ws, b0, nb, ctx);
      }
      return;
    }
    tensor_grad_sample<T>(g, xt, wr, off, mask, go_tma, go_tma_lo, ld, goff, gmask, nullptr, 0, g.B, ctx);
    tensor_grad_input<T>(g, w, go, gx, data_ws, ctx);
    return;
  }
  generic_backward(g, dtype, x, w, off, mask, go, gx, gw, gb, goff, gmask, data_ws, ctx, GENERIC_BWD_DATA);
}

template void tensor_backward_t<__half>(const DevGeom&, const void*, const void*, const void*, const void*,
                                        const void*, void*, void*, void*, void*, void*, void*, const void*, LaunchCtx&);
template void tensor_backward_t<__nv_bfloat16>(const DevGeom&, const void*, const void*, const void*, const void*,
                                               const void*, void*, void*, void*, void*, void*, void*, const void*,
                                               LaunchCtx&);
template void tensor_backward_t<float>(const DevGeom&, const void*, const void*, const void*, const void*,
                                       const void*, void*, void*, void*, void*, void*, void*, const void*, LaunchCtx&);

}  // namespace tc
}  // namespace mdconv
