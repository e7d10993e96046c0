// tensor_bwd_sample.cu -- tcgen05 kernel for grad_offset and grad_mask (fp16 / bf16).
//
// Reference: the col2im_coord / mask-gradient half of the fused gradient kernels
// (src/mdeformable_conv.cu:295-315, src/deformable_conv.cu:262-286, 3D twins), which run one
// thread per column element and resolve the sum over channels with 2-3 atomicAdd per element onto
// C/dg-way contended addresses, fed by a materialised grad_columns buffer (G2, :417-419).
//
// Here:   gcol[pos, kk] = sum_co go[pos, co] * W[co, kk]           (tcgen05.mma, accumulators in TMEM)
//         S_c[pos, tap] = sum_ch gcol[pos, (ch, tap)] * x[corner_c(pos, tap), ch]
//         grad_mask     = sum_c w_c * S_c
//         grad_offset_a = mask * sum_c (d w_c / d p_a) * S_c
// so the columns never exist in HBM, there are no atomics, and every output element is written exactly
// once in a fixed summation order (deterministic).  gcol is consumed from shared memory here; for 16-bit
// callers it is ALSO streamed to HBM once (the reference's grad_columns), for the col2im gather of grad_input.
//
// Work decomposition: one item = (tile of 128 positions, (deformable group, tap) phase); every phase writes its
// own outputs, so the m_tiles * dg * K items are dealt to the persistent CTAs as equal contiguous ranges (a whole
// number of tiles per CTA would leave the last wave 30 % full at cfg2: 784 tiles on 148 SMs).
//
//   warp 0       TMA producer: grad_output tile (MN-major A operand, straight from NCHW) and weight
//                slices (MN-major B operand), 32 output channels per K-block -> 4-slot smem ring
//   warp 1       MMA issuer, M=128 positions x N=64*nb channels of one tap, K = output channels
//   warp 2       TMEM allocator (2 accumulator stages x 256 columns); 16-bit callers: also the grad_columns storer
//                (every [128 positions x 64 channels] block the movers lay out leaves for HBM by cp.async.bulk.tensor,
//                what k_col2im_gather turns into grad_input: tensor_bwd_input.cu)
//   warps 4-7    "movers": compute the per-(position, tap) sampling entry (corner indices, fractions,
//                mask) for the next phase; tcgen05.ld the accumulator, round it to the storage dtype and
//                re-lay it out in shared memory as [position][64 channels] (128B-swizzled rows)
//   warps 8-23   dot warps: 8 lanes per position, 16-byte channels-last loads of the corner values
//                (all in flight together), dot products against gcol (HFMA2 quads folded into fp32 for
//                fp16, fp32 FMAs for bf16), warp-shuffle reduction over the 8 lanes, final combine and
//                store of grad_offset / grad_mask
#include <string.h>

#include "tc_common.cuh"

namespace mdconv {
namespace tc {

namespace {

constexpr int GS_DOT_WARP0 = 8;
constexpr int GS_DOT_WARPS = 16;
constexpr int GS_DOT_THREADS = GS_DOT_WARPS * 32;
constexpr int GS_THREADS = (GS_DOT_WARP0 + GS_DOT_WARPS) * 32;  // 768
#ifndef MDCONV_GS_STAGES
#define MDCONV_GS_STAGES 3
#endif
constexpr int GS_STAGES = MDCONV_GS_STAGES;
constexpr int GS_KCO = 32;                // output channels per K-block
constexpr int GS_A_BYTES = 2 * GS_KCO * 128;  // two [32 co x 64 pos] blocks
constexpr int GS_B_BYTES = 4 * GS_KCO * 128;  // up to four [32 co x 64 kk] blocks
constexpr int GS_T_BYTES = 128 * 128;     // one [128 pos x 64 ch] block in the storage dtype
// registers after setmaxnreg (the CTA owns GS_THREADS x 80).  3D: the dot warps get all there is (their two rows x 8
// corner vectors are 64 registers on their own; at 80 they spill).  2D: nobody is short at 80, and the split costs
// (control warps spill at 24, movers below 80): every warp keeps its launch-time registers.
template <int ND> struct GsRegs {
  static constexpr bool REBALANCE = ND == 3;   // 2D: every warp keeps its launch-time 80 (the split only costs there)
  static constexpr uint32_t CTL = 24, MOVER = 72, DOT = 96;   // measured at cfg4: 598 us (dot 88: 671 us)
  static_assert(128 * (CTL + MOVER) + GS_DOT_THREADS * DOT <= GS_THREADS * 80, "setmaxnreg.inc would wait forever");
};

// fp32 callers: split bf16 operands (hi / lo grad_output tiles and filter slices, three products per K step, see
// tensor_path.cu FwdCfg); the accumulator goes to shared memory as fp32 (a [128 x 64] block is 32 KB) and the dot
// products against the fp32 corner values run in fp32.
template <typename T> struct GsCfg {
  static constexpr bool F32 = false;
  static constexpr int A_STAGE = GS_A_BYTES, T_BYTES = GS_T_BYTES, MAX_NB = 4;
  // slots of the T ring (accumulator blocks in shared memory, 16 KB each): deep enough that a block can wait for its
  // grad_columns store (save_gcol: the TMA engine reads it ~1 us after the dot warps are done with it) without stalling
  // the movers.  Measured: 2D 3 / 4 / 5 / 6 / 7 / 8 slots = 287 / 283 / 277 / 274 / 277 / 277 us at cfg2; 3D wants the
  // shared memory as L1 instead (4 / 5 / 6 / 8 slots = 603 / 618 / 617 / 632 us at cfg4)
  static constexpr int t_slots(int nd) { return nd == 2 ? 6 : 4; }
};
template <> struct GsCfg<float> {
  static constexpr bool F32 = true;
  static constexpr int A_STAGE = 2 * GS_A_BYTES, T_BYTES = 2 * GS_T_BYTES, MAX_NB = 2;   // B stage: hi blocks, then lo
  static constexpr int t_slots(int) { return 2; }
};
// byte offset of 16-byte chunk `c` (0..15) of row `row` in an fp32 T block: 256-byte rows, each 128-byte half
// swizzled like sw128
__device__ __forceinline__ uint32_t t32_off(int row, int c) {
  return (uint32_t)(row * 256 + ((c & 8) << 4) + (((c & 7) ^ (row & 7)) << 4));
}

// Offset / mask slab ring (north_star: "TMA bulk loads of ... offset/mask slabs into shared memory"): per work item
// (128-position tile = two 64-position blocks, one tap) the ND offset rows and the mask row of each block arrive by
// cp.async.bulk.tensor straight from the caller's NCHW tensors, OM_SLOTS items ahead of the movers, so the sampling
// entries are built from shared memory instead of behind 3-4 dependent global loads per item.
constexpr int GS_OM_SLOTS = 4;
constexpr int GS_OM_ROW_BYTES = 128;                       // 64 positions x 2 bytes
template <int ND> constexpr int gs_om_slot_bytes() { return 2 * (ND + 1) * GS_OM_ROW_BYTES; }   // 2 blocks x (ND offsets + mask)

template <int ND>
struct __align__(16) SampleEntry {
  int idx[1 << ND];  // uint4 index of the corner's 8-channel vector (channel offset excluded), -1: value is 0
  float fr[ND];
  float m;
  uint32_t flags;    // bit 0: grad_offset is produced, bit 1: row is a real output position
};

struct TcGsParams {
  DevGeom g;
  const void* xt;    // channels-last input
  const void* off;
  const void* mask;
  void* goff;
  void* gmask;
  int nb;            // 64-channel blocks per accumulator stage (1, 2 or 4)
  int nsub;          // stages per (deformable group, tap) = Cdg / 64 / nb
  int kblocks;       // K blocks per stage = Cog / 32
  int pbi;           // 64-position blocks per sample
  int pb_base;       // first 64-position block of this launch (a whole number of samples: sample0 * pbi)
  int total_pb;      // one past its last block: (sample0 + samples) * pbi
  int m_tiles;       // ceil((total_pb - pb_base) / 2)
  long long items;   // m_tiles * dg * K (tile, phase) work items, split evenly over the CTAs
  int om_tma;        // 1: offset / mask slabs come by TMA (needs P % 8 == 0: 16-byte global strides), 0: global loads
  int zero;          // always 0 at run time (tc::drain_point)
  int save_gcol;     // 1: every gcol block also goes to HBM ([samples][K][P][C], tmap_gcol) for the col2im gather of grad_input
};

// dot product of two 8-element vectors stored in the tensor dtype
__device__ __forceinline__ float dot8(const uint4& a, const uint4& b, __half) {
  const __half2* x = reinterpret_cast<const __half2*>(&a);
  const __half2* y = reinterpret_cast<const __half2*>(&b);
  __half2 p = __hmul2(x[0], y[0]);
  p = __hfma2(x[1], y[1], p);
  p = __hfma2(x[2], y[2], p);
  p = __hfma2(x[3], y[3], p);
  const float2 f = __half22float2(p);
  return f.x + f.y;
}
__device__ __forceinline__ float dot8(const uint4& a, const uint4& b, __nv_bfloat16) {
  const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* y = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162 p = __hmul2(x[0], y[0]);
  p = __hfma2(x[1], y[1], p);
  p = __hfma2(x[2], y[2], p);
  p = __hfma2(x[3], y[3], p);
  const float2 f = __bfloat1622float2(p);
  return f.x + f.y;
}
__device__ __forceinline__ float cvt_from_storage(float v) { return v; }
__device__ __forceinline__ float cvt_from_storage(__half v) { return __half2float(v); }
__device__ __forceinline__ float cvt_from_storage(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float dot8(const uint4&, const uint4&, float) { return 0.f; }        // (16-bit paths only)
__device__ __forceinline__ uint32_t pack2(float, float, float) { return 0u; }                   // (16-bit paths only)
__device__ __forceinline__ uint32_t pack2(float a, float b, __half) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack2(float a, float b, __nv_bfloat16) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <typename T, int ND, bool MOD>
__global__ void __launch_bounds__(GS_THREADS, 1)
k_tc_grad_sample(const __grid_constant__ CUtensorMap tmap_go, const __grid_constant__ CUtensorMap tmap_w,
                 const __grid_constant__ CUtensorMap tmap_off, const __grid_constant__ CUtensorMap tmap_mask,
                 const __grid_constant__ CUtensorMap tmap_go2, const __grid_constant__ CUtensorMap tmap_w2,
                 const __grid_constant__ CUtensorMap tmap_gcol, const TcGsParams prm) {
  constexpr int NC = 1 << ND;
  constexpr bool F32 = GsCfg<T>::F32;
  constexpr int A_STAGE = GsCfg<T>::A_STAGE;
  constexpr int T_BYTES = GsCfg<T>::T_BYTES;
  constexpr int T_SLOTS = GsCfg<T>::t_slots(ND);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = sA + GS_STAGES * A_STAGE;
  uint8_t* sT = sB + GS_STAGES * GS_B_BYTES;
  SampleEntry<ND>* sPlan = reinterpret_cast<SampleEntry<ND>*>(sT + T_SLOTS * T_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sPlan) + 2 * 128 * sizeof(SampleEntry<ND>));
  uint64_t* full = bars;                    // [GS_STAGES]  TMA -> MMA
  uint64_t* empty = full + GS_STAGES;       // [GS_STAGES]  MMA -> TMA
  uint64_t* tmem_full = empty + GS_STAGES;  // [2]          MMA -> movers
  uint64_t* tmem_empty = tmem_full + 2;     // [2]          movers -> MMA
  uint64_t* t_full = tmem_empty + 2;        // [T_SLOTS]    movers -> dot warps (and the gcol storer)
  uint64_t* t_empty = t_full + T_SLOTS;     // [T_SLOTS]    dot warps (and the gcol storer) -> movers
  uint64_t* plan_full = t_empty + T_SLOTS;  // [2]          movers -> dot warps
  uint64_t* plan_empty = plan_full + 2;     // [2]          dot warps -> movers
  uint64_t* om_full = plan_empty + 2;       // [GS_OM_SLOTS] TMA -> movers (offset / mask slabs)
  uint64_t* om_empty = om_full + GS_OM_SLOTS;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(om_empty + GS_OM_SLOTS);
  uint8_t* sOM = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tmem_ptr + 4) + 127) & ~(uintptr_t)127);

  const DevGeom& g = prm.g;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_go);
    ptx::prefetch_tensormap(&tmap_w);
    if (F32) { ptx::prefetch_tensormap(&tmap_go2); ptx::prefetch_tensormap(&tmap_w2); }
    for (int s = 0; s < GS_STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], ptx::arrivals(128));
      ptx::mbar_init(&plan_full[s], ptx::arrivals(128));
      ptx::mbar_init(&plan_empty[s], ptx::arrivals(GS_DOT_THREADS));
    }
    for (int s = 0; s < T_SLOTS; ++s) {
      ptx::mbar_init(&t_full[s], ptx::arrivals(128));
      ptx::mbar_init(&t_empty[s], ptx::arrivals(GS_DOT_THREADS) + (prm.save_gcol ? 1u : 0u));   // (+ the gcol storer)
    }
    for (int s = 0; s < GS_OM_SLOTS; ++s) {
      ptx::mbar_init(&om_full[s], 1);
      ptx::mbar_init(&om_empty[s], ptx::arrivals(128));
    }
    if (prm.om_tma) {
      ptx::prefetch_tensormap(&tmap_off);
      if (MOD) ptx::prefetch_tensormap(&tmap_mask);
    }
    if (prm.save_gcol) ptx::prefetch_tensormap(&tmap_gcol);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<512>(tmem_ptr);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  const int nb = prm.nb, nsub = prm.nsub;
  const int bpd = g.Cdg >> 6;            // 64-channel blocks per deformable group
  const int bpg = g.Cg >> 6;             // 64-channel blocks per conv group
  const int phases = g.dg * g.K;         // (deformable group, tap) phases per tile
  // this CTA's contiguous range of (tile, phase) items; all roles walk it in the same order
  const long long item0 = prm.items * blockIdx.x / gridDim.x;
  const long long item1 = prm.items * (blockIdx.x + 1) / gridDim.x;
  const int t_first = (int)(item0 / phases), phs_first = (int)(item0 - (long long)t_first * phases);
  const int n_items = (int)(item1 - item0);

  // Register re-balancing (the CTA owns GS_THREADS x 80): the control warpgroup and the movers give registers to the
  // dot warps, whose corner-load buffers (2 rows x 4 / 8 corners x 16 bytes in flight) set the pace.
  if (warp < 4) {
  if constexpr (GsRegs<ND>::REBALANCE) ptx::reg_dealloc<GsRegs<ND>::CTL>();
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int t = t_first, phs = phs_first;
      int pb_b[2] = {0, 0}, pb_p0[2] = {0, 0};
      bool new_tile = true;
      for (int it = 0; it < n_items; ++it) {
        if (new_tile) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int pb = prm.pb_base + 2 * t + h;
            pb_b[h] = pb < prm.total_pb ? pb / prm.pbi : g.B;  // b == B: fully out of bounds -> zero fill
            pb_p0[h] = pb < prm.total_pb ? (pb % prm.pbi) * 64 : 0;
          }
          new_tile = false;
        }
        {
          const int dgi = phs / g.K, tap = phs - dgi * g.K;
          for (int sub = 0; sub < nsub; ++sub) {
            const int blk0 = dgi * bpd + sub * nb;  // first 64-channel block (global numbering)
            const int grp = blk0 / bpg;
            const int ccl0 = blk0 - grp * bpg;      // block index inside the conv group
            for (int kb = 0; kb < prm.kblocks; ++kb) {
              const int co0 = grp * g.Cog + kb * GS_KCO;
              ptx::mbar_wait(&empty[s], ph ^ 1);
              ptx::mbar_arrive_expect_tx(&full[s], (uint32_t)(GS_A_BYTES + nb * GS_KCO * 128) * (F32 ? 2u : 1u));
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                ptx::tma_load_3d(sA + s * A_STAGE + h * (GS_KCO * 128), &tmap_go, &full[s], pb_p0[h], co0, pb_b[h]);
                if (F32)
                  ptx::tma_load_3d(sA + s * A_STAGE + GS_A_BYTES + h * (GS_KCO * 128), &tmap_go2, &full[s], pb_p0[h], co0,
                                   pb_b[h]);
              }
              for (int j = 0; j < nb; ++j) {
                ptx::tma_load_2d(sB + s * GS_B_BYTES + j * (GS_KCO * 128), &tmap_w, &full[s],
                                 ((ccl0 + j) * g.K + tap) * 64, co0);
                if (F32)
                  ptx::tma_load_2d(sB + s * GS_B_BYTES + GS_B_BYTES / 2 + j * (GS_KCO * 128), &tmap_w2, &full[s],
                                   ((ccl0 + j) * g.K + tap) * 64, co0);
              }
              if (++s == GS_STAGES) { s = 0; ph ^= 1; }
            }
          }
        }
        if (++phs == phases) { phs = 0; ++t; new_tile = true; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_f16(F32 ? 1 : umma_fmt<T>(), /*A MN-major*/ 1, /*B MN-major*/ 1, 128, nb * 64);
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      {
        for (int st = 0; st < n_items * nsub; ++st) {
          ptx::mbar_wait(&tmem_empty[as], aph ^ 1);
          ptx::tc_fence_after_sync();
          const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
          for (int kb = 0; kb < prm.kblocks; ++kb) {
            ptx::mbar_wait(&full[s], ph);
            ptx::tc_fence_after_sync();
            const uint32_t a_addr = ptx::smem_u32(sA + s * A_STAGE);
            const uint32_t b_addr = ptx::smem_u32(sB + s * GS_B_BYTES);
#pragma unroll
            for (int k = 0; k < GS_KCO / 16; ++k) {  // 16 output channels per MMA
              const uint64_t ad = ptx::umma_smem_desc_sw128(a_addr + k * 2048, GS_KCO * 128, 1024);
              const uint64_t bd = ptx::umma_smem_desc_sw128(b_addr + k * 2048, GS_KCO * 128, 1024);
              ptx::umma_f16(d_tmem, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              if (F32) {   // + hi*lo + lo*hi
                const uint64_t al = ptx::umma_smem_desc_sw128(a_addr + GS_A_BYTES + k * 2048, GS_KCO * 128, 1024);
                const uint64_t bl = ptx::umma_smem_desc_sw128(b_addr + GS_B_BYTES / 2 + k * 2048, GS_KCO * 128, 1024);
                ptx::umma_f16(d_tmem, ad, bl, idesc, 1u);
                ptx::umma_f16(d_tmem, al, bd, idesc, 1u);
              }
            }
            ptx::umma_commit(&empty[s]);
            if (++s == GS_STAGES) { s = 0; ph ^= 1; }
          }
          ptx::umma_commit(&tmem_full[as]);
          if (++as == 2) { as = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== gcol storer =====================
    // grad_columns for the col2im gather of grad_input (reference src/mdeformable_conv.cu:396-419 materialises the
    // same buffer through a GEMM): every [128 positions x 64 channels] block the movers lay out for the dot warps also
    // leaves for HBM, straight from that shared-memory tile by TMA (two 64-position boxes; rows past the end of a
    // sample are clipped by the tensor map).  The storer is one more reader of the T ring.
    if (lane == 0 && prm.save_gcol && !F32) {
      int ts = 0, prev = -1;                    // prev: slot whose store may still be reading shared memory
      const uint64_t pol = ptx::l2_policy_evict_first();   // a write-once stream: it should not displace the operands
      uint32_t tph = 0;
      int t = t_first, phs = phs_first;
      for (int it = 0; it < n_items; ++it) {
        const int dgi = phs / g.K, tap = phs - dgi * g.K;
        int sb[2], sp0[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int pb = prm.pb_base + 2 * t + h;
          sb[h] = pb < prm.total_pb ? pb / prm.pbi - prm.pb_base / prm.pbi : -1;   // sample inside this launch's buffer
          sp0[h] = pb < prm.total_pb ? (pb % prm.pbi) * 64 : 0;
        }
        for (int sub = 0; sub < nsub; ++sub) {
          for (int j = 0; j < nb; ++j) {
            const int c0 = (dgi * bpd + sub * nb + j) * 64;   // grad_columns plane of this tap: [P][C]
            ptx::mbar_wait(&t_full[ts], tph);
            const uint8_t* src = sT + ts * T_BYTES;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#ifdef MDCONV_EXP_NOSTORE   // timing experiment only (wrong grad_input): the ring protocol without the stores
              if (false) ptx::tma_store_4d_hint(&tmap_gcol, src + h * 8192, c0, sp0[h], tap, sb[h], pol);
#else
              if (sb[h] >= 0) ptx::tma_store_4d_hint(&tmap_gcol, src + h * 8192, c0, sp0[h], tap, sb[h], pol);
#endif
            ptx::bulk_commit_group();
            if (prev >= 0) {                    // the store before this one has read its tile: the movers may overwrite it
              ptx::bulk_wait_group_read<1>();
              ptx::mbar_arrive(&t_empty[prev]);
            }
            prev = ts;
            if (++ts == T_SLOTS) { ts = 0; tph ^= 1; }
          }
        }
        if (++phs == phases) { phs = 0; ++t; }
      }
      if (prev >= 0) {
        ptx::bulk_wait_group_read<0>();
        ptx::mbar_arrive(&t_empty[prev]);
      }
      ptx::bulk_wait_group<0>();                // all stores complete before the CTA retires
    }
  } else if (warp == 3) {
    // ===================== TMA producer: offset / mask slabs =====================
    if (lane == 0 && prm.om_tma) {
      int os = 0;
      uint32_t oph = 0;
      int t = t_first, phs = phs_first;
      for (int it = 0; it < n_items; ++it) {
        const int dgi = phs / g.K, tap = phs - dgi * g.K;
        ptx::mbar_wait_relaxed(&om_empty[os], oph ^ 1, 64);
        ptx::mbar_arrive_expect_tx(&om_full[os], (uint32_t)(2 * (ND + (MOD ? 1 : 0)) * GS_OM_ROW_BYTES));
        uint8_t* slot = sOM + os * gs_om_slot_bytes<ND>();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int pb = prm.pb_base + 2 * t + h;
          const int b = pb < prm.total_pb ? pb / prm.pbi : g.B;        // b == B: out of bounds -> zero fill
          const int p0 = pb < prm.total_pb ? (pb % prm.pbi) * 64 : 0;
          uint8_t* dst = slot + h * (ND + 1) * GS_OM_ROW_BYTES;
          ptx::tma_load_3d(dst, &tmap_off, &om_full[os], p0, (dgi * g.K + tap) * ND, b);
          if (MOD) ptx::tma_load_3d(dst + ND * GS_OM_ROW_BYTES, &tmap_mask, &om_full[os], p0, dgi * g.K + tap, b);
        }
        if (++os == GS_OM_SLOTS) { os = 0; oph ^= 1; }
        if (++phs == phases) { phs = 0; ++t; }
      }
    }
  }
  } else if (warp < 8) {
    // ===================== movers: sampling entries + TMEM -> smem re-layout =====================
    if constexpr (GsRegs<ND>::REBALANCE) ptx::reg_dealloc<GsRegs<ND>::MOVER>();
    const int row = threadIdx.x - 128;  // TMEM lane == tile row
    const T* __restrict__ off = reinterpret_cast<const T*>(prm.off);
    const T* __restrict__ mask = reinterpret_cast<const T*>(prm.mask);
    const int C8 = F32 ? g.C >> 2 : g.C >> 3;   // 16-byte vectors per position
    const bool rule_b = g.rule != 0;
    int as = 0, ts = 0, ps = 0, os = 0;
    uint32_t aph = 0, tph = 0, pph = 0, oph = 0;
    int t = t_first, phs = phs_first;
    bool new_tile = true;
    bool valid = false;
    RowCoord<ND> rc;
    for (int it = 0; it < n_items; ++it) {
      if (new_tile) {
        const int pb = prm.pb_base + 2 * t + (row >> 6);
        const int p = (pb % prm.pbi) * 64 + (row & 63);
        valid = pb < prm.total_pb && p < g.P;
        decode_row<ND>(g, valid ? pb / prm.pbi : 0, p, valid, rc);
        new_tile = false;
      }
      {
        const int dgi = phs / g.K, tap = phs - dgi * g.K;
        {
          // ---- sampling entry of (row, dgi, tap) ----
          SampleEntry<ND> e;
#pragma unroll
          for (int c = 0; c < NC; ++c) e.idx[c] = -1;
#pragma unroll
          for (int a = 0; a < ND; ++a) e.fr[a] = 0.f;
          e.m = 0.f;
          e.flags = 0u;
          RawSample<ND> raw;
          if (prm.om_tma) {
            // this row's offsets / mask of the item, from the slab the TMA warp fetched ahead
            ptx::mbar_wait(&om_full[os], oph);
            const T* srow = reinterpret_cast<const T*>(sOM + os * gs_om_slot_bytes<ND>() +
                                                       (row >> 6) * (ND + 1) * GS_OM_ROW_BYTES) + (row & 63);
#pragma unroll
            for (int a = 0; a < ND; ++a) raw.o[a] = cvt_from_storage(srow[a * 64]);
            raw.m = MOD ? cvt_from_storage(srow[ND * 64]) : 1.f;
            ptx::mbar_arrive_warp(&om_empty[os]);
            if (++os == GS_OM_SLOTS) { os = 0; oph ^= 1; }
          } else if (valid) {
            raw = fetch_sample<T, ND, MOD>(g, off, mask, rc, dgi, tap);
          }
          if (valid) {
            SamplePos<ND> sp;
            build_pos<ND>(g, rc, tap, raw, sp);
            bool lo_ok[ND], hi_ld[ND];
#pragma unroll
            for (int a = 0; a < ND; ++a) {
              const int lo = sp.lo[a], hi = lo + 1;
              lo_ok[a] = lo >= 0 && lo <= g.in[a] - 1;
              const bool hi_in = hi >= 0 && hi <= g.in[a] - 1;
              // rule B (reference src/deformable_conv.cu:254-261): a "high" corner on an axis whose
              // fraction is <= EPS is absent; rule A loads it whenever it is inside the input.
              hi_ld[a] = rule_b ? (hi_in && fabsf(sp.fr[a]) > kEps) : hi_in;
              e.fr[a] = sp.fr[a];
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              bool ok = true;
              int pix = 0;
#pragma unroll
              for (int a = 0; a < ND; ++a) {
                const int hi = (c >> a) & 1;
                ok = ok && (hi ? hi_ld[a] : lo_ok[a]);
                pix = pix * g.in[a] + sp.lo[a] + hi;
              }
              if (ok) e.idx[c] = (rc.b * g.IN + pix) * C8;
            }
            e.m = sp.m;
            e.flags = 2u | ((rule_b || sp.in_range) ? 1u : 0u);
          }
          ptx::mbar_wait(&plan_empty[ps], pph ^ 1);
          const uint32_t pa = ptx::smem_u32(sPlan + ps * 128 + row);
          const uint4* q = reinterpret_cast<const uint4*>(&e);
#pragma unroll
          for (int i = 0; i < (int)(sizeof(SampleEntry<ND>) / 16); ++i) sts128(pa + 16 * i, q[i]);
          ptx::mbar_arrive_warp(&plan_full[ps]);
          if (++ps == 2) { ps = 0; pph ^= 1; }
        }
        for (int sub = 0; sub < nsub; ++sub) {
          ptx::mbar_wait(&tmem_full[as], aph);
          ptx::tc_fence_after_sync();
          const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(as * 256);
          for (int j = 0; j < nb; ++j) {
            ptx::mbar_wait(&t_empty[ts], tph ^ 1);
            const uint32_t tb = ptx::smem_u32(sT + ts * T_BYTES);
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
              uint32_t v[32];
              ptx::tmem_ld_32x32(taddr + j * 64 + hlf * 32, v);
              ptx::tmem_ld_wait();
              if constexpr (F32) {
#pragma unroll
                for (int q = 0; q < 8; ++q)   // 4 channels = one 16-byte chunk, kept in fp32
                  sts128(tb + t32_off(row, hlf * 8 + q), make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
                continue;
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) {  // 8 channels = one 16-byte chunk
                uint4 o;
                o.x = pack2(__uint_as_float(v[8 * q]), __uint_as_float(v[8 * q + 1]), T());
                o.y = pack2(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]), T());
                o.z = pack2(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]), T());
                o.w = pack2(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]), T());
                sts128(tb + sw128(row, hlf * 4 + q), o);
              }
            }
            if (prm.save_gcol) ptx::fence_proxy_async_smem();   // the TMA store reads the tile through the async proxy
            ptx::mbar_arrive_warp(&t_full[ts]);
            if (++ts == T_SLOTS) { ts = 0; tph ^= 1; }
          }
          ptx::tc_fence_before_sync();
          ptx::mbar_arrive_warp(&tmem_empty[as]);
          if (++as == 2) { as = 0; aph ^= 1; }
        }
      }
      if (++phs == phases) { phs = 0; ++t; new_tile = true; }
    }
  } else {
    // ===================== dot warps =====================
    if constexpr (GsRegs<ND>::REBALANCE) ptx::reg_alloc<GsRegs<ND>::DOT>();
    const int dw = warp - GS_DOT_WARP0;
    const int rq = lane >> 3;  // row inside a pass
    const int j = lane & 7;    // 8-channel chunk
    const int row0 = dw * 8 + rq;  // pass 0 row (pass 1: +4)
    const uint4* __restrict__ xq = reinterpret_cast<const uint4*>(prm.xt);
    T* __restrict__ goff = reinterpret_cast<T*>(prm.goff);
    T* __restrict__ gmask = reinterpret_cast<T*>(prm.gmask);
    int ts = 0, ps = 0;
    uint32_t tph = 0, pph = 0;
    int t = t_first, phs = phs_first;
    for (int it = 0; it < n_items; ++it) {
      {
        const int dgi = phs / g.K, tap = phs - dgi * g.K;
        ptx::mbar_wait(&plan_full[ps], pph);
        const uint32_t plan_addr = ptx::smem_u32(sPlan + ps * 128);
        // corner indices of this thread's two rows (fixed for the whole phase)
        int idx[2][NC];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const uint32_t ea = plan_addr + (row0 + 4 * i) * (uint32_t)sizeof(SampleEntry<ND>);
#pragma unroll
          for (int c4 = 0; c4 < NC / 4; ++c4) {
            const uint4 iv = lds128(ea + 16 * c4);
            idx[i][4 * c4] = (int)iv.x; idx[i][4 * c4 + 1] = (int)iv.y;
            idx[i][4 * c4 + 2] = (int)iv.z; idx[i][4 * c4 + 3] = (int)iv.w;
          }
        }
        float S[2][NC];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int c = 0; c < NC; ++c) S[i][c] = 0.f;
        // Software pipeline over the (64-channel block, row pass) steps of the item: while one row's dot products
        // run, the corner loads of the NEXT step are in flight (two register buffers; the corner values do not
        // depend on the accumulator, only their use waits for the T tile).
        const int nblk = nsub * nb;   // 64-channel blocks of this (deformable group, tap)
        if constexpr (F32) {
          // fp32: this lane's 8 channels are two 16-byte vectors of the fp32 input and of the fp32 accumulator block;
          // per block and row: corner loads (4 corners at a time), wait for the block, fp32 dot products
          for (int blk = 0; blk < nblk; ++blk) {
            const uint4* xc = xq + (((dgi * bpd + blk) * 64) >> 2) + 2 * j;
            bool waited = false;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
#pragma unroll
              for (int c4 = 0; c4 < NC; c4 += 4) {
                uint4 xa[4], xb[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const int id = idx[i][c4 + c];
                  xa[c] = id >= 0 ? __ldg(xc + id) : make_uint4(0u, 0u, 0u, 0u);
                  xb[c] = id >= 0 ? __ldg(xc + id + 1) : make_uint4(0u, 0u, 0u, 0u);
                }
                if (!waited) { ptx::mbar_wait(&t_full[ts], tph); waited = true; }
                const uint32_t tb = ptx::smem_u32(sT + ts * T_BYTES);
                const uint4 g0 = lds128(tb + t32_off(row0 + 4 * i, 2 * j));
                const uint4 g1 = lds128(tb + t32_off(row0 + 4 * i, 2 * j + 1));
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  float d = __uint_as_float(g0.x) * __uint_as_float(xa[c].x);
                  d = fmaf(__uint_as_float(g0.y), __uint_as_float(xa[c].y), d);
                  d = fmaf(__uint_as_float(g0.z), __uint_as_float(xa[c].z), d);
                  d = fmaf(__uint_as_float(g0.w), __uint_as_float(xa[c].w), d);
                  d = fmaf(__uint_as_float(g1.x), __uint_as_float(xb[c].x), d);
                  d = fmaf(__uint_as_float(g1.y), __uint_as_float(xb[c].y), d);
                  d = fmaf(__uint_as_float(g1.z), __uint_as_float(xb[c].z), d);
                  d = fmaf(__uint_as_float(g1.w), __uint_as_float(xb[c].w), d);
                  S[i][c4 + c] += d;
                }
              }
            }
            ptx::mbar_arrive_warp(&t_empty[ts]);
            if (++ts == T_SLOTS) { ts = 0; tph ^= 1; }
          }
        }
        auto load_row = [&](int blk, int i, uint4 (&v)[NC], uint32_t after = 0u) {
          const uint4* xc = xq + (((dgi * bpd + blk) * 64) >> 3) + j + after;
#pragma unroll
          for (int c = 0; c < NC; ++c) v[c] = idx[i][c] >= 0 ? __ldg(xc + idx[i][c]) : make_uint4(0u, 0u, 0u, 0u);
        };
        if constexpr (F32) {
          // (handled above)
        } else if constexpr (ND == 2) {
          // (the wait for the row about to be used comes BEFORE the next row's loads are issued: every LDG of the loop
          // retires through one scoreboard, see the forward kernel's gather loop, tensor_path.cu)
          uint4 va[NC], vb[NC];
          load_row(0, 0, va);
          for (int blk = 0; blk < nblk; ++blk) {
            load_row(blk, 1, vb, tc::drain_point(va[NC - 1].w, prm.zero));
            ptx::mbar_wait(&t_full[ts], tph);
            const uint32_t tb = ptx::smem_u32(sT + ts * T_BYTES);
            {
              const uint4 gc = lds128(tb + sw128(row0, j));
#pragma unroll
              for (int c = 0; c < NC; ++c) S[0][c] += dot8(gc, va[c], T());
            }
            const uint32_t db = tc::drain_point(vb[NC - 1].w, prm.zero);
            if (blk + 1 < nblk) load_row(blk + 1, 0, va, db);
            {
              const uint4 gc = lds128(tb + sw128(row0 + 4, j) + db);   // (+ db: the wait stays ahead of the branch)
#pragma unroll
              for (int c = 0; c < NC; ++c) S[1][c] += dot8(gc, vb[c], T());
            }
            ptx::mbar_arrive_warp(&t_empty[ts]);
            if (++ts == T_SLOTS) { ts = 0; tph ^= 1; }
          }
        } else {
          // 3D: 8 corners per row -- the rolling form needs more live registers than the dot warps have (it
          // spills and runs a third slower); both rows' loads are issued together ahead of the wait instead
          for (int blk = 0; blk < nblk; ++blk) {
            uint4 xv[2][NC];
            load_row(blk, 0, xv[0]);
            load_row(blk, 1, xv[1]);
            ptx::mbar_wait(&t_full[ts], tph);
            const uint32_t tb = ptx::smem_u32(sT + ts * T_BYTES);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const uint4 gc = lds128(tb + sw128(row0 + 4 * i, j));
#pragma unroll
              for (int c = 0; c < NC; ++c) S[i][c] += dot8(gc, xv[i][c], T());
            }
            ptx::mbar_arrive_warp(&t_empty[ts]);
            if (++ts == T_SLOTS) { ts = 0; tph ^= 1; }
          }
        }
        // ---- reduce over the 8 lanes of each row, combine, store ----
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            float v = S[i][c];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            S[i][c] = v;
          }
        if (j < 2) {
          // lane j of a row group finishes pass j (all 8 lanes hold the same totals)
          float Sc[NC];
#pragma unroll
          for (int c = 0; c < NC; ++c) Sc[c] = j == 1 ? S[1][c] : S[0][c];
          const int row = row0 + 4 * j;
          const uint32_t ea = plan_addr + row * (uint32_t)sizeof(SampleEntry<ND>) + 16 * (NC / 4);
          const uint4 tail = lds128(ea);  // fr[ND], m, flags (2D)  /  fr[3], m (3D)
          float fr[3];
          float m;
          uint32_t flags;
          fr[0] = __uint_as_float(tail.x);
          fr[1] = __uint_as_float(tail.y);
          if (ND == 2) {
            fr[2] = 0.f;
            m = __uint_as_float(tail.z);
            flags = tail.w;
          } else {
            fr[2] = __uint_as_float(tail.z);
            m = __uint_as_float(tail.w);
            const uint4 tail2 = lds128(ea + 16);
            flags = tail2.x;
          }
          if (flags & 2u) {
            const int pb = prm.pb_base + 2 * t + (row >> 6);
            const int b = pb / prm.pbi;
            const int p = (pb % prm.pbi) * 64 + (row & 63);
            float gm = 0.f;
            float go_a[ND];
#pragma unroll
            for (int a = 0; a < ND; ++a) go_a[a] = 0.f;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              float w = 1.f;
#pragma unroll
              for (int a = 0; a < ND; ++a) w *= ((c >> a) & 1) ? fr[a] : (1.f - fr[a]);
              gm = fmaf(w, Sc[c], gm);
#pragma unroll
              for (int a = 0; a < ND; ++a) {
                float d = ((c >> a) & 1) ? 1.f : -1.f;
#pragma unroll
                for (int bb = 0; bb < ND; ++bb)
                  if (bb != a) d *= ((c >> bb) & 1) ? fr[bb] : (1.f - fr[bb]);
                go_a[a] = fmaf(d, Sc[c], go_a[a]);
              }
            }
            const size_t ch = (size_t)b * g.dg * g.K + (size_t)dgi * g.K + tap;
#pragma unroll
            for (int a = 0; a < ND; ++a)
              goff[(ch * ND + a) * g.P + p] = cvt_to<T, float>((flags & 1u) ? m * go_a[a] : 0.f);
            if (MOD) gmask[ch * g.P + p] = cvt_to<T, float>(gm);
          }
        }
        ptx::mbar_arrive_warp(&plan_empty[ps]);
        if (++ps == 2) { ps = 0; pph ^= 1; }
      }
      if (++phs == phases) { phs = 0; ++t; }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<512>(tmem_base);
}

template <typename T, int ND>
size_t gs_smem_bytes() {
  return 1024 + (size_t)GS_STAGES * (GsCfg<T>::A_STAGE + GS_B_BYTES) + GsCfg<T>::t_slots(ND) * GsCfg<T>::T_BYTES +
         2 * 128 * sizeof(SampleEntry<ND>) +
         (2 * GS_STAGES + 8 + 2 * GsCfg<T>::t_slots(ND) + 2 * GS_OM_SLOTS) * sizeof(uint64_t) + 160 + (size_t)GS_OM_SLOTS * gs_om_slot_bytes<ND>();
}

struct GsMaps {
  CUtensorMap go, w, off, mask, go2, w2;   // go2 / w2: the "lo" halves of fp32 callers
  CUtensorMap gcol;                        // grad_columns [samples][K][P][C] (save_gcol)
};
template <typename T, int ND, bool MOD>
void launch_grad_sample(const GsMaps& m, const TcGsParams& prm, LaunchCtx& ctx) {
  const size_t smem = gs_smem_bytes<T, ND>();
  static std::atomic<unsigned long long> configured{0};
  opt_in_dynamic_smem(k_tc_grad_sample<T, ND, MOD>, smem, configured, ctx);
  const int sms = sm_count() - ctx.spare_sms > 0 ? sm_count() - ctx.spare_sms : 1;
  const int grid = prm.items < sms ? (int)prm.items : sms;
  ProfScope prof(ctx, "tc_grad_sample");
  k_tc_grad_sample<T, ND, MOD><<<grid, GS_THREADS, smem, ctx.stream>>>(m.go, m.w, m.off, m.mask, m.go2, m.w2, m.gcol,
                                                                       prm);
  MDCONV_CHECK_LAUNCH(ctx);
}

}  // namespace

bool grad_sample_supported(const DevGeom& g) {
  return g.Cog % GS_KCO == 0 && g.Cg % 64 == 0 && g.Cdg % 64 == 0;
}

// xt: channels-last input; wr: weight re-ordered to [Co][(channel-block, tap, 64 channels)] (fp32 callers: the bf16
// hi bank followed by the lo bank); go: grad_output rows [B][Co][go_ld] (go_ld = P, or P rounded up to 8 elements for
// the padded copy; fp32 callers: the bf16 hi copy, go_lo the lo copy).  One launch covers samples [sample0, sample0 +
// samples); gcol (16-bit callers, may be NULL) receives their grad_columns as [samples][K][P][C].
template <typename T>
void tensor_grad_sample(const DevGeom& g, const T* xt, const void* wr, const void* off, const void* mask, const void* go,
                        const void* go_lo, int go_ld, void* goff, void* gmask, void* gcol, int sample0, int samples,
                        LaunchCtx& ctx) {
  constexpr bool F32 = std::is_same<T, float>::value;
  using E = typename std::conditional<F32, __nv_bfloat16, T>::type;   // element type of the TMA-read operands
  TcGsParams prm;
  prm.g = g;
  prm.zero = 0;
  prm.xt = xt; prm.off = off; prm.mask = mask; prm.goff = goff; prm.gmask = gmask;
  const int bpd = g.Cdg / 64, bpg = g.Cg / 64;
  int nb = GsCfg<T>::MAX_NB;
  while (nb > 1 && (bpd % nb != 0 || bpg % nb != 0)) nb >>= 1;
  prm.nb = nb;
  prm.nsub = bpd / nb;
  prm.kblocks = g.Cog / GS_KCO;
  prm.pbi = (g.P + 63) / 64;
  prm.pb_base = sample0 * prm.pbi;
  prm.total_pb = (sample0 + samples) * prm.pbi;
  prm.m_tiles = (prm.total_pb - prm.pb_base + 1) / 2;
  prm.items = (long long)prm.m_tiles * g.dg * g.K;

  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    if (ctx.err == cudaSuccess) ctx.err = cudaErrorNotSupported;
    return;
  }
  GsMaps m;
  memset(&m, 0, sizeof(m));
  auto enc_go = [&](CUtensorMap* tm, const void* base) {
    const cuuint64_t dims[3] = {(cuuint64_t)g.P, (cuuint64_t)g.Co, (cuuint64_t)g.B};
    const cuuint64_t strides[2] = {(cuuint64_t)go_ld * sizeof(E), (cuuint64_t)g.Co * go_ld * sizeof(E)};
    const cuuint32_t box[3] = {64, GS_KCO, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(tm, tmap_dtype<E>(), 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  auto enc_w = [&](CUtensorMap* tm, const void* base) {
    const cuuint64_t dims[2] = {(cuuint64_t)g.Kg, (cuuint64_t)g.Co};
    const cuuint64_t strides[1] = {(cuuint64_t)g.Kg * sizeof(E)};
    const cuuint32_t box[2] = {64, GS_KCO};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, tmap_dtype<E>(), 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  bool ok = enc_go(&m.go, go) && enc_w(&m.w, wr);
  if (ok && F32) ok = enc_go(&m.go2, go_lo) && enc_w(&m.w2, reinterpret_cast<const char*>(wr) + wr_bytes(g) / 2);
  if (!ok) {
    if (ctx.err == cudaSuccess) ctx.err = cudaErrorInvalidValue;
    return;
  }
  // offset [B][dg*K*nd][P] and mask [B][dg*K][P] as 3D maps, un-swizzled boxes of 64 positions x (nd | 1) channels
  // (16-bit callers with 16-byte row strides; everything else builds the sampling entries from global loads)
  prm.om_tma = (!F32 && g.P % 8 == 0 && ((uintptr_t)off & 15) == 0 && (!g.modulated || ((uintptr_t)mask & 15) == 0)) ? 1 : 0;
  if (prm.om_tma) {
    const cuuint32_t estr[3] = {1, 1, 1};
    const cuuint64_t odims[3] = {(cuuint64_t)g.P, (cuuint64_t)g.dg * g.K * g.nd, (cuuint64_t)g.B};
    const cuuint64_t ostr[2] = {(cuuint64_t)g.P * sizeof(T), (cuuint64_t)g.dg * g.K * g.nd * g.P * sizeof(T)};
    const cuuint32_t obox[3] = {64, (cuuint32_t)g.nd, 1};
    CUresult r = enc(&m.off, tmap_dtype<E>(), 3, const_cast<void*>(off), odims, ostr, obox, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS && g.modulated) {
      const cuuint64_t mdims[3] = {(cuuint64_t)g.P, (cuuint64_t)g.dg * g.K, (cuuint64_t)g.B};
      const cuuint64_t mstr[2] = {(cuuint64_t)g.P * sizeof(T), (cuuint64_t)g.dg * g.K * g.P * sizeof(T)};
      const cuuint32_t mbox[3] = {64, 1, 1};
      r = enc(&m.mask, tmap_dtype<E>(), 3, const_cast<void*>(mask), mdims, mstr, mbox, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) prm.om_tma = 0;   // fall back to global loads
  }
  // grad_columns [samples][K][P][C] in the storage dtype (one [P][C] plane per tap: neighbouring positions of a tap
  // are neighbouring 2*C-byte rows, for the stores here and for the gather that follows): boxes of 64 positions x 64
  // channels, written from the swizzled T tile
  prm.save_gcol = (gcol != nullptr && !F32) ? 1 : 0;
  if (prm.save_gcol) {
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const cuuint64_t cdims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.P, (cuuint64_t)g.K, (cuuint64_t)samples};
    const cuuint64_t cstr[3] = {(cuuint64_t)g.C * sizeof(T), (cuuint64_t)g.P * g.C * sizeof(T),
                                (cuuint64_t)g.K * g.P * g.C * sizeof(T)};
    const cuuint32_t cbox[4] = {64, 64, 1, 1};
    if (enc(&m.gcol, tmap_dtype<E>(), 4, gcol, cdims, cstr, cbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      if (ctx.err == cudaSuccess) ctx.err = cudaErrorInvalidValue;
      return;
    }
  }
  if (g.nd == 2) {
    if (g.modulated) launch_grad_sample<T, 2, true>(m, prm, ctx);
    else launch_grad_sample<T, 2, false>(m, prm, ctx);
  } else {
    if (g.modulated) launch_grad_sample<T, 3, true>(m, prm, ctx);
    else launch_grad_sample<T, 3, false>(m, prm, ctx);
  }
}

template void tensor_grad_sample<__half>(const DevGeom&, const __half*, const void*, const void*, const void*,
                                         const void*, const void*, int, void*, void*, void*, int, int, LaunchCtx&);
template void tensor_grad_sample<__nv_bfloat16>(const DevGeom&, const __nv_bfloat16*, const void*, const void*,
                                                const void*, const void*, const void*, int, void*, void*, void*, int, int,
                                                LaunchCtx&);
template void tensor_grad_sample<float>(const DevGeom&, const float*, const void*, const void*, const void*,
                                        const void*, const void*, int, void*, void*, void*, int, int, LaunchCtx&);

}  // namespace tc
}  // namespace mdconv
