// tensor_bwd_input.cu -- grad_input, atomics-free, two routes over one inverse index.
//
// Reference: the col2im half of the fused gradient kernels (src/mdeformable_conv.cu:282-293 and
// twins): one thread per column element, 4 (2D) / 8 (3D) atomicAdd into grad_input each, fed by a
// materialised grad_columns buffer (G2, :417-419).
//
// Here the scatter is turned into a gather.  A small inverse index (CSR) is built per call:
//     bin (sample, deformable group, input position, tap)  ->  list of (output position, weight)
// with weight = interpolation weight x mask of the corner that touches that input position.
//
// Route 1 (16-bit callers, default): k_col2im_gather.  k_tc_grad_sample has streamed grad_columns
//     gcol[pos, (tap, c)] = sum_co go[pos, co] * W[co, c, tap]
// into the workspace; grad_input[ipos, c] = sum over the entries of ipos's bins of weight * gcol[entry, c] is a
// plain gather-reduce (no contraction).  Bins are position-major here, so one input position owns one run of entries.
//
// Route 2 (fp32 callers, or a grad_columns budget of 0): k_tc_grad_input, a third contraction:
//     G[pos, (tap, co)] = sum over the bin's entries of weight * go[entry position, co]
//     grad_input[pos, c] = sum over (tap, co) of G[pos, (tap, co)] * W[co, c, tap]
// is one implicit GEMM: G tiles are gathered straight into the swizzled UMMA A operand (never in
// HBM), W^T slices come by TMA, the accumulator lives in TMEM.  Bins are tap-major for this route.
// Either way every grad_input element is written exactly once (deterministic: bin entries are sorted).
//
// Warp roles of k_tc_grad_input:
//   warp 0      TMA producer: transposed-weight slices [BN channels x 64 co] -> smem ring
//   warp 1      MMA issuer (M = 128 input positions, N = BN channels, K = 64 co per block)
//   warp 2      TMEM allocator
//   warps 4-7   epilogue: tcgen05.ld -> convert -> NCHW stores
//   warps 8-15  gather producers: 8 lanes per input position walk the bin, 16-byte channels-last
//               loads of grad_output, weighted accumulate, 16-byte stores into the swizzled A tile
#include <string.h>

#include <cub/device/device_scan.cuh>

#include "tc_common.cuh"

namespace mdconv {
namespace tc {

namespace {

constexpr int GI_BM = 128;
#ifndef MDCONV_GI_STAGES
#define MDCONV_GI_STAGES 3
#endif
constexpr int GI_STAGES = MDCONV_GI_STAGES;
constexpr int GI_PLAN_SLOTS = 2;
constexpr int GI_MAX_BN = 256;
constexpr int GI_PLAN_WARP0 = 8;      // warps 8-11: plan producers (bin heads -> smem)
constexpr int GI_GATHER_WARP0 = 12;   // warps 12-27: gather producers
constexpr int GI_GATHER_WARPS = 16;
constexpr int GI_GATHER_THREADS = GI_GATHER_WARPS * 32;
constexpr int GI_THREADS = (GI_GATHER_WARP0 + GI_GATHER_WARPS) * 32;  // 896
constexpr int GI_A_BYTES = GI_BM * 128;
constexpr int GI_B_BYTES = GI_MAX_BN * 128;
constexpr int GI_NE = 8;  // bin entries staged per (position, tap); longer bins continue from global memory

// One (input position, tap) of the tile: the first GI_NE entries of its bin, ready for the gather.
struct __align__(16) GiPlanRow {
  uint32_t idx[GI_NE]; // byte offset of the entry's row in the channels-last grad_output
  uint16_t w[GI_NE];   // interpolation weight x mask in the storage dtype (fp16 / bf16 bits)
  int n;               // entries in the bin
  int e0;              // index of the bin's first entry (for bins longer than GI_NE)
};
static_assert(sizeof(GiPlanRow) == 64, "GiPlanRow must be 64 bytes");
// fp32 callers: the same row with fp32 weights (the gather accumulates in fp32 and splits the sum into bf16 hi / lo)
struct __align__(16) GiPlanRowF32 {
  uint32_t idx[GI_NE];
  float w[GI_NE];
  int n;
  int e0;
  int pad[2];
};
static_assert(sizeof(GiPlanRowF32) == 80, "GiPlanRowF32 must be 80 bytes");
// fp32 callers run the contraction as split bf16 (see tensor_path.cu, FwdCfg): two A tiles and two B tiles per stage,
// three products per K step, at most 128 channels per tile.
template <typename T> struct GiCfg {
  static constexpr bool F32 = false;
  static constexpr int A_STAGE = GI_A_BYTES, MAX_BN_T = GI_MAX_BN;
  using Row = GiPlanRow;
};
template <> struct GiCfg<float> {
  static constexpr bool F32 = true;
  static constexpr int A_STAGE = 2 * GI_A_BYTES, MAX_BN_T = 128;
  using Row = GiPlanRowF32;
};
__device__ __forceinline__ void gi_split8(const float (&v)[8], uint4& hi, uint4& lo) {
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
// registers after setmaxnreg (the CTA owns GI_THREADS x 72)
constexpr uint32_t GI_REGS_CTL = 32, GI_REGS_EPI = 40, GI_REGS_PLAN = 48, GI_REGS_GATHER = 96;
static_assert(128 * (GI_REGS_CTL + GI_REGS_EPI + GI_REGS_PLAN) + GI_GATHER_THREADS * GI_REGS_GATHER <= GI_THREADS * 72,
              "setmaxnreg.inc would wait forever");

struct BinEntry {
  int row;   // b * P + output position: row of the channels-last grad_output
  float w;   // interpolation weight x mask
};

// ---- inverse index ----------------------------------------------------------------------------
// One thread per (sample, deformable group, output position, tap).  FILL == false: count the entries
// of each bin; FILL == true: place them (slot taken from the remaining count).
// pos_major: the layout the col2im gather reads -- the K bins of one input position are neighbours (bin =
// ((sample, deformable group), input position, tap)), so its entries of ALL taps are one contiguous run, and an entry
// is ready to use: `row` = element offset of the grad_columns block it scales, ((b * K + tap) * P + output position) * C,
// `w` = the weight as a packed pair in the storage dtype.  Otherwise tap-major (bin = ((sample, deformable group),
// tap, input position), row = b * P + output position, fp32 weight), what k_tc_grad_input walks.
__device__ __forceinline__ float pack_weight(float w, __half) {
  const __half2 h = __float2half2_rn(w);
  return __uint_as_float(*reinterpret_cast<const uint32_t*>(&h));
}
__device__ __forceinline__ float pack_weight(float w, __nv_bfloat16) {
  const __nv_bfloat162 h = __float2bfloat162_rn(w);
  return __uint_as_float(*reinterpret_cast<const uint32_t*>(&h));
}
__device__ __forceinline__ float pack_weight(float w, float) { return w; }
template <typename T, int ND, bool MOD, bool FILL>
__global__ void __launch_bounds__(256) k_bin_entries(const DevGeom g, const T* __restrict__ off,
                                                     const T* __restrict__ mask, int* __restrict__ count,
                                                     const int* __restrict__ start, BinEntry* __restrict__ entries,
                                                     int pos_major) {
  const long long n = (long long)g.B * g.dg * g.K * g.P;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    // thread order follows the bin layout, so the bins (and entry sectors) a wave of threads touches are a compact,
    // L2-resident region that is complete when it is written back: positions fastest for tap-major bins, taps
    // fastest for position-major bins (3D fill 482 us with positions fastest there)
    int p, tap;
    long long r;
    if (pos_major) {
      tap = (int)(i % g.K); r = i / g.K;
      p = (int)(r % g.P); r /= g.P;
    } else {
      p = (int)(i % g.P); r = i / g.P;
      tap = (int)(r % g.K); r /= g.K;
    }
    const int dgi = (int)(r % g.dg);
    const int b = (int)(r / g.dg);
    RowCoord<ND> rc;
    decode_row<ND>(g, b, p, true, rc);
    SamplePos<ND> sp;
    sample_pos<T, ND, MOD>(g, off, mask, rc, dgi, tap, sp);
    if (!sp.in_range) continue;
    // all corners first, then all atomics, then all stores: the 2^ND (atomic -> store) chains are independent and
    // this keeps them in flight together (the fill pass is latency-bound: 8 returning atomics per thread in 3D)
    size_t bin[1 << ND];
    float wgt[1 << ND];
    bool okc[1 << ND];
#pragma unroll
    for (int c = 0; c < (1 << ND); ++c) {
      float w = sp.m;
      bool ok = true;
      int pix = 0;
#pragma unroll
      for (int a = 0; a < ND; ++a) {
        const int hi = (c >> a) & 1;
        const int ia = sp.lo[a] + hi;
        // zero-weight "high" corners are skipped, as the reference skips their adds
        // (src/mdeformable_conv.cu:285-293, src/deformable_conv.cu:254-261)
        ok = ok && (hi ? (ia <= g.in[a] - 1 && sp.fr[a] > kEps) : (ia >= 0));
        w *= hi ? sp.fr[a] : (1.f - sp.fr[a]);
        pix = pix * g.in[a] + ia;
      }
      okc[c] = ok;
      wgt[c] = w;
      const size_t bd = (size_t)b * g.dg + dgi;
      bin[c] = !ok ? 0 : pos_major ? (bd * g.IN + pix) * g.K + tap
                                   : (bd * g.K + tap) * g.IN + pix;   // tap-major: neighbours in p hit neighbouring bins
    }
    if (!FILL) {
#pragma unroll
      for (int c = 0; c < (1 << ND); ++c)
        if (okc[c]) atomicAdd(count + bin[c], 1);
    } else {
      int slot[1 << ND];
#pragma unroll
      for (int c = 0; c < (1 << ND); ++c) slot[c] = okc[c] ? atomicSub(count + bin[c], 1) - 1 : 0;
#pragma unroll
      for (int c = 0; c < (1 << ND); ++c) slot[c] += okc[c] ? __ldg(start + bin[c]) : 0;
#pragma unroll
      for (int c = 0; c < (1 << ND); ++c) {
        if (okc[c]) {
          BinEntry e;
          e.row = pos_major ? ((b * g.K + tap) * g.P + p) * g.C : b * g.P + p;
          e.w = pos_major ? pack_weight(wgt[c], T()) : wgt[c];
          entries[slot[c]] = e;
        }
      }
    }
  }
}

// Sort each bin by output position so the summation order (and the result) is run-to-run identical.
// Bins of up to 8 entries (almost all) are sorted in registers with a fixed compare-exchange network
// on 64-bit keys (row in the high word; rows are unique inside a bin); longer ones by insertion.
__device__ __forceinline__ void cswap(unsigned long long& a, unsigned long long& b) {
  const unsigned long long lo = a < b ? a : b, hi = a < b ? b : a;
  a = lo; b = hi;
}
// Bins longer than GI_SORT_LOCAL entries (strongly clustered offsets) are not sorted here: their ids go to a work
// list (`big`: [0] = count, [1..] = bin ids; it reuses the count array, which is all zero again after the fill
// pass) and k_bin_sort_big sorts each with a whole block.
constexpr int GI_SORT_LOCAL = 64;
__global__ void __launch_bounds__(256) k_bin_sort(const int* __restrict__ start, BinEntry* __restrict__ entries,
                                                  size_t nbins, int* __restrict__ big) {
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(entries);  // {row (low word), w} -> swap on use
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbins; i += (size_t)gridDim.x * blockDim.x) {
    const int s = start[i], e = start[i + 1];
    const int n = e - s;
    if (n < 2) continue;
    if (n <= 8) {
      unsigned long long k[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        if (a < n) {
          const unsigned long long v = keys[s + a];
          k[a] = (v << 32) | (v >> 32);  // row into the high word
        } else {
          k[a] = ~0ull;
        }
      }
      // 19-comparator sorting network for 8 keys
      cswap(k[0], k[1]); cswap(k[2], k[3]); cswap(k[4], k[5]); cswap(k[6], k[7]);
      cswap(k[0], k[2]); cswap(k[1], k[3]); cswap(k[4], k[6]); cswap(k[5], k[7]);
      cswap(k[1], k[2]); cswap(k[5], k[6]); cswap(k[0], k[4]); cswap(k[3], k[7]);
      cswap(k[1], k[5]); cswap(k[2], k[6]);
      cswap(k[1], k[4]); cswap(k[3], k[6]);
      cswap(k[2], k[4]); cswap(k[3], k[5]);
      cswap(k[3], k[4]);
#pragma unroll
      for (int a = 0; a < 8; ++a)
        if (a < n) keys[s + a] = (k[a] << 32) | (k[a] >> 32);
    } else if (n <= 16) {
      unsigned long long k[16];
#pragma unroll
      for (int a = 0; a < 16; ++a) {
        if (a < n) {
          const unsigned long long v = keys[s + a];
          k[a] = (v << 32) | (v >> 32);
        } else {
          k[a] = ~0ull;
        }
      }
      // Batcher odd-even merge sort, 16 keys, 63 comparators (typical 3D bins hold ~8 entries)
      cswap(k[0], k[1]); cswap(k[2], k[3]); cswap(k[0], k[2]); cswap(k[1], k[3]); cswap(k[1], k[2]);
      cswap(k[4], k[5]); cswap(k[6], k[7]); cswap(k[4], k[6]); cswap(k[5], k[7]); cswap(k[5], k[6]);
      cswap(k[0], k[4]); cswap(k[2], k[6]); cswap(k[2], k[4]); cswap(k[1], k[5]); cswap(k[3], k[7]);
      cswap(k[3], k[5]); cswap(k[1], k[2]); cswap(k[3], k[4]); cswap(k[5], k[6]); cswap(k[8], k[9]);
      cswap(k[10], k[11]); cswap(k[8], k[10]); cswap(k[9], k[11]); cswap(k[9], k[10]); cswap(k[12], k[13]);
      cswap(k[14], k[15]); cswap(k[12], k[14]); cswap(k[13], k[15]); cswap(k[13], k[14]); cswap(k[8], k[12]);
      cswap(k[10], k[14]); cswap(k[10], k[12]); cswap(k[9], k[13]); cswap(k[11], k[15]); cswap(k[11], k[13]);
      cswap(k[9], k[10]); cswap(k[11], k[12]); cswap(k[13], k[14]); cswap(k[0], k[8]); cswap(k[4], k[12]);
      cswap(k[4], k[8]); cswap(k[2], k[10]); cswap(k[6], k[14]); cswap(k[6], k[10]); cswap(k[2], k[4]);
      cswap(k[6], k[8]); cswap(k[10], k[12]); cswap(k[1], k[9]); cswap(k[5], k[13]); cswap(k[5], k[9]);
      cswap(k[3], k[11]); cswap(k[7], k[15]); cswap(k[7], k[11]); cswap(k[3], k[5]); cswap(k[7], k[9]);
      cswap(k[11], k[13]); cswap(k[1], k[2]); cswap(k[3], k[4]); cswap(k[5], k[6]); cswap(k[7], k[8]);
      cswap(k[9], k[10]); cswap(k[11], k[12]); cswap(k[13], k[14]);
#pragma unroll
      for (int a = 0; a < 16; ++a)
        if (a < n) keys[s + a] = (k[a] << 32) | (k[a] >> 32);
    } else if (n <= GI_SORT_LOCAL) {
      for (int a = s + 1; a < e; ++a) {
        const BinEntry v = entries[a];
        int q = a - 1;
        while (q >= s && entries[q].row > v.row) {
          entries[q + 1] = entries[q];
          --q;
        }
        entries[q + 1] = v;
      }
    } else {
      big[1 + atomicAdd(big, 1)] = (int)i;   // at most nbins - 1 ids fit behind the counter; a bin id is < nbins
    }
  }
}

// One block per long bin: bitonic sorting network with every comparator ascending (partner i ^ (k-1) in the first
// step of a merge, i ^ j afterwards), so the virtual padding up to a power of two is +infinity at the END and
// comparators that touch it are no-ops.  O(n log^2 n / 256) steps per bin, straight on global memory.
__global__ void __launch_bounds__(256) k_bin_sort_big(const int* __restrict__ start, BinEntry* __restrict__ entries,
                                                      int* __restrict__ big) {
  const int nbig = big[0];
  for (int wi = blockIdx.x; wi < nbig; wi += gridDim.x) {
    const int bin = big[1 + wi];
    const int s = start[bin], n = start[bin + 1] - s;
    BinEntry* e = entries + s;
    unsigned np2 = 1;
    while (np2 < (unsigned)n) np2 <<= 1;
    for (unsigned k = 2; k <= np2; k <<= 1) {
      for (unsigned j = k >> 1; j > 0; j >>= 1) {
        for (unsigned i = threadIdx.x; i < np2; i += blockDim.x) {
          const unsigned l = (j == (k >> 1)) ? (i ^ (k - 1)) : (i ^ j);
          if (l > i && l < (unsigned)n) {
            const BinEntry a = e[i], b = e[l];
            if (a.row > b.row) { e[i] = b; e[l] = a; }
          }
        }
        __syncthreads();
      }
    }
  }
}

struct TcGiParams {
  DevGeom g;
  const void* got;           // channels-last grad_output [B][P][Co]
  const int* bin_start;      // [nbins + 1]
  const BinEntry* entries;
  void* gx;                  // grad_input, NCHW
  int BN;                    // channels per tile
  int m_tiles, n_tiles, total_tiles;
  int cobs;                  // 64-wide output-channel blocks per group = Cog / 64
  int kblocks;               // K * cobs
  int zero;                  // always 0 at run time (tc::drain_point)
};

// 8-channel accumulator of the gather: packed pairs in the storage dtype (one HFMA2 / HFMA2.BF16 per channel pair,
// as in the forward gather).
template <typename T> struct Acc8;
template <> struct Acc8<__half> {
  __half2 a[4];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __float2half2_rn(0.f);
  }
  __device__ __forceinline__ void fma(const uint4& v, float w) { fma(v, __float2half_rn(w)); }
  __device__ __forceinline__ void fma(const uint4& v, __half w) {
    const __half2 wh = __half2half2(w);
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __hfma2(wh, h[i], a[i]);
  }
  __device__ __forceinline__ uint4 pack() const {
    uint4 r;
    uint32_t* o = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = *reinterpret_cast<const uint32_t*>(&a[i]);
    return r;
  }
};
template <> struct Acc8<__nv_bfloat16> {
  __nv_bfloat162 a[4];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __float2bfloat162_rn(0.f);
  }
  __device__ __forceinline__ void fma(const uint4& v, float w) { fma(v, __float2bfloat16_rn(w)); }
  __device__ __forceinline__ void fma(const uint4& v, __nv_bfloat16 w) {
    const __nv_bfloat162 wh = __bfloat162bfloat162(w);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __hfma2(wh, h[i], a[i]);
  }
  __device__ __forceinline__ uint4 pack() const {
    uint4 r;
    uint32_t* o = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = *reinterpret_cast<const uint32_t*>(&a[i]);
    return r;
  }
};

template <> struct Acc8<float> {   // placeholder: the fp32 gather has its own accumulation (never instantiated paths)
  __device__ __forceinline__ void clear() {}
  __device__ __forceinline__ void fma(const uint4&, float) {}
  __device__ __forceinline__ uint4 pack() const { return make_uint4(0u, 0u, 0u, 0u); }
};

// ---- grad_input as a col2im GATHER over grad_columns (16-bit callers, default) --------------------
// k_tc_grad_sample already holds gcol[pos, (tap, c)] = sum_co go[pos, co] * W[co, c, tap] block by block in shared
// memory for its dot products; with save_gcol it also streams those blocks to HBM ([samples][K][P][C], the reference's
// grad_columns buffer, src/mdeformable_conv.cu:396-419).  grad_input is then
//     gx[ipos, c] = sum over taps, over the entries e of bin (sample, deformable group of c, tap, ipos):  w_e * gcol[row_e, (tap, c)]
// -- the reference's col2im (:282-293) read through the inverse index instead of scattered with atomics: no third
// contraction, no channels-last copy of grad_output, no transposed filter bank, and still one write per element in a
// fixed order.  One thread = 8 channels of one input position; its entries (all taps: one contiguous run of the
// position-major index) are walked C2I_B at a time, so that many 16-byte loads are in flight per thread; fp32
// accumulation.  NCHW callers: a [64 positions x C] tile is turned
// through shared memory so the stores run along the positions; channels-last callers store their rows directly.
constexpr int C2I_MAX_POS = 64;    // input positions per tile: ~512 (position, 8-channel chunk) pairs, 8 .. 64 positions
constexpr int C2I_CS = 512;        // channels per tile (wider layers take several channel slabs per position tile)
constexpr int C2I_THREADS = 256;
#ifndef MDCONV_C2I_B
#define MDCONV_C2I_B 8
#endif
#ifndef MDCONV_C2I_MINB
#define MDCONV_C2I_MINB 4
#endif
constexpr int C2I_B = MDCONV_C2I_B;        // entries per batch of the gather loop (16-byte loads in flight per thread)
constexpr int C2I_MINB = MDCONV_C2I_MINB;  // blocks per SM the register allocation aims at

// packed 8-channel partial sum of one batch: sum_k w_k * v_k with w_k already a packed pair (entry.y)
template <int NB, typename H2>
__device__ __forceinline__ void c2i_batch_t(const uint4 (&v)[NB], const int2 (&ee)[NB], H2 (&a)[4]) {
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    const uint32_t wb = (uint32_t)ee[k].y;
    const H2 w = *reinterpret_cast<const H2*>(&wb);
    const H2* h = reinterpret_cast<const H2*>(&v[k]);
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = k == 0 ? __hmul2(w, h[i]) : __hfma2(w, h[i], a[i]);
  }
}
template <int NB>
__device__ __forceinline__ uint4 c2i_batch(const uint4 (&v)[NB], const int2 (&ee)[NB], __half) {
  __half2 a[4];
  c2i_batch_t<NB, __half2>(v, ee, a);
  return *reinterpret_cast<const uint4*>(a);
}
template <int NB>
__device__ __forceinline__ uint4 c2i_batch(const uint4 (&v)[NB], const int2 (&ee)[NB], __nv_bfloat16) {
  __nv_bfloat162 a[4];
  c2i_batch_t<NB, __nv_bfloat162>(v, ee, a);
  return *reinterpret_cast<const uint4*>(a);
}
// fp32 += a packed partial sum
__device__ __forceinline__ void c2i_flush_t(float (&acc)[8], const uint4& p, __half) {
  const __half2* h = reinterpret_cast<const __half2*>(&p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    acc[2 * i] += f.x;
    acc[2 * i + 1] += f.y;
  }
}
__device__ __forceinline__ void c2i_flush_t(float (&acc)[8], const uint4& p, __nv_bfloat16) {
  const uint32_t* u = reinterpret_cast<const uint32_t*>(&p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {   // bf16 -> fp32 is a shift
    acc[2 * i] += __uint_as_float(u[i] << 16);
    acc[2 * i + 1] += __uint_as_float(u[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 c2i_pack(const float (&acc)[8], __half) {
  uint4 r;
  uint32_t* o = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 h = __floats2half2_rn(acc[2 * i], acc[2 * i + 1]);
    o[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return r;
}
__device__ __forceinline__ uint4 c2i_pack(const float (&acc)[8], __nv_bfloat16) {
  uint4 r;
  uint32_t* o = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
    o[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return r;
}

template <typename T, bool CL>
__global__ void __launch_bounds__(C2I_THREADS, C2I_MINB) k_col2im_gather(const DevGeom g, const T* __restrict__ gcol,
                                                                  const int* __restrict__ bin_start,
                                                                  const BinEntry* __restrict__ entries,
                                                                  T* __restrict__ gx, long long pos_base,
                                                                  long long total_pos, int n_tiles, int slabs,
                                                                  int tile_pos) {
  extern __shared__ __align__(16) uint8_t c2i_smem[];   // NCHW: [tile_pos][slab channels + 8] elements
  const int2* __restrict__ ent2 = reinterpret_cast<const int2*>(entries);
  T* tile = reinterpret_cast<T*>(c2i_smem);
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const long long pos0 = pos_base + (long long)(t / slabs) * tile_pos;   // [pos_base, total_pos): this launch's samples
    const int c_base = (t % slabs) * C2I_CS;      // first channel of this tile's slab
    const int cs = min(C2I_CS, g.C - c_base);     // channels in it
    const int ldt = cs + 8;                       // padded row of the staging tile (elements)
    const int CH = cs >> 3;                       // 8-channel chunks per position
    const int items = tile_pos * CH;              // (position, chunk) pairs per tile
    // bounds [e0, e1) of a (position, chunk) pair's run of entries: position-major bins, so the entries of all K taps
    // of the input position are contiguous (~4 * K of them)
    auto run_bounds = [&](int it, int& e0, int& e1) {
      const int chunk = it % CH, pl = it / CH;
      const long long gpos = pos0 + pl;
      e0 = 0; e1 = 0;
      if (it < items && gpos < total_pos) {
        const int b = (int)(gpos / g.IN), ipos = (int)(gpos - (long long)b * g.IN);
        const int dgi = (c_base + chunk * 8) / g.Cdg;
        const size_t bin0 = (((size_t)b * g.dg + dgi) * g.IN + ipos) * g.K;
        e0 = __ldg(bin_start + bin0);
        e1 = __ldg(bin_start + bin0 + g.K);
      }
    };
    int e0, e1;
    run_bounds(threadIdx.x, e0, e1);
    for (int it = threadIdx.x; it < items; it += C2I_THREADS) {
      const int chunk = it % CH, pl = it / CH;    // a warp = consecutive chunks of one position (>= 256 channels), or of a few
      const long long gpos = pos0 + pl;
      // the bounds of this thread's NEXT pair are fetched now, and the head of its run is pulled into L1 once they
      // are here, so the next pair starts on its grad_columns loads instead of behind three dependent misses
      int n0, n1;
      run_bounds(it + C2I_THREADS, n0, n1);
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      {
        const T* colc = gcol + c_base + chunk * 8;   // an entry's block: colc + entry.x
        // C2I_B entries at a time: their (L1-resident) records, then their grad_columns vectors all in flight
        // together; the products of a batch are summed in packed pairs of the storage dtype, as the gather of
        // k_tc_grad_input does, and batches are folded in fp32.
        int e = e0;
        for (; e + C2I_B <= e1; e += C2I_B) {
          int2 ee[C2I_B];
          uint4 v[C2I_B];
#pragma unroll
          for (int k = 0; k < C2I_B; ++k) ee[k] = __ldg(ent2 + e + k);
#pragma unroll
          for (int k = 0; k < C2I_B; ++k) v[k] = __ldg(reinterpret_cast<const uint4*>(colc + ee[k].x));
          c2i_flush_t(acc, c2i_batch<C2I_B>(v, ee, T()), T());
        }
        if (n1 > n0) {
          asm volatile("prefetch.global.L1 [%0];" ::"l"(ent2 + n0));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(ent2 + n0 + 16));
        }
        if (e < e1) {                             // the last, partial batch
          int2 ee[C2I_B];
          uint4 v[C2I_B];
#pragma unroll
          for (int k = 0; k < C2I_B; ++k) ee[k] = e + k < e1 ? __ldg(ent2 + e + k) : make_int2(0, 0);   // weight 0
#pragma unroll
          for (int k = 0; k < C2I_B; ++k) v[k] = e + k < e1 ? __ldg(reinterpret_cast<const uint4*>(colc + ee[k].x)) : make_uint4(0u, 0u, 0u, 0u);
          c2i_flush_t(acc, c2i_batch<C2I_B>(v, ee, T()), T());
        }
      }
      e0 = n0; e1 = n1;
      const uint4 out = c2i_pack(acc, T());
      if (CL) {
        if (gpos < total_pos) *reinterpret_cast<uint4*>(gx + (size_t)gpos * g.C + c_base + chunk * 8) = out;
      } else {
        *reinterpret_cast<uint4*>(tile + (size_t)pl * ldt + chunk * 8) = out;   // 16-byte rows-of-channels store
      }
    }
    if (!CL) {
      __syncthreads();
      // store: thread -> (8-position chunk, channel), lanes along the channels (conflict-free 2-byte reads of the
      // staging tile); a tile may straddle samples, and IN need not be a multiple of 8
      const int vec_ok = (g.IN % 8 == 0) && ((reinterpret_cast<uintptr_t>(gx) & 15) == 0);
      for (int q = threadIdx.x; q < cs * (tile_pos / 8); q += C2I_THREADS) {
        const int cl = q % cs, pc = (q / cs) * 8;
        const int c = c_base + cl;
        const long long gp = pos0 + pc;
        if (gp >= total_pos) continue;
        __align__(16) T vals[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) vals[k] = tile[(size_t)(pc + k) * ldt + cl];
        const int b = (int)(gp / g.IN), ipos = (int)(gp - (long long)b * g.IN);
        if (vec_ok && gp + 8 <= total_pos) {      // (IN % 8 == 0: the 8 positions are in one sample, 16-byte aligned)
          *reinterpret_cast<uint4*>(gx + ((size_t)b * g.C + c) * g.IN + ipos) = *reinterpret_cast<const uint4*>(vals);
        } else {
          for (int k = 0; k < 8 && gp + k < total_pos; ++k) {
            const long long gk = gp + k;
            const int bk = (int)(gk / g.IN), ik = (int)(gk - (long long)bk * g.IN);
            gx[((size_t)bk * g.C + c) * g.IN + ik] = vals[k];
          }
        }
      }
      __syncthreads();
    }
  }
}

// CL: the caller's activations are channels-last (grad_output is gathered in place, grad_input leaves through TMA
// stores); a compile-time switch so the NCHW instantiation carries none of that epilogue's registers or code.
// STAGES: depth of the A/B operand ring.  3 for 2D; 2 for 3D, whose bins are twice as long (8 corners per sample):
// there the 48 KB a third stage takes are worth more as L1 for the inverse-index gather (measured at cfg4).
template <typename T, bool CL, int STAGES>
__global__ void __launch_bounds__(GI_THREADS, 1)
k_tc_grad_input(const __grid_constant__ CUtensorMap tmap_wt, const __grid_constant__ CUtensorMap tmap_gx,
                const __grid_constant__ CUtensorMap tmap_wt2, const TcGiParams prm) {
  constexpr bool F32 = GiCfg<T>::F32;
  constexpr int A_STAGE = GiCfg<T>::A_STAGE;
  using PlanRow = typename GiCfg<T>::Row;
  static_assert(!(F32 && CL), "no channels-last variant of the fp32 split path");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * A_STAGE;
  PlanRow* sPlan = reinterpret_cast<PlanRow*>(sB + STAGES * GI_B_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPlan + GI_PLAN_SLOTS * GI_BM);
  uint64_t* full_a = bars;
  uint64_t* full_b = full_a + STAGES;
  uint64_t* empty = full_b + STAGES;
  uint64_t* plan_full = empty + STAGES;
  uint64_t* plan_empty = plan_full + GI_PLAN_SLOTS;
  uint64_t* tmem_full = plan_empty + GI_PLAN_SLOTS;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const DevGeom& g = prm.g;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_wt);
    if (F32) ptx::prefetch_tensormap(&tmap_wt2);
    (void)tmap_gx;
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_a[s], ptx::arrivals(GI_GATHER_THREADS));
      ptx::mbar_init(&full_b[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < GI_PLAN_SLOTS; ++s) {
      ptx::mbar_init(&plan_full[s], ptx::arrivals(GI_BM));
      ptx::mbar_init(&plan_empty[s], ptx::arrivals(GI_GATHER_THREADS));
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
  const long long total_pos = (long long)g.B * g.IN;

  if (warp < 4) {
    ptx::reg_dealloc<GI_REGS_CTL>();
  if (warp == 0) {
    // ===================== TMA producer: W^T slices =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < prm.total_tiles; t += gridDim.x) {
        const int nt = t / prm.m_tiles;
        const int c0 = nt * BN;
        for (int kb = 0; kb < prm.kblocks; ++kb) {
          ptx::mbar_wait_relaxed(&empty[s], ph ^ 1, 64);
          ptx::mbar_arrive_expect_tx(&full_b[s], (uint32_t)BN * 128u * (F32 ? 2u : 1u));
          ptx::tma_load_2d(sB + s * GI_B_BYTES, &tmap_wt, &full_b[s], kb * 64, c0);
          if (F32) ptx::tma_load_2d(sB + s * GI_B_BYTES + GI_B_BYTES / 2, &tmap_wt2, &full_b[s], kb * 64, c0);   // lo halves
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_f16(F32 ? 1 : umma_fmt<T>(), 0, 0, GI_BM, BN);
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int t = blockIdx.x; t < prm.total_tiles; t += gridDim.x) {
        ptx::mbar_wait(&tmem_empty[as], aph ^ 1);
        ptx::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * GI_MAX_BN);
        for (int kb = 0; kb < prm.kblocks; ++kb) {
          ptx::mbar_wait(&full_b[s], ph);
          ptx::mbar_wait(&full_a[s], ph);
          ptx::tc_fence_after_sync();
          const uint32_t a_addr = ptx::smem_u32(sA + s * A_STAGE);
          const uint32_t b_addr = ptx::smem_u32(sB + s * GI_B_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = ptx::umma_smem_desc_sw128(a_addr + k * 32, 0, 1024);
            const uint64_t bd = ptx::umma_smem_desc_sw128(b_addr + k * 32, 0, 1024);
            ptx::umma_f16(d_tmem, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            if (F32) {   // + hi*lo + lo*hi
              const uint64_t al = ptx::umma_smem_desc_sw128(a_addr + GI_A_BYTES + k * 32, 0, 1024);
              const uint64_t bl = ptx::umma_smem_desc_sw128(b_addr + GI_B_BYTES / 2 + k * 32, 0, 1024);
              ptx::umma_f16(d_tmem, ad, bl, idesc, 1u);
              ptx::umma_f16(d_tmem, al, bd, idesc, 1u);
            }
          }
          ptx::umma_commit(&empty[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        ptx::umma_commit(&tmem_full[as]);
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  }
  } else if (warp < 8) {
    // ===================== epilogue: TMEM -> grad_input (NCHW) =====================
    ptx::reg_dealloc<GI_REGS_EPI>();
    const int ew = warp & 3;
    T* __restrict__ gx = reinterpret_cast<T*>(prm.gx);
    int as = 0;
    uint32_t aph = 0;
    for (int t = blockIdx.x; t < prm.total_tiles; t += gridDim.x) {
      const int mt = t % prm.m_tiles;
      const int nt = t / prm.m_tiles;
      const long long gp = (long long)mt * GI_BM + ew * 32 + lane;
      const bool valid = gp < total_pos;
      const int b = valid ? (int)(gp / g.IN) : 0;
      const int pix = valid ? (int)(gp % g.IN) : 0;
      T* xrow = gx + ((size_t)b * g.C + (size_t)nt * BN) * g.IN + pix;
      ptx::mbar_wait_relaxed(&tmem_full[as], aph);
      ptx::tc_fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * GI_MAX_BN);
      if constexpr (CL) {
        // channels-last grad_input [B*IN][C]: the thread's row is contiguous in memory -- 16-byte stores straight
        // from the converted accumulator (no staging tile: the 32 KB it would take halve the L1 left to the
        // inverse-index gather, which costs more than the TMA store saves; the forward's epilogue, whose gather is
        // less L1-sensitive, does go through shared memory and cp.async.bulk.tensor)
        T* crow = gx + (size_t)(valid ? gp : 0) * g.C + (size_t)nt * BN;
        for (int cc = 0; cc < BN; cc += 16) {
          uint32_t v[16];
          ptx::tmem_ld_32x16(taddr + cc, v);
          ptx::tmem_ld_wait();
          if (valid) {
            uint4 o0, o1;
            o0.x = pack2_storage(__uint_as_float(v[0]), __uint_as_float(v[1]), T());
            o0.y = pack2_storage(__uint_as_float(v[2]), __uint_as_float(v[3]), T());
            o0.z = pack2_storage(__uint_as_float(v[4]), __uint_as_float(v[5]), T());
            o0.w = pack2_storage(__uint_as_float(v[6]), __uint_as_float(v[7]), T());
            o1.x = pack2_storage(__uint_as_float(v[8]), __uint_as_float(v[9]), T());
            o1.y = pack2_storage(__uint_as_float(v[10]), __uint_as_float(v[11]), T());
            o1.z = pack2_storage(__uint_as_float(v[12]), __uint_as_float(v[13]), T());
            o1.w = pack2_storage(__uint_as_float(v[14]), __uint_as_float(v[15]), T());
            *reinterpret_cast<uint4*>(crow + cc) = o0;
            *reinterpret_cast<uint4*>(crow + cc + 8) = o1;
          }
        }
      } else {
        for (int cc = 0; cc < BN; cc += 16) {
          uint32_t v[16];
          ptx::tmem_ld_32x16(taddr + cc, v);
          ptx::tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 16; ++j) xrow[(size_t)(cc + j) * g.IN] = cvt_to<T, float>(__uint_as_float(v[j]));
          }
        }
      }
      ptx::tc_fence_before_sync();
      ptx::mbar_arrive_warp(&tmem_empty[as]);
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  } else if (warp < GI_GATHER_WARP0) {
    // ===================== plan producers =====================
    ptx::reg_dealloc<GI_REGS_PLAN>();
    // thread = tile row (input position): per tap, copy the head of the bin into shared memory; the
    // bin bounds of the next tap are fetched while this tap's entries are in flight.
    const int row = threadIdx.x - GI_PLAN_WARP0 * 32;
    const int* __restrict__ bin_start = prm.bin_start;
    const int2* __restrict__ entries2 = reinterpret_cast<const int2*>(prm.entries);
    const int Co8 = g.Co >> 3;
    int ps = 0;
    uint32_t pph = 0;
    for (int t = blockIdx.x; t < prm.total_tiles; t += gridDim.x) {
      const int mt = t % prm.m_tiles;
      const int nt = t / prm.m_tiles;
      const int dgi = (nt * BN) / g.Cdg;
      const long long gp = (long long)mt * GI_BM + row;
      // bins are tap-major: bin(tap) = bin0 + tap * IN; the bounds of the next tap's bin are fetched one tap ahead
      long long bin0 = -1;
      if (gp < total_pos) {
        const long long b = gp / g.IN, pix = gp - b * g.IN;
        bin0 = (b * g.dg + dgi) * (long long)g.K * g.IN + pix;
      }
      int e_lo = 0, e_hi = 0, n_lo = 0, n_hi = 0;
      if (bin0 >= 0) { e_lo = __ldg(bin_start + bin0); e_hi = __ldg(bin_start + bin0 + 1); }
      if (bin0 >= 0 && g.K > 1) { n_lo = __ldg(bin_start + bin0 + g.IN); n_hi = __ldg(bin_start + bin0 + g.IN + 1); }
      for (int tap = 0; tap < g.K; ++tap) {
        const int e0 = e_lo, n = e_hi - e_lo;
        e_lo = n_lo; e_hi = n_hi;
        if (bin0 >= 0 && tap + 2 < g.K) {
          const long long nb = bin0 + (long long)(tap + 2) * g.IN;
          n_lo = __ldg(bin_start + nb); n_hi = __ldg(bin_start + nb + 1);
        }
        PlanRow r;
#pragma unroll
        for (int i = 0; i < GI_NE; ++i) { r.idx[i] = 0u; r.w[i] = 0; }
        r.n = n;
        r.e0 = e0;
#pragma unroll
        for (int i = 0; i < GI_NE; ++i) {
          if (i < n) {
            const int2 ev = __ldg(entries2 + e0 + i);
            r.idx[i] = (uint32_t)ev.x * (uint32_t)(g.Co * sizeof(T));
            if constexpr (F32) {
              r.w[i] = __int_as_float(ev.y);
            } else {
              const T wt = cvt_to<T, float>(__int_as_float(ev.y));
              r.w[i] = *reinterpret_cast<const uint16_t*>(&wt);
            }
          }
        }
        ptx::mbar_wait_relaxed(&plan_empty[ps], pph ^ 1, 128);
        const uint32_t pa = ptx::smem_u32(sPlan + ps * GI_BM + row);
        const uint4* q = reinterpret_cast<const uint4*>(&r);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(PlanRow) / 16); ++i) sts128(pa + 16 * i, q[i]);
        ptx::mbar_arrive_warp(&plan_full[ps]);
        if (++ps == GI_PLAN_SLOTS) { ps = 0; pph ^= 1; }
      }
    }
  } else {
    // ===================== gather producers =====================
    ptx::reg_alloc<GI_REGS_GATHER>();
    const int gw = warp - GI_GATHER_WARP0;
    const int chunk = lane & 7;
    const int row0 = gw * 8 + (lane >> 3);
    const uint4* __restrict__ gq = reinterpret_cast<const uint4*>(prm.got);
    const int2* __restrict__ entries2 = reinterpret_cast<const int2*>(prm.entries);
    const int Co8 = g.Co >> 3;
    int s = 0, ps = 0;
    uint32_t ph = 0, pph = 0;
    // Software pipeline over the (output-channel block, pass) steps of one tap: the plan rows of this thread's
    // two passes (bin heads: row offsets, weights, lengths) are read once per tap -- the cobs K-blocks of the tap
    // share them -- and the loads of the next step are in flight while the current one is accumulated
    // (ping-pong register buffers; the pipeline drains only at tap boundaries).
    for (int t = blockIdx.x; t < prm.total_tiles; t += gridDim.x) {
      const int nt = t / prm.m_tiles;
      const int grp = (nt * BN) / g.Cg;
      for (int tap = 0; tap < g.K; ++tap) {
        ptx::mbar_wait(&plan_full[ps], pph);
        const uint32_t plan_addr = ptx::smem_u32(sPlan + ps * GI_BM + row0);
        // byte address of this lane's 8 channels in row 0 of output-channel block 0; entries add a 32-bit
        // row offset, every further block 128 bytes
        const char* gcb = reinterpret_cast<const char*>(prm.got) + ((size_t)grp * g.Cog + chunk * 8) * sizeof(T);
        if constexpr (F32) {
          // fp32: per (output-channel block, pass) the bin head's rows are loaded (2 x 16 bytes = 8 channels each),
          // accumulated in fp32, split into bf16 hi / lo and stored into the two A tiles; no ping-pong (the loads of
          // one pass already fill the registers)
          for (int cob = 0; cob < prm.cobs; ++cob) {
            ptx::mbar_wait(&empty[s], ph ^ 1);
            const uint32_t a_addr = ptx::smem_u32(sA + s * A_STAGE);
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
              const uint32_t pa = plan_addr + pass * 4 * (uint32_t)sizeof(PlanRow);
              const uint4 i0 = lds128(pa), i1 = lds128(pa + 16), w0 = lds128(pa + 32), w1 = lds128(pa + 48),
                          meta = lds128(pa + 64);
              const uint32_t idx[GI_NE] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
              const uint32_t wb[GI_NE] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
              const int n = (int)meta.x;
              uint4 va[GI_NE], vb[GI_NE];
#pragma unroll
              for (int i = 0; i < GI_NE; ++i)
                if (i < n) {
                  va[i] = __ldg(reinterpret_cast<const uint4*>(gcb + idx[i]));
                  vb[i] = __ldg(reinterpret_cast<const uint4*>(gcb + idx[i]) + 1);
                }
              float acc[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) acc[k] = 0.f;
              auto fma8 = [&](const uint4& a, const uint4& b, float w) {
                acc[0] = fmaf(w, __uint_as_float(a.x), acc[0]); acc[1] = fmaf(w, __uint_as_float(a.y), acc[1]);
                acc[2] = fmaf(w, __uint_as_float(a.z), acc[2]); acc[3] = fmaf(w, __uint_as_float(a.w), acc[3]);
                acc[4] = fmaf(w, __uint_as_float(b.x), acc[4]); acc[5] = fmaf(w, __uint_as_float(b.y), acc[5]);
                acc[6] = fmaf(w, __uint_as_float(b.z), acc[6]); acc[7] = fmaf(w, __uint_as_float(b.w), acc[7]);
              };
#pragma unroll
              for (int i = 0; i < GI_NE; ++i)
                if (i < n) fma8(va[i], vb[i], __uint_as_float(wb[i]));
              for (int e = GI_NE; e < n; ++e) {   // bins longer than the staged head
                const int2 ee = __ldg(entries2 + (int)meta.y + e);
                const uint4* src = reinterpret_cast<const uint4*>(gcb + (size_t)ee.x * (size_t)(g.Co * sizeof(T)));
                fma8(__ldg(src), __ldg(src + 1), __int_as_float(ee.y));
              }
              uint4 hi, lo;
              gi_split8(acc, hi, lo);
              sts128(a_addr + sw128(row0 + pass * 4, chunk), hi);
              sts128(a_addr + GI_A_BYTES + sw128(row0 + pass * 4, chunk), lo);
            }
            ptx::fence_proxy_async_smem();
            ptx::mbar_arrive_warp(&full_a[s]);
            if (++s == STAGES) { s = 0; ph ^= 1; }
            gcb += 64 * sizeof(T);
          }
          ptx::mbar_arrive_warp(&plan_empty[ps]);
          if (++ps == GI_PLAN_SLOTS) { ps = 0; pph ^= 1; }
          continue;
        }
        uint4 va[GI_NE], vb[GI_NE];
        // the row offsets are re-read from the plan row at every issue (two LDS.128) rather than held for the
        // tap: two load buffers + both passes' offsets do not fit the gather warps' registers
        auto issue = [&](uint4 (&v)[GI_NE], int pass, const char* base, uint32_t after) {
          const uint32_t pa = plan_addr + pass * 4 * (uint32_t)sizeof(PlanRow) + after;
          const uint4 i0 = lds128(pa), i1 = lds128(pa + 16), meta = lds128(pa + 48);
          const uint32_t idx[GI_NE] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
          const int n = (int)meta.x;
#pragma unroll
          for (int i = 0; i < GI_NE; ++i)
            if (i < n) v[i] = __ldg(reinterpret_cast<const uint4*>(base + idx[i]));
        };
        auto finish = [&](const uint4 (&v)[GI_NE], int pass, const char* base, uint32_t a_addr) {
          const uint32_t pa = plan_addr + pass * 4 * (uint32_t)sizeof(PlanRow);
          const uint4 wv = lds128(pa + 32), meta = lds128(pa + 48);
          const T* wh = reinterpret_cast<const T*>(&wv);
          const int n = (int)meta.x;
          Acc8<T> acc;
          acc.clear();
#pragma unroll
          for (int i = 0; i < GI_NE; ++i)
            if (i < n) acc.fma(v[i], wh[i]);
          for (int e = GI_NE; e < n; ++e) {   // bins longer than the staged head (kept serial: unrolling it spills
            const int2 ee = __ldg(entries2 + (int)meta.y + e);   // the ping-pong load buffers of the common path)
            acc.fma(__ldg(reinterpret_cast<const uint4*>(base + (size_t)ee.x * (size_t)(g.Co * 2))),
                    __int_as_float(ee.y));
          }
          sts128(a_addr + sw128(row0 + pass * 4, chunk), acc.pack());
        };
        // the wait for the buffer about to be accumulated comes BEFORE the next loads are issued: every LDG of the
        // loop retires through one scoreboard (see the forward kernel's gather loop, tensor_path.cu)
        issue(va, 0, gcb, 0u);
        for (int cob = 0; cob < prm.cobs; ++cob) {
          const uint32_t da = tc::drain_point(va[0].x, prm.zero);
          issue(vb, 1, gcb, da);
          ptx::mbar_wait(&empty[s], ph ^ 1);
          const uint32_t a_addr = ptx::smem_u32(sA + s * A_STAGE);
          finish(va, 0, gcb, a_addr);
          const uint32_t db = tc::drain_point(vb[0].x, prm.zero);
          if (cob + 1 < prm.cobs) issue(va, 0, gcb + 128, db);
          finish(vb, 1, gcb, a_addr + db);
          ptx::fence_proxy_async_smem();
          ptx::mbar_arrive_warp(&full_a[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
          gcb += 128;
        }
        ptx::mbar_arrive_warp(&plan_empty[ps]);
        if (++ps == GI_PLAN_SLOTS) { ps = 0; pph ^= 1; }
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<512>(tmem_base);
}

// W [Co][Cg][K] -> Wt [C][tap * Cog + col]   (c = grp * Cg + cl, co = grp * Cog + col)
template <typename T>
__global__ void k_transpose_weight(const T* __restrict__ w, T* __restrict__ wt, int Co, int Cg, int Cog, int K) {
  const size_t n = (size_t)Co * Cg * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    // destination-major so the writes coalesce
    const int KC = K * Cog;
    const int c = (int)(i / KC);
    const int r = (int)(i % KC);
    const int tap = r / Cog, col = r - tap * Cog;
    const int grp = c / Cg, cl = c - grp * Cg;
    wt[i] = w[((size_t)(grp * Cog + col) * Cg + cl) * K + tap];
  }
}

// fp32 filters: the same transposition, split into bf16 hi / lo banks
__global__ void k_transpose_weight_split(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi,
                                         __nv_bfloat16* __restrict__ lo, int Co, int Cg, int Cog, int K) {
  const size_t n = (size_t)Co * Cg * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int KC = K * Cog;
    const int c = (int)(i / KC);
    const int r = (int)(i % KC);
    const int tap = r / Cog, col = r - tap * Cog;
    const int grp = c / Cg, cl = c - grp * Cg;
    const float v = w[((size_t)(grp * Cog + col) * Cg + cl) * K + tap];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

template <typename T>
size_t smem_bytes_gi(int stages) {
  return 1024 + (size_t)stages * (GiCfg<T>::A_STAGE + GI_B_BYTES) +
         (size_t)GI_PLAN_SLOTS * GI_BM * sizeof(typename GiCfg<T>::Row) +
         (3 * stages + 2 * GI_PLAN_SLOTS + 4) * sizeof(uint64_t) + 16;
}

template <typename T, bool CL, int STAGES>
void launch_grad_input(const CUtensorMap& tmap, const CUtensorMap& tmap_gx, const CUtensorMap& tmap2,
                       const TcGiParams& prm, int grid, LaunchCtx& ctx) {
  const size_t smem = smem_bytes_gi<T>(STAGES);
  static std::atomic<unsigned long long> configured{0};
  opt_in_dynamic_smem(k_tc_grad_input<T, CL, STAGES>, smem, configured, ctx);
  ProfScope prof(ctx, "tc_grad_input");
  k_tc_grad_input<T, CL, STAGES><<<grid, GI_THREADS, smem, ctx.stream>>>(tmap, tmap_gx, tmap2, prm);
  MDCONV_CHECK_LAUNCH(ctx);
}

int pick_bn_gi(const DevGeom& g) {
  int bn = g.esize == 4 ? GiCfg<float>::MAX_BN_T : GI_MAX_BN;
  while (bn > 64 && (g.Cg % bn != 0 || g.Cdg % bn != 0)) bn >>= 1;
  return bn;
}

size_t num_bins(const DevGeom& g) { return (size_t)g.B * g.dg * g.IN * g.K; }
size_t max_entries(const DevGeom& g) { return ((size_t)g.B * g.dg * g.K * g.P) << g.nd; }

struct GiLayout {
  size_t count, start, entries, scan_tmp, scan_tmp_bytes, total;
};
GiLayout gi_layout(const DevGeom& g) {
  GiLayout L;
  const size_t nb = num_bins(g);
  size_t o = 0;
  L.count = o; o += align256((nb + 1) * sizeof(int));
  L.start = o; o += align256((nb + 1) * sizeof(int));
  L.entries = o; o += align256(max_entries(g) * sizeof(BinEntry));
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const int*)nullptr, (int*)nullptr, (int)(nb + 1));
  L.scan_tmp_bytes = tmp;
  L.scan_tmp = o; o += align256(tmp);
  L.total = o;
  return L;
}

template <typename T, int ND, bool MOD>
void build_bins(const DevGeom& g, const void* off, const void* mask, char* ws, const GiLayout& L, int pos_major,
                LaunchCtx& ctx) {
  int* count = reinterpret_cast<int*>(ws + L.count);
  int* start = reinterpret_cast<int*>(ws + L.start);
  BinEntry* entries = reinterpret_cast<BinEntry*>(ws + L.entries);
  const size_t nb = num_bins(g);
  cudaError_t e = cudaMemsetAsync(count, 0, (nb + 1) * sizeof(int), ctx.stream);
  if (e != cudaSuccess && ctx.err == cudaSuccess) ctx.err = e;
  const long long n = (long long)g.B * g.dg * g.K * g.P;
  const int blocks = (int)((n + 255) / 256 > 148 * 16 ? 148 * 16 : (n + 255) / 256);
  {
    ProfScope prof(ctx, "bin_count");
    k_bin_entries<T, ND, MOD, false><<<blocks, 256, 0, ctx.stream>>>(g, (const T*)off, (const T*)mask, count, start, entries,
                                                                     pos_major);
    MDCONV_CHECK_LAUNCH(ctx);
  }
  {
    ProfScope prof(ctx, "bin_scan");
    size_t tmp = L.scan_tmp_bytes;
    e = cub::DeviceScan::ExclusiveSum(ws + L.scan_tmp, tmp, count, start, (int)(nb + 1), ctx.stream);
    if (e != cudaSuccess && ctx.err == cudaSuccess) ctx.err = e;
    ctx.launches++;
  }
  {
    ProfScope prof(ctx, "bin_fill");
    k_bin_entries<T, ND, MOD, true><<<blocks, 256, 0, ctx.stream>>>(g, (const T*)off, (const T*)mask, count, start, entries,
                                                                    pos_major);
    MDCONV_CHECK_LAUNCH(ctx);
  }
  {
    ProfScope prof(ctx, "bin_sort");
    const int sb = (int)((nb + 255) / 256 > 148 * 16 ? 148 * 16 : (nb + 255) / 256);
    // `count` is all zero again here (the fill pass took every slot): it doubles as the long-bin work list
    k_bin_sort<<<sb, 256, 0, ctx.stream>>>(start, entries, nb, count);
    MDCONV_CHECK_LAUNCH(ctx);
    k_bin_sort_big<<<148, 256, 0, ctx.stream>>>(start, entries, count);
    MDCONV_CHECK_LAUNCH(ctx);
  }
}

}  // namespace

bool grad_input_supported(const DevGeom& g) {
  if (g.Cog % 64 != 0 || g.Cg % 64 != 0 || g.Cdg % 64 != 0 || g.Co % 8 != 0) return false;
  if (num_bins(g) + 1 > 2147483647ULL || max_entries(g) > 2147483647ULL) return false;
  if ((long long)g.B * g.P * g.Co * g.esize > 4294967295LL) return false;  // 32-bit byte offsets into grad_output
  return true;
}

// channels-last copy of grad_output (NCHW callers only: channels-last callers' grad_output is gathered in place)
static size_t got_bytes(const DevGeom& g) { return g.cl ? 0 : align256((size_t)g.B * g.Co * g.P * g.esize); }
// transposed filter bank(s): one 16-bit bank, or bf16 hi + lo for fp32 callers
static size_t wt_bytes(const DevGeom& g) { return align256((size_t)g.Co * g.Kg * 2) * (g.esize == 4 ? 2 : 1); }

// ---- col2im-gather path ---------------------------------------------------------------------------
// grad_columns ([samples][K][P][C], storage dtype) live in ONE workspace buffer capped by the budget
// (MDCONV_B200_GCOL_MAX_MB / mdconv_set_gcol_budget_mb(), default 1024 MiB).  A batch that fits goes through in one
// (k_tc_grad_sample, k_col2im_gather) pair; a larger one is walked in chunks of whole samples that fit, re-using the
// buffer (measured at cfg2: 4 / 8 chunks cost +9 % / +27 % of the pair -- smaller launches fill the machine worse, and
// keeping the buffer L2-resident buys nothing).  Budget 0 -- and fp32 callers, and batches whose offsets do not fit 31
// bits -- take the contraction path (k_tc_grad_input), which needs no buffer.
static std::atomic<long long> g_gcol_budget_mb{-1};
long long set_gcol_budget_mb(long long mb) {
  const long long prev = g_gcol_budget_mb.load();
  g_gcol_budget_mb.store(mb < 0 ? 0 : mb);
  return prev;
}
static long long gcol_budget_bytes() {
  long long mb = g_gcol_budget_mb.load();
  if (mb < 0) {
    const char* e = getenv("MDCONV_B200_GCOL_MAX_MB");
    mb = e ? atoll(e) : 1024;
    if (mb < 0) mb = 0;
    g_gcol_budget_mb.store(mb);
  }
  return mb << 20;
}
static long long gcol_sample_bytes(const DevGeom& g) { return (long long)g.P * g.K * g.C * 2; }
bool grad_input_uses_gcol(const DevGeom& g) {
  if (g.esize != 2 || gcol_budget_bytes() <= 0) return false;
  if ((long long)g.B * g.P * g.K * g.C > 2147483647LL) return false;   // an entry holds a 31-bit element offset
  return gcol_sample_bytes(g) <= (512LL << 20);                        // (one sample is the smallest chunk)
}
int gcol_chunk_samples(const DevGeom& g) {
  long long n = gcol_budget_bytes() / gcol_sample_bytes(g);
  if (n < 1) n = 1;
  if (n > g.B) n = g.B;
  return (int)n;
}
size_t gcol_bytes(const DevGeom& g) { return align256((size_t)gcol_chunk_samples(g) * (size_t)gcol_sample_bytes(g)); }

size_t grad_input_workspace(const DevGeom& g) {
  // gather path: grad_columns buffer + inverse index;  contraction path: channels-last grad_output + transposed weight + index
  if (grad_input_uses_gcol(g)) return gcol_bytes(g) + gi_layout(g).total;
  return got_bytes(g) + wt_bytes(g) + gi_layout(g).total;
}
// offset of the inverse index inside the grad_input workspace
static size_t gi_index_offset(const DevGeom& g) { return grad_input_uses_gcol(g) ? gcol_bytes(g) : got_bytes(g) + wt_bytes(g); }
void* grad_input_gcol(const DevGeom& g, void* ws) { return grad_input_uses_gcol(g) ? ws : nullptr; }

// grad_input of samples [sample0, sample0 + samples) from the buffer k_tc_grad_sample has just filled for them
template <typename T>
void tensor_grad_input_from_gcol(const DevGeom& g, void* gx, void* ws, int sample0, int samples, LaunchCtx& ctx) {
  if constexpr (std::is_same<T, float>::value) {
    (void)g; (void)gx; (void)ws; (void)sample0; (void)samples;
    if (ctx.err == cudaSuccess) ctx.err = cudaErrorNotSupported;
    return;
  } else {
    char* base = reinterpret_cast<char*>(ws);
    // entries hold offsets from the start of the WHOLE batch's grad_columns; the buffer holds this chunk's part
    const T* gcol = reinterpret_cast<const T*>(base) - (size_t)sample0 * g.P * g.K * g.C;
    base += gcol_bytes(g);
    const GiLayout L = gi_layout(g);
    if (ctx.side_join != nullptr) {   // inverse index is ready
      cudaStreamWaitEvent(ctx.stream, ctx.side_join, 0);
      ctx.side_join = nullptr;
    }
    const long long pos_base = (long long)sample0 * g.IN, total_pos = (long long)(sample0 + samples) * g.IN;
    const int slabs = (g.C + C2I_CS - 1) / C2I_CS;
    const int cs = g.C < C2I_CS ? g.C : C2I_CS;
    int tile_pos = (512 / (cs >> 3)) & ~7;          // ~512 (position, chunk) pairs per tile, two per thread (256 / 1024: +5 %)
    if (tile_pos < 8) tile_pos = 8;
    if (tile_pos > C2I_MAX_POS) tile_pos = C2I_MAX_POS;
    const int n_tiles = (int)((total_pos - pos_base + tile_pos - 1) / tile_pos) * slabs;
    const size_t smem_max = (size_t)C2I_MAX_POS * (C2I_CS + 8) * sizeof(T);
    const size_t smem = g.cl ? 0 : (size_t)tile_pos * (cs + 8) * sizeof(T);
    const int sms = sm_count() - ctx.spare_sms > 0 ? sm_count() - ctx.spare_sms : 1;
    int per_sm = smem > 0 ? (int)((200 * 1024) / smem) : C2I_MINB;   // (blocks per SM by registers: __launch_bounds__)
    if (per_sm > C2I_MINB) per_sm = C2I_MINB;
    if (per_sm < 1) per_sm = 1;
    const int grid = n_tiles < sms * per_sm ? n_tiles : sms * per_sm;
    const int* start = reinterpret_cast<const int*>(base + L.start);
    const BinEntry* entries = reinterpret_cast<const BinEntry*>(base + L.entries);
    ProfScope prof(ctx, "col2im_gather");
    if (g.cl) {
      k_col2im_gather<T, true><<<grid, C2I_THREADS, 0, ctx.stream>>>(g, gcol, start, entries, (T*)gx, pos_base, total_pos,
                                                                     n_tiles, slabs, tile_pos);
    } else {
      static std::atomic<unsigned long long> configured{0};
      opt_in_dynamic_smem(k_col2im_gather<T, false>, smem_max, configured, ctx);
      k_col2im_gather<T, false><<<grid, C2I_THREADS, smem, ctx.stream>>>(g, gcol, start, entries, (T*)gx, pos_base,
                                                                         total_pos, n_tiles, slabs, tile_pos);
    }
    MDCONV_CHECK_LAUNCH(ctx);
  }
}
template void tensor_grad_input_from_gcol<float>(const DevGeom&, void*, void*, int, int, LaunchCtx&);
template void tensor_grad_input_from_gcol<__half>(const DevGeom&, void*, void*, int, int, LaunchCtx&);
template void tensor_grad_input_from_gcol<__nv_bfloat16>(const DevGeom&, void*, void*, int, int, LaunchCtx&);

// Stage 1 (can be issued early): the inverse index, on the side stream when there is one.  Returns with
// the side stream forked from ctx.stream; tensor_grad_input() joins it before the GEMM kernel.
template <typename T>
void tensor_grad_input_prepare(const DevGeom& g, const void* off, const void* mask, void* ws, LaunchCtx& ctx) {
  char* base = reinterpret_cast<char*>(ws) + gi_index_offset(g);
  const GiLayout L = gi_layout(g);
  // no fork while the caller's stream is being captured into a CUDA graph: the whole call then stays on that stream
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(ctx.stream, &cap) != cudaSuccess) { (void)cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
  SideStream* ss = cap == cudaStreamCaptureStatusNone ? side_stream() : nullptr;
  cudaStream_t main_stream = ctx.stream;
  if (ss) {
    cudaEventRecord(ss->fork, main_stream);
    cudaStreamWaitEvent(ss->stream, ss->fork, 0);
    ctx.stream = ss->stream;
  }
  const int pos_major = grad_input_uses_gcol(g) ? 1 : 0;   // the layout the consumer of this call reads
  if (g.nd == 2) {
    if (g.modulated) build_bins<T, 2, true>(g, off, mask, base, L, pos_major, ctx);
    else build_bins<T, 2, false>(g, off, mask, base, L, pos_major, ctx);
  } else {
    if (g.modulated) build_bins<T, 3, true>(g, off, mask, base, L, pos_major, ctx);
    else build_bins<T, 3, false>(g, off, mask, base, L, pos_major, ctx);
  }
  if (ss) {
    cudaEventRecord(ss->join, ss->stream);
    ctx.stream = main_stream;
    ctx.side_join = ss->join;
  }
}

template <typename T>
void tensor_grad_input(const DevGeom& g, const void* w, const void* go, void* gx, void* ws, LaunchCtx& ctx) {
  char* base = reinterpret_cast<char*>(ws);
  const T* got = g.cl ? reinterpret_cast<const T*>(go) : reinterpret_cast<const T*>(base);
  base += got_bytes(g);
  constexpr bool F32 = std::is_same<T, float>::value;
  using E = typename std::conditional<F32, __nv_bfloat16, T>::type;   // element type of the filter bank(s)
  char* wt = base;
  base += wt_bytes(g);
  const GiLayout L = gi_layout(g);

  if (!g.cl) launch_to_channels_last<T>((const T*)go, const_cast<T*>(got), g.B, g.Co, g.P, ctx);
  {
    ProfScope prof(ctx, "transpose_weight");
    const size_t n = (size_t)g.Co * g.Kg;
    const int blocks = (int)((n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256);
    if constexpr (F32)
      k_transpose_weight_split<<<blocks, 256, 0, ctx.stream>>>(
          (const float*)w, reinterpret_cast<__nv_bfloat16*>(wt), reinterpret_cast<__nv_bfloat16*>(wt + wt_bytes(g) / 2),
          g.Co, g.Cg, g.Cog, g.K);
    else
      k_transpose_weight<T><<<blocks, 256, 0, ctx.stream>>>((const T*)w, reinterpret_cast<T*>(wt), g.Co, g.Cg, g.Cog, g.K);
    MDCONV_CHECK_LAUNCH(ctx);
  }
  if (ctx.side_join != nullptr) {   // inverse index is ready
    cudaStreamWaitEvent(ctx.stream, ctx.side_join, 0);
    ctx.side_join = nullptr;
  }

  TcGiParams prm;
  prm.g = g;
  prm.got = got;
  prm.bin_start = reinterpret_cast<const int*>(base + L.start);
  prm.entries = reinterpret_cast<const BinEntry*>(base + L.entries);
  prm.gx = gx;
  prm.BN = pick_bn_gi(g);
  prm.zero = 0;
  const long long total_pos = (long long)g.B * g.IN;
  prm.m_tiles = (int)((total_pos + GI_BM - 1) / GI_BM);
  prm.n_tiles = g.C / prm.BN;
  prm.total_tiles = prm.m_tiles * prm.n_tiles;
  prm.cobs = g.Cog / 64;
  prm.kblocks = g.K * prm.cobs;

  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    if (ctx.err == cudaSuccess) ctx.err = cudaErrorNotSupported;
    return;
  }
  CUtensorMap tmap, tmap2, tmap_gx;
  memset(&tmap2, 0, sizeof(tmap2));
  memset(&tmap_gx, 0, sizeof(tmap_gx));
  const cuuint64_t dims[2] = {(cuuint64_t)g.K * g.Cog, (cuuint64_t)g.C};
  const cuuint64_t strides[1] = {(cuuint64_t)g.K * g.Cog * sizeof(E)};
  const cuuint32_t box[2] = {64, (cuuint32_t)prm.BN};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&tmap, tmap_dtype<E>(), 2, wt, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS && F32)
    r = enc(&tmap2, tmap_dtype<E>(), 2, wt + wt_bytes(g) / 2, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (ctx.err == cudaSuccess) ctx.err = cudaErrorInvalidValue;
    return;
  }
  const int sms = sm_count() - ctx.spare_sms > 0 ? sm_count() - ctx.spare_sms : 1;
  const int grid = prm.total_tiles < sms ? prm.total_tiles : sms;
  if constexpr (F32) {
    launch_grad_input<T, false, 2>(tmap, tmap_gx, tmap2, prm, grid, ctx);
  } else {
    if (g.cl && !encode_rows_by_channels<T>(&tmap_gx, gx, (size_t)g.B * g.IN, g.C)) {
      if (ctx.err == cudaSuccess) ctx.err = cudaErrorInvalidValue;
      return;
    }
    if (g.nd == 3) {
      if (g.cl) launch_grad_input<T, true, 2>(tmap, tmap_gx, tmap2, prm, grid, ctx);
      else launch_grad_input<T, false, 2>(tmap, tmap_gx, tmap2, prm, grid, ctx);
    } else {
      if (g.cl) launch_grad_input<T, true, GI_STAGES>(tmap, tmap_gx, tmap2, prm, grid, ctx);
      else launch_grad_input<T, false, GI_STAGES>(tmap, tmap_gx, tmap2, prm, grid, ctx);
    }
  }
}

template void tensor_grad_input<float>(const DevGeom&, const void*, const void*, void*, void*, LaunchCtx&);
template void tensor_grad_input_prepare<float>(const DevGeom&, const void*, const void*, void*, LaunchCtx&);
template void tensor_grad_input<__half>(const DevGeom&, const void*, const void*, void*, void*, LaunchCtx&);
template void tensor_grad_input<__nv_bfloat16>(const DevGeom&, const void*, const void*, void*, void*, LaunchCtx&);
template void tensor_grad_input_prepare<__half>(const DevGeom&, const void*, const void*, void*, LaunchCtx&);
template void tensor_grad_input_prepare<__nv_bfloat16>(const DevGeom&, const void*, const void*, void*, LaunchCtx&);

}  // namespace tc
}  // namespace mdconv
