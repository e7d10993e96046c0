// generic_kernels.cu -- fused SIMT implicit-GEMM kernels: every dtype (f32/f16/bf16/f64), 2D/3D,
// plain/modulated, any groups / deformable groups / stride / dilation / kernel shape.
//
// No column buffer is ever written to HBM: the offset-driven gather is evaluated inside the tile
// loaders of three implicit GEMMs
//   forward      y[co, pix]      = sum_kk W[co, kk]      * col[kk, pix]        (kk = c*K + tap)
//   bwd-data     gcol[kk, pix]   = sum_co W[co, kk]      * go[co, pix]   -> scattered in the epilogue
//   bwd-weight   gW[co, kk]      = sum_pix go[co, pix]   * col[kk, pix]
// This is the path used for shapes/dtypes the tcgen05 path does not take (fp32 needs true fp32
// accumulation of fp32 products for the 1e-4 bar, fp64 has no tensor path on sm_100a).
//
// Replaces (reference, src/): K1-K8 im2col / gradient kernels and the G1-G4 addmm_ call sites,
// e.g. mdeformable_conv.cu:38-87,172-184,203-318,412-452.
#include "common.cuh"

namespace mdconv {

namespace {

constexpr int BM = 64;   // output positions per tile
constexpr int BN = 64;   // output channels (or kk rows) per tile
constexpr int BK = 16;   // reduction chunk
constexpr int NT = 256;  // threads per block

template <typename F> __device__ __forceinline__ void atomic_add(F* p, F v) { atomicAdd(p, v); }

// ------------------------------------------------------------------------------------------------
// forward:  grid (B * ceil(P/BM), ceil(Cog/BN), groups)
// ------------------------------------------------------------------------------------------------
// BNT output channels per tile (64 or 128): each thread owns 4 pixels x BNT/16 channels; the wider tile does 32
// FMAs per three 16-byte shared-memory reads instead of 16 per two, which is what paces this kernel.
template <typename T, typename F, int ND, bool MOD, int BNT>
__global__ void __launch_bounds__(NT) k_generic_forward(const DevGeom g, const T* __restrict__ x,
                                                        const T* __restrict__ w, const T* __restrict__ bias,
                                                        const T* __restrict__ off, const T* __restrict__ mask,
                                                        T* __restrict__ y) {
  constexpr int NC = 1 << ND;
  __shared__ F As[BK][BM + 4];  // col chunk   [channel][pix]
  constexpr int CJ = BNT / 16;  // output channels per thread
  __shared__ F Bs[BK][BNT + 4];  // weight chunk [channel][co]
  // sampling plan of the current (tap, deformable group): per pixel the corner indices and the
  // interpolation weight x mask, computed ONCE and reused by every channel of the group (the
  // reference recomputes it per channel, src/mdeformable_conv.cu:59-83)
  __shared__ int Pi[NC][BM];
  __shared__ F Pw[NC][BM];

  // pixel tiles run over the flattened (sample, position) index, so a tile may straddle two samples and no
  // tile is partly empty except the very last one (P = 196 would otherwise waste 23 % of every 4th tile)
  const long long total_pix = (long long)g.B * g.P;
  const long long gp0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BNT;
  const int grp = blockIdx.z;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;

  // the plan threads (tid < BM) own one pixel each: its sample, position and per-sample offset / mask bases
  const long long my_gp = gp0 + (tid < BM ? tid : 0);
  const bool my_valid = tid < BM && my_gp < total_pix;
  const int my_b = my_valid ? (int)(my_gp / g.P) : 0;
  const int my_p = my_valid ? (int)(my_gp - (long long)my_b * g.P) : 0;
  const T* off_b = off + (size_t)my_b * g.dg * ND * g.K * g.P;
  const T* mask_b = MOD ? mask + (size_t)my_b * g.dg * g.K * g.P : nullptr;
  const int x_base = my_b * g.C * g.IN;   // folded into the corner indices (all tensors are < 2^31 elements)

  F acc[4][CJ];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < CJ; ++j) acc[i][j] = (F)0;

  const int lm = tid % BM;       // pixel handled in the col loader
  const int lk0 = tid / BM;      // first channel row handled (stride 4)
  const int wk = tid % BK;       // weight loader: channel
  const int wn0 = tid / BK;      // weight loader: first co (stride 16)

  // channels of this conv group, walked in runs that share a deformable group
  int cl0 = 0;
  while (cl0 < g.Cg) {
    const int cglob = grp * g.Cg + cl0;
    const int dgi = cglob / g.Cdg;
    const int room = (dgi + 1) * g.Cdg - cglob;
    const int cl1 = cl0 + room < g.Cg ? cl0 + room : g.Cg;
    for (int tap = 0; tap < g.K; ++tap) {
      // ---- plan of (pixel, tap, dgi) ----
      __syncthreads();  // previous users of Pi / Pw are done
      if (tid < BM) {
        const int p = my_p;
#pragma unroll
        for (int c = 0; c < NC; ++c) { Pi[c][tid] = -1; Pw[c][tid] = (F)0; }   // -1: absent corner, never loaded
        if (my_valid) {
          SamplePoint<ND, F> sp;
          make_sample_point<ND, T, F>(g, off_b, dgi, tap, p, sp);
          if (sp.in_range) {
            F m = (F)1;
            if (MOD) m = ld_as<F>(mask_b + (size_t)(dgi * g.K + tap) * g.P + p);
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              F wgt = m;
              bool ok = true;
              int idx = 0;
#pragma unroll
              for (int a = 0; a < ND; ++a) {
                const int hi = (c >> a) & 1;
                const int ia = sp.lo[a] + hi;
                ok = ok && (hi ? (ia <= g.in[a] - 1) : (ia >= 0));
                wgt *= hi ? sp.fr[a] : ((F)1 - sp.fr[a]);
                idx = idx * g.in[a] + ia;
              }
              if (ok) { Pi[c][tid] = x_base + idx; Pw[c][tid] = wgt; }
            }
          }
        }
      }
      __syncthreads();
      int pi[NC];
      F pw[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) { pi[c] = Pi[c][lm]; pw[c] = Pw[c][lm]; }

      for (int c0 = cl0; c0 < cl1; c0 += BK) {
        // ---- gather the col chunk: BK channels x BM pixels of this tap ----
#pragma unroll
        for (int i = 0; i < BK / 4; ++i) {
          const int cl = c0 + lk0 + 4 * i;
          F v = (F)0;
          if (cl < cl1) {
            const T* plane = x + (size_t)(grp * g.Cg + cl) * g.IN;
#pragma unroll
            for (int c = 0; c < NC; ++c)
              if (pi[c] >= 0) v += pw[c] * ld_as<F>(plane + pi[c]);   // absent corners contribute an exact 0
          }
          As[lk0 + 4 * i][lm] = v;
        }
        // ---- weight chunk ----
#pragma unroll
        for (int i = 0; i < BNT / 16; ++i) {
          const int n = wn0 + 16 * i;
          const int cl = c0 + wk;
          F v = (F)0;
          if (cl < cl1 && n0 + n < g.Cog) v = ld_as<F>(w + (size_t)(grp * g.Cog + n0 + n) * g.Kg + (size_t)cl * g.K + tap);
          Bs[wk][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
          F a[4], bq[CJ];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = As[kk][tx * 4 + i];
#pragma unroll
          for (int j = 0; j < CJ; ++j) bq[j] = Bs[kk][ty * CJ + j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < CJ; ++j) acc[i][j] += a[i] * bq[j];
        }
        __syncthreads();
      }
    }
    cl0 = cl1;
  }

#pragma unroll
  for (int j = 0; j < CJ; ++j) {
    const int n = n0 + ty * CJ + j;
    if (n >= g.Cog) continue;
    const int co = grp * g.Cog + n;
    const F bv = g.with_bias ? ld_as<F>(bias + co) : (F)0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long gp = gp0 + tx * 4 + i;
      if (gp < total_pix) {
        const int bb = (int)(gp / g.P);
        const int p = (int)(gp - (long long)bb * g.P);
        y[((size_t)bb * g.Co + co) * g.P + p] = cvt_to<T, F>(acc[i][j] + bv);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward data:  grid (B * ceil(P/BM), ceil(Kg/BN), groups)
//   phase 1  gcol tile [BN kk][BM pix] = W^T * go           (reduction over Cog)
//   phase 2  per (pix, tap): walk the tile's channels, reuse one sampling plan per deformable
//            group, scatter grad_input (atomics), and sum grad_offset / grad_mask over the tile's
//            channels in registers -> one atomic per (pix, tap, axis) per tile.
// Accumulators gx/goff/gmask are F-typed buffers (the outputs themselves for f32/f64, fp32
// workspace for f16/bf16), zeroed by the host before launch.
// ------------------------------------------------------------------------------------------------
template <typename T, typename F, int ND, bool MOD>
__global__ void __launch_bounds__(NT) k_generic_bwd_data(const DevGeom g, const T* __restrict__ x,
                                                         const T* __restrict__ w, const T* __restrict__ off,
                                                         const T* __restrict__ mask, const T* __restrict__ go,
                                                         F* __restrict__ gx, F* __restrict__ goff,
                                                         F* __restrict__ gmask) {
  // The gcol tile [kk][pix] outlives the two operand chunks, which alias its storage
  // (keeps the f64 instantiation under the 48 KB static limit).
  __shared__ F Cs[BN][BM + 1];
  static_assert(2 * BK * (BN + 4) <= BN * (BM + 1), "operand chunks must fit inside the gcol tile");
  F(*Ws)[BN + 4] = reinterpret_cast<F(*)[BN + 4]>(&Cs[0][0]);  // [co][kk]
  F(*Gs)[BM + 4] = Ws + BK;                                     // [co][pix]

  const int tilesP = (g.P + BM - 1) / BM;
  const int b = blockIdx.x / tilesP;
  const int p0 = (blockIdx.x % tilesP) * BM;
  const int kk0 = blockIdx.y * BN;
  const int grp = blockIdx.z;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;  // tx -> 4 pixels, ty -> 4 kk rows

  const T* go_b = go + ((size_t)b * g.Co + (size_t)grp * g.Cog) * g.P;
  const T* w_g = w + (size_t)grp * g.Cog * g.Kg;

  F acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = (F)0;

  const int l64 = tid % 64, l4 = tid / 64;
  for (int c0 = 0; c0 < g.Cog; c0 += BK) {
#pragma unroll
    for (int i = 0; i < BK / 4; ++i) {
      const int co = c0 + l4 + 4 * i;
      F wv = (F)0, gv = (F)0;
      if (co < g.Cog) {
        if (kk0 + l64 < g.Kg) wv = ld_as<F>(w_g + (size_t)co * g.Kg + kk0 + l64);
        if (p0 + l64 < g.P) gv = ld_as<F>(go_b + (size_t)co * g.P + p0 + l64);
      }
      Ws[l4 + 4 * i][l64] = wv;
      Gs[l4 + 4 * i][l64] = gv;
    }
    __syncthreads();
#pragma unroll
    for (int cc = 0; cc < BK; ++cc) {
      F a[4], bq[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Gs[cc][tx * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bq[j] = Ws[cc][ty * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * bq[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) Cs[ty * 4 + j][tx * 4 + i] = acc[i][j];
  __syncthreads();

  // ---- phase 2: scatter ----
  const int m = tid % BM;
  const int q = tid / BM;  // 0..3, taps are strided by 4 over q
  const int p = p0 + m;
  if (p >= g.P) return;
  const bool rule_b = g.rule != 0;
  const T* x_b = x + (size_t)b * g.C * g.IN;
  const T* off_b = off + (size_t)b * g.dg * ND * g.K * g.P;
  const T* mask_b = MOD ? mask + (size_t)b * g.dg * g.K * g.P : nullptr;
  F* gx_b = gx + (size_t)b * g.C * g.IN;
  F* goff_b = goff + (size_t)b * g.dg * ND * g.K * g.P;
  F* gmask_b = MOD ? gmask + (size_t)b * g.dg * g.K * g.P : nullptr;
  const int kk_end = min(kk0 + BN, g.Kg);

  for (int tap = q; tap < g.K; tap += 4) {
    // channels cl with kk0 <= cl*K + tap < kk_end
    int cl = (kk0 - tap + g.K - 1) / g.K;
    if (cl < 0) cl = 0;
    int cur_dg = -1;
    SamplePoint<ND, F> s;
    CornerPlan<ND, F> cp;
    F mval = (F)1;
    F so[ND];
    F sm = (F)0;
#pragma unroll
    for (int a = 0; a < ND; ++a) so[a] = (F)0;

    for (; cl * g.K + tap < kk_end; ++cl) {
      const int kk = cl * g.K + tap;
      const int c = grp * g.Cg + cl;
      const int dgi = fdiv(c, g.fd_Cdg);
      if (dgi != cur_dg) {
        if (cur_dg >= 0) {  // flush sums of the previous deformable group
          if (cp.offset_grad) {
#pragma unroll
            for (int a = 0; a < ND; ++a)
              atomic_add(goff_b + ((size_t)(cur_dg * g.K + tap) * ND + a) * g.P + p, so[a]);
          }
          if (MOD) atomic_add(gmask_b + (size_t)(cur_dg * g.K + tap) * g.P + p, sm);
#pragma unroll
          for (int a = 0; a < ND; ++a) so[a] = (F)0;
          sm = (F)0;
        }
        cur_dg = dgi;
        make_sample_point<ND, T, F>(g, off_b, dgi, tap, p, s);
        make_corner_plan<ND, F>(g, s, rule_b, cp);
        if (MOD) mval = ld_as<F>(mask_b + (size_t)(dgi * g.K + tap) * g.P + p);
      }
      const F gc = Cs[kk - kk0][m];
      const F dval = mval * gc;
      const T* plane = x_b + (size_t)c * g.IN;
      F* gplane = gx_b + (size_t)c * g.IN;
      F v[1 << ND];
      F val = (F)0;
#pragma unroll
      for (int cr = 0; cr < (1 << ND); ++cr) {
        v[cr] = ((cp.load_mask >> cr) & 1u) ? ld_as<F>(plane + cp.idx[cr]) : (F)0;
        val += cp.wgt[cr] * v[cr];
      }
#pragma unroll
      for (int cr = 0; cr < (1 << ND); ++cr)
        if ((cp.add_mask >> cr) & 1u) atomic_add(gplane + cp.idx[cr], cp.wgt[cr] * dval);
      if (cp.offset_grad) {
#pragma unroll
        for (int a = 0; a < ND; ++a) so[a] += coord_derivative<ND, F>(s, v, a) * dval;
      }
      if (MOD) sm += gc * val;
    }
    if (cur_dg >= 0) {
      if (cp.offset_grad) {
#pragma unroll
        for (int a = 0; a < ND; ++a)
          atomic_add(goff_b + ((size_t)(cur_dg * g.K + tap) * ND + a) * g.P + p, so[a]);
      }
      if (MOD) atomic_add(gmask_b + (size_t)(cur_dg * g.K + tap) * g.P + p, sm);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward weight:  grid (ceil(Kg/BN), ceil(Cog/BN) * groups, splits)
//   gW[co, kk] += sum over this split's (sample, 32-position chunk) units of go * col
// ------------------------------------------------------------------------------------------------
constexpr int BP = 32;  // positions per reduction chunk

template <typename T, typename F, int ND, bool MOD>
__global__ void __launch_bounds__(NT) k_generic_bwd_weight(const DevGeom g, const T* __restrict__ x,
                                                           const T* __restrict__ off, const T* __restrict__ mask,
                                                           const T* __restrict__ go, F* __restrict__ gw) {
  __shared__ F Gs[BP][BN + 4];  // [pix][co]
  __shared__ F As[BP][BN + 4];  // [pix][kk]

  const int kk0 = blockIdx.x * BN;
  const int tilesN = (g.Cog + BN - 1) / BN;
  const int grp = blockIdx.y / tilesN;
  const int n0 = (blockIdx.y % tilesN) * BN;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;  // tx -> 4 kk, ty -> 4 co
  const int chunksP = (g.P + BP - 1) / BP;
  const int units = g.B * chunksP;

  F acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = (F)0;

  const int lm = tid % BP, lr0 = tid / BP;  // loader: position, first row (stride 8)

  for (int u = blockIdx.z; u < units; u += gridDim.z) {
    const int b = u / chunksP;
    const int p = (u % chunksP) * BP + lm;
    const T* x_b = x + (size_t)b * g.C * g.IN;
    const T* off_b = off + (size_t)b * g.dg * ND * g.K * g.P;
    const T* mask_b = MOD ? mask + (size_t)b * g.dg * g.K * g.P : nullptr;
    const T* go_b = go + ((size_t)b * g.Co + (size_t)grp * g.Cog) * g.P;
#pragma unroll
    for (int i = 0; i < BN / 8; ++i) {
      const int r = lr0 + 8 * i;
      F gv = (F)0, cv = (F)0;
      if (p < g.P) {
        if (n0 + r < g.Cog) gv = ld_as<F>(go_b + (size_t)(n0 + r) * g.P + p);
        const int kk = kk0 + r;
        if (kk < g.Kg) {
          const int cl = fdiv(kk, g.fd_K), tap = kk - cl * g.K;
          const int c = grp * g.Cg + cl;
          const int dgi = fdiv(c, g.fd_Cdg);
          SamplePoint<ND, F> s;
          make_sample_point<ND, T, F>(g, off_b, dgi, tap, p, s);
          cv = interp_forward<ND, T, F>(g, x_b + (size_t)c * g.IN, s);
          if (MOD) cv *= ld_as<F>(mask_b + (size_t)(dgi * g.K + tap) * g.P + p);
        }
      }
      Gs[lm][r] = gv;
      As[lm][r] = cv;
    }
    __syncthreads();
#pragma unroll 8
    for (int pp = 0; pp < BP; ++pp) {
      F a[4], bq[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[pp][tx * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bq[j] = Gs[pp][ty * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * bq[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + ty * 4 + j;
    if (n >= g.Cog) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = kk0 + tx * 4 + i;
      if (kk < g.Kg) atomic_add(gw + (size_t)(grp * g.Cog + n) * g.Kg + kk, acc[i][j]);
    }
  }
}

// grad_bias[co] = sum over samples and positions of grad_output; one block per channel,
// fixed reduction order (deterministic).
template <typename T, typename F>
__global__ void __launch_bounds__(NT) k_grad_bias(const DevGeom g, const T* __restrict__ go, T* __restrict__ gb,
                                                  float* __restrict__ gb32, int ld) {
  __shared__ F red[NT];
  const int co = blockIdx.x;
  F acc = (F)0;
  for (int b = 0; b < g.B; ++b) {
    const T* row = go + ((size_t)b * g.Co + co) * ld;
    for (int p = threadIdx.x; p < g.P; p += NT) acc += ld_as<F>(row + p);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int sft = NT / 2; sft > 0; sft >>= 1) {
    if ((int)threadIdx.x < sft) red[threadIdx.x] += red[threadIdx.x + sft];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    gb[co] = cvt_to<T, F>(red[0]);
    if (gb32 != nullptr) gb32[co] = (float)red[0];
  }
}

template <typename T>
__global__ void k_convert_f32(const float* __restrict__ src, T* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = cvt_to<T, float>(src[i]);
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

struct BwdSizes {
  size_t n_gx, n_goff, n_gmask, n_gw;
};
BwdSizes bwd_sizes(const DevGeom& g) {
  BwdSizes s;
  s.n_gx = (size_t)g.B * g.C * g.IN;
  s.n_goff = (size_t)g.B * g.dg * g.nd * g.K * g.P;
  s.n_gmask = g.modulated ? (size_t)g.B * g.dg * g.K * g.P : 0;
  s.n_gw = (size_t)g.Co * g.Kg;
  return s;
}

template <typename T, typename F, int ND, bool MOD>
void launch_forward(const DevGeom& g, const void* x, const void* w, const void* bias, const void* off,
                    const void* mask, void* y, LaunchCtx& ctx) {
  const unsigned tiles = (unsigned)(((long long)g.B * g.P + BM - 1) / BM);
  ProfScope prof(ctx, "generic_forward");
  if (g.Cog >= 96 && sizeof(F) == 4) {   // wide tile: 128 output channels (the fp64 build keeps the 64-wide one)
    dim3 grid(tiles, (unsigned)((g.Cog + 127) / 128), (unsigned)g.g);
    k_generic_forward<T, F, ND, MOD, 128><<<grid, NT, 0, ctx.stream>>>(g, (const T*)x, (const T*)w, (const T*)bias,
                                                                       (const T*)off, (const T*)mask, (T*)y);
  } else {
    dim3 grid(tiles, (unsigned)((g.Cog + BN - 1) / BN), (unsigned)g.g);
    k_generic_forward<T, F, ND, MOD, BN><<<grid, NT, 0, ctx.stream>>>(g, (const T*)x, (const T*)w, (const T*)bias,
                                                                      (const T*)off, (const T*)mask, (T*)y);
  }
  MDCONV_CHECK_LAUNCH(ctx);
}

template <typename T>
void convert_from_f32(const float* src, void* dst, size_t n, LaunchCtx& ctx) {
  if (n == 0) return;
  const int blocks = (int)((n + 1023) / 1024 > 148 * 8 ? 148 * 8 : (n + 1023) / 1024);
  k_convert_f32<T><<<blocks, 256, 0, ctx.stream>>>(src, (T*)dst, n);
  MDCONV_CHECK_LAUNCH(ctx);
}

template <typename T, typename F, int ND, bool MOD>
void launch_backward(const DevGeom& g, const void* x, const void* w, const void* off, const void* mask,
                     const void* go, void* gx, void* gw, void* gb, void* goff, void* gmask, void* ws,
                     LaunchCtx& ctx, int parts) {
  const bool do_data = parts & GENERIC_BWD_DATA, do_weight = parts & GENERIC_BWD_WEIGHT,
             do_bias = parts & GENERIC_BWD_BIAS;
  const BwdSizes sz = bwd_sizes(g);
  constexpr bool direct = sizeof(T) == sizeof(F);  // f32 / f64 accumulate straight into the outputs
  F *a_gx, *a_goff, *a_gmask, *a_gw;
  if (direct) {
    a_gx = (F*)gx; a_goff = (F*)goff; a_gmask = (F*)gmask; a_gw = (F*)gw;
  } else {
    char* base = (char*)ws;
    a_gx = (F*)base;          base += align256(sz.n_gx * sizeof(F));
    a_goff = (F*)base;        base += align256(sz.n_goff * sizeof(F));
    a_gmask = (F*)base;       base += align256(sz.n_gmask * sizeof(F));
    a_gw = (F*)base;
  }
  auto zero = [&](void* p, size_t bytes) {
    if (bytes == 0) return;
    cudaError_t e = cudaMemsetAsync(p, 0, bytes, ctx.stream);
    if (e != cudaSuccess && ctx.err == cudaSuccess) ctx.err = e;
  };
  if (do_data) {
    zero(a_gx, sz.n_gx * sizeof(F));
    zero(a_goff, sz.n_goff * sizeof(F));
    if (MOD) zero(a_gmask, sz.n_gmask * sizeof(F));
  }
  if (do_weight) zero(a_gw, sz.n_gw * sizeof(F));

  const int tilesP = (g.P + BM - 1) / BM;
  if (do_data) {
    dim3 grid((unsigned)(g.B * tilesP), (unsigned)((g.Kg + BN - 1) / BN), (unsigned)g.g);
    ProfScope prof(ctx, "generic_bwd_data");
    k_generic_bwd_data<T, F, ND, MOD><<<grid, NT, 0, ctx.stream>>>(g, (const T*)x, (const T*)w, (const T*)off,
                                                                   (const T*)mask, (const T*)go, a_gx, a_goff,
                                                                   a_gmask);
    MDCONV_CHECK_LAUNCH(ctx);
  }
  if (do_weight) {
    const int tilesK = (g.Kg + BN - 1) / BN;
    const int tilesN = (g.Cog + BN - 1) / BN;
    const int units = g.B * ((g.P + BP - 1) / BP);
    int splits = (148 * 4 + tilesK * tilesN * g.g - 1) / (tilesK * tilesN * g.g);
    if (splits > units) splits = units;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    dim3 grid((unsigned)tilesK, (unsigned)(tilesN * g.g), (unsigned)splits);
    ProfScope prof(ctx, "generic_bwd_weight");
    k_generic_bwd_weight<T, F, ND, MOD><<<grid, NT, 0, ctx.stream>>>(g, (const T*)x, (const T*)off,
                                                                     (const T*)mask, (const T*)go, a_gw);
    MDCONV_CHECK_LAUNCH(ctx);
  }
  if (do_bias && g.with_bias && gb != nullptr) {
    ProfScope prof(ctx, "grad_bias");
    k_grad_bias<T, F><<<g.Co, NT, 0, ctx.stream>>>(g, (const T*)go, (T*)gb, ctx.gb32, ctx.go_ld > 0 ? ctx.go_ld : g.P);
    MDCONV_CHECK_LAUNCH(ctx);
  }
  if (!direct) {
    if constexpr (!direct) {
      if (do_data) {
        convert_from_f32<T>((const float*)a_gx, gx, sz.n_gx, ctx);
        convert_from_f32<T>((const float*)a_goff, goff, sz.n_goff, ctx);
        if (MOD) convert_from_f32<T>((const float*)a_gmask, gmask, sz.n_gmask, ctx);
      }
      if (do_weight) convert_from_f32<T>((const float*)a_gw, gw, sz.n_gw, ctx);
    }
  }
  if (do_weight && ctx.gw32 != nullptr && sizeof(F) == 4) {   // fp32 accumulator (fp16 / bf16 / fp32 calls): plain copy
    cudaError_t e = cudaMemcpyAsync(ctx.gw32, a_gw, sz.n_gw * sizeof(float), cudaMemcpyDeviceToDevice, ctx.stream);
    if (e != cudaSuccess && ctx.err == cudaSuccess) ctx.err = e;
  }
}

template <typename T, typename F>
void dispatch_forward(const DevGeom& g, const void* x, const void* w, const void* bias, const void* off,
                      const void* mask, void* y, LaunchCtx& ctx) {
  if (g.nd == 2) {
    if (g.modulated) launch_forward<T, F, 2, true>(g, x, w, bias, off, mask, y, ctx);
    else launch_forward<T, F, 2, false>(g, x, w, bias, off, mask, y, ctx);
  } else {
    if (g.modulated) launch_forward<T, F, 3, true>(g, x, w, bias, off, mask, y, ctx);
    else launch_forward<T, F, 3, false>(g, x, w, bias, off, mask, y, ctx);
  }
}

template <typename T, typename F>
void dispatch_backward(const DevGeom& g, const void* x, const void* w, const void* off, const void* mask,
                       const void* go, void* gx, void* gw, void* gb, void* goff, void* gmask, void* ws,
                       LaunchCtx& ctx, int parts) {
  if (g.nd == 2) {
    if (g.modulated) launch_backward<T, F, 2, true>(g, x, w, off, mask, go, gx, gw, gb, goff, gmask, ws, ctx, parts);
    else launch_backward<T, F, 2, false>(g, x, w, off, mask, go, gx, gw, gb, goff, gmask, ws, ctx, parts);
  } else {
    if (g.modulated) launch_backward<T, F, 3, true>(g, x, w, off, mask, go, gx, gw, gb, goff, gmask, ws, ctx, parts);
    else launch_backward<T, F, 3, false>(g, x, w, off, mask, go, gx, gw, gb, goff, gmask, ws, ctx, parts);
  }
}

}  // namespace

size_t generic_workspace_bytes(const DevGeom& g, int dtype, int backward) {
  if (!backward) return 0;
  if (dtype == MDCONV_F32 || dtype == MDCONV_F64) return 0;
  const BwdSizes sz = bwd_sizes(g);
  return align256(sz.n_gx * 4) + align256(sz.n_goff * 4) + align256(sz.n_gmask * 4) + align256(sz.n_gw * 4);
}

void generic_forward(const DevGeom& g, int dtype, const void* x, const void* w, const void* bias, const void* off,
                     const void* mask, void* y, LaunchCtx& ctx) {
  switch (dtype) {
    case MDCONV_F32: dispatch_forward<float, float>(g, x, w, bias, off, mask, y, ctx); break;
    case MDCONV_F16: dispatch_forward<__half, float>(g, x, w, bias, off, mask, y, ctx); break;
    case MDCONV_BF16: dispatch_forward<__nv_bfloat16, float>(g, x, w, bias, off, mask, y, ctx); break;
    case MDCONV_F64: dispatch_forward<double, double>(g, x, w, bias, off, mask, y, ctx); break;
    default: break;
  }
}

void generic_backward(const DevGeom& g, int dtype, const void* x, const void* w, const void* off, const void* mask,
                      const void* go, void* gx, void* gw, void* gb, void* goff, void* gmask, void* ws,
                      LaunchCtx& ctx, int parts) {
  switch (dtype) {
    case MDCONV_F32: dispatch_backward<float, float>(g, x, w, off, mask, go, gx, gw, gb, goff, gmask, ws, ctx, parts); break;
    case MDCONV_F16: dispatch_backward<__half, float>(g, x, w, off, mask, go, gx, gw, gb, goff, gmask, ws, ctx, parts); break;
    case MDCONV_BF16:
      dispatch_backward<__nv_bfloat16, float>(g, x, w, off, mask, go, gx, gw, gb, goff, gmask, ws, ctx, parts);
      break;
    case MDCONV_F64: dispatch_backward<double, double>(g, x, w, off, mask, go, gx, gw, gb, goff, gmask, ws, ctx, parts); break;
    default: break;
  }
}

}  // namespace mdconv
