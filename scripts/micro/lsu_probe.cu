// LSU / L1 wavefront-cost probe for the corner gather (sm_100a).  Not part of the product.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o lsu_probe lsu_probe.cu
// Every mode issues warp-wide loads of random 128-byte lines out of a window (L1-resident: 32 KB, L2: 16 MB)
// with 16 warps per SM, 8 independent loads in flight per thread, and reports cycles per warp instruction
// and per 128-byte line (SM-wide, i.e. the LSU throughput the gather warps see).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k_probe(const uint4* __restrict__ buf, uint32_t lines_mask, int iters,
                                                   unsigned long long* cyc, uint32_t* sink) {
    extern __shared__ uint4 sm[];
    const int lane = threadIdx.x & 31;
    if (MODE == 5 || MODE == 7) {
        for (int i = threadIdx.x; i < 256 * 8; i += blockDim.x) sm[i] = buf[i];
    }
    __syncthreads();
    uint32_t seed = (blockIdx.x * 512 + (threadIdx.x & ~31)) * 2654435761u + 12345u;   // one stream per warp
    uint32_t acc = 0;
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint32_t r = lcg(seed);
            if (MODE == 0) {          // LDG.128, 8 lanes per line, 4 random lines
                uint32_t line = __shfl_sync(0xffffffffu, lcg(seed) ^ (lane * 7919u), lane & 24) & lines_mask;
                uint4 v = buf[line * 8 + (lane & 7)];
                acc ^= v.x ^ v.y ^ v.z ^ v.w;
            } else if (MODE == 1) {   // LDG.64, 16 lanes per line, 2 random lines
                uint32_t line = __shfl_sync(0xffffffffu, lcg(seed) ^ (lane * 7919u), lane & 16) & lines_mask;
                uint2 v = reinterpret_cast<const uint2*>(buf)[line * 16 + (lane & 15)];
                acc ^= v.x ^ v.y;
            } else if (MODE == 2) {   // LDG.32, one line per instruction
                uint32_t line = r & lines_mask;
                uint32_t v = reinterpret_cast<const uint32_t*>(buf)[line * 32 + lane];
                acc ^= v;
            } else if (MODE == 3) {   // LDG.128, 4 contiguous lines
                uint32_t line = (r & lines_mask) & ~3u;
                uint4 v = buf[line * 8 + lane];
                acc ^= v.x ^ v.y ^ v.z ^ v.w;
            } else if (MODE == 4) {   // LDG.128, two random PAIRS of adjacent lines (corner (w, w+1) pairs)
                uint32_t line = (__shfl_sync(0xffffffffu, lcg(seed) ^ (lane * 7919u), lane & 16) & lines_mask) & ~1u;
                uint4 v = buf[line * 8 + (lane & 15)];
                acc ^= v.x ^ v.y ^ v.z ^ v.w;
            } else if (MODE == 5) {   // LDS.128, 8 lanes per row, 4 random rows of a 32 KB window
                uint32_t line = __shfl_sync(0xffffffffu, lcg(seed) ^ (lane * 7919u), lane & 24) & 255u;
                uint4 v = sm[line * 8 + (lane & 7)];
                acc ^= v.x ^ v.y ^ v.z ^ v.w;
            } else if (MODE == 6) {   // LDG.128, 8 lanes per line, all four groups read the SAME line
                uint32_t line = r & lines_mask;
                uint4 v = buf[line * 8 + (lane & 7)];
                acc ^= v.x ^ v.y ^ v.z ^ v.w;
            } else if (MODE == 7) {   // LDS.64, 16 lanes per row, 2 random rows
                uint32_t line = __shfl_sync(0xffffffffu, lcg(seed) ^ (lane * 7919u), lane & 16) & 255u;
                uint2 v = reinterpret_cast<const uint2*>(sm)[line * 16 + (lane & 15)];
                acc ^= v.x ^ v.y;
            } else if (MODE == 8) {   // LDG.128, 8 lanes per line, 2 distinct random lines (each read by 2 groups)
                uint32_t line = __shfl_sync(0xffffffffu, lcg(seed) ^ (lane * 7919u), lane & 16) & lines_mask;
                uint4 v = buf[line * 8 + (lane & 7)];
                acc ^= v.x ^ v.y ^ v.z ^ v.w;
            }
        }
    }
    unsigned long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 0x12345u) sink[0] = acc;
}

static const char* names[] = {"LDG.128 4 random lines", "LDG.64  2 random lines", "LDG.32  1 line", "LDG.128 4 contiguous lines",
                              "LDG.128 2 random line pairs", "LDS.128 4 random rows", "LDG.128 same line x4 groups",
                              "LDS.64  2 random rows", "LDG.128 2 random lines (x2 groups)"};
static const int lines_per[] = {4, 2, 1, 4, 4, 4, 1, 2, 2};

template <int MODE>
void run(const uint4* buf, uint32_t mask, const char* tier, unsigned long long* cyc, uint32_t* sink, int sms) {
    const int iters = 2000;
    cudaFuncSetAttribute(k_probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int rep = 0; rep < 2; ++rep) k_probe<MODE><<<sms, 512, 32768>>>(buf, mask, iters, cyc, sink);
    cudaDeviceSynchronize();
    unsigned long long h[256]; cudaMemcpy(h, cyc, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < sms; ++i) s += (double)h[i];
    s /= sms;
    double instr = 16.0 * iters * 8;   // warp instructions per SM
    printf("%-38s %-4s  %7.2f cyc/warp-instr  %6.2f cyc/line  %6.1f B/cyc/SM\n", names[MODE], tier, s / instr,
           s / instr / lines_per[MODE], (MODE == 2 ? 128.0 : MODE == 1 || MODE == 7 ? 256.0 : 512.0) * instr / s);
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t big = 16u << 20;
    uint4* buf; cudaMalloc(&buf, big); cudaMemset(buf, 1, big);
    unsigned long long* cyc; cudaMalloc(&cyc, 256 * 8);
    uint32_t* sink; cudaMalloc(&sink, 4);
    struct { const char* n; uint32_t mask; } tiers[] = {{"L1", 255u}, {"L2", (uint32_t)(big / 128 - 1)}};
    for (auto& t : tiers) {
        run<0>(buf, t.mask, t.n, cyc, sink, sms); run<1>(buf, t.mask, t.n, cyc, sink, sms);
        run<2>(buf, t.mask, t.n, cyc, sink, sms); run<3>(buf, t.mask, t.n, cyc, sink, sms);
        run<4>(buf, t.mask, t.n, cyc, sink, sms); run<6>(buf, t.mask, t.n, cyc, sink, sms);
        run<8>(buf, t.mask, t.n, cyc, sink, sms);
    }
    run<5>(buf, 255u, "smem", cyc, sink, sms); run<7>(buf, 255u, "smem", cyc, sink, sms);
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
