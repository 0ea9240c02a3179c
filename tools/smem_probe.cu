// smem_probe.cu — micro-measurements behind the EM kernel's design choices (B200, sm_100a).  Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_probe tools/smem_probe.cu && ./smem_probe
// Prints, per access pattern, shared-memory load instructions per clock and SM and useful GB/s:
//   lds128 all lanes / lanes < 8 / lanes < 16 / every 4th lane   — does a partly predicated LDS.128 cost fewer wavefronts?
//   lds64 all lanes, lds64 stride-2 (2-way conflict)  — candidates for the M-step's score reads
// (the clock64 column runs at half the SM clock on this part: read the GB/s column)
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kThreads = 1024, kSlots = 4096;



template <int MODE>
__global__ void __launch_bounds__(kThreads, 2) probe(uint4* sink, int iters, long long* cycles) {
    extern __shared__ __align__(16) uint4 buf[];
    for (int i = threadIdx.x; i < kSlots; i += kThreads) buf[i] = make_uint4(i, i * 3u, i * 5u, i * 7u);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint4 acc = make_uint4(0, 0, 0, 0);
    unsigned at = threadIdx.x;
    bool on = true;
    if (MODE == 1) on = lane < 8;
    if (MODE == 2) on = lane < 16;
    if (MODE == 3) on = (lane & 3) == 0;
    if (MODE == 7) on = lane < 5;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (MODE <= 3 || MODE == 7) {
                if (on) {
                    const uint4 v = buf[(at + u * 96u) & (kSlots - 1)];
                    acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
                }
            } else if (MODE == 4) {            // LDS.64, consecutive lanes consecutive 8-byte words
                const uint2 v = reinterpret_cast<const uint2*>(buf)[(at + u * 96u) & (2 * kSlots - 1)];
                acc.x ^= v.x; acc.y ^= v.y;
            } else if (MODE == 5) {            // LDS.64, lanes 16 bytes apart (2-way bank conflict)
                const uint2 v = reinterpret_cast<const uint2*>(buf)[(2 * at + u * 96u) & (2 * kSlots - 1)];
                acc.x ^= v.x; acc.y ^= v.y;
            }
        }
        at += 1536u + (acc.x & 32u);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    if (acc.x == 0x12345u && acc.y == 0x54321u) sink[blockIdx.x * kThreads + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, int sms, double bytes_per_lane_load, double active_frac) {
    const size_t smem = sizeof(uint4) * kSlots;
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = 2 * sms, iters = 2048;
    uint4* sink;
    long long* cyc;
    cudaMalloc(&sink, sizeof(uint4) * (size_t)grid * kThreads);
    cudaMalloc(&cyc, 8);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(a);
        probe<MODE><<<grid, kThreads, smem>>>(sink, iters, cyc);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (r && ms < best) best = ms;
    }
    long long c = 0;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double warp_instr = (double)grid * (kThreads / 32) * iters * 16.0;        // load (or shuffle-pair) instructions, all SMs
    const double per_sm_clk = warp_instr / sms / (double)c;                          // warp instructions per clock and SM (CTA 0's clock)
    const double gbs = warp_instr * 32 * active_frac * bytes_per_lane_load / (best * 1e-3) / 1e9;
    printf("%-34s %8.3f ms  %6.3f warp-instr/clk/SM  (%5.2f clk each)  useful %9.1f GB/s  err=%s\n", name, best, per_sm_clk, 1.0 / per_sm_clk, gbs,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink);
    cudaFree(cyc);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d\n", sms);
    run<0>("LDS.128 all 32 lanes", sms, 16, 1.0);
    run<1>("LDS.128 lanes 0-7 only", sms, 16, 0.25);
    run<7>("LDS.128 lanes 0-4 only", sms, 16, 5 / 32.0);
    run<2>("LDS.128 lanes 0-15 only", sms, 16, 0.5);
    run<3>("LDS.128 every 4th lane", sms, 16, 0.25);
    run<4>("LDS.64 all lanes, contiguous", sms, 8, 1.0);
    run<5>("LDS.64 all lanes, 16 B apart", sms, 8, 1.0);
    return 0;
}
