// fb_microbench.cu — the measured ceiling of the roof that binds the EM kernel: shared-memory read bandwidth.
//
// The E-step reads one 16-byte {P,Q} pair per lane and base term (LDS.128, consecutive lanes at consecutive rows: no bank
// conflicts, four 128-byte wavefronts per warp instruction).  This kernel issues exactly that access pattern and nothing
// else of weight — 128-bit loads from conflict-free addresses, folded with integer XORs so that neither the FP64 pipe nor
// memory outside the SM is involved — on every SM at full occupancy.  bench.py reports the EM kernel's useful
// shared-memory bytes per second against this number instead of against the data-sheet product
// (SMs x 128 B/clk x max clock), which assumes a clock the chip does not hold under load.
#include "fb_device.h"

namespace {

constexpr int kMbThreads = 1024;
constexpr int kMbSlots = 4096;          // 16-byte slots per CTA (64 KB)

__global__ void __launch_bounds__(kMbThreads, 2) fb_smem_read_kernel(uint4* __restrict__ sink, int iters) {
    extern __shared__ __align__(16) uint4 mb_buf[];
    for (int i = threadIdx.x; i < kMbSlots; i += kMbThreads) mb_buf[i] = make_uint4(i, i * 3u, i * 5u, i * 7u);
    __syncthreads();
    uint4 acc = make_uint4(0, 0, 0, 0);
    unsigned at = threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const uint4 v = mb_buf[(at + u * 96u) & (kMbSlots - 1)];     // lanes stay consecutive: conflict-free
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
        at += 1536u + (acc.x & 32u);      // data-dependent, warp-aligned step: the loads cannot be hoisted or merged
    }
    if (acc.x == 0x12345u && acc.y == 0x54321u) sink[blockIdx.x * kMbThreads + threadIdx.x] = acc;   // never true in practice
}

}  // namespace

// Shared-memory read bandwidth of the whole device in GB/s (best of `reps` timed launches after one warm-up).
cudaError_t fb_microbench_smem(double* gbs, int reps) {
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    const size_t smem = sizeof(uint4) * kMbSlots;
    e = cudaFuncSetAttribute(fb_smem_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = 2 * sms, iters = 4096;
    uint4* sink = nullptr;
    e = cudaMalloc((void**)&sink, sizeof(uint4) * (size_t)grid * kMbThreads);
    if (e != cudaSuccess) return e;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    double best = 0;
    for (int r = 0; r <= reps; r++) {
        cudaEventRecord(a);
        fb_smem_read_kernel<<<grid, kMbThreads, smem>>>(sink, iters);
        cudaEventRecord(b);
        e = cudaEventSynchronize(b);
        if (e != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        const double bytes = (double)grid * kMbThreads * iters * 16.0 * 16.0;
        if (r > 0 && ms > 0) best = bytes / (ms * 1e-3) / 1e9 > best ? bytes / (ms * 1e-3) / 1e9 : best;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(sink);
    if (e == cudaSuccess) e = cudaGetLastError();
    *gbs = best;
    return e;
}
