// Micro-benchmark: how fast can one SM pull L2-resident data into shared memory with bulk async copies (the TMA
// engine), and does it matter whether neighbouring SMs read the SAME lines (L2-side limit) or different ones
// (SM-ingest limit)?  Decides whether cluster multicast can lift the GEMM main loop (which ingests 64-75 B/clk/SM).
//   nvcc -arch=sm_100a -O3 -o tma_ingest tma_ingest.cu && ./tma_ingest
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CHUNK = 16384, STAGES = 8;

__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// group: CTAs [g*group, (g+1)*group) read the same region (group = 1: all distinct; 148: all the same)
__global__ void __launch_bounds__(128, 1) ingest(const uint8_t* __restrict__ src, size_t region_bytes, int group, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * CHUNK);
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint8_t* base = src + static_cast<size_t>(blockIdx.x / group) * region_bytes;
        const int nchunk = static_cast<int>(region_bytes / CHUNK);
        auto issue = [&](int it) {
            const int s = it % STAGES;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bars[s])), "r"(CHUNK) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(smem + s * CHUNK)),
                         "l"(base + static_cast<size_t>(it % nchunk) * CHUNK), "r"(CHUNK), "r"(s32(&bars[s]))
                         : "memory");
        };
        const long long t0 = clock64();
        for (int it = 0; it < STAGES && it < iters; ++it) issue(it);
        for (int it = 0; it < iters; ++it) {
            const int s = it % STAGES;
            const uint32_t parity = (it / STAGES) & 1;
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bars[s])), "r"(parity) : "memory");
            if (it + STAGES < iters) issue(it + STAGES);
        }
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
}

int main() {
    const size_t region = 384 * 1024;                 // per-group footprint; 148 x 384 KB = 57 MB stays in L2
    uint8_t* d;
    cudaMalloc(&d, 148 * region);
    cudaMemset(d, 1, 148 * region);
    long long* o;
    cudaMalloc(&o, 148 * sizeof(long long));
    const int smem = STAGES * CHUNK + 256;
    cudaFuncSetAttribute(ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2000;                           // 32 MB per CTA
    for (int nblk : {148, 74, 16, 1})
        for (int group : {1, 2, 4, 9, 148}) {
            if (group > nblk && group != 148) continue;
            float ms = 0;
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(a);
                ingest<<<nblk, 128, smem>>>(d, region, group, iters, o);
                cudaEventRecord(b);
                cudaDeviceSynchronize();
                cudaEventElapsedTime(&ms, a, b);
            }
            long long h[148];
            cudaMemcpy(h, o, nblk * sizeof(long long), cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int i = 0; i < nblk; ++i) avg += h[i];
            avg /= nblk;
            const double bytes = static_cast<double>(iters) * CHUNK;
            printf("CTAs %3d  sharing group %3d : %6.1f B/clk/SM (clock64)   %7.2f TB/s aggregate (events, %.3f ms)\n", nblk, group, bytes / avg,
                   bytes * nblk / (ms * 1e-3) / 1e12, ms);
        }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    return 0;
}
