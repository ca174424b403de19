// Micro-benchmark: MUFU.EX2 / FFMA / F2FP issue throughput per SM on sm_100a (cycles per warp-instruction per SMSP).
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
    float a[16];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i * 0.01f;
    unsigned acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
            if (MODE == 2) { unsigned r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %1;" : "=r"(r) : "f"(a[i])); acc ^= r; }
            if (MODE == 3) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[(i + 8) & 15])); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[(i + 4) & 15])); }
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = (float)(t1 - t0); }
    if (s == 123.456f) out[1] = s + acc;
}
int main() {
    float* d; cudaMalloc(&d, 64);
    const char* names[4] = {"MUFU.EX2", "FFMA", "F2FP pack", "EX2 + 2 FFMA"};
    for (int warps = 4; warps <= 32; warps *= 2)
        for (int mode = 0; mode < 4; ++mode) {
            int iters = 2000; float h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, warps * 32>>>(d, iters);
                if (mode == 1) k<1><<<148, warps * 32>>>(d, iters);
                if (mode == 2) k<2><<<148, warps * 32>>>(d, iters);
                if (mode == 3) k<3><<<148, warps * 32>>>(d, iters);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
            double per = h / (double)(iters * 16) / (warps / 4.0);   // cycles per warp-instruction-group per SMSP
            printf("%-14s warps/SM=%2d : %.2f cycles per warp-instr(group) per SMSP\n", names[mode], warps, per);
        }
    return 0;
}
