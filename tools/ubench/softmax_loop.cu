// Micro-benchmark of the attention softmax inner loop (one thread per 128-element score row held in registers), isolated
// from TMA / tcgen05: which resource bounds it — MUFU throughput shared by the warps of an SM sub-partition, or the single
// warp's own instruction stream?  Variants are run with 1 and 2 warps per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_loop softmax_loop.cu && ./softmax_loop
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -126.0f);
    const float fl = x + 12582912.0f;
    const float r = x - (fl - 12582912.0f);
    float p = fmaf(0.0551716648f, r, 0.2426111251f);
    p = fmaf(p, r, 0.6932609677f);
    p = fmaf(p, r, 0.9999280572f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(fl) << 23));
}
__device__ __forceinline__ void fma2(float& a, float& b, float x0, float x1, float sc, float c) {
    unsigned long long d, x, s, cc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(s) : "f"(sc));
    asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(s), "l"(cc));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(d));
}

// VAR bits: 1 = pack+store P, 2 = max3 on the fly, 4 = FFMA2, 8 = poly every 4th, 16 = row sum by FADD
template <int VAR>
__global__ void __launch_bounds__(256, 1) loop_kernel(const float* __restrict__ in, float* __restrict__ out, int iters, long long* __restrict__ cycles) {
    extern __shared__ uint4 smem[];
    float sv[128];
#pragma unroll
    for (int i = 0; i < 128; ++i) sv[i] = in[(blockIdx.x * blockDim.x + threadIdx.x) * 128 + i];
    const uint32_t pS = static_cast<uint32_t>(__cvta_generic_to_shared(smem)) + threadIdx.x * 256;
    const int sw = threadIdx.x & 7;
    const float scale = 0.17f;
    float msc = 1.0f, sink = 0.f, m0 = -1e30f, m1 = -1e30f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        msc += 0.001f;
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                float x0, x1, x2, x3;
                if (VAR & 4) {
                    fma2(x0, x1, sv[c * 32 + i], sv[c * 32 + i + 1], scale, -msc);
                    fma2(x2, x3, sv[c * 32 + i + 2], sv[c * 32 + i + 3], scale, -msc);
                } else {
                    x0 = fmaf(sv[c * 32 + i], scale, -msc); x1 = fmaf(sv[c * 32 + i + 1], scale, -msc);
                    x2 = fmaf(sv[c * 32 + i + 2], scale, -msc); x3 = fmaf(sv[c * 32 + i + 3], scale, -msc);
                }
                if (VAR & 2) { m0 = fmaxf(m0, fmaxf(x0, x2)); m1 = fmaxf(m1, fmaxf(x1, x3)); }
                const float p0 = (VAR & 8) ? ex2_poly(x0) : ex2(x0);
                const float p1 = ex2(x1), p2 = ex2(x2), p3 = ex2(x3);
                if (VAR & 16) { rs0 += p0 + p2; rs1 += p1 + p3; }
                if (VAR & 1) {
                    __half2 ha = __floats2half2_rn(p0, p1), hb = __floats2half2_rn(p2, p3);
                    pk[i >> 1] = *reinterpret_cast<uint32_t*>(&ha);
                    pk[(i >> 1) + 1] = *reinterpret_cast<uint32_t*>(&hb);
                } else {
                    sink += (p0 + p1) + (p2 + p3);      // 3 FADD / 4 elements when nothing else consumes P
                }
            }
            if (VAR & 1) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pS + ((((c * 4) + q) ^ sw) << 4)), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                                 "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
            }
        }
        sink += rs0 + rs1;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = sink + m0 + m1;
}

template <int VAR>
void run(const char* name, const float* in, float* out, long long* cyc) {
    for (int threads : {128, 256}) {
        const int iters = 200;
        cudaFuncSetAttribute(loop_kernel<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        loop_kernel<VAR><<<148, threads, 65536>>>(in, out, iters, cyc);
        loop_kernel<VAR><<<148, threads, 65536>>>(in, out, iters, cyc);
        cudaDeviceSynchronize();
        long long h = 0;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-44s warps/SMSP=%d : %7.0f cycles per 128-element row pass\n", name, threads / 128, double(h) / iters);
    }
}

int main() {
    float *in, *out; long long* cyc;
    cudaMalloc(&in, 148 * 256 * 128 * 4); cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 8);
    cudaMemset(in, 0, 148 * 256 * 128 * 4);
    run<0>("FFMA + EX2 (+sum sink)", in, out, cyc);
    run<1>("FFMA + EX2 + pack + STS", in, out, cyc);
    run<3>("FFMA + EX2 + pack + STS + max3", in, out, cyc);
    run<19>("FFMA + EX2 + pack + STS + max3 + rowsum", in, out, cyc);
    run<5>("FFMA2 + EX2 + pack + STS", in, out, cyc);
    run<7>("FFMA2 + EX2 + pack + STS + max3", in, out, cyc);
    run<11>("FFMA + 3/4 EX2 + 1/4 poly + pack + STS + max3", in, out, cyc);
    run<15>("FFMA2 + 3/4 EX2 + 1/4 poly + pack + STS + max3", in, out, cyc);
    if (cudaGetLastError() != cudaSuccess) { printf("CUDA error\n"); return 1; }
    return 0;
}
