// VAE-decoder kernels that are not GEMM shaped: the 1->C input convolution, GroupNorm(+SiLU) on
// channels-last fp16 volumes, and the weight repacks into the implicit-GEMM [Cout, 27*Cin] layout.
//   reference: models/vae3d_dib.py:93-145 (ResnetBlock), :344 (conv_in), :366-367 (norm_out / conv_out), :429
#include "kernels.cuh"

namespace tpx {

// ---------------------------------------------------------------------------------------------------------
// post_quant_conv (1x1x1, 1->1) + conv_in (3x3x3, pad 1, 1->C) : z [P,1,4,4,4] -> channels-last fp16 [P,64,C]
// One CTA per primitive, one thread per output channel; the 6^3 zero-padded latent lives in smem.
// ---------------------------------------------------------------------------------------------------------
template <typename ZT>
__global__ void __launch_bounds__(256) vae_conv_in_kernel(const ZT* __restrict__ z, const __half* __restrict__ w_pq, const __half* __restrict__ b_pq,
                                                          const __half* __restrict__ W /*[C,27]*/, const __half* __restrict__ bias, int C,
                                                          __half* __restrict__ out) {
    __shared__ float s_z[6 * 6 * 6];
    const int p = blockIdx.x;
    for (int i = threadIdx.x; i < 216; i += blockDim.x) s_z[i] = 0.f;
    __syncthreads();
    if (threadIdx.x < 64) {
        const int v = threadIdx.x;
        const float zq = h2f_round(static_cast<float>(z[static_cast<size_t>(p) * 64 + v]) * __half2float(w_pq[0]) + __half2float(b_pq[0]));
        s_z[(((v >> 4) + 1) * 6 + ((v >> 2) & 3) + 1) * 6 + (v & 3) + 1] = zq;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float w[27];
#pragma unroll
        for (int t = 0; t < 27; ++t) w[t] = __half2float(W[c * 27 + t]);
        const float b = __half2float(bias[c]);
        for (int v = 0; v < 64; ++v) {
            const int zz = v >> 4, yy = (v >> 2) & 3, xx = v & 3;
            float acc = b;
#pragma unroll
            for (int t = 0; t < 27; ++t) acc = fmaf(w[t], s_z[((zz + t / 9) * 6 + yy + (t / 3) % 3) * 6 + xx + t % 3], acc);
            out[(static_cast<size_t>(p) * 64 + v) * C + c] = __float2half_rn(acc);
        }
    }
}

int launch_vae_conv_in(const void* z, int z_dtype, const __half* w_pq, const __half* b_pq, const __half* W, const __half* bias, int P, int C,
                       __half* out, cudaStream_t st) {
    if (P <= 0) return TPX_OK;
    ProfScope prof(PROF_VAE_MISC, st);
    if (z_dtype == TPX_DTYPE_F32) vae_conv_in_kernel<float><<<P, 256, 0, st>>>(static_cast<const float*>(z), w_pq, b_pq, W, bias, C, out);
    else if (z_dtype == TPX_DTYPE_F16) vae_conv_in_kernel<__half><<<P, 256, 0, st>>>(static_cast<const __half*>(z), w_pq, b_pq, W, bias, C, out);
    else { set_error("vae_conv_in: unsupported latent dtype %d", z_dtype); return TPX_ERR_ARG; }
    TPX_LAUNCH_CHECK();
    return TPX_OK;
}

// ---------------------------------------------------------------------------------------------------------
// GroupNorm(groups, eps, affine) [+ SiLU] on channels-last fp16 [P, S3, C]; statistics per (primitive, group)
// in fp32.  One CTA per primitive; each thread streams 16-byte channel octets (coalesced), per-channel
// partial sums meet in shared memory, then a second streaming pass normalises (the re-read hits L2).
// ---------------------------------------------------------------------------------------------------------
constexpr int GN_MAXC = 256;
// x * sigmoid(x) with MUFU.EX2 + MUFU.RCP (relative error ~2e-7, far below the fp16 rounding of the result)
__device__ __forceinline__ float silu_fast(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return x * r;
}
// RES > 0: the primitive's S3 * C / 8 octets are exactly RES per thread and stay in registers between the statistics and the
// normalisation (one read, RES loads in flight per thread: the 32 KB volumes of the 4^3 stage and of the 32-channel 8^3 stage);
// RES == 0: streaming two-pass version for larger volumes.
template <int RES>
__global__ void __launch_bounds__(256, RES > 0 ? 3 : 1) groupnorm_silu_kernel(const __half* __restrict__ x, const __half* __restrict__ gamma,
                                                             const __half* __restrict__ beta, int S3, int C, int groups, float eps, int apply_silu,
                                                             __half* __restrict__ out) {
    __shared__ float s_sum[GN_MAXC], s_sq[GN_MAXC], s_scale[GN_MAXC], s_shift[GN_MAXC];
    __shared__ float s_part[256 * 16];
    const int p = blockIdx.x;
    const int oct = C >> 3;                    // 16-byte octets per voxel
    const int total = S3 * oct;
    const uint4* xp = reinterpret_cast<const uint4*>(x + static_cast<size_t>(p) * S3 * C);
    // blockDim (256) is a multiple of oct (<= 32), so a thread always sees the same octet
    const int my_oct = threadIdx.x % oct;
    float ls[8], lq[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { ls[i] = 0.f; lq[i] = 0.f; }
    Pack8 keep[RES > 0 ? RES : 1];
    if constexpr (RES > 0) {
#pragma unroll
        for (int j = 0; j < RES; ++j) keep[j].u = xp[threadIdx.x + j * 256];
#pragma unroll
        for (int j = 0; j < RES; ++j) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float f = __half2float(keep[j].h[k]);
                ls[k] += f;
                lq[k] = fmaf(f, f, lq[k]);
            }
        }
    } else {
#pragma unroll 4
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            Pack8 v;
            v.u = xp[i];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float f = __half2float(v.h[k]);
                ls[k] += f;
                lq[k] = fmaf(f, f, lq[k]);
            }
        }
    }
    // deterministic cross-thread reduction (fixed order; run-to-run bit-identical): per-thread partials -> smem,
    // then channel c sums the partials of the 256/oct threads that own its octet.
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        s_part[threadIdx.x * 16 + k] = ls[k];
        s_part[threadIdx.x * 16 + 8 + k] = lq[k];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int o = c >> 3, k = c & 7;
        float a = 0.f, b = 0.f;
        for (int t = o; t < 256; t += oct) {
            a += s_part[t * 16 + k];
            b += s_part[t * 16 + 8 + k];
        }
        s_sum[c] = a;
        s_sq[c] = b;
    }
    __syncthreads();
    const int cpg = C / groups;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g0 = (c / cpg) * cpg;
        float gs = 0.f, gq = 0.f;
        for (int k = 0; k < cpg; ++k) { gs += s_sum[g0 + k]; gq += s_sq[g0 + k]; }
        const float n = static_cast<float>(S3 * cpg);
        const float mean = gs / n;
        const float var = fmaxf(gq / n - mean * mean, 0.f);
        const float rstd = rsqrtf(var + eps);
        const float sc = rstd * __half2float(gamma[c]);
        s_scale[c] = sc;
        s_shift[c] = __half2float(beta[c]) - mean * sc;
    }
    __syncthreads();
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = s_scale[my_oct * 8 + k]; sh[k] = s_shift[my_oct * 8 + k]; }
    uint4* op = reinterpret_cast<uint4*>(out + static_cast<size_t>(p) * S3 * C);
    if constexpr (RES > 0) {
#pragma unroll
        for (int j = 0; j < RES; ++j) {
            Pack8 o;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float f = fmaf(__half2float(keep[j].h[k]), sc[k], sh[k]);
                if (apply_silu) f = silu_fast(f);
                o.h[k] = __float2half_rn(f);
            }
            op[threadIdx.x + j * 256] = o.u;
        }
    } else {
#pragma unroll 4
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            Pack8 v, o;
            v.u = xp[i];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float f = fmaf(__half2float(v.h[k]), sc[k], sh[k]);
                if (apply_silu) f = silu_fast(f);
                o.h[k] = __float2half_rn(f);
            }
            op[i] = o.u;
        }
    }
}

int launch_groupnorm_silu(const __half* x, const __half* gamma, const __half* beta, int P, int S3, int C, int groups, float eps, int apply_silu,
                          __half* out, cudaStream_t st) {
    TPX_CHECK(C % 8 == 0 && C <= GN_MAXC && 256 % (C / 8) == 0 && groups > 0 && C % groups == 0, TPX_ERR_SHAPE,
              "groupnorm: channels %d / groups %d unsupported (C%%8==0, C<=%d, C/8 | 256)", C, groups, GN_MAXC);
    if (P <= 0) return TPX_OK;
    ProfScope prof(PROF_GROUPNORM, st);
    if (S3 * (C / 8) == 8 * 256) groupnorm_silu_kernel<8><<<P, 256, 0, st>>>(x, gamma, beta, S3, C, groups, eps, apply_silu, out);
    else groupnorm_silu_kernel<0><<<P, 256, 0, st>>>(x, gamma, beta, S3, C, groups, eps, apply_silu, out);
    TPX_LAUNCH_CHECK();
    return TPX_OK;
}

// ---------------------------------------------------------------------------------------------------------
// weight repacks (run once at load time)
// ---------------------------------------------------------------------------------------------------------
// Conv3d weight [Cout,Cin,3,3,3] -> [CoutPad, 27*Cin] with k = tap*Cin + ci, tap = (kz*3+ky)*3+kx.
// transposed != 0: source is a ConvTranspose3d(k3,s1,p1) weight [Cin,Cout,3,3,3]; the equivalent correlation
// kernel is w'[co,ci,kz,ky,kx] = w[ci,co,2-kz,2-ky,2-kx]  (vae3d_dib.py:367).
template <typename T>
__global__ void pack_conv3_kernel(const T* __restrict__ src, __half* __restrict__ dst, int Cout, int Cin, int transposed) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long n = static_cast<long long>(Cout) * 27 * Cin;
    if (i >= n) return;
    const int co = static_cast<int>(i / (27 * Cin));
    const int r = static_cast<int>(i - static_cast<long long>(co) * 27 * Cin);
    const int tap = r / Cin, ci = r - tap * Cin;
    size_t s;
    if (!transposed) s = (static_cast<size_t>(co) * Cin + ci) * 27 + tap;
    else s = (static_cast<size_t>(ci) * Cout + co) * 27 + (26 - tap);
    dst[i] = __float2half_rn(static_cast<float>(src[s]));
}
// ConvTranspose3d(k2,s2) weight [Cin,Cout,2,2,2] -> [8*Cout, Cin]: row (abc*Cout + co), abc = (a*2+b)*2+c.
template <typename T>
__global__ void pack_convt2_kernel(const T* __restrict__ src, __half* __restrict__ dst, int Cin, int Cout) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long n = 8LL * Cout * Cin;
    if (i >= n) return;
    const int row = static_cast<int>(i / Cin), ci = static_cast<int>(i - static_cast<long long>(row) * Cin);
    const int abc = row / Cout, co = row - abc * Cout;
    dst[i] = __float2half_rn(static_cast<float>(src[(static_cast<size_t>(ci) * Cout + co) * 8 + abc]));
}

int launch_pack_conv3(const void* src, int dtype, __half* dst, int Cout, int Cin, int transposed, cudaStream_t st) {
    const long long n = static_cast<long long>(Cout) * 27 * Cin;
    const unsigned grid = static_cast<unsigned>((n + 255) / 256);
    if (dtype == TPX_DTYPE_F32) pack_conv3_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(src), dst, Cout, Cin, transposed);
    else if (dtype == TPX_DTYPE_F16) pack_conv3_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half*>(src), dst, Cout, Cin, transposed);
    else { set_error("pack_conv3: unsupported dtype %d", dtype); return TPX_ERR_ARG; }
    TPX_LAUNCH_CHECK();
    return TPX_OK;
}
int launch_pack_convt2(const void* src, int dtype, __half* dst, int Cin, int Cout, cudaStream_t st) {
    const long long n = 8LL * Cout * Cin;
    const unsigned grid = static_cast<unsigned>((n + 255) / 256);
    if (dtype == TPX_DTYPE_F32) pack_convt2_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(src), dst, Cin, Cout);
    else if (dtype == TPX_DTYPE_F16) pack_convt2_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half*>(src), dst, Cin, Cout);
    else { set_error("pack_convt2: unsupported dtype %d", dtype); return TPX_ERR_ARG; }
    TPX_LAUNCH_CHECK();
    return TPX_OK;
}

}  // namespace tpx
