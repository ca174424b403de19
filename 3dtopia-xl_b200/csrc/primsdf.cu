// PrimSDF point query (the consumer of the VAE output layout):
//   reference: models/primsdf.py:52-109 (forward, grid_sample_feat, prim_weight); called on 256^3 = 16.8 M points in
//   8192-point chunks by extract_texmesh (inference.py:108-116) through a dense [points x prims] weight matrix.
// Here: one thread per point, the primitives' (1/scale, pos) streamed through shared memory; a point visits every
// primitive once (box test in registers), samples only the few that cover it (trilinear, align_corners=True, from the
// L2-resident feature volumes) and never materialises the weight matrix.  FP32 SIMT: this is gather/compare work,
// not tensor-core work.
#include "kernels.cuh"

namespace tpx {

namespace {
constexpr int PS_THREADS = 256;
constexpr int PS_CHUNK = 1024;   // primitives staged per pass (16 KB of float4)

template <int DF>
__global__ void __launch_bounds__(PS_THREADS) primsdf_query_kernel(const float* __restrict__ x, const float* __restrict__ srt,
                                                                   const float* __restrict__ feat, long long n, int K, int S, int inference,
                                                                   float* __restrict__ out) {
    __shared__ float4 s_prim[PS_CHUNK];   // (1/scale, tx, ty, tz)
    const long long i = static_cast<long long>(blockIdx.x) * PS_THREADS + threadIdx.x;
    const bool live = i < n;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (live) { px = x[3 * i]; py = x[3 * i + 1]; pz = x[3 * i + 2]; }
    float acc[DF];
#pragma unroll
    for (int c = 0; c < DF; ++c) acc[c] = 0.f;
    float wsum = 0.f, best_d2 = INFINITY;
    int best_k = 0;
    const int S3 = S * S * S;
    const float half_span = 0.5f * static_cast<float>(S - 1);
    for (int k0 = 0; k0 < K; k0 += PS_CHUNK) {
        const int kc = min(PS_CHUNK, K - k0);
        __syncthreads();
        for (int t = threadIdx.x; t < kc; t += PS_THREADS) {
            const float4 v = *reinterpret_cast<const float4*>(srt + 4 * static_cast<size_t>(k0 + t));
            s_prim[t] = make_float4(1.0f / v.x, v.y, v.z, v.w);
        }
        __syncthreads();
        if (!live) continue;
        for (int t = 0; t < kc; ++t) {
            const float4 p = s_prim[t];
            const float dx = px - p.y, dy = py - p.z, dz = pz - p.w;
            const float d2 = dx * dx + dy * dy + dz * dz;
            if (d2 < best_d2) { best_d2 = d2; best_k = k0 + t; }
            float lx = dx * p.x, ly = dy * p.x, lz = dz * p.x;
            float w = 1.0f - fmaxf(fabsf(lx), fmaxf(fabsf(ly), fabsf(lz)));
            if (fabsf(w) < 1e-5f) {   // on the box boundary: decide with the reference's exact division, not the reciprocal
                const float sc = srt[4 * static_cast<size_t>(k0 + t)];
                lx = dx / sc; ly = dy / sc; lz = dz / sc;
                w = 1.0f - fmaxf(fabsf(lx), fmaxf(fabsf(ly), fabsf(lz)));
            }
            if (w > 0.f) {
                wsum += w;
                // trilinear sample of feat[k] viewed as [DF, D(z), H(y), W(x)], align_corners=True; |l| < 1 so no padding case
                const float fx = (lx + 1.0f) * half_span, fy = (ly + 1.0f) * half_span, fz = (lz + 1.0f) * half_span;
                const int x0 = min(static_cast<int>(fx), S - 2), y0 = min(static_cast<int>(fy), S - 2), z0 = min(static_cast<int>(fz), S - 2);
                const float ax = fx - x0, ay = fy - y0, az = fz - z0;
                const float* f = feat + static_cast<size_t>(k0 + t) * DF * S3 + (z0 * S + y0) * S + x0;
#pragma unroll
                for (int c = 0; c < DF; ++c) {
                    const float* fc = f + c * S3;
                    const float c00 = fc[0] * (1.f - ax) + fc[1] * ax;
                    const float c01 = fc[S] * (1.f - ax) + fc[S + 1] * ax;
                    const float c10 = fc[S * S] * (1.f - ax) + fc[S * S + 1] * ax;
                    const float c11 = fc[S * S + S] * (1.f - ax) + fc[S * S + S + 1] * ax;
                    const float v = (c00 * (1.f - ay) + c01 * ay) * (1.f - az) + (c10 * (1.f - ay) + c11 * ay) * az;
                    acc[c] = fmaf(w, v, acc[c]);
                }
            }
        }
    }
    if (!live) return;
    float* o = out + static_cast<size_t>(i) * DF;
    if (wsum > 0.f) {
        const float inv = 1.0f / (wsum + 1e-6f);
        o[0] = acc[0] * inv;
#pragma unroll
        for (int c = 1; c < DF; ++c) o[c] = fminf(fmaxf(acc[c] * inv, 0.f), 1.f);
    } else {
#pragma unroll
        for (int c = 0; c < DF; ++c) o[c] = 0.f;
        if (inference && K > 0) {
            // SDF of an uncovered point: nearest voxel of the nearest primitive, same sign, plus the L2 distance to it
            const float4 v = *reinterpret_cast<const float4*>(srt + 4 * static_cast<size_t>(best_k));
            const float sc = v.x;
            const float lx = (px - v.y) / sc, ly = (py - v.z) / sc, lz = (pz - v.w) / sc;
            // The nearest voxel is one of the 2x2x2 lattice neighbours of the (clamped) local position.  They are ranked the
            // way the reference ranks all S^3 candidates (primsdf.py:93-96): fp32 L2 distance to pos + scale * linspace
            // grid, first flat index wins a tie — a far point often has two candidates whose fp32 distances coincide.
            const int x0 = min(max(static_cast<int>(floorf((lx + 1.f) * half_span)), 0), S - 2);
            const int y0 = min(max(static_cast<int>(floorf((ly + 1.f) * half_span)), 0), S - 2);
            const int z0 = min(max(static_cast<int>(floorf((lz + 1.f) * half_span)), 0), S - 2);
            const float step = 2.0f / static_cast<float>(S - 1);
            auto lin = [&](int i) { return i < S / 2 ? -1.f + step * static_cast<float>(i) : 1.f - step * static_cast<float>(S - 1 - i); };  // torch.linspace
            float dist = INFINITY;
            int xi = x0, yi = y0, zi = z0;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int cz = z0 + (c >> 2), cy = y0 + ((c >> 1) & 1), cx = x0 + (c & 1);
                const float ex = px - __fadd_rn(v.y, __fmul_rn(sc, lin(cx)));
                const float ey = py - __fadd_rn(v.z, __fmul_rn(sc, lin(cy)));
                const float ez = pz - __fadd_rn(v.w, __fmul_rn(sc, lin(cz)));
                const float d = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez)));
                if (d < dist) { dist = d; xi = cx; yi = cy; zi = cz; }
            }
            const float sdf = feat[static_cast<size_t>(best_k) * DF * S3 + (zi * S + yi) * S + xi];
            const float sgn = sdf > 0.f ? 1.f : (sdf < 0.f ? -1.f : 0.f);
            o[0] = sdf + dist * sgn;
        }
    }
}
}  // namespace

int launch_primsdf_query(const float* x, const float* srt, const float* feat, long long n, int K, int S, int dim_feat, int inference, float* out,
                         cudaStream_t st) {
    TPX_CHECK(dim_feat == 6, TPX_ERR_SHAPE, "primsdf_query: dim_feat %d (kernel covers the released 6-channel layout)", dim_feat);
    TPX_CHECK(S >= 2 && S <= 32 && K >= 0, TPX_ERR_SHAPE, "primsdf_query: bad primitive geometry (S %d, K %d)", S, K);
    TPX_CHECK((reinterpret_cast<uintptr_t>(srt) & 15) == 0, TPX_ERR_ARG, "primsdf_query: srt must be 16-B aligned");
    if (n <= 0) return TPX_OK;
    ProfScope prof(PROF_VAE_MISC, st);
    const long long blocks = (n + PS_THREADS - 1) / PS_THREADS;
    TPX_CHECK(blocks < (1LL << 31), TPX_ERR_SHAPE, "primsdf_query: too many points");
    primsdf_query_kernel<6><<<static_cast<unsigned>(blocks), PS_THREADS, 0, st>>>(x, srt, feat, n, K, S, inference, out);
    TPX_LAUNCH_CHECK();
    return TPX_OK;
}

}  // namespace tpx
