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

// ---- pieces shared by the exhaustive and the grid-binned kernels (identical arithmetic => identical results) ----------------------
template <int DF>
struct PointAcc {
    float acc[DF];
    float wsum;
};

// visit primitive k for the point (px,py,pz): box weight, and — if it covers the point — its trilinear sample
template <int DF>
__device__ __forceinline__ void visit_prim(PointAcc<DF>& a, float px, float py, float pz, const float4 v /* scale, tx, ty, tz */, int k,
                                           const float* __restrict__ feat, int S, int S3, float half_span) {
    const float inv = 1.0f / v.x;
    const float dx = px - v.y, dy = py - v.z, dz = pz - v.w;
    float lx = dx * inv, ly = dy * inv, lz = dz * inv;
    float w = 1.0f - fmaxf(fabsf(lx), fmaxf(fabsf(ly), fabsf(lz)));
    if (fabsf(w) < 1e-5f) {   // on the box boundary: decide with the reference's exact division, not the reciprocal
        lx = dx / v.x; ly = dy / v.x; lz = dz / v.x;
        w = 1.0f - fmaxf(fabsf(lx), fmaxf(fabsf(ly), fabsf(lz)));
    }
    if (w > 0.f) {
        a.wsum += w;
        // trilinear sample of feat[k] viewed as [DF, D(z), H(y), W(x)], align_corners=True; |l| < 1 so no padding case
        const float fx = (lx + 1.0f) * half_span, fy = (ly + 1.0f) * half_span, fz = (lz + 1.0f) * half_span;
        const int x0 = min(static_cast<int>(fx), S - 2), y0 = min(static_cast<int>(fy), S - 2), z0 = min(static_cast<int>(fz), S - 2);
        const float ax = fx - x0, ay = fy - y0, az = fz - z0;
        const float* f = feat + static_cast<size_t>(k) * DF * S3 + (z0 * S + y0) * S + x0;
#pragma unroll
        for (int c = 0; c < DF; ++c) {
            const float* fc = f + c * S3;
            const float c00 = fc[0] * (1.f - ax) + fc[1] * ax;
            const float c01 = fc[S] * (1.f - ax) + fc[S + 1] * ax;
            const float c10 = fc[S * S] * (1.f - ax) + fc[S * S + 1] * ax;
            const float c11 = fc[S * S + S] * (1.f - ax) + fc[S * S + S + 1] * ax;
            const float val = (c00 * (1.f - ay) + c01 * ay) * (1.f - az) + (c10 * (1.f - ay) + c11 * ay) * az;
            a.acc[c] = fmaf(w, val, a.acc[c]);
        }
    }
}

template <int DF>
__device__ __forceinline__ void finish_point(const PointAcc<DF>& a, float px, float py, float pz, int best_k, const float* __restrict__ srt,
                                             const float* __restrict__ feat, int K, int S, int S3, float half_span, int inference, float* __restrict__ o) {
    if (a.wsum > 0.f) {
        const float inv = 1.0f / (a.wsum + 1e-6f);
        o[0] = a.acc[0] * inv;
#pragma unroll
        for (int c = 1; c < DF; ++c) o[c] = fminf(fmaxf(a.acc[c] * inv, 0.f), 1.f);
        return;
    }
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

// exhaustive version: every point visits every primitive (kept as the exactness reference of the binned kernel and as its fallback)
template <int DF>
__global__ void __launch_bounds__(PS_THREADS) primsdf_query_kernel(const float* __restrict__ x, const float* __restrict__ srt,
                                                                   const float* __restrict__ feat, long long n, int K, int S, int inference,
                                                                   float* __restrict__ out) {
    __shared__ float4 s_prim[PS_CHUNK];   // (scale, tx, ty, tz)
    const long long i = static_cast<long long>(blockIdx.x) * PS_THREADS + threadIdx.x;
    const bool live = i < n;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (live) { px = x[3 * i]; py = x[3 * i + 1]; pz = x[3 * i + 2]; }
    PointAcc<DF> a;
#pragma unroll
    for (int c = 0; c < DF; ++c) a.acc[c] = 0.f;
    a.wsum = 0.f;
    float best_d2 = INFINITY;
    int best_k = 0;
    const int S3 = S * S * S;
    const float half_span = 0.5f * static_cast<float>(S - 1);
    for (int k0 = 0; k0 < K; k0 += PS_CHUNK) {
        const int kc = min(PS_CHUNK, K - k0);
        __syncthreads();
        for (int t = threadIdx.x; t < kc; t += PS_THREADS) s_prim[t] = *reinterpret_cast<const float4*>(srt + 4 * static_cast<size_t>(k0 + t));
        __syncthreads();
        if (!live) continue;
        for (int t = 0; t < kc; ++t) {
            const float4 p = s_prim[t];
            const float dx = px - p.y, dy = py - p.z, dz = pz - p.w;
            const float d2 = dx * dx + dy * dy + dz * dz;
            if (d2 < best_d2) { best_d2 = d2; best_k = k0 + t; }
            visit_prim<DF>(a, px, py, pz, p, k0 + t, feat, S, S3, half_span);
        }
    }
    if (!live) return;
    finish_point<DF>(a, px, py, pz, best_k, srt, feat, K, S, S3, half_span, inference, out + static_cast<size_t>(i) * DF);
}

// =====================================================================================================================
// Grid-binned query.  A G^3 grid over the bounding box of the primitives' boxes (united with [-1,1]^3) carries two
// ascending index lists per cell, built on the device from srt:
//   cover[c] : primitives whose box [pos - scale, pos + scale] meets the cell            (a superset of those covering a point of c)
//   near[c]  : primitives that can be the nearest centre for SOME point of the cell:      mindist(c, pos_k) <= min_j maxdist(c, pos_j)
// A point walks cover[cell] (same visit_prim, same ascending order => the same sums as the exhaustive kernel) and, only if nothing
// covers it, near[cell] for the nearest centre (same strict-< scan in ascending index => the same argmin).  Points outside
// the grid, or a grid whose lists overflowed the workspace, take the exhaustive loop.
// Layout of the workspace: GridHdr | counts[2 C] | offsets[2 C + 2] | entries[cap]      (C = G^3)
// =====================================================================================================================
constexpr int PG_G = 32;
constexpr int PG_CELLS = PG_G * PG_G * PG_G;
struct GridHdr {
    float ox, oy, oz, inv_h, h;
    int total_cover, total_near, overflow, cap, K;
};

__global__ void __launch_bounds__(256) grid_bounds_kernel(const float* __restrict__ srt, int K, GridHdr* __restrict__ hdr, int cap) {
    __shared__ float s_lo[3][256], s_hi[3][256];
    float lo[3] = {-1.f, -1.f, -1.f}, hi[3] = {1.f, 1.f, 1.f};
    for (int k = threadIdx.x; k < K; k += 256) {
        const float4 v = *reinterpret_cast<const float4*>(srt + 4 * static_cast<size_t>(k));
        const float s = fabsf(v.x);
        const float c[3] = {v.y, v.z, v.w};
#pragma unroll
        for (int d = 0; d < 3; ++d) { lo[d] = fminf(lo[d], c[d] - s); hi[d] = fmaxf(hi[d], c[d] + s); }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) { s_lo[d][threadIdx.x] = lo[d]; s_hi[d][threadIdx.x] = hi[d]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = 1; t < 256; ++t)
            for (int d = 0; d < 3; ++d) { lo[d] = fminf(lo[d], s_lo[d][t]); hi[d] = fmaxf(hi[d], s_hi[d][t]); }
        const float ext = fmaxf(hi[0] - lo[0], fmaxf(hi[1] - lo[1], hi[2] - lo[2])) * 1.0001f + 1e-6f;
        hdr->ox = lo[0]; hdr->oy = lo[1]; hdr->oz = lo[2];
        hdr->h = ext / PG_G;
        hdr->inv_h = PG_G / ext;
        hdr->total_cover = 0; hdr->total_near = 0; hdr->overflow = 0; hdr->cap = cap; hdr->K = K;
    }
}

// one warp per cell; FILL = 0: count both lists, FILL = 1: write them (ascending k, ordered warp compaction)
template <int FILL>
__global__ void __launch_bounds__(256) grid_lists_kernel(const float* __restrict__ srt, int K, const GridHdr* __restrict__ hdr, int* __restrict__ counts,
                                                         const int* __restrict__ offsets, int* __restrict__ entries) {
    extern __shared__ float4 s_srt[];      // all K primitives
    for (int k = threadIdx.x; k < K; k += blockDim.x) s_srt[k] = *reinterpret_cast<const float4*>(srt + 4 * static_cast<size_t>(k));
    __syncthreads();
    const int cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (cell >= PG_CELLS) return;
    const int lane = threadIdx.x & 31;
    const float h = hdr->h;
    const int cz = cell / (PG_G * PG_G), cy = (cell / PG_G) % PG_G, cx = cell % PG_G;
    const float lo[3] = {hdr->ox + cx * h, hdr->oy + cy * h, hdr->oz + cz * h};
    const float hi[3] = {lo[0] + h, lo[1] + h, lo[2] + h};
    // pass A: upper bound of the nearest-centre distance over the cell = min_k maxdist(cell, pos_k)
    float ub2 = INFINITY;
    for (int k = lane; k < K; k += 32) {
        const float4 v = s_srt[k];
        const float c[3] = {v.y, v.z, v.w};
        float m2 = 0.f;
#pragma unroll
        for (int d = 0; d < 3; ++d) { const float m = fmaxf(fabsf(c[d] - lo[d]), fabsf(c[d] - hi[d])); m2 += m * m; }
        ub2 = fminf(ub2, m2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ub2 = fminf(ub2, __shfl_xor_sync(0xffffffffu, ub2, o));
    ub2 = ub2 * 1.0001f + 1e-9f;            // margin over fp32 rounding of the per-point distances
    int n_cover = 0, n_near = 0;
    const int off_c = FILL ? offsets[cell] : 0, off_n = FILL ? offsets[PG_CELLS + cell] : 0;
    const int cap = hdr->cap;
    for (int k0 = 0; k0 < K; k0 += 32) {
        const int k = k0 + lane;
        bool cover = false, near = false;
        if (k < K) {
            const float4 v = s_srt[k];
            const float s = fabsf(v.x) * 1.0001f + 1e-6f;
            const float c[3] = {v.y, v.z, v.w};
            cover = true;
            float d2 = 0.f;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                cover = cover && (c[d] + s >= lo[d]) && (c[d] - s <= hi[d]);
                const float g = fmaxf(fmaxf(lo[d] - c[d], c[d] - hi[d]), 0.f);
                d2 += g * g;
            }
            near = d2 <= ub2;
        }
        const unsigned mc = __ballot_sync(0xffffffffu, cover), mn = __ballot_sync(0xffffffffu, near);
        if (FILL) {
            const unsigned below = (1u << lane) - 1u;
            if (cover) { const int pos = off_c + n_cover + __popc(mc & below); if (pos < cap) entries[pos] = k; }
            if (near) { const int pos = off_n + n_near + __popc(mn & below); if (pos < cap) entries[pos] = k; }
        }
        n_cover += __popc(mc);
        n_near += __popc(mn);
    }
    if (!FILL && lane == 0) { counts[cell] = n_cover; counts[PG_CELLS + cell] = n_near; }
}

// exclusive scan of the 2 C counts (single block), totals and the overflow flag
__global__ void __launch_bounds__(1024) grid_scan_kernel(const int* __restrict__ counts, int* __restrict__ offsets, GridHdr* __restrict__ hdr) {
    __shared__ int s_tot[1024];
    constexpr int N = 2 * PG_CELLS, PER = N / 1024;
    const int t = threadIdx.x;
    int local = 0;
    for (int i = 0; i < PER; ++i) local += counts[t * PER + i];
    s_tot[t] = local;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = t >= o ? s_tot[t - o] : 0;
        __syncthreads();
        s_tot[t] += v;
        __syncthreads();
    }
    int run = s_tot[t] - local;
    for (int i = 0; i < PER; ++i) { offsets[t * PER + i] = run; run += counts[t * PER + i]; }
    if (t == 1023) {
        offsets[N] = run;
        hdr->total_cover = offsets[PG_CELLS];
        hdr->total_near = run - offsets[PG_CELLS];
        hdr->overflow = run > hdr->cap ? 1 : 0;
    }
}

template <int DF>
__global__ void __launch_bounds__(PS_THREADS) primsdf_query_grid_kernel(const float* __restrict__ x, const float* __restrict__ srt,
                                                                        const float* __restrict__ feat, const GridHdr* __restrict__ hdr,
                                                                        const int* __restrict__ offsets, const int* __restrict__ entries, long long n, int K,
                                                                        int S, int inference, float* __restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * PS_THREADS + threadIdx.x;
    if (i >= n) return;
    const float px = x[3 * i], py = x[3 * i + 1], pz = x[3 * i + 2];
    PointAcc<DF> a;
#pragma unroll
    for (int c = 0; c < DF; ++c) a.acc[c] = 0.f;
    a.wsum = 0.f;
    float best_d2 = INFINITY;
    int best_k = 0;
    const int S3 = S * S * S;
    const float half_span = 0.5f * static_cast<float>(S - 1);
    const float gx = (px - hdr->ox) * hdr->inv_h, gy = (py - hdr->oy) * hdr->inv_h, gz = (pz - hdr->oz) * hdr->inv_h;
    const bool inside = gx >= 0.f && gy >= 0.f && gz >= 0.f && gx < static_cast<float>(PG_G) && gy < static_cast<float>(PG_G) && gz < static_cast<float>(PG_G);
    if (!inside || hdr->overflow != 0) {          // exhaustive loop (rare: points outside the primitives' bounding box / oversized lists)
        for (int k = 0; k < K; ++k) {
            const float4 p = __ldg(reinterpret_cast<const float4*>(srt + 4 * static_cast<size_t>(k)));
            const float dx = px - p.y, dy = py - p.z, dz = pz - p.w;
            const float d2 = dx * dx + dy * dy + dz * dz;
            if (d2 < best_d2) { best_d2 = d2; best_k = k; }
            visit_prim<DF>(a, px, py, pz, p, k, feat, S, S3, half_span);
        }
    } else {
        const int cell = (static_cast<int>(gz) * PG_G + static_cast<int>(gy)) * PG_G + static_cast<int>(gx);
        const int c0 = offsets[cell], c1 = offsets[cell + 1];
        for (int e = c0; e < c1; ++e) {
            const int k = entries[e];
            const float4 p = __ldg(reinterpret_cast<const float4*>(srt + 4 * static_cast<size_t>(k)));
            visit_prim<DF>(a, px, py, pz, p, k, feat, S, S3, half_span);
        }
        if (!(a.wsum > 0.f) && inference) {
            const int n0 = offsets[PG_CELLS + cell], n1 = offsets[PG_CELLS + cell + 1];
            for (int e = n0; e < n1; ++e) {
                const int k = entries[e];
                const float4 p = __ldg(reinterpret_cast<const float4*>(srt + 4 * static_cast<size_t>(k)));
                const float dx = px - p.y, dy = py - p.z, dz = pz - p.w;
                const float d2 = dx * dx + dy * dy + dz * dz;
                if (d2 < best_d2) { best_d2 = d2; best_k = k; }
            }
        }
    }
    finish_point<DF>(a, px, py, pz, best_k, srt, feat, K, S, S3, half_span, inference, out + static_cast<size_t>(i) * DF);
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

size_t primsdf_grid_bytes(long long cap_entries) {
    return sizeof(GridHdr) + 256 + static_cast<size_t>(2 * PG_CELLS) * 4 + static_cast<size_t>(2 * PG_CELLS + 2) * 4 + static_cast<size_t>(cap_entries) * 4;
}

namespace {
struct GridPtrs { GridHdr* hdr; int* counts; int* offsets; int* entries; long long cap; };
GridPtrs grid_carve(void* ws, size_t bytes) {
    uint8_t* b = static_cast<uint8_t*>(ws);
    GridPtrs g;
    g.hdr = reinterpret_cast<GridHdr*>(b);
    g.counts = reinterpret_cast<int*>(b + 256);
    g.offsets = g.counts + 2 * PG_CELLS;
    g.entries = g.offsets + 2 * PG_CELLS + 2;
    const size_t fixed = 256 + static_cast<size_t>(4 * PG_CELLS + 2) * 4;
    g.cap = bytes > fixed ? static_cast<long long>((bytes - fixed) / 4) : 0;
    return g;
}
}  // namespace

int launch_primsdf_grid_build(const float* srt, int K, void* ws, size_t ws_bytes, cudaStream_t st) {
    TPX_CHECK(K > 0 && K <= 4096, TPX_ERR_SHAPE, "primsdf_grid_build: K %d (1..4096 primitives fit the builder's shared memory)", K);
    TPX_CHECK((reinterpret_cast<uintptr_t>(srt) & 15) == 0 && (reinterpret_cast<uintptr_t>(ws) & 255) == 0, TPX_ERR_ARG, "primsdf_grid_build: alignment");
    const GridPtrs g = grid_carve(ws, ws_bytes);
    TPX_CHECK(g.cap >= 1024, TPX_ERR_ARG, "primsdf_grid_build: workspace too small");
    ProfScope prof(PROF_VAE_MISC, st);
    const int cap = g.cap > 0x7fffffffLL ? 0x7fffffff : static_cast<int>(g.cap);
    grid_bounds_kernel<<<1, 256, 0, st>>>(srt, K, g.hdr, cap);
    TPX_LAUNCH_CHECK();
    const size_t smem = static_cast<size_t>(K) * 16;
    static bool attr_set = false;
    if (!attr_set) {
        TPX_CUDA(cudaFuncSetAttribute(grid_lists_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 16));
        TPX_CUDA(cudaFuncSetAttribute(grid_lists_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 16));
        attr_set = true;
    }
    grid_lists_kernel<0><<<PG_CELLS / 8, 256, smem, st>>>(srt, K, g.hdr, g.counts, g.offsets, g.entries);
    TPX_LAUNCH_CHECK();
    grid_scan_kernel<<<1, 1024, 0, st>>>(g.counts, g.offsets, g.hdr);
    TPX_LAUNCH_CHECK();
    grid_lists_kernel<1><<<PG_CELLS / 8, 256, smem, st>>>(srt, K, g.hdr, g.counts, g.offsets, g.entries);
    TPX_LAUNCH_CHECK();
    return TPX_OK;
}

int launch_primsdf_query_grid(const float* x, const float* srt, const float* feat, const void* ws, size_t ws_bytes, long long n, int K, int S, int dim_feat,
                              int inference, float* out, cudaStream_t st) {
    TPX_CHECK(dim_feat == 6, TPX_ERR_SHAPE, "primsdf_query: dim_feat %d (kernel covers the released 6-channel layout)", dim_feat);
    TPX_CHECK(S >= 2 && S <= 32 && K > 0, TPX_ERR_SHAPE, "primsdf_query: bad primitive geometry (S %d, K %d)", S, K);
    if (n <= 0) return TPX_OK;
    const GridPtrs g = grid_carve(const_cast<void*>(ws), ws_bytes);
    ProfScope prof(PROF_VAE_MISC, st);
    const long long blocks = (n + PS_THREADS - 1) / PS_THREADS;
    TPX_CHECK(blocks < (1LL << 31), TPX_ERR_SHAPE, "primsdf_query: too many points");
    primsdf_query_grid_kernel<6><<<static_cast<unsigned>(blocks), PS_THREADS, 0, st>>>(x, srt, feat, g.hdr, g.offsets, g.entries, n, K, S, inference, out);
    TPX_LAUNCH_CHECK();
    return TPX_OK;
}

}  // namespace tpx
