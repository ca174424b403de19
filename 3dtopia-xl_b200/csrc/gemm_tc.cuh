// Persistent, warp-specialised tcgen05 GEMM for sm_100a:   C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
//   warp 0   : TMA producer  (cp.async.bulk.tensor -> 128B/64B-swizzled smem ring, mbarrier complete_tx)
//   warp 1   : MMA issuer    (one thread: tcgen05.mma cta_group::1 kind::f16, 128 x BN x 16, fp32 accum in TMEM)
//   warp 2   : TMEM allocator
//   warps 4-7: epilogue      (tcgen05.ld 32x32b, one accumulator row per thread; fused bias / GELU / gated
//                             residual / head split / conv scatter; 16-byte stores)
//   Two TMEM accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// A-operand modes: AMODE_LINEAR (2-D tensor map over a row-major [M,K] fp16 matrix) and AMODE_CONV3 (5-D tensor
// map over a channels-last [P,S,S,S,C] fp16 volume: the 27 taps of a 3x3x3 / pad 1 convolution are 27 shifted
// box loads, the halo comes from TMA out-of-bounds zero fill — implicit GEMM with no im2col buffer).
#pragma once
#include "tpx_common.cuh"

namespace tpx {

enum GemmEpiMode : int {
    EPI_STORE = 0,       // out0[row*ldo+col] = h( h(acc+bias) * post_scale )
    EPI_GELU = 1,        // out0 = h( gelu_tanh( h(acc+bias) ) )
    EPI_HEADS = 2,       // split columns into (which, head, d) -> out{which}[b,head,n,DhP]  (zero pad written)
    EPI_GATED = 3,       // xres[row*ldx+col] += float( h( gate[b,col] * h(acc+bias) ) )
    EPI_RESID_SCALE = 4, // out0 = h( (acc + bias + resid[row*ldo+col]) * alpha )
    EPI_CONVT2 = 5,      // k2s2 transposed conv scatter: col=(abc,co), row=(p,z,y,x)@4^3 -> channels-last 8^3
    EPI_NCDHW = 6,       // out[(p*Cout+co)*S3 + vox] = acc + bias   (fp32 or fp16 out, co < n_valid)
};
enum GemmAMode : int { AMODE_LINEAR = 0, AMODE_CONV3 = 1 };

struct GemmArgs {
    int M, N, num_kb;
    int conv_S, chunks_per_tap;
    const __half* bias;
    float post_scale;
    __half* out0;
    __half* out1;
    __half* out2;
    int ldo;
    int split_cols, Dh, DhP, H, Nseq;
    float* xres;
    int ldx;
    const __half* gate;
    int gate_bstride, rows_per_batch, gate_batches;
    const __half* resid;
    float alpha;
    float* out32;
    int n_valid, S3;
};

template <int BN, int BK>
struct GemmCfg {
    static constexpr int BM = 128;
    static constexpr int SW = BK * 2;
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES_RAW = (196 * 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    static_assert(B_BYTES % 1024 == 0 && A_BYTES % 1024 == 0, "stage operands must stay 1024-B aligned");
    static_assert(BN % 16 == 0 && BN <= 256, "UMMA N");
};

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

union Pack8 {
    uint4 u;
    __half2 h2[4];
    __half h[8];
};

// ---- epilogue for one 8-column group of one row ---------------------------------------------------------
template <int EPI>
__device__ __forceinline__ void epi_group8(const GemmArgs& g, int row, int col, const uint32_t* acc /*8 fp32 bit patterns*/) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(acc[i]);
    if (g.bias != nullptr) {
        Pack8 b;
        b.u = *reinterpret_cast<const uint4*>(g.bias + col);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += __half2float(b.h[i]);
    }
    if constexpr (EPI == EPI_STORE || EPI == EPI_GELU) {
        Pack8 o;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float x = h2f_round(v[i]);
            if constexpr (EPI == EPI_GELU) x = gelu_tanh(x);
            else if (g.post_scale != 1.0f) x = x * g.post_scale;
            o.h[i] = __float2half_rn(x);
        }
        *reinterpret_cast<uint4*>(g.out0 + static_cast<size_t>(row) * g.ldo + col) = o.u;
    } else if constexpr (EPI == EPI_HEADS) {
        const int which = col / g.split_cols;
        const int c = col - which * g.split_cols;
        const int head = c / g.Dh;
        const int d = c - head * g.Dh;
        const int b = row / g.Nseq;
        const int n = row - b * g.Nseq;
        __half* base = which == 0 ? g.out0 : (which == 1 ? g.out1 : g.out2);
        __half* dst = base + (static_cast<size_t>(b * g.H + head) * g.Nseq + n) * g.DhP + d;
        Pack8 o;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float x = h2f_round(v[i]);
            if (g.post_scale != 1.0f && which == 0) x = x * g.post_scale;
            o.h[i] = __float2half_rn(x);
        }
        *reinterpret_cast<uint4*>(dst) = o.u;
        if (d + 8 == g.Dh) {  // last real group of this head: write the zero padding d in [Dh, DhP)
            for (int p = g.Dh; p < g.DhP; p += 8) *reinterpret_cast<uint4*>(dst + (p - d)) = make_uint4(0, 0, 0, 0);
        }
    } else if constexpr (EPI == EPI_GATED) {
        const int b = (row / g.rows_per_batch) % g.gate_batches;
        Pack8 gt;
        gt.u = *reinterpret_cast<const uint4*>(g.gate + static_cast<size_t>(b) * g.gate_bstride + col);
        float* xp = g.xres + static_cast<size_t>(row) * g.ldx + col;
        float4 x0 = *reinterpret_cast<float4*>(xp);
        float4 x1 = *reinterpret_cast<float4*>(xp + 4);
        float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) xs[i] += h2f_round(__half2float(gt.h[i]) * h2f_round(v[i]));
        *reinterpret_cast<float4*>(xp) = make_float4(xs[0], xs[1], xs[2], xs[3]);
        *reinterpret_cast<float4*>(xp + 4) = make_float4(xs[4], xs[5], xs[6], xs[7]);
    } else if constexpr (EPI == EPI_RESID_SCALE) {
        const size_t off = static_cast<size_t>(row) * g.ldo + col;
        Pack8 o;
        if (g.resid != nullptr) {
            Pack8 r;
            r.u = *reinterpret_cast<const uint4*>(g.resid + off);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += __half2float(r.h[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) o.h[i] = __float2half_rn(v[i] * g.alpha);
        *reinterpret_cast<uint4*>(g.out0 + off) = o.u;
    } else if constexpr (EPI == EPI_CONVT2) {
        const int abc = col / g.split_cols;
        const int co = col - abc * g.split_cols;
        const int p = row >> 6, vox = row & 63;
        const int z = vox >> 4, y = (vox >> 2) & 3, x = vox & 3;
        const int oz = 2 * z + (abc >> 2), oy = 2 * y + ((abc >> 1) & 1), ox = 2 * x + (abc & 1);
        Pack8 o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.h[i] = __float2half_rn(v[i]);
        *reinterpret_cast<uint4*>(g.out0 + (static_cast<size_t>(p) * 512 + (oz * 8 + oy) * 8 + ox) * g.split_cols + co) = o.u;
    } else if constexpr (EPI == EPI_NCDHW) {
        const int p = row / g.S3, vox = row - p * g.S3;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int co = col + i;
            if (co < g.n_valid) {
                const size_t off = (static_cast<size_t>(p) * g.n_valid + co) * g.S3 + vox;
                if (g.out32 != nullptr) g.out32[off] = v[i];
                else g.out0[off] = __float2half_rn(v[i]);
            }
        }
    }
}

template <int BN, int BK, int AMODE, int EPI>
__global__ void __launch_bounds__(256, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
    using Cfg = GemmCfg<BN, BK>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 128);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_n = (g.N + BN - 1) / BN;
    const int tiles_m = (g.M + 127) / 128;
    const int num_tiles = tiles_m * tiles_n;
    const int num_kb = g.num_kb;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile / tiles_n, n_blk = tile - m_blk * tiles_n;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    if constexpr (AMODE == AMODE_LINEAR) {
                        tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m_blk * 128);
                    } else {
                        const int tap = kb / g.chunks_per_tap, chunk = kb - tap * g.chunks_per_tap;
                        const int dz = tap / 9, dy = (tap / 3) % 3, dx = tap % 3;
                        if (g.conv_S == 4) {  // 128 rows = 2 primitives x 4^3 voxels
                            tma_load_5d(sa, &tmA, &full_bar[stage], chunk * BK, dx - 1, dy - 1, dz - 1, m_blk * 2);
                        } else {              // S == 8: 128 rows = 2 z-slices of one primitive
                            tma_load_5d(sa, &tmA, &full_bar[stage], chunk * BK, dx - 1, dy - 1, (m_blk & 3) * 2 + dz - 1, m_blk >> 2);
                        }
                    }
                    tma_load_2d(sa + Cfg::A_BYTES, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(128, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t adesc = umma_desc_kmajor<Cfg::SW>(a_addr);
                    const uint64_t bdesc = umma_desc_kmajor<Cfg::SW>(a_addr + Cfg::A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        umma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);   // frees the smem slot when these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[acc]);         // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        const int quad = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile / tiles_n, n_blk = tile - m_blk * tiles_n;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const int row = m_blk * 128 + quad * 32 + lane;
            const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16);
            constexpr int CH = BN >= 32 ? 32 : 16;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += CH) {
                uint32_t r[32];
                if constexpr (CH == 32) tmem_ld_32x32(taddr + c0, r);
                else tmem_ld_32x16(taddr + c0, r);
                tmem_ld_wait();
                if (row < g.M) {
#pragma unroll
                    for (int j = 0; j < CH; j += 8) {
                        const int col = n_blk * BN + c0 + j;
                        if (col < g.N) epi_group8<EPI>(g, row, col, &r[j]);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---- host side ----------------------------------------------------------------------------------------
struct GemmProblem {
    // A operand
    const __half* A;
    int a_mode;         // AMODE_*
    int lda;            // linear: row stride in elements
    int conv_S, conv_C; // conv: spatial size (4|8) and channel count of the channels-last volume; P = M / S^3
    // W operand [N, K] row-major fp16 (K-major)
    const __half* W;
    int M, N, K;
    int BN;             // tile N (128/192/256/64/32/16); BK chosen from K (64, or 32 when conv_C == 32)
    int epi;
    GemmArgs args;      // epilogue fields (M,N,num_kb filled by the launcher)
};
int launch_gemm(const GemmProblem& p, cudaStream_t stream);
int gemm_num_sms();

}  // namespace tpx
