// tcgen05 flash attention for head_dim 72 or 64 (padded to 80):  out = softmax(q k^T * scale) v
//   (contract of xformers.ops.memory_efficient_attention as called at models/attention.py:54,109)
//
// One CTA = 256 query rows (two 128-row tiles A/B) of one (batch, head); K / V^T stream through a 2-stage TMA ring
// in 128-key tiles shared by both query tiles.
//   warp 0      TMA producer (Q once, then K and V^T tiles; mbarrier complete_tx)
//   warp 1      MMA issuer   (one thread):  S_X = Q_X K^T  (128x128x80, fp32 in TMEM),  O_X(j) = P_X V  (128x80x128)
//   warp 2      TMEM allocator (512 columns: S_A, S_B, O_A, O_B)
//   warps 4-7   softmax warpgroup of tile A, warps 8-11 of tile B: ONE THREAD PER QUERY ROW (TMEM lane == row), so the
//               row max / row sum need no shuffles.  The whole 128-wide score row is pulled into registers with four
//               back-to-back tcgen05.ld (S is released to the tensor core immediately), P is written to shared memory
//               as fp16 in the 128B-swizzled K-major layout the PV MMA reads, and O accumulates IN TMEM across key
//               tiles.  The running max is allowed to go stale by up to 2^8 (p <= 256 fits fp16 comfortably); only
//               when a row's max grows by more than that is O rescaled in TMEM (tcgen05.ld / mul / tcgen05.st), which
//               is rare after the first tiles — no per-tile correction pass.
//   setmaxnreg moves registers from the control warps (56) to the softmax warps (200); ptxas does NOT bound a branch's
//   register use by its setmaxnreg value, so the control budget must cover what the control code really uses for the 128-register score row.
// While warpgroup A does softmax on tile j the tensor core runs S_B(j) / PV; S_X(j+1) is issued as soon as warpgroup X
// has pulled S_X(j) into registers, so MMA, TMA and the exponentials overlap.
//
// Layouts: q, k  [B,H,N,80] fp16 (d >= 72 zero);  vT [B,H,80,NkPad] fp16 (V transposed: keys contiguous, NkPad % 8 == 0,
// rows d >= 72 and key columns >= Nk finite/zero);  out [B,Nq,H*72] fp16.
#include <cstdlib>

#include "kernels.cuh"

namespace tpx {

namespace {

constexpr int TA_DHP = 80;
// row Dh of V^T (the first padding row; Dh = 72 or 64) / column Dh of O carries the softmax row sums
constexpr int TA_BQ = 128, TA_BKV = 128;
constexpr int TA_Q_BYTES = 128 * 128 + 128 * 32;        // 64-wide SW128 part + 16-wide SW32 part
constexpr int TA_K_BYTES = TA_Q_BYTES;
constexpr int TA_VBOX_BYTES = TA_DHP * 128;             // 80 rows x 64 keys
constexpr int TA_V_BYTES = 2 * TA_VBOX_BYTES;
constexpr int TA_P_BYTES = 2 * 128 * 128;
constexpr int TA_KV_STAGE = TA_K_BYTES + TA_V_BYTES;
constexpr int TA_OFF_Q = 0;
constexpr int TA_OFF_KV = 2 * TA_Q_BYTES;
constexpr int TA_KV_STAGES = 3;
constexpr int TA_OFF_P = TA_OFF_KV + TA_KV_STAGES * TA_KV_STAGE;
constexpr int TA_OFF_BAR = TA_OFF_P + 2 * TA_P_BYTES;
constexpr int TA_SMEM = TA_OFF_BAR + 256 + 1024;
constexpr int TA_THREADS = 384;
static_assert(TA_Q_BYTES % 1024 == 0 && TA_VBOX_BYTES % 1024 == 0, "operand tiles must stay 1024-B aligned");

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 2^x on the FMA pipe: x = n + r with n = round(x), r in [-0.5, 0.5]; cubic minimax for 2^r (max relative error 7.5e-5, below
// the fp16 rounding P gets anyway); n goes straight into the exponent field.  Takes a share of the exponentials off the
// MUFU unit, which is the bound of the softmax warps (16 ex2/clk/SM: 2048 cycles per 2x128x128 tile pair).
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -126.0f);
    const float fl = x + 12582912.0f;                 // 1.5 * 2^23: the low mantissa bits now hold n
    const float r = x - (fl - 12582912.0f);
    float p = fmaf(0.0551716648f, r, 0.2426111251f);
    p = fmaf(p, r, 0.6932609677f);
    p = fmaf(p, r, 0.9999280572f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(fl) << 23));
}

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// timeline probe (only when a debug buffer is passed): block 0, one thread per role writes (tag, clock) pairs
#define TA_DBG(slot, tag)                                                                    \
    do {                                                                                     \
        if (dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {       \
            const int _n = static_cast<int>(dbg[(slot) * 1024]);                             \
            if (_n < 500) { dbg[(slot) * 1024 + 1 + 2 * _n] = (tag); dbg[(slot) * 1024 + 2 + 2 * _n] = clock64(); dbg[(slot) * 1024] = _n + 1; } \
        }                                                                                    \
    } while (0)

// POLY: every POLY-th exponential of a full tile is evaluated by ex2_poly instead of MUFU.EX2 (0 = none).
template <int POLY>
__global__ void __launch_bounds__(TA_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQa, const __grid_constant__ CUtensorMap tmQb, const __grid_constant__ CUtensorMap tmKa,
                    const __grid_constant__ CUtensorMap tmKb, const __grid_constant__ CUtensorMap tmV, __half* __restrict__ out, int H, int Nq, int Nk,
                    int Dh, float scale_log2, long long* __restrict__ dbg, unsigned stagger_cycles, int flags) {
    extern __shared__ __align__(1024) uint8_t ta_smem_raw[];
    // 1024-B alignment by pointer arithmetic on the __shared__ array (an integer round trip would demote every later
    // access to generic LD/ST instead of LDS/STS)
    uint8_t* smem = ta_smem_raw + ((1024u - (smem_u32(ta_smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TA_OFF_BAR);
    uint64_t* q_full = bars;            // 1
    uint64_t* kv_full = bars + 1;       // TA_KV_STAGES
    uint64_t* kv_empty = bars + 4;      // TA_KV_STAGES
    uint64_t* s_full = bars + 7;        // 2 (per query tile)
    uint64_t* s_free = bars + 9;        // 2
    uint64_t* p_full = bars + 11;       // 2
    uint64_t* o_full = bars + 13;       // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (2 * TA_BQ);
    const int h = blockIdx.y, b = blockIdx.z;
    const int bh = b * H + h;
    const int nkt = (Nk + TA_BKV - 1) / TA_BKV;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQa); tma_prefetch_desc(&tmQb); tma_prefetch_desc(&tmKa); tma_prefetch_desc(&tmKb); tma_prefetch_desc(&tmV);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < TA_KV_STAGES; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_empty[i], 2);      // one commit from each of the two MMA issuer warps
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_free[i], 128);
            mbar_init(&p_full[i], 128);
            mbar_init(&o_full[i], 1);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();
    // TMEM columns: S_A [0,128) S_B [128,256) O_A [256,336) O_B [384,464)

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(q_full, 2 * TA_Q_BYTES);
            for (int X = 0; X < 2; ++X) {
                uint8_t* qs = smem + TA_OFF_Q + X * TA_Q_BYTES;
                const int row = bh * Nq + q0 + X * TA_BQ;
                tma_load_2d(qs, &tmQa, q_full, 0, row);
                tma_load_2d(qs + 128 * 128, &tmQb, q_full, 64, row);
            }
        }
        __syncwarp();
        for (int j = 0; j < nkt; ++j) {
            const int s = j % TA_KV_STAGES;
            mbar_wait(&kv_empty[s], ((j / TA_KV_STAGES) & 1) ^ 1);
            if (elect_one()) {
                uint8_t* ks = smem + TA_OFF_KV + s * TA_KV_STAGE;
                uint8_t* vs = ks + TA_K_BYTES;
                mbar_arrive_expect_tx(&kv_full[s], TA_KV_STAGE);
                const int krow = bh * Nk + j * TA_BKV;
                tma_load_2d(ks, &tmKa, &kv_full[s], 0, krow);
                tma_load_2d(ks + 128 * 128, &tmKb, &kv_full[s], 64, krow);
                tma_load_2d(vs, &tmV, &kv_full[s], j * TA_BKV, bh * TA_DHP);
                tma_load_2d(vs + TA_VBOX_BYTES, &tmV, &kv_full[s], j * TA_BKV + 64, bh * TA_DHP);
            }
            __syncwarp();
        }
    } else if (warp == 1 || warp == 3) {
        // Two warp-converged issuer warps (uniform control flow, one elected lane issues; descriptors stay in uniform
        // registers): warp 1 drives query tile A, warp 3 tile B, so the two softmax warpgroups run as independent pipelines
        // and can de-phase instead of being re-synchronised by one in-order issuer.
        const int X = warp == 1 ? 0 : 1;
        constexpr uint32_t idesc_qk = umma_idesc_f16(128, 128);
        constexpr uint32_t idesc_pv = umma_idesc_f16(128, TA_DHP);
        const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
        const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
        auto issue_qk = [&](int X, int s) {
            const uint32_t qa = sbase + TA_OFF_Q + X * TA_Q_BYTES, ka = sbase + TA_OFF_KV + s * TA_KV_STAGE;
            const uint64_t dq = umma_desc_kmajor<128>(qa), dk = umma_desc_kmajor<128>(ka);
            const uint64_t dq2 = umma_desc_kmajor<32>(qa + 128 * 128), dk2 = umma_desc_kmajor<32>(ka + 128 * 128);
            const uint32_t ts = tbase + X * 128;
            if (elect_one()) {
#pragma unroll
                for (int i = 0; i < 4; ++i) umma_f16(ts, dq + 2 * i, dk + 2 * i, idesc_qk, i > 0 ? 1u : 0u);
                umma_f16(ts, dq2, dk2, idesc_qk, 1u);
                umma_commit(&s_full[X]);
            }
            __syncwarp();
        };
        auto issue_pv = [&](int X, int s, uint32_t acc_first) {
            const uint32_t pa = sbase + TA_OFF_P + X * TA_P_BYTES, va = sbase + TA_OFF_KV + s * TA_KV_STAGE + TA_K_BYTES;
            const uint32_t to = tbase + 256 + X * 128;
            const uint64_t dp0 = umma_desc_kmajor<128>(pa), dv0 = umma_desc_kmajor<128>(va);
            const uint64_t dp1 = umma_desc_kmajor<128>(pa + 128 * 128), dv1 = umma_desc_kmajor<128>(va + TA_VBOX_BYTES);
            if (elect_one()) {
#pragma unroll
                for (int i = 0; i < 4; ++i) umma_f16(to, dp0 + 2 * i, dv0 + 2 * i, idesc_pv, i != 0 ? 1u : acc_first);
#pragma unroll
                for (int i = 0; i < 4; ++i) umma_f16(to, dp1 + 2 * i, dv1 + 2 * i, idesc_pv, 1u);
                umma_commit(&o_full[X]);
            }
            __syncwarp();
        };
        // Row sums on the tensor core: row Dh of every V^T tile (a padding row, Dh < 80) is overwritten with ones after the TMA
        // lands, so column Dh of O accumulates sum_k P[r,k] — rescaled together with O — and the softmax threads carry no sum.
        // Both issuer warps write the same 2 x 128 bytes (row 72 is swizzle row 0: identity chunk order).
        auto write_ones = [&](int s) {
            const uint32_t a = sbase + TA_OFF_KV + s * TA_KV_STAGE + TA_K_BYTES + Dh * 128 + lane * 4;
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(0x3C003C00u) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(a + TA_VBOX_BYTES), "r"(0x3C003C00u) : "memory");
            fence_proxy_async();
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        mbar_wait(&kv_full[0], 0);
        write_ones(0);
        tc_fence_after();
        issue_qk(X, 0);
        for (int j = 0; j < nkt; ++j) {
            const int s = j % TA_KV_STAGES;
            if (j + 1 < nkt) {           // next scores first: the warpgroup gets S(j+1) while it exponentiates tile j
                const int sn = (j + 1) % TA_KV_STAGES;
                mbar_wait(&kv_full[sn], ((j + 1) / TA_KV_STAGES) & 1);
                write_ones(sn);
                mbar_wait(&s_free[X], j & 1);
                tc_fence_after();
                issue_qk(X, sn);
            }
            mbar_wait(&p_full[X], j & 1);
            tc_fence_after();
            issue_pv(X, s, j > 0 ? 1u : 0u);      // O_X accumulates in TMEM across key tiles
            if (elect_one()) umma_commit(&kv_empty[s]);
            __syncwarp();
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        const int X = (warp - 4) >> 2;               // query tile of this warpgroup
        const int quad = warp & 3;
        const int r = quad * 32 + lane;              // row in the tile == TMEM lane
        const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t tS = tmem_base + X * 128 + lane_off;
        const uint32_t tO = tmem_base + 256 + X * 128 + lane_off;
        const uint32_t pS = smem_u32(smem) + TA_OFF_P + X * TA_P_BYTES + r * 128;   // shared-window address of this row of P
        const int sw = r & 7;
        float m_ref = 0.f;
        const bool stale_max = (flags & 1) != 0;
        // De-phase the two softmax warpgroups: they share the four MUFU units, so running their exponential phases in
        // lockstep halves each one's rate while the XU idles during their (also simultaneous) load / max / sync phases.
        for (int j = 0; j < nkt; ++j) {
            const int nvalid = Nk - j * TA_BKV;      // keys of this tile that exist (>= 1)
            if (r == 0) TA_DBG(X, 1);
            mbar_wait(&s_full[X], j & 1);
            tc_fence_after();
            if (r == 0) TA_DBG(X, 2);
            if (X == 1 && j == 0 && stagger_cycles != 0) {
                // De-phase the two softmax warpgroups once, by a fixed number of cycles.  Warp q of each warpgroup sits on
                // sub-partition q and the two share its MUFU unit (16 ex2/clk/SM in all; tools/ubench/softmax_loop.cu: 1376 cycles
                // per 128-exponential row pass for one warp alone, 2436 for two together).  In lockstep both exponentiate at half
                // rate and then both leave the unit idle through their load / wait phases; half a tile period apart the phases
                // interleave and the unit is nearly saturated (2 x 1400 of a ~3000-cycle period): 68 -> 62 us.  The window is
                // narrow (1400-1800 cycles; nothing at 1000 or 2200) and has to be re-measured when this kernel changes
                // (tools/attn_perf.py).  Locking the phase on every tile — smem counter, one-way or two-way mbarrier ping-pong,
                // or starting B when A's first tile is done — cost more than the drift it removes (72-79 us).
                const long long t_end = clock64() + stagger_cycles;
                while (clock64() < t_end) {}
            }
            uint32_t sv[128];
            tmem_ld_32x32(tS, reinterpret_cast<uint32_t(&)[32]>(sv[0]));
            tmem_ld_32x32(tS + 32, reinterpret_cast<uint32_t(&)[32]>(sv[32]));
            tmem_ld_32x32(tS + 64, reinterpret_cast<uint32_t(&)[32]>(sv[64]));
            tmem_ld_32x32(tS + 96, reinterpret_cast<uint32_t(&)[32]>(sv[96]));
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&s_free[X]);                 // the score row is in registers: S_X may be overwritten by Q K^T(j+1)
            if (r == 0) TA_DBG(X, 3);
            // O_X rescale by 2^((m_ref - mx) * scale) for the rows whose running maximum grew past the lazy threshold
            auto rescale_o = [&](bool need, float mx) {
                const float a = need ? ex2((m_ref - mx) * scale_log2) : 1.0f;
                if (need) m_ref = mx;
                uint32_t t[32];
                tmem_ld_32x32(tO, t);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * a);
                tmem_st_32x32(tO, t);
                tmem_ld_32x32(tO + 32, t);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * a);
                tmem_st_32x32(tO + 32, t);
                tmem_ld_32x16(tO + 64, t);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * a);
                tmem_st_32x16(tO + 64, t);
                tmem_st_wait();
            };
            if (j > 0 && nvalid >= TA_BKV && stale_max) {
                // Fast path (every full tile after the first).  The lazy-rescale rule already tolerates a reference maximum that
                // is stale by up to 2^8, so the exponentials do not have to wait for this tile's row maximum: they run against
                // m_ref while the maximum is reduced in the same instruction stream (its ~600-cycle dependent chain hides
                // under the MUFU-bound exponentials).  Only if some row's maximum did jump past the threshold is the tile redone.
                mbar_wait(&o_full[X], (j - 1) & 1);              // P V(j-1) retired: P_X may be rewritten, O_X is quiescent
                tc_fence_after();
                if (r == 0) TA_DBG(X, 5);
                const float msc = m_ref * scale_log2;
                float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float s0 = __uint_as_float(sv[c * 32 + i]), s1 = __uint_as_float(sv[c * 32 + i + 1]);
                        const float s2 = __uint_as_float(sv[c * 32 + i + 2]), s3 = __uint_as_float(sv[c * 32 + i + 3]);
                        m0 = fmaxf(m0, fmaxf(s0, s2)); m1 = fmaxf(m1, fmaxf(s1, s3));      // 3-input FMNMX3, two chains
                        // element (i + k) of the row goes to the FMA-pipe polynomial when its index is a multiple of POLY
                        const float p0 = (POLY > 0 && (c * 32 + i) % POLY == 0) ? ex2_poly(fmaf(s0, scale_log2, -msc)) : ex2(fmaf(s0, scale_log2, -msc));
                        const float p1 = (POLY > 0 && (c * 32 + i + 1) % POLY == 0) ? ex2_poly(fmaf(s1, scale_log2, -msc)) : ex2(fmaf(s1, scale_log2, -msc));
                        const float p2 = (POLY > 0 && (c * 32 + i + 2) % POLY == 0) ? ex2_poly(fmaf(s2, scale_log2, -msc)) : ex2(fmaf(s2, scale_log2, -msc));
                        const float p3 = (POLY > 0 && (c * 32 + i + 3) % POLY == 0) ? ex2_poly(fmaf(s3, scale_log2, -msc)) : ex2(fmaf(s3, scale_log2, -msc));
                        __half2 ha = __floats2half2_rn(p0, p1), hb = __floats2half2_rn(p2, p3);
                        pk[i >> 1] = *reinterpret_cast<uint32_t*>(&ha);
                        pk[(i >> 1) + 1] = *reinterpret_cast<uint32_t*>(&hb);
                    }
                    const uint32_t dst = pS + (c >> 1) * (128 * 128);
                    const int cc0 = (c & 1) * 4;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((cc0 + q) ^ sw) << 4)), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                                     "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                                     : "memory");
                }
                const float mx = fmaxf(m0, m1);
                const bool need = (mx - m_ref) * scale_log2 > 8.0f;
                if (__any_sync(0xffffffffu, need)) {             // rare: redo this tile against the new maximum
                    rescale_o(need, mx);
                    const float msc2 = m_ref * scale_log2;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t pk[16];
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const float p0 = ex2(fmaf(__uint_as_float(sv[c * 32 + i]), scale_log2, -msc2));
                            const float p1 = ex2(fmaf(__uint_as_float(sv[c * 32 + i + 1]), scale_log2, -msc2));
                            __half2 hh = __floats2half2_rn(p0, p1);
                            pk[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
                        }
                        const uint32_t dst = pS + (c >> 1) * (128 * 128);
                        const int cc0 = (c & 1) * 4;
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((cc0 + q) ^ sw) << 4)), "r"(pk[4 * q]),
                                         "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                                         : "memory");
                    }
                }
            } else {
                // First tile, ragged last tile (or stale_max off): maximum first, then the exponentials.
                float mx = -INFINITY;
                if (nvalid >= TA_BKV) {
                    float m0 = __uint_as_float(sv[0]), m1 = __uint_as_float(sv[1]), m2 = __uint_as_float(sv[2]), m3 = __uint_as_float(sv[3]);
#pragma unroll
                    for (int i = 4; i < 128; i += 4) {
                        m0 = fmaxf(m0, __uint_as_float(sv[i]));
                        m1 = fmaxf(m1, __uint_as_float(sv[i + 1]));
                        m2 = fmaxf(m2, __uint_as_float(sv[i + 2]));
                        m3 = fmaxf(m3, __uint_as_float(sv[i + 3]));
                    }
                    mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                } else {
#pragma unroll
                    for (int i = 0; i < 128; ++i)
                        if (i < nvalid) mx = fmaxf(mx, __uint_as_float(sv[i]));
                }
                bool waited_o = false;
                if (j == 0) {
                    m_ref = mx;
                } else {
                    const bool need = (mx - m_ref) * scale_log2 > 8.0f;     // stale max tolerated up to 2^8
                    if (__any_sync(0xffffffffu, need)) {
                        mbar_wait(&o_full[X], (j - 1) & 1);              // P V(j-1) retired: O_X is quiescent
                        tc_fence_after();
                        waited_o = true;
                        rescale_o(need, mx);
                    }
                }
                const float msc = m_ref * scale_log2;
                if (r == 0) TA_DBG(X, 4);
                if (j > 0 && !waited_o) mbar_wait(&o_full[X], (j - 1) & 1);   // P_X buffer free (normally already true)
                if (r == 0) TA_DBG(X, 5);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float p0 = ex2(fmaf(__uint_as_float(sv[c * 32 + i]), scale_log2, -msc));
                        float p1 = ex2(fmaf(__uint_as_float(sv[c * 32 + i + 1]), scale_log2, -msc));
                        if (c * 32 + i >= nvalid) p0 = 0.f;
                        if (c * 32 + i + 1 >= nvalid) p1 = 0.f;
                        __half2 hh = __floats2half2_rn(p0, p1);
                        pk[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
                    }
                    const uint32_t dst = pS + (c >> 1) * (128 * 128);
                    const int cc0 = (c & 1) * 4;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((cc0 + q) ^ sw) << 4)), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                                     "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                                     : "memory");
                }
            }
            if (r == 0) TA_DBG(X, 6);
            fence_proxy_async();            // make the P stores visible to the tensor core (async proxy)
            tc_fence_before();
            mbar_arrive(&p_full[X]);
            if (r == 0) TA_DBG(X, 7);
        }
        mbar_wait(&o_full[X], (nkt - 1) & 1);
        tc_fence_after();
        const int row = q0 + X * TA_BQ + r;
        __half* orow = out + (static_cast<size_t>(b) * Nq + (row < Nq ? row : 0)) * (H * Dh) + h * Dh;
        uint32_t t2[32];
        tmem_ld_32x16(tO + 64, t2);                  // columns 64..79: the last 8 value columns and, at column Dh, the row sum
        tmem_ld_wait();
        const float inv = 1.0f / __uint_as_float((Dh == 64 ? t2[0] : t2[8]));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            uint32_t t[32];
            if (c < 2) {
                tmem_ld_32x32(tO + c * 32, t);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) t[i] = t2[i];
            }
#pragma unroll
            for (int d8 = 0; d8 < (c < 2 ? 32 : 16); d8 += 8) {
                const int d = c * 32 + d8;
                if (row < Nq && d < Dh) {
                    Pack8 v;
#pragma unroll
                    for (int i = 0; i < 8; ++i) v.h[i] = __float2half_rn(__uint_as_float(t[d8 + i]) * inv);
                    *reinterpret_cast<uint4*>(orow + d) = v.u;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}


// =====================================================================================================================
// Software-pipelined variant of attention_tc_kernel (same roles, layouts, barriers).  Measured (tools/attn_timeline.py): a
// softmax warp spends ~1000 cycles of every ~3300-cycle tile outside its exponentials — barrier polls, the tcgen05.ld of the
// score row, the wait for the single P buffer, fence + arrive — and the two warpgroups drift into phase, so those cycles
// are not covered by the other warp's MUFU work.  Here the warp covers them itself:
//   * the first 32 scores of tile j+1 are pulled into spare registers while tile j's last chunk is exponentiated, and the
//     other 96 are loaded under the first chunk's exponentials (no exposed tcgen05.ld / s_full poll);
//   * the first chunk is exponentiated into registers BEFORE the wait for P V(j-1) (the P buffer), which hides that wait.
// =====================================================================================================================
// POLY: every POLY-th exponential of a full tile is evaluated by ex2_poly instead of MUFU.EX2 (0 = none).
template <int POLY>
__global__ void __launch_bounds__(TA_THREADS, 1)
attention_tc_p_kernel(const __grid_constant__ CUtensorMap tmQa, const __grid_constant__ CUtensorMap tmQb, const __grid_constant__ CUtensorMap tmKa,
                    const __grid_constant__ CUtensorMap tmKb, const __grid_constant__ CUtensorMap tmV, __half* __restrict__ out, int H, int Nq, int Nk,
                    int Dh, float scale_log2, long long* __restrict__ dbg, unsigned stagger_cycles, int flags) {
    extern __shared__ __align__(1024) uint8_t ta_smem_raw[];
    // 1024-B alignment by pointer arithmetic on the __shared__ array (an integer round trip would demote every later
    // access to generic LD/ST instead of LDS/STS)
    uint8_t* smem = ta_smem_raw + ((1024u - (smem_u32(ta_smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TA_OFF_BAR);
    uint64_t* q_full = bars;            // 1
    uint64_t* kv_full = bars + 1;       // TA_KV_STAGES
    uint64_t* kv_empty = bars + 4;      // TA_KV_STAGES
    uint64_t* s_full = bars + 7;        // 2 (per query tile)
    uint64_t* s_free = bars + 9;        // 2
    uint64_t* p_full = bars + 11;       // 2
    uint64_t* o_full = bars + 13;       // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (2 * TA_BQ);
    const int h = blockIdx.y, b = blockIdx.z;
    const int bh = b * H + h;
    const int nkt = (Nk + TA_BKV - 1) / TA_BKV;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQa); tma_prefetch_desc(&tmQb); tma_prefetch_desc(&tmKa); tma_prefetch_desc(&tmKb); tma_prefetch_desc(&tmV);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < TA_KV_STAGES; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_empty[i], 2);      // one commit from each of the two MMA issuer warps
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_free[i], 128);
            mbar_init(&p_full[i], 128);
            mbar_init(&o_full[i], 1);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();
    // TMEM columns: S_A [0,128) S_B [128,256) O_A [256,336) O_B [384,464)

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(q_full, 2 * TA_Q_BYTES);
            for (int X = 0; X < 2; ++X) {
                uint8_t* qs = smem + TA_OFF_Q + X * TA_Q_BYTES;
                const int row = bh * Nq + q0 + X * TA_BQ;
                tma_load_2d(qs, &tmQa, q_full, 0, row);
                tma_load_2d(qs + 128 * 128, &tmQb, q_full, 64, row);
            }
        }
        __syncwarp();
        for (int j = 0; j < nkt; ++j) {
            const int s = j % TA_KV_STAGES;
            mbar_wait(&kv_empty[s], ((j / TA_KV_STAGES) & 1) ^ 1);
            if (elect_one()) {
                uint8_t* ks = smem + TA_OFF_KV + s * TA_KV_STAGE;
                uint8_t* vs = ks + TA_K_BYTES;
                mbar_arrive_expect_tx(&kv_full[s], TA_KV_STAGE);
                const int krow = bh * Nk + j * TA_BKV;
                tma_load_2d(ks, &tmKa, &kv_full[s], 0, krow);
                tma_load_2d(ks + 128 * 128, &tmKb, &kv_full[s], 64, krow);
                tma_load_2d(vs, &tmV, &kv_full[s], j * TA_BKV, bh * TA_DHP);
                tma_load_2d(vs + TA_VBOX_BYTES, &tmV, &kv_full[s], j * TA_BKV + 64, bh * TA_DHP);
            }
            __syncwarp();
        }
    } else if (warp == 1 || warp == 3) {
        // Two warp-converged issuer warps (uniform control flow, one elected lane issues; descriptors stay in uniform
        // registers): warp 1 drives query tile A, warp 3 tile B, so the two softmax warpgroups run as independent pipelines
        // and can de-phase instead of being re-synchronised by one in-order issuer.
        const int X = warp == 1 ? 0 : 1;
        constexpr uint32_t idesc_qk = umma_idesc_f16(128, 128);
        constexpr uint32_t idesc_pv = umma_idesc_f16(128, TA_DHP);
        const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
        const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
        auto issue_qk = [&](int X, int s) {
            const uint32_t qa = sbase + TA_OFF_Q + X * TA_Q_BYTES, ka = sbase + TA_OFF_KV + s * TA_KV_STAGE;
            const uint64_t dq = umma_desc_kmajor<128>(qa), dk = umma_desc_kmajor<128>(ka);
            const uint64_t dq2 = umma_desc_kmajor<32>(qa + 128 * 128), dk2 = umma_desc_kmajor<32>(ka + 128 * 128);
            const uint32_t ts = tbase + X * 128;
            if (elect_one()) {
#pragma unroll
                for (int i = 0; i < 4; ++i) umma_f16(ts, dq + 2 * i, dk + 2 * i, idesc_qk, i > 0 ? 1u : 0u);
                umma_f16(ts, dq2, dk2, idesc_qk, 1u);
                umma_commit(&s_full[X]);
            }
            __syncwarp();
        };
        auto issue_pv = [&](int X, int s, uint32_t acc_first) {
            const uint32_t pa = sbase + TA_OFF_P + X * TA_P_BYTES, va = sbase + TA_OFF_KV + s * TA_KV_STAGE + TA_K_BYTES;
            const uint32_t to = tbase + 256 + X * 128;
            const uint64_t dp0 = umma_desc_kmajor<128>(pa), dv0 = umma_desc_kmajor<128>(va);
            const uint64_t dp1 = umma_desc_kmajor<128>(pa + 128 * 128), dv1 = umma_desc_kmajor<128>(va + TA_VBOX_BYTES);
            if (elect_one()) {
#pragma unroll
                for (int i = 0; i < 4; ++i) umma_f16(to, dp0 + 2 * i, dv0 + 2 * i, idesc_pv, i != 0 ? 1u : acc_first);
#pragma unroll
                for (int i = 0; i < 4; ++i) umma_f16(to, dp1 + 2 * i, dv1 + 2 * i, idesc_pv, 1u);
                umma_commit(&o_full[X]);
            }
            __syncwarp();
        };
        // Row sums on the tensor core: row Dh of every V^T tile (a padding row, Dh < 80) is overwritten with ones after the TMA
        // lands, so column Dh of O accumulates sum_k P[r,k] — rescaled together with O — and the softmax threads carry no sum.
        // Both issuer warps write the same 2 x 128 bytes (row 72 is swizzle row 0: identity chunk order).
        auto write_ones = [&](int s) {
            const uint32_t a = sbase + TA_OFF_KV + s * TA_KV_STAGE + TA_K_BYTES + Dh * 128 + lane * 4;
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(0x3C003C00u) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(a + TA_VBOX_BYTES), "r"(0x3C003C00u) : "memory");
            fence_proxy_async();
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        mbar_wait(&kv_full[0], 0);
        write_ones(0);
        tc_fence_after();
        issue_qk(X, 0);
        for (int j = 0; j < nkt; ++j) {
            const int s = j % TA_KV_STAGES;
            if (j + 1 < nkt) {           // next scores first: the warpgroup gets S(j+1) while it exponentiates tile j
                const int sn = (j + 1) % TA_KV_STAGES;
                mbar_wait(&kv_full[sn], ((j + 1) / TA_KV_STAGES) & 1);
                write_ones(sn);
                mbar_wait(&s_free[X], j & 1);
                tc_fence_after();
                issue_qk(X, sn);
            }
            mbar_wait(&p_full[X], j & 1);
            tc_fence_after();
            issue_pv(X, s, j > 0 ? 1u : 0u);      // O_X accumulates in TMEM across key tiles
            if (elect_one()) umma_commit(&kv_empty[s]);
            __syncwarp();
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        const int X = (warp - 4) >> 2;               // query tile of this warpgroup
        const int quad = warp & 3;
        const int r = quad * 32 + lane;              // row in the tile == TMEM lane
        const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t tS = tmem_base + X * 128 + lane_off;
        const uint32_t tO = tmem_base + 256 + X * 128 + lane_off;
        const uint32_t pS = smem_u32(smem) + TA_OFF_P + X * TA_P_BYTES + r * 128;   // shared-window address of this row of P
        const int sw = r & 7;
        float m_ref = 0.f;
        const bool stale_max = (flags & 1) != 0;
        auto rescale_o = [&](bool need, float mx) {
            const float a = need ? ex2((m_ref - mx) * scale_log2) : 1.0f;
            if (need) m_ref = mx;
            uint32_t t[32];
            tmem_ld_32x32(tO, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * a);
            tmem_st_32x32(tO, t);
            tmem_ld_32x32(tO + 32, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * a);
            tmem_st_32x32(tO + 32, t);
            tmem_ld_32x16(tO + 64, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * a);
            tmem_st_32x16(tO + 64, t);
            tmem_st_wait();
        };
        // exponentials of one 32-score chunk against msc -> 16 packed fp16 pairs; optional running max (two chains)
        auto exp_chunk = [&](const uint32_t* sc, int c, float msc, uint32_t (&pk)[16], bool track, float& m0, float& m1) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float s0 = __uint_as_float(sc[i]), s1 = __uint_as_float(sc[i + 1]);
                const float s2 = __uint_as_float(sc[i + 2]), s3 = __uint_as_float(sc[i + 3]);
                if (track) { m0 = fmaxf(m0, fmaxf(s0, s2)); m1 = fmaxf(m1, fmaxf(s1, s3)); }
                const float p0 = (POLY > 0 && (c * 32 + i) % POLY == 0) ? ex2_poly(fmaf(s0, scale_log2, -msc)) : ex2(fmaf(s0, scale_log2, -msc));
                const float p1 = (POLY > 0 && (c * 32 + i + 1) % POLY == 0) ? ex2_poly(fmaf(s1, scale_log2, -msc)) : ex2(fmaf(s1, scale_log2, -msc));
                const float p2 = (POLY > 0 && (c * 32 + i + 2) % POLY == 0) ? ex2_poly(fmaf(s2, scale_log2, -msc)) : ex2(fmaf(s2, scale_log2, -msc));
                const float p3 = (POLY > 0 && (c * 32 + i + 3) % POLY == 0) ? ex2_poly(fmaf(s3, scale_log2, -msc)) : ex2(fmaf(s3, scale_log2, -msc));
                __half2 ha = __floats2half2_rn(p0, p1), hb = __floats2half2_rn(p2, p3);
                pk[i >> 1] = *reinterpret_cast<uint32_t*>(&ha);
                pk[(i >> 1) + 1] = *reinterpret_cast<uint32_t*>(&hb);
            }
        };
        auto store_chunk = [&](int c, const uint32_t (&pk)[16]) {
            const uint32_t dst = pS + (c >> 1) * (128 * 128);
            const int cc0 = (c & 1) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((cc0 + q) ^ sw) << 4)), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                             "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                             : "memory");
        };

        // prologue: the first 32 scores of tile 0
        uint32_t nxt[32];
        mbar_wait(&s_full[X], 0);
        tc_fence_after();
        if (X == 1 && stagger_cycles != 0) {
            const long long t_end = clock64() + stagger_cycles;
            while (clock64() < t_end) {}
        }
        tmem_ld_32x32(tS, nxt);
        tmem_ld_wait();
        for (int j = 0; j < nkt; ++j) {
            const int nvalid = Nk - j * TA_BKV;      // keys of this tile that exist (>= 1)
            uint32_t sv[128];
#pragma unroll
            for (int i = 0; i < 32; ++i) sv[i] = nxt[i];
            if (r == 0) TA_DBG(X, 2);
            // the other 96 scores load under the first chunk's exponentials
            tmem_ld_32x32(tS + 32, reinterpret_cast<uint32_t(&)[32]>(sv[32]));
            tmem_ld_32x32(tS + 64, reinterpret_cast<uint32_t(&)[32]>(sv[64]));
            tmem_ld_32x32(tS + 96, reinterpret_cast<uint32_t(&)[32]>(sv[96]));
            const bool fast = j > 0 && nvalid >= TA_BKV && stale_max;
            if (fast) {
                const float msc = m_ref * scale_log2;
                float m0 = -INFINITY, m1 = -INFINITY;
                uint32_t pk[16];
                exp_chunk(&sv[0], 0, msc, pk, true, m0, m1);                 // into registers: the P buffer may still be read by P V(j-1)
                mbar_wait(&o_full[X], (j - 1) & 1);                          // P V(j-1) retired: P_X may be rewritten, O_X is quiescent
                tc_fence_after();
                if (r == 0) TA_DBG(X, 5);
                store_chunk(0, pk);
                tmem_ld_wait();                                              // chunks 1..3 have been in flight for a whole chunk of exponentials
                tc_fence_before();
                mbar_arrive(&s_free[X]);                                     // the score row is in registers: S_X may be overwritten by Q K^T(j+1)
                if (r == 0) TA_DBG(X, 3);
                exp_chunk(&sv[32], 1, msc, pk, true, m0, m1);
                store_chunk(1, pk);
                exp_chunk(&sv[64], 2, msc, pk, true, m0, m1);
                store_chunk(2, pk);
                if (j + 1 < nkt) {                                           // Q K^T(j+1) was issued at s_free: normally complete by now
                    mbar_wait(&s_full[X], (j + 1) & 1);
                    tc_fence_after();
                    tmem_ld_32x32(tS, nxt);                                  // first chunk of the next tile, lands under chunk 3's exponentials
                }
                exp_chunk(&sv[96], 3, msc, pk, true, m0, m1);
                store_chunk(3, pk);
                const float mx = fmaxf(m0, m1);
                const bool need = (mx - m_ref) * scale_log2 > 8.0f;
                if (__any_sync(0xffffffffu, need)) {             // rare: redo this tile against the new maximum
                    tmem_ld_wait();
                    rescale_o(need, mx);
                    const float msc2 = m_ref * scale_log2;
                    float d0 = 0.f, d1 = 0.f;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        exp_chunk(&sv[c * 32], c, msc2, pk, false, d0, d1);
                        store_chunk(c, pk);
                    }
                }
            } else {
                // first tile, ragged last tile (or stale_max off): maximum first, then the exponentials
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&s_free[X]);
                if (r == 0) TA_DBG(X, 3);
                float mx = -INFINITY;
#pragma unroll
                for (int i = 0; i < 128; ++i)
                    if (i < nvalid) mx = fmaxf(mx, __uint_as_float(sv[i]));
                if (j > 0) {
                    mbar_wait(&o_full[X], (j - 1) & 1);          // P V(j-1) retired: O_X quiescent, P_X free
                    tc_fence_after();
                }
                if (j == 0) {
                    m_ref = mx;
                } else {
                    const bool need = (mx - m_ref) * scale_log2 > 8.0f;
                    if (__any_sync(0xffffffffu, need)) rescale_o(need, mx);
                }
                const float msc = m_ref * scale_log2;
                if (r == 0) TA_DBG(X, 5);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float p0 = ex2(fmaf(__uint_as_float(sv[c * 32 + i]), scale_log2, -msc));
                        float p1 = ex2(fmaf(__uint_as_float(sv[c * 32 + i + 1]), scale_log2, -msc));
                        if (c * 32 + i >= nvalid) p0 = 0.f;
                        if (c * 32 + i + 1 >= nvalid) p1 = 0.f;
                        __half2 hh = __floats2half2_rn(p0, p1);
                        pk[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
                    }
                    store_chunk(c, pk);
                }
                if (j + 1 < nkt) {
                    mbar_wait(&s_full[X], (j + 1) & 1);
                    tc_fence_after();
                    tmem_ld_32x32(tS, nxt);
                }
            }
            if (r == 0) TA_DBG(X, 6);
            fence_proxy_async();            // make the P stores visible to the tensor core (async proxy)
            tc_fence_before();
            mbar_arrive(&p_full[X]);
            tmem_ld_wait();                 // nxt (first chunk of tile j+1) is valid
            if (r == 0) TA_DBG(X, 7);
            if (r == 0) TA_DBG(X, 1);
        }
        mbar_wait(&o_full[X], (nkt - 1) & 1);
        tc_fence_after();
        const int row = q0 + X * TA_BQ + r;
        __half* orow = out + (static_cast<size_t>(b) * Nq + (row < Nq ? row : 0)) * (H * Dh) + h * Dh;
        uint32_t t2[32];
        tmem_ld_32x16(tO + 64, t2);                  // columns 64..79: the last 8 value columns and, at column Dh, the row sum
        tmem_ld_wait();
        const float inv = 1.0f / __uint_as_float((Dh == 64 ? t2[0] : t2[8]));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            uint32_t t[32];
            if (c < 2) {
                tmem_ld_32x32(tO + c * 32, t);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) t[i] = t2[i];
            }
#pragma unroll
            for (int d8 = 0; d8 < (c < 2 ? 32 : 16); d8 += 8) {
                const int d = c * 32 + d8;
                if (row < Nq && d < Dh) {
                    Pack8 v;
#pragma unroll
                    for (int i = 0; i < 8; ++i) v.h[i] = __float2half_rn(__uint_as_float(t[d8 + i]) * inv);
                    *reinterpret_cast<uint4*>(orow + d) = v.u;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}


}  // namespace

int launch_attention_tc(const __half* q, const __half* k, const __half* vT, __half* out, int B, int H, int Nq, int Nk, int NkPad, int Dh, float scale,
                        cudaStream_t st, long long* dbg) {
    TPX_CHECK(B > 0 && H > 0 && Nq > 0 && Nk > 0, TPX_ERR_SHAPE, "attention_tc: empty problem");
    TPX_CHECK(Dh == 72 || Dh == 64, TPX_ERR_SHAPE, "attention_tc: head dim %d (kernel covers Dh = 72 and 64: 80-wide tiles whose first padding row carries the row sums)", Dh);
    TPX_CHECK(NkPad % 8 == 0 && NkPad >= Nk, TPX_ERR_SHAPE, "attention_tc: NkPad %d must be a multiple of 8 and >= Nk %d", NkPad, Nk);
    TPX_CHECK(H <= 65535 && B <= 65535, TPX_ERR_SHAPE, "attention_tc: grid too large");
    CUtensorMap mQa, mQb, mKa, mKb, mV;
    int rc;
    const long long qrows = static_cast<long long>(B) * H * Nq, krows = static_cast<long long>(B) * H * Nk;
    if ((rc = make_tensor_map_2d(q, qrows, TA_DHP, TA_DHP, 128, 64, &mQa)) != TPX_OK) return rc;
    if ((rc = make_tensor_map_2d(q, qrows, TA_DHP, TA_DHP, 128, 16, &mQb)) != TPX_OK) return rc;
    if ((rc = make_tensor_map_2d(k, krows, TA_DHP, TA_DHP, 128, 64, &mKa)) != TPX_OK) return rc;
    if ((rc = make_tensor_map_2d(k, krows, TA_DHP, TA_DHP, 128, 16, &mKb)) != TPX_OK) return rc;
    if ((rc = make_tensor_map_2d(vT, static_cast<long long>(B) * H * TA_DHP, NkPad, NkPad, TA_DHP, 64, &mV)) != TPX_OK) return rc;
    static const int poly = getenv("TPX_ATT_POLY") ? atoi(getenv("TPX_ATT_POLY")) : 0;   // 4: every 4th exponential as a cubic on the FMA pipe
    static const int variant = getenv("TPX_ATT_VARIANT") ? atoi(getenv("TPX_ATT_VARIANT")) : 1;   // 1: software-pipelined softmax warps (default); 0: the plain loop
    static bool attr_set = false;
    if (!attr_set) {
        TPX_CUDA(cudaFuncSetAttribute(attention_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM));
        TPX_CUDA(cudaFuncSetAttribute(attention_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM));
        TPX_CUDA(cudaFuncSetAttribute(attention_tc_p_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM));
        TPX_CUDA(cudaFuncSetAttribute(attention_tc_p_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM));
        attr_set = true;
    }
    ProfScope prof(PROF_ATTENTION, st);
    static const unsigned stagger = getenv("TPX_ATT_STAGGER") ? static_cast<unsigned>(atoi(getenv("TPX_ATT_STAGGER"))) : 1600u;   // cycles; see the de-phasing note in the kernel
    static const int stale_max = getenv("TPX_ATT_STALE_MAX") ? atoi(getenv("TPX_ATT_STALE_MAX")) : 1;   // 0: always reduce the maximum first
    dim3 grid((Nq + 2 * TA_BQ - 1) / (2 * TA_BQ), H, B);
    auto kern = variant == 0 ? (poly == 0 ? attention_tc_kernel<0> : attention_tc_kernel<4>) : (poly == 0 ? attention_tc_p_kernel<0> : attention_tc_p_kernel<4>);
    TPX_CUDA(launch_pdl(kern, grid, dim3(TA_THREADS), TA_SMEM, st, mQa, mQb, mKa, mKb, mV, out, H, Nq, Nk, Dh, scale * 1.4426950408889634f, dbg, stagger,
                        stale_max ? 1 : 0));
    TPX_LAUNCH_CHECK();
    return TPX_OK;
}

}  // namespace tpx
