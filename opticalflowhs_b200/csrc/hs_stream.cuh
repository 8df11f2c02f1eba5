// hs_stream.cuh -- temporally blocked Horn-Schunck iteration for sm_100a (B200).
//
// Replaces T consecutive runCLKernels() calls (HSOpticalFlowOpenCL.cpp:476-679: H2D u,v ->
// u_v_avgKernel -> u_v_updateKernel -> D2H u,v; Kernels.cl:43-90) by ONE launch that reads
// u, v and the three coefficient planes once and writes u, v once.
//
// Design (B200-first, HBM-bound fp32 stencil -- no tensor cores on purpose):
//  * work unit = one WARP x (128-column strip, row chunk, frame pair).  Warps are autonomous:
//    own shared-memory rings, own mbarriers, no __syncthreads anywhere; one warp per CTA by default.
//  * the fp32 planes are row-interleaved in HBM ([row][u|v][pitch], [row][a|b|c][pitch]), so ONE TMA
//    operation (cp.async.bulk.tensor.4d, box = 128 columns x all planes x 2 rows, issued by one
//    elected lane) refills a ring slot.  Out-of-image columns are zero-filled by the TMA unit and
//    re-clamped in registers (Neumann border, Tex2D Kernels.cl:2-9): no bounds checks on loads.
//  * the warp streams DOWN the rows.  Each lane owns 4 adjacent columns.  Time step s+1 of row
//    r-1 is produced as soon as time step s of row r exists, so T time steps are in flight as a
//    register pipeline: per stage and field only two partial sums per pixel are kept
//    (p = G(r-1) + 2h(r), g = G(r); see hs_common.cuh), 16 registers per stage.
//  * left/right neighbours come from warp shuffles (one __shfl_up + one __shfl_down per field
//    and stage-row); the strip carries a halo of HL >= T columns on each side that absorbs the
//    shrinking valid region, the chunk carries T warm-up rows above and below.
//  * three code paths: a generic tick with run-time predicates (pipeline fill, bottom-edge drain,
//    tiny frames) and two branch-free steady-state loops (interior strips / strips touching the
//    left or right image edge), two ticks = one TMA row group per trip, ring positions and
//    barrier phases kept incrementally.
//  * results leave through coalesced 16-byte stores of the strip's valid columns.
// Per pixel-iteration: 14 FP32 instructions, 1 shuffle, 12 B of shared-memory reads; HBM
// traffic 28 B / T per pixel-iteration (+ halo overhead; measured 29 B per pixel and launch at T=4).
// tests/stream_model.py is the numpy model of exactly this bookkeeping.
#include <type_traits>
#include <utility>

#pragma once
#include "hs_common.cuh"
#include "hs_launch.h"

namespace hs {

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1u << 22)) __trap();   // a lost TMA transaction must fail loudly, never hang the GPU
    } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int x, int y, int z, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(z), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int x, int pl, int y, int z, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(pl), "r"(y), "r"(z), "r"(bar)
        : "memory");
}

// ---- geometry ------------------------------------------------------------------------------------
template <int T, int RG, int NGC, int NGUV> struct StreamCfg {
    static constexpr int HL = (T + 3) / 4 * 4;              // column halo per side, multiple of 4
    static constexpr int VALIDW = kStripW - 2 * HL;         // columns a strip produces
    static constexpr int NRC = NGC * RG;                    // coefficient ring rows
    static constexpr int NRUV = NGUV * RG;                  // u/v ring rows
    static constexpr int ROWB = kStripW * 4;                // bytes per ring row
    static constexpr int SMEM_WARP = (3 * NRC + 2 * NRUV) * ROWB + 128;   // coefficient ring + u/v ring + mbarriers
    // a coefficient group is refilled at the end of the RG-tick body that retires it; it must
    // have been issued at least one body before it is needed
    static_assert((NGC - 1) * RG >= T + 2, "coefficient ring too short for the stage lag");
    static_assert(NGC + NGUV <= 16, "barrier block is 128 bytes");
    static_assert(RG == 2, "the steady-state body is written for 2-row TMA boxes");
};
template <int T> struct DefaultCfg {
    static constexpr int RG = kStreamRowsPerBox;
    static constexpr int NGC = T <= 4 ? 4 : 6;              // 8 or 12 coefficient rows
    static constexpr int NGUV = 2;                          // 4 u/v rows
    // Main loop of interior strips unrolled over one whole coefficient-ring period (all shared-memory offsets become
    // immediates: 56 instead of 84 instructions per stage-row at T = 4).  Measured SLOWER on B200 (T = 4: 731 k vs
    // 774 k Mpixel-iterations/s sustained, T = 8: 566 k vs 773 k): the kernel is bound by dependency latency at two
    // warps per scheduler, not by issue slots, and the 29-79 KB loop bodies fall out of the instruction caches.
    static constexpr bool PERIOD_UNROLL = false;
    using type = StreamCfg<T, RG, NGC, NGUV>;
};

// ring index arithmetic: power-of-two rings wrap with one AND
template <int SIZE_BYTES> __device__ __forceinline__ int wrap_down(int off) {   // off in (-SIZE, SIZE)
    if ((SIZE_BYTES & (SIZE_BYTES - 1)) == 0) return off & (SIZE_BYTES - 1);
    return off < 0 ? off + SIZE_BYTES : off;
}
template <int SIZE_BYTES> __device__ __forceinline__ int wrap_up(int off) {     // off in [0, 2*SIZE)
    if ((SIZE_BYTES & (SIZE_BYTES - 1)) == 0) return off & (SIZE_BYTES - 1);
    return off >= SIZE_BYTES ? off - SIZE_BYTES : off;
}

// ---- the kernel -----------------------------------------------------------------------------------
#ifndef HS_STREAM_MIN_CTAS
#define HS_STREAM_MIN_CTAS 1     // experiments: 3 caps the kernel at 168 registers (12 warps per SM)
#endif
template <int T, int ST, bool PEER>
__global__ void __launch_bounds__(128, HS_STREAM_MIN_CTAS)
k_jacobi_stream(const __grid_constant__ CUtensorMap tm_uv, const __grid_constant__ CUtensorMap tm_c, const StreamArgs A) {
    using C = typename DefaultCfg<T>::type;
    constexpr int RG = DefaultCfg<T>::RG, NGC = DefaultCfg<T>::NGC, NGUV = DefaultCfg<T>::NGUV;
    constexpr int NRC = C::NRC, NRUV = C::NRUV, ROWB = C::ROWB;
    // rings are row-interleaved like the planes in HBM: a|b|c of one row are adjacent (CROW bytes), u|v likewise
    constexpr int CROW = 3 * ROWB, UROW = 2 * ROWB;
    constexpr int CB = NRC * CROW, UB = NRUV * UROW;        // ring sizes in bytes
    extern __shared__ __align__(128) uint8_t smem_raw[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long unit = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (unit >= A.total_units) return;
    const int sx = (int)(unit % A.nsx);
    const long long tt = unit / A.nsx;
    const int cy = (int)(tt % A.ncy);
    const int z = (int)(tt / A.ncy);
    const int W = A.W, H = A.H;
    const int R0 = A.out_lo + cy * A.chunk_rows;
    const int R1 = min(R0 + A.chunk_rows, A.out_hi);
    const int x0 = sx * C::VALIDW - C::HL;
    const int col0 = x0 + lane * 4;
    const int rs = max(R0 - T, 0);                 // first input row streamed
    const int last_tick = R1 - 1 + T;              // tick at which row R1-1 of time T is produced
    const int last_in = min(last_tick, H - 1);     // last real input row
    const int g0 = rs / RG, glast = last_in / RG;  // TMA row groups (absolute, RG-aligned)
    // The steady state starts at the first group boundary after the T fill ticks (group gs).  The rings are
    // addressed from a VIRTUAL origin g0v <= g0 chosen so that gs falls on ring slot 0: the main loop then
    // covers one whole ring period per trip with every shared-memory offset a compile-time constant.
    const int fill_end = (rs + T + RG - 1) / RG * RG;
    const int gs = fill_end / RG;
    const int g0v = gs - NGC * ((gs - g0 + NGC - 1) / NGC);
    const int s0 = g0 - g0v;                       // ring slots below s0 see their first box one period later
    const int base = g0v * RG;
    const bool edge = (sx == 0) || (x0 + kStripW - 1 >= W - 1);
    const bool wmis = (W & 3) != 0;

    uint8_t* wsm = smem_raw + (size_t)warp * C::SMEM_WARP;
    // byte layout per warp: coefficient ring | u/v ring | mbarriers
    const uint8_t* sa_l = wsm + lane * 16;                  // this lane's 16-byte column group
    const uint8_t* su_l = wsm + CB + lane * 16;
    const uint32_t sa32 = smem_u32(wsm), su32 = sa32 + CB;
    const uint32_t bar0 = su32 + UB;                        // cbar[NGC] then uvbar[NGUV]

    auto issue_coef_slot = [&](int g, int slot) {  // lane 0 only: one box = RG rows x {a,b,c} x 128 columns
        const uint32_t bar = bar0 + 8u * slot;
        mbar_expect_tx(bar, (uint32_t)RG * CROW);
        tma_load_4d(sa32 + (uint32_t)slot * RG * CROW, &tm_c, x0, 0, g * RG, A.z_c0 + z, bar);
    };
    auto issue_uv_slot = [&](int g, int slot) {    // lane 0 only: one box = RG rows x {u,v} x 128 columns
        const uint32_t bar = bar0 + 8u * (NGC + slot);
        mbar_expect_tx(bar, (uint32_t)RG * UROW);
        tma_load_4d(su32 + (uint32_t)slot * RG * UROW, &tm_uv, x0, 0, g * RG, A.z_in0 + z, bar);
    };
    auto issue_coef = [&](int g) { issue_coef_slot(g, (g - g0v) % NGC); };
    auto issue_uv = [&](int g) { issue_uv_slot(g, (g - g0v) % NGUV); };

    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NGC + NGUV; ++i) mbar_init(bar0 + 8u * i, 1);
        fence_mbar_init();
        // A slot whose first box arrives in an odd virtual round gets one empty phase up front, so that
        // "parity = round & 1" holds for every slot.
#pragma unroll
        for (int k = 0; k < NGC; ++k)
            if (k < s0) mbar_arrive(bar0 + 8u * k);
#pragma unroll
        for (int k = 0; k < NGUV; ++k)
            if (s0 > k && (((s0 - k + NGUV - 1) / NGUV) & 1)) mbar_arrive(bar0 + 8u * (NGC + k));
#pragma unroll
        for (int k = 0; k < NGUV; ++k)
            if (g0 + k <= glast) issue_uv(g0 + k);
#pragma unroll
        for (int k = 0; k < NGC; ++k)
            if (g0 + k <= glast) issue_coef(g0 + k);
    }
    __syncwarp();

    // register pipeline: per stage and field two partial sums per pixel, held as packed pixel pairs
    // (index 0: u of pixels 0,1; 1: u of pixels 2,3; 2, 3: v likewise) for FFMA2 / FADD2 / FMUL2
    f32x2 p[T][4], g[T][4];
#pragma unroll
    for (int s = 0; s < T; ++s)
#pragma unroll
        for (int j = 0; j < 4; ++j) { p[s][j] = 0ull; g[s][j] = 0ull; }

    const bool lane_out = (lane >= C::HL / 4) && (lane < 32 - C::HL / 4) && (col0 < W);
    float* uo = A.u_out + (size_t)z * A.out_pair_pitch + col0;
    float* vo = A.v_out + (size_t)z * A.out_pair_pitch + col0;

    // Peer transport (row strips over several GPUs): rows a neighbour keeps as ghost rows are stored a second time,
    // straight into that neighbour's destination buffer over NVLink.  Decided per chunk: only the units at the top
    // and bottom of the strip ever take the branch.
    const bool push_up = PEER && A.peer_up != nullptr && R0 < A.up_hi && R1 > A.up_lo;
    const bool push_dn = PEER && A.peer_dn != nullptr && R0 < A.dn_hi && R1 > A.dn_lo;
    const bool push_any = push_up || push_dn;
    const long long voff = A.v_out - A.u_out;
    auto push_row = [&](const int ro, const float (&cu)[4], const float (&cv)[4]) {
        if (push_up && ro >= A.up_lo && ro < A.up_hi) {
            float* q = A.peer_up + (size_t)(ro + A.up_delta) * A.row_pitch + col0;
            *reinterpret_cast<float4*>(q) = make_float4(cu[0], cu[1], cu[2], cu[3]);
            *reinterpret_cast<float4*>(q + voff) = make_float4(cv[0], cv[1], cv[2], cv[3]);
        }
        if (push_dn && ro >= A.dn_lo && ro < A.dn_hi) {
            float* q = A.peer_dn + (size_t)(ro + A.dn_delta) * A.row_pitch + col0;
            *reinterpret_cast<float4*>(q) = make_float4(cu[0], cu[1], cu[2], cu[3]);
            *reinterpret_cast<float4*>(q + voff) = make_float4(cv[0], cv[1], cv[2], cv[3]);
        }
    };

    // time step S+1 of one row from its averages and coefficients (two packed pixel pairs per field)
    auto update_rows = [&](const f32x2 (&ub)[2], const f32x2 (&vb)[2], const float4& ka, const float4& kb, const float4& kc,
                           float (&cu)[4], float (&cv)[4]) {
        f32x2 un, vn;
        update_fast2(ub[0], vb[0], pk2(ka.x, ka.y), pk2(kb.x, kb.y), pk2(kc.x, kc.y), pk2(-ka.x, -ka.y), pk2(-kb.x, -kb.y), un, vn);
        unpk2(un, cu[0], cu[1]); unpk2(vn, cv[0], cv[1]);
        update_fast2(ub[1], vb[1], pk2(ka.z, ka.w), pk2(kb.z, kb.w), pk2(kc.z, kc.w), pk2(-ka.z, -ka.w), pk2(-kb.z, -kb.w), un, vn);
        unpk2(un, cu[2], cu[3]); unpk2(vn, cv[2], cv[3]);
    };

    // one stage-row in steady state: time step S of one row (cu, cv) -> time step S+1 of the row
    // above it, written back into cu, cv.  coff = byte offset of that row's coefficients.
    auto stage_row = [&](auto edge_tag, auto s_tag, float (&cu)[4], float (&cv)[4], const int coff) {
        constexpr bool EDGE = decltype(edge_tag)::value;
        constexpr int S = decltype(s_tag)::value;
        if (EDGE && wmis) { sanitize_right(cu, col0, W); sanitize_right(cv, col0, W); }
        float lu = __shfl_up_sync(kFull, cu[3], 1), ru = __shfl_down_sync(kFull, cu[0], 1);
        float lv = __shfl_up_sync(kFull, cv[3], 1), rv = __shfl_down_sync(kFull, cv[0], 1);
        if (EDGE) { clamp_lr(cu, col0, W, lu, ru); clamp_lr(cv, col0, W, lv, rv); }
        const float4 ka = *reinterpret_cast<const float4*>(sa_l + coff);
        const float4 kb = *reinterpret_cast<const float4*>(sa_l + ROWB + coff);
        const float4 kc = *reinterpret_cast<const float4*>(sa_l + 2 * ROWB + coff);
        f32x2 hu[2], hv[2], ub[2], vb[2];
        hsum4(cu, lu, ru, hu);
        hsum4(cv, lv, rv, hv);
        const f32x2 c2u[2] = {pk2(cu[0], cu[1]), pk2(cu[2], cu[3])}, c2v[2] = {pk2(cv[0], cv[1]), pk2(cv[2], cv[3])};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const f32x2 Gu = rowG2<ST>(c2u[j], hu[j]), Gv = rowG2<ST>(c2v[j], hv[j]);
            ub[j] = combine2<ST>(p[S][j], Gu);
            vb[j] = combine2<ST>(p[S][2 + j], Gv);
            p[S][j] = pOf2<ST>(g[S][j], hu[j]); g[S][j] = Gu;
            p[S][2 + j] = pOf2<ST>(g[S][2 + j], hv[j]); g[S][2 + j] = Gv;
        }
        update_rows(ub, vb, ka, kb, kc, cu, cv);
    };

    // ---- generic tick: pipeline fill, bottom-edge drain, tiny frames (runtime predicates) ----------
    auto tick_gen = [&](const int r) {
        const int rr = r - base;                   // row relative to the first TMA group
        float cu[4], cv[4];
        bool have = false;
        if (r <= H - 1) {
            if ((rr % RG) == 0 || r == rs) {       // first row consumed from this group
                const int gr = rr / RG;
                mbar_wait(bar0 + 8u * (NGC + gr % NGUV), (gr / NGUV) & 1);
                mbar_wait(bar0 + 8u * (gr % NGC), (gr / NGC) & 1);
            }
            const int q = (rr % NRUV) * UROW;
            const float4 tu = *reinterpret_cast<const float4*>(su_l + q);
            const float4 tv = *reinterpret_cast<const float4*>(su_l + ROWB + q);
            cu[0] = tu.x; cu[1] = tu.y; cu[2] = tu.z; cu[3] = tu.w;
            cv[0] = tv.x; cv[1] = tv.y; cv[2] = tv.z; cv[3] = tv.w;
            have = true;
        }
#pragma unroll
        for (int S = 0; S < T; ++S) {
            const int rho = r - S;                 // row of time step S this stage receives
            bool virt = false;
            if (rho < rs || rho > H) have = false;
            else if (rho == H) { virt = true; have = true; }   // row H == row H-1 (clamp)
            if (!have) continue;
            f32x2 ub[2], vb[2];
            if (virt) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    ub[j] = combine2<ST>(p[S][j], g[S][j]);
                    vb[j] = combine2<ST>(p[S][2 + j], g[S][2 + j]);
                }
            } else {
                if (edge && wmis) { sanitize_right(cu, col0, W); sanitize_right(cv, col0, W); }
                float lu = __shfl_up_sync(kFull, cu[3], 1), ru = __shfl_down_sync(kFull, cu[0], 1);
                float lv = __shfl_up_sync(kFull, cv[3], 1), rv = __shfl_down_sync(kFull, cv[0], 1);
                if (edge) { clamp_lr(cu, col0, W, lu, ru); clamp_lr(cv, col0, W, lv, rv); }
                f32x2 hu[2], hv[2];
                hsum4(cu, lu, ru, hu);
                hsum4(cv, lv, rv, hv);
                const f32x2 c2u[2] = {pk2(cu[0], cu[1]), pk2(cu[2], cu[3])}, c2v[2] = {pk2(cv[0], cv[1]), pk2(cv[2], cv[3])};
                if (rho == rs) {                   // first row of this stage: replicate upwards
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const f32x2 Gu = rowG2<ST>(c2u[j], hu[j]), Gv = rowG2<ST>(c2v[j], hv[j]);
                        p[S][j] = pOf2<ST>(Gu, hu[j]); g[S][j] = Gu;
                        p[S][2 + j] = pOf2<ST>(Gv, hv[j]); g[S][2 + j] = Gv;
                    }
                    have = false;
                    continue;
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const f32x2 Gu = rowG2<ST>(c2u[j], hu[j]), Gv = rowG2<ST>(c2v[j], hv[j]);
                    ub[j] = combine2<ST>(p[S][j], Gu);
                    vb[j] = combine2<ST>(p[S][2 + j], Gv);
                    p[S][j] = pOf2<ST>(g[S][j], hu[j]); g[S][j] = Gu;
                    p[S][2 + j] = pOf2<ST>(g[S][2 + j], hv[j]); g[S][2 + j] = Gv;
                }
            }
            // time step S+1 of row rho-1 (coefficients of that row)
            const int q = ((rr - S - 1 + 2 * NRC) % NRC) * CROW;
            const float4 ka = *reinterpret_cast<const float4*>(sa_l + q);
            const float4 kb = *reinterpret_cast<const float4*>(sa_l + ROWB + q);
            const float4 kc = *reinterpret_cast<const float4*>(sa_l + 2 * ROWB + q);
            update_rows(ub, vb, ka, kb, kc, cu, cv);
        }
        const int ro = r - T;
        if (have && lane_out && ro >= R0 && ro < R1) {
            const size_t o = (size_t)ro * A.row_pitch;
            *reinterpret_cast<float4*>(uo + o) = make_float4(cu[0], cu[1], cu[2], cu[3]);
            *reinterpret_cast<float4*>(vo + o) = make_float4(cv[0], cv[1], cv[2], cv[3]);
            if constexpr (PEER) { if (push_any) push_row(ro, cu, cv); }
        }
        // Ring refills, after every lane consumed its shared-memory reads of this tick:
        //  * the u/v group whose last row was read by stage 0 in this tick,
        //  * the coefficient group whose last row (r-T) was used by the last stage in this tick.
        const int qq = r - T - g0 * RG;            // rows of groups before g0 were never loaded
        const bool refill_uv = (r <= H - 1) && (rr % RG) == RG - 1;
        const bool refill_c = qq >= 0 && (qq % RG) == RG - 1;
        if (refill_uv || refill_c) {
            __syncwarp();
            if (lane == 0) {
                if (refill_uv) { const int gn = g0v + rr / RG + NGUV; if (gn <= glast) issue_uv(gn); }
                if (refill_c) { const int gn = g0 + qq / RG + NGC; if (gn <= glast) issue_coef(gn); }
            }
        }
    };

    // ---- steady state -------------------------------------------------------------------------------
    // Both loops run with all T stages active and no run-time predicates on rows.  Output row of tick r is
    // r - T; it is stored when 0 <= r - T - R0 < R1 - R0 (one unsigned compare, lanes of the halo excluded).
    const unsigned out_rows = lane_out ? (unsigned)(R1 - R0) : 0u;
    float* uo_row = uo;                                         // advanced by one row per tick
    float* vo_row = vo;
    auto tick_body = [&](auto edge_tag, const int ubyte, auto coff_of, const int ro) {
        float cu[4], cv[4];
        const float4 tu = *reinterpret_cast<const float4*>(su_l + ubyte);
        const float4 tv = *reinterpret_cast<const float4*>(su_l + ubyte + ROWB);
        cu[0] = tu.x; cu[1] = tu.y; cu[2] = tu.z; cu[3] = tu.w;
        cv[0] = tv.x; cv[1] = tv.y; cv[2] = tv.z; cv[3] = tv.w;
        // stage S consumes row (r-S) and needs the coefficients of row (r-S-1)
        [&]<int... S>(std::integer_sequence<int, S...>) {
            (stage_row(edge_tag, std::integral_constant<int, S>{}, cu, cv, coff_of(S + 1)), ...);
        }(std::make_integer_sequence<int, T>{});
        if ((unsigned)(ro - R0) < out_rows) {
            *reinterpret_cast<float4*>(uo_row) = make_float4(cu[0], cu[1], cu[2], cu[3]);
            *reinterpret_cast<float4*>(vo_row) = make_float4(cv[0], cv[1], cv[2], cv[3]);
            if constexpr (PEER) { if (push_any) push_row(ro, cu, cv); }
        }
        uo_row += A.row_pitch; vo_row += A.row_pitch;
    };

    // (a) main loop, strips that touch no image edge: one whole coefficient-ring period (NGC row groups = NRC
    //     ticks) per trip, fully unrolled -- ring rows, barrier slots and the u/v barrier parities are constants.
    constexpr int DRET = (T + RG - 1) / RG;                     // a coefficient group retires DRET groups after its first use
    auto steady_period = [&](int r, const int r_end) {          // (r - base) is a multiple of NRC; returns the next tick
        int grow = (r - base) / RG;                             // virtual group index, multiple of NGC
        uint32_t cpar = (grow / NGC) & 1, upar = (grow / NGUV) & 1;
        uo_row = uo + (size_t)(r - T) * A.row_pitch;
        vo_row = vo + (size_t)(r - T) * A.row_pitch;
#pragma unroll 1
        for (; r + NRC - 1 <= r_end; r += NRC) {
#pragma unroll
            for (int q = 0; q < NGC; ++q) {
                mbar_wait(bar0 + 8u * (NGC + q % NGUV), upar ^ ((q / NGUV) & 1));
                mbar_wait(bar0 + 8u * q, cpar);
#pragma unroll
                for (int j = 0; j < RG; ++j) {
                    const int cr = q * RG + j;                  // ring row of this tick (compile-time)
                    tick_body(std::false_type{}, (cr % NRUV) * UROW, [&](int lag) { return ((cr - lag + NRC) % NRC) * CROW; }, r + cr - T);
                }
                __syncwarp();
                if (lane == 0) {                                // refill the u/v slot just drained and the coefficient slot just retired
                    const int gabs = g0v + grow + q;
                    if (gabs + NGUV <= glast) issue_uv_slot(gabs + NGUV, q % NGUV);
                    if (gabs - DRET + NGC <= glast) issue_coef_slot(gabs - DRET + NGC, (q - DRET + NGC) % NGC);
                }
            }
            grow += NGC;
            cpar ^= 1;
            if ((NGC / NGUV) & 1) upar ^= 1;
        }
        return r;
    };

    // (b) rolled loop, one row group per trip: edge strips, and the groups left over by (a)
    auto steady = [&](auto edge_tag, int r, const int r_end) {   // r group-aligned; returns the next tick
        const int rr0 = r - base;
        int grow = rr0 / RG;                                   // virtual group index of the current tick
        int urow = rr0 % NRUV, crow = rr0 % NRC;              // ring row of the current tick
        int uslot = grow % NGUV, upar = (grow / NGUV) & 1, cslot = grow % NGC, cpar = (grow / NGC) & 1;
        int gfin = grow - DRET;                                // coefficient group retired when the current group ends
        uo_row = uo + (size_t)(r - T) * A.row_pitch;
        vo_row = vo + (size_t)(r - T) * A.row_pitch;
#pragma unroll 1
        for (; r + RG - 1 <= r_end; r += RG) {
            mbar_wait(bar0 + 8u * (NGC + uslot), upar);
            mbar_wait(bar0 + 8u * cslot, cpar);
#pragma unroll
            for (int j = 0; j < RG; ++j) {
                const int ur = wrap_up<NRUV>(urow + j), cr = wrap_up<NRC>(crow + j);
                tick_body(edge_tag, ur * UROW, [&](int lag) { return wrap_down<NRC>(cr - lag) * CROW; }, r + j - T);
            }
            urow = wrap_up<NRUV>(urow + RG);
            crow = wrap_up<NRC>(crow + RG);
            __syncwarp();
            if (lane == 0) {
                if (g0v + grow + NGUV <= glast) issue_uv(g0v + grow + NGUV);
                if (gfin >= s0 && g0v + gfin + NGC <= glast) issue_coef(g0v + gfin + NGC);
            }
            ++grow; ++gfin;
            if (++uslot == NGUV) { uslot = 0; upar ^= 1; }
            if (++cslot == NGC) { cslot = 0; cpar ^= 1; }
        }
        return r;
    };

    int r = rs;
    int gen_end = min(fill_end, last_tick + 1);                  // pipeline fill (and tiny frames)
    const int steady_end = min(H - 1, last_tick);                // last tick with a real input row
    for (int pass = 0; pass < 2; ++pass) {
        for (; r < gen_end; ++r) tick_gen(r);
        if (pass == 1 || r > last_tick) break;
        // whole groups only, so that the generic ticks that follow see consistent ring bookkeeping
        const int st_end = r + ((steady_end - r + 1) / RG) * RG - 1;
        if (edge) {
            r = steady(std::true_type{}, r, st_end);
        } else {
            if constexpr (DefaultCfg<T>::PERIOD_UNROLL) r = steady_period(r, st_end);
            r = steady(std::false_type{}, r, st_end);
        }
        gen_end = last_tick + 1;                                 // bottom edge / remainder: generic ticks
    }

    // Halo-exchange signal: every unit counts itself done after its stores (local and peer) are visible system-wide;
    // the last one resets the counter and publishes the epoch to both neighbours, whose streams wait on that word
    // (cuStreamWaitValue32) before they launch the next block.  No kernel ever spins on it.
    if (PEER && A.done_counter != nullptr) {
        __syncwarp();
        if (lane == 0) {
            __threadfence_system();
            const unsigned prev = atomicAdd(A.done_counter, 1u);
            if (prev == (unsigned)(A.total_units - 1)) {
                *A.done_counter = 0u;
                __threadfence_system();
                if (A.flag_up) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(A.flag_up), "r"(A.epoch) : "memory");
                if (A.flag_dn) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(A.flag_dn), "r"(A.epoch) : "memory");
            }
        }
    }
}


}  // namespace hs
