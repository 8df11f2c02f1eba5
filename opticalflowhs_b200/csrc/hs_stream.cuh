// hs_stream.cuh -- temporally blocked Horn-Schunck iteration for sm_100a (B200).
//
// Replaces T consecutive runCLKernels() calls (HSOpticalFlowOpenCL.cpp:476-679: H2D u,v ->
// u_v_avgKernel -> u_v_updateKernel -> D2H u,v; Kernels.cl:43-90) by ONE launch that reads
// u, v and the three coefficient planes once and writes u, v once.
//
// Design (B200-first, HBM-bound fp32 stencil -- no tensor cores on purpose):
//  * work unit = one WARP x (128-column strip, row chunk, frame pair).  Warps are autonomous: own shared-memory
//    rings, own mbarriers, no __syncthreads anywhere.  A CTA is exactly one warp, so the unit index is blockIdx.x and
//    the whole geometry is warp-uniform by construction (uniform registers, uniform datapath).
//  * the fp32 planes are row-interleaved in HBM ([row][u|v][pitch], [row][a|b|c][pitch]), so ONE TMA operation
//    (cp.async.bulk.tensor.4d, box = 128 columns x all planes x RG rows, issued by one elected lane) refills a ring
//    slot.  Out-of-image columns are zero-filled by the TMA unit and re-clamped in registers (Neumann border, Tex2D
//    Kernels.cl:2-9): no bounds checks on loads.
//  * the warp streams DOWN the rows.  Each lane owns 4 adjacent columns.  Time step s+1 of row r-1 is produced as
//    soon as time step s of row r exists, so T time steps are in flight as a register pipeline: per stage and
//    field only two partial sums per pixel are kept (p = G(r-1) + 2h(r), g = G(r); see hs_common.cuh), held as
//    packed pixel pairs for FFMA2 / FADD2 / FMUL2: 16 registers per stage.
//  * left/right neighbours come from warp shuffles (one __shfl_up + one __shfl_down per field and stage-row); the
//    strip carries a halo of HL >= T columns on each side that absorbs the shrinking valid region, the chunk carries
//    T warm-up rows above and below.
//  * two code paths: a generic tick with run-time predicates (pipeline fill, bottom-edge drain, tiny frames) and a
//    branch-free steady-state loop (interior strips / strips touching the left or right image edge), RG ticks = one
//    TMA row group per trip, every shared-memory address "register + immediate", ring pointers, barrier phases and
//    TMA coordinates advanced once per trip.
//  * results leave through coalesced 16-byte stores of the strip's valid columns -- into the partner buffer, into a
//    planar staging slot (host pipeline), and a second time into the neighbour strips over NVLink (PEER).
// Per stage-row (128 pixels x 1 iteration) a lane issues 49-50 instructions at T = 4..6, of which 32 are FP32 math, 4
// shuffles and 3 LDS.128; HBM traffic is 28 B / T per pixel-iteration (+ halo overhead: 29.2 B per pixel and launch
// measured at T = 6).  tests/stream_model.py is the numpy model of exactly this bookkeeping.
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
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// A lost TMA transaction must fail loudly, never hang the GPU.  The bound is WALL TIME (%globaltimer, 10 s), not a
// poll count: under a profiler's kernel replay, a sanitizer or heavy co-tenancy a healthy wait can take arbitrarily
// many polls, but never seconds.  The clock is only read on the slow path (first try_wait failed).
constexpr uint64_t kMbarTimeoutNs = 10ull * 1000 * 1000 * 1000;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = globaltimer_ns();
    while (!mbar_try_wait(bar, parity))
        if (globaltimer_ns() - t0 > kMbarTimeoutNs) __trap();
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int x, int pl, int y, int z, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(pl), "r"(y), "r"(z), "r"(bar)
        : "memory");
}

// 16 bytes of shared memory at `base + OFF` (OFF a compile-time constant: the address is "register + immediate").
// Volatile on purpose: consecutive ticks read the same coefficient rows (stage S+1 of tick J+1 uses the row stage S of
// tick J used); once every address of a trip is a constant expression the compiler would merge those loads and keep
// 12 registers per row alive across whole ticks -- measured 17-34 % slower (spills, no ILP left) than re-reading.
template <int OFF> __device__ __forceinline__ float4 lds128(uint32_t base) {
    float4 v;
    asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4 + %5];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base), "n"(OFF));
    return v;
}

// ---- geometry ------------------------------------------------------------------------------------
template <int T, int RG_, int NGC_, int NGUV_> struct StreamCfg {
    static constexpr int RG = RG_, NGC = NGC_, NGUV = NGUV_;
    // Column halo per side, a multiple of 4: the strip origin x0 = sx * VALIDW - HL is the innermost coordinate of
    // every TMA box, and the unit faults (the mbarrier never completes) unless that is 16-byte aligned -- a halo of
    // 6 for T = 5, 6 (34 instead of 35 strips on a 3840-wide frame) was tried and traps.
    static constexpr int HL = (T + 3) / 4 * 4;
    static constexpr int VALIDW = kStripW - 2 * HL;         // columns a strip produces
    static constexpr int NRC = NGC * RG;                    // coefficient ring rows
    static constexpr int NRUV = NGUV * RG;                  // u/v ring rows
    static constexpr int ROWB = kStripW * 4;                // bytes per ring row
    static constexpr int SMEM_WARP = (3 * NRC + 2 * NRUV) * ROWB + 128;   // coefficient ring + u/v ring + mbarriers
    // A coefficient group is in use for DRET + 1 trips (one trip = RG ticks = one row group) and is refilled at the
    // end of the trip that retires it; one more slot keeps a whole trip between the refill and its first use.
    static constexpr int DRET = (T + RG - 1) / RG;
    static_assert(NGC >= DRET + 2, "coefficient ring too short for the stage lag");
    static_assert(NGUV == 2, "the steady-state loop toggles between two u/v slots");
    static_assert(NGC + NGUV <= 16, "barrier block is 128 bytes");
};
// Rows per TMA box (= ticks per steady-state trip).  Taller boxes spread the per-trip overhead (barrier waits, TMA
// issue, ring bookkeeping, the register moves ptxas leaves on the loop back edge) over more stage-rows, but the
// coefficient ring needs (ceil(T/RG) + 2) x RG rows and the u/v ring 2 x RG rows per warp, eight warps have to fit
// the 227 KB of an SM, and the unrolled trip has to stay inside the 32 KB instruction cache:
//   T <= 3: 2 rows (occupancy of the shallow blocks)      T = 4: 4 rows (12 + 8 ring rows, 26 KB)
//   T = 5, 6: 3 rows (12 + 6 ring rows, 24 KB)            T = 7, 8: 2 rows (12 + 4 ring rows, 22 KB)
// -DHS_STREAM_RG_T<k>=<rows> overrides one entry for experiments.
#ifndef HS_STREAM_RG_T4
#define HS_STREAM_RG_T4 4
#endif
#ifndef HS_STREAM_RG_T5
#define HS_STREAM_RG_T5 3
#endif
#ifndef HS_STREAM_RG_T6
#define HS_STREAM_RG_T6 3
#endif
#ifndef HS_STREAM_RG_T7
#define HS_STREAM_RG_T7 2
#endif
#ifndef HS_STREAM_RG_T8
#define HS_STREAM_RG_T8 2
#endif
constexpr int stream_rows_per_box(int T) {
    return T == 4 ? HS_STREAM_RG_T4 : T == 5 ? HS_STREAM_RG_T5 : T == 6 ? HS_STREAM_RG_T6 : T == 7 ? HS_STREAM_RG_T7
         : T == 8 ? HS_STREAM_RG_T8 : 2;
}
template <int T> struct DefaultCfg {
    static constexpr int RG = stream_rows_per_box(T);
    static constexpr int NGC = (T + RG - 1) / RG + 2;
    static constexpr int NGUV = 2;
    using type = StreamCfg<T, RG, NGC, NGUV>;
};

// Experiment switch (never in the product build): -DHS_EXP_COEF_L2 makes every work unit fetch its coefficient boxes
// from the same 60 rows of pair 0, so the coefficient stream comes out of L2 instead of HBM.  The results are
// meaningless; the launch time is an UPPER BOUND for what recomputing a, b, c in-tile from the two 8-bit frames
// (12 B -> 2 B per pixel and launch) could buy, with the extra arithmetic priced at zero.
#ifdef HS_EXP_COEF_L2
#define HS_COEF_Y(y) ((y) % 60)
#define HS_COEF_Z(z) (0)
#else
#define HS_COEF_Y(y) (y)
#define HS_COEF_Z(z) (z)
#endif

// Peer store of one output row (row strips over several GPUs): rows a neighbour keeps as ghost rows are stored a second
// time, straight into that neighbour's destination buffer over NVLink.  Deliberately NOT inlined and fed through a pointer
// to the kernel's (grid-constant) argument block: inlined, the compiler hoists the address arithmetic out of the tick
// loop and keeps eight more registers alive through the steady-state loop, which at T = 6 means spills and 10 % more
// instructions for every unit of the launch.  Only the few ticks at a strip seam ever call it.
static __device__ __noinline__ void peer_push_row(const StreamArgs* __restrict__ Ap, int ro, int col0, float4 u, float4 v) {
    const long long voff = Ap->v_out - Ap->u_out;
    if (Ap->peer_up != nullptr && ro >= Ap->up_lo && ro < Ap->up_hi) {
        float* q = Ap->peer_up + (size_t)(ro + Ap->up_delta) * Ap->row_pitch + col0;
        *reinterpret_cast<float4*>(q) = u;
        *reinterpret_cast<float4*>(q + voff) = v;
    }
    if (Ap->peer_dn != nullptr && ro >= Ap->dn_lo && ro < Ap->dn_hi) {
        float* q = Ap->peer_dn + (size_t)(ro + Ap->dn_delta) * Ap->row_pitch + col0;
        *reinterpret_cast<float4*>(q) = u;
        *reinterpret_cast<float4*>(q + voff) = v;
    }
}

#ifndef HS_SCALAR_STATE_MAX_T
#define HS_SCALAR_STATE_MAX_T 7    // deepest block whose pipeline state is kept as scalar halves (see k_jacobi_stream)
#endif
#ifndef HS_TRIP_UNROLL
#define HS_TRIP_UNROLL 1           // experiments: 2 = two trips of the steady-state loop per iteration of the compiled loop
#endif
constexpr int kTripUnroll = HS_TRIP_UNROLL;

// ---- the kernel -----------------------------------------------------------------------------------
#ifndef HS_STREAM_MIN_CTAS
#define HS_STREAM_MIN_CTAS 1     // experiments: 3 caps the kernel at 168 registers (12 warps per SM)
#endif
#ifdef HS_STREAM_MAXNREG           // experiments: cap the registers directly (e.g. 224 = nine warps per SM)
#define HS_STREAM_BOUNDS __maxnreg__(HS_STREAM_MAXNREG)
#else
#define HS_STREAM_BOUNDS __launch_bounds__(32, HS_STREAM_MIN_CTAS)
#endif
template <int T, int ST, bool PEER, bool TRACK = false>
__global__ void HS_STREAM_BOUNDS
k_jacobi_stream(const __grid_constant__ CUtensorMap tm_uv, const __grid_constant__ CUtensorMap tm_c, const __grid_constant__ StreamArgs A) {
    using C = typename DefaultCfg<T>::type;
    constexpr int RG = C::RG, NGC = C::NGC, NGUV = C::NGUV, DRET = C::DRET;
    constexpr int NRC = C::NRC, NRUV = C::NRUV, ROWB = C::ROWB;
    // rings are row-interleaved like the planes in HBM: a|b|c of one row are adjacent (CROW bytes), u|v likewise
    constexpr int CROW = 3 * ROWB, UROW = 2 * ROWB;
    constexpr int CB = NRC * CROW, UB = NRUV * UROW;        // ring sizes in bytes
    extern __shared__ __align__(128) uint8_t smem_raw[];

    // One warp per CTA, always: the unit index is then a function of blockIdx alone, so the whole geometry (strip,
    // chunk, pair, row range, TMA coordinates, ring phases) is warp-uniform BY CONSTRUCTION and the compiler keeps it
    // in uniform registers -- with warps sharing a CTA it cannot prove that, and the bookkeeping competes with the
    // pipeline state for vector registers (T = 6: 255 registers with spills -> 247 without, 10 % fewer instructions).
    // Programmatic dependent launch: the next launch of the stream may start its CTAs as soon as SM resources free up,
    // run everything that does not touch memory written by this one (index arithmetic, barrier initialisation) and
    // then block in griddepcontrol.wait below until this grid has completed and flushed.  Hides the launch latency
    // and the prologue of each of the ceil(N / T) launches of a computation, and fills the tail of the last wave.
    asm volatile("griddepcontrol.launch_dependents;");
    const int lane = threadIdx.x;
    const long long unit = blockIdx.x;
    if (unit >= A.total_units) return;
    const int sx = (int)(unit % A.nsx);
    const long long tt = unit / A.nsx;
    int cy = (int)(tt % A.ncy);
    const int z = (int)(tt / A.ncy);
    if (PEER && A.seam_first && A.ncy > 2) cy = cy == 0 ? 0 : (cy == 1 ? A.ncy - 1 : cy - 1);   // seam chunks run first
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

    uint8_t* wsm = smem_raw;
    // byte layout: coefficient ring | u/v ring | mbarriers
    const uint8_t* sa_l = wsm + lane * 16;                  // this lane's 16-byte column group
    const uint8_t* su_l = wsm + CB + lane * 16;
    const uint32_t sa32 = smem_u32(wsm), su32 = sa32 + CB;
    const uint32_t bar0 = su32 + UB;                        // uvbar[NGUV] then cbar[NGC], 8 bytes each
    auto ubar_of = [&](int slot) { return bar0 + 8u * (uint32_t)slot; };
    auto cbar_of = [&](int slot) { return bar0 + 8u * (uint32_t)(NGUV + slot); };

    auto issue_coef_slot = [&](int g, int slot) {  // lane 0 only: one box = RG rows x {a,b,c} x 128 columns
        const uint32_t bar = cbar_of(slot);
        mbar_expect_tx(bar, (uint32_t)RG * CROW);
        tma_load_4d(sa32 + (uint32_t)slot * RG * CROW, &tm_c, x0, 0, HS_COEF_Y(g * RG), HS_COEF_Z(A.z_c0 + z), bar);
    };
    auto issue_uv_slot = [&](int g, int slot) {    // lane 0 only: one box = RG rows x {u,v} x 128 columns
        const uint32_t bar = ubar_of(slot);
        mbar_expect_tx(bar, (uint32_t)RG * UROW);
        tma_load_4d(su32 + (uint32_t)slot * RG * UROW, &tm_uv, x0, 0, g * RG, A.z_in0 + z, bar);
    };
    auto issue_coef = [&](int g) { issue_coef_slot(g, (g - g0v) % NGC); };
    auto issue_uv = [&](int g) { issue_uv_slot(g, (g - g0v) % NGUV); };

    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NGC + NGUV; ++i) mbar_init(bar0 + 8u * i, 1);
        fence_mbar_init();
    }
    // everything above is independent of the previous launch; from here on its results (u, v, stop words) are read
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // EPS criterion (StreamArgs::stop ... eps): `lim` = stages that really iterate; the others pass their input through
    int lim = T;
    const int zt = TRACK ? A.z_trk0 + z : 0;
    if constexpr (TRACK) {
        const int st = A.stop[zt];
        if (A.trk_mode == 0) {
            if (st) return;                        // this pair met the criterion in an earlier block
            lim = A.trk_t;
        } else {
            if (sx == 0 && cy == 0 && lane < kMaxT) A.emax_next[zt * kMaxT + lane] = 0u;
            if (st && (st >> 1) <= A.trk_base) return;
            int s = 0;                             // first sweep of the block whose max-norm fell below eps
            while (s < A.trk_t && !((double)__uint_as_float(A.emax[zt * kMaxT + s]) < A.eps)) ++s;
            if (s == A.trk_t) return;              // none: the main launch's result stands
            lim = s + 1;
            if (lane == 0) A.stop[zt] = ((A.trk_base + lim) << 1) | A.trk_dst_parity;   // every unit of the pair writes the same word
        }
    }
    if (lane == 0) {
        // A slot whose first box arrives in an odd virtual round gets one empty phase up front, so that
        // "parity = round & 1" holds for every slot.
#pragma unroll
        for (int k = 0; k < NGC; ++k)
            if (k < s0) mbar_arrive(cbar_of(k));
#pragma unroll
        for (int k = 0; k < NGUV; ++k)
            if (s0 > k && (((s0 - k + NGUV - 1) / NGUV) & 1)) mbar_arrive(ubar_of(k));
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
    // The loop-carried state is held as SCALAR halves and packed where an FFMA2 / FADD2 / FMUL2 consumes it: ptxas cannot
    // coalesce the copies of 64-bit register PAIRS across the loop back edge (alignment of the pair), which left 65 (T = 6)
    // to 99 (T = 8) register moves per trip; with scalar halves 8 remain and the trip shrinks from 895 to 838 instructions
    // (T = 6: 990 -> 1 011 k Mpixel-iterations/s sustained).  At T = 8 the scalar form needs more than 255 registers
    // (120 bytes of spills inside the loop, 1 012 -> 990 k), so the deepest block keeps the packed form.
    struct F2 {
        float lo, hi;
        __device__ __forceinline__ operator f32x2() const { return pk2(lo, hi); }
        __device__ __forceinline__ F2& operator=(f32x2 v) { unpk2(v, lo, hi); return *this; }
    };
    using State = typename std::conditional<(T <= HS_SCALAR_STATE_MAX_T), F2, f32x2>::type;
    State p[T][4], g[T][4];
#pragma unroll
    for (int s = 0; s < T; ++s)
#pragma unroll
        for (int j = 0; j < 4; ++j) { p[s][j] = (f32x2)0ull; g[s][j] = (f32x2)0ull; }

    const bool lane_out = (lane >= C::HL / 4) && (lane < 32 - C::HL / 4) && (col0 < W);
    // TRACK: what each stage received one tick ago (= the old value of the row it puts out now) and its running
    // max |new - old| over the pixels this unit owns (rows [R0, R1), the strip's valid columns, columns < W)
    constexpr int TS = TRACK ? T : 1;
    float pu[TS][4], pv[TS][4], em[TS];
#pragma unroll
    for (int s = 0; s < TS; ++s) {
        em[s] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) { pu[s][j] = 0.f; pv[s][j] = 0.f; }
    }
    const int wj = min(max(W - col0, 0), 4);       // this lane's columns inside the image
    auto pass_through = [&](const int S, float (&cu)[4], float (&cv)[4]) {   // stage S >= lim: out = previous input
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float tu = pu[S][j], tv = pv[S][j];
            pu[S][j] = cu[j]; pv[S][j] = cv[j];
            cu[j] = tu; cv[j] = tv;
        }
    };
    auto track_delta = [&](const int S, const float (&cu)[4], const float (&cv)[4], const bool all_cols) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (all_cols || j < wj)
                em[S] = fmaxf(em[S], fmaxf(fabsf(__fsub_rn(cu[j], pu[S][j])), fabsf(__fsub_rn(cv[j], pv[S][j]))));
    };
    float* uo = A.u_out + (size_t)z * A.out_pair_pitch + col0;
    float* vo = A.v_out + (size_t)z * A.out_pair_pitch + col0;

    // Peer transport (row strips over several GPUs): rows a neighbour keeps as ghost rows are stored a second time,
    // straight into that neighbour's destination buffer over NVLink.  Decided per chunk: only the units at the top
    // and bottom of the strip ever take the branch.
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
    // above it, written back into cu, cv.  ka, kb, kc = this lane's coefficients of that row.
    auto stage_row = [&](auto edge_tag, auto s_tag, float (&cu)[4], float (&cv)[4], const float4& ka, const float4& kb, const float4& kc,
                         const bool own) {
        constexpr bool EDGE = decltype(edge_tag)::value;
        constexpr int S = decltype(s_tag)::value;
        [[maybe_unused]] float iu[4], iv[4];
        if constexpr (TRACK) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { iu[j] = cu[j]; iv[j] = cv[j]; }
        }
        if (EDGE && wmis) { sanitize_right(cu, col0, W); sanitize_right(cv, col0, W); }
        float lu = __shfl_up_sync(kFull, cu[3], 1), ru = __shfl_down_sync(kFull, cu[0], 1);
        float lv = __shfl_up_sync(kFull, cv[3], 1), rv = __shfl_down_sync(kFull, cv[0], 1);
        if (EDGE) { clamp_lr(cu, col0, W, lu, ru); clamp_lr(cv, col0, W, lv, rv); }
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
        if constexpr (TRACK) {
            if (own) track_delta(S, cu, cv, !EDGE);
#pragma unroll
            for (int j = 0; j < 4; ++j) { pu[S][j] = iu[j]; pv[S][j] = iv[j]; }
        }
    };

    // ---- generic tick: pipeline fill, bottom-edge drain, tiny frames (runtime predicates) ----------
    auto tick_gen = [&](const int r) {
        const int rr = r - base;                   // row relative to the first TMA group
        float cu[4], cv[4];
        bool have = false;
        if (r <= H - 1) {
            if ((rr % RG) == 0 || r == rs) {       // first row consumed from this group
                const int gr = rr / RG;
                mbar_wait(ubar_of(gr % NGUV), (gr / NGUV) & 1);
                mbar_wait(cbar_of(gr % NGC), (gr / NGC) & 1);
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
            if constexpr (TRACK) {
                if (S >= lim) {                    // pass-through stage
                    if (virt) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) { cu[j] = pu[S][j]; cv[j] = pv[S][j]; }
                    } else {
                        pass_through(S, cu, cv);
                        if (rho == rs) have = false;   // first row of the stage: nothing to put out yet
                    }
                    continue;
                }
            }
            [[maybe_unused]] float iu[4], iv[4];
            if constexpr (TRACK) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { iu[j] = cu[j]; iv[j] = cv[j]; }
            }
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
                    if constexpr (TRACK) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) { pu[S][j] = iu[j]; pv[S][j] = iv[j]; }
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
            if constexpr (TRACK) {                 // time step S+1 of row rho-1 against its time step S
                if (lane_out && rho - 1 >= R0 && rho - 1 < R1) track_delta(S, cu, cv, false);
                if (!virt) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { pu[S][j] = iu[j]; pv[S][j] = iv[j]; }
                }
            }
        }
        const int ro = r - T;
        if (have && lane_out && ro >= R0 && ro < R1) {
            const size_t o = (size_t)ro * A.row_pitch;
            *reinterpret_cast<float4*>(uo + o) = make_float4(cu[0], cu[1], cu[2], cu[3]);
            *reinterpret_cast<float4*>(vo + o) = make_float4(cv[0], cv[1], cv[2], cv[3]);
            if constexpr (PEER) {
                if ((A.peer_up != nullptr && ro < A.up_hi) || (A.peer_dn != nullptr && ro >= A.dn_lo))
                    peer_push_row(&A, ro, col0, make_float4(cu[0], cu[1], cu[2], cu[3]), make_float4(cv[0], cv[1], cv[2], cv[3]));
            }
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
    // All T stages active, no run-time predicates on rows.  One trip = RG ticks = one TMA row group of u/v consumed.
    // Every shared-memory address of a trip is "register + compile-time constant": the loop keeps one pointer per
    // coefficient group in use (kg[k] = group of the trip minus k, k = 0..DRET) and one for the current u/v slot, and
    // rotates them once per trip; barrier addresses, phases and TMA coordinates advance incrementally.
    // Output row of tick r is r - T; it is stored when 0 <= r - T - R0 < R1 - R0 (one unsigned compare, lanes of
    // the halo excluded).
    const unsigned out_rows = lane_out ? (unsigned)(R1 - R0) : 0u;
    auto steady = [&](auto edge_tag, int r, const int r_end) {   // r group-aligned; returns the next tick
        const int grow = (r - base) / RG;                      // virtual group index of the first trip (>= DRET)
        const uint32_t sa_l32 = sa32 + lane * 16, su_l32 = su32 + lane * 16;
        uint32_t kg[DRET + 1];                                 // shared-memory address (this lane's 16 bytes) of group -k
#pragma unroll
        for (int k = 0; k <= DRET; ++k) kg[k] = sa_l32 + ((grow - k) % NGC) * (RG * CROW);
        uint32_t up = su_l32 + (grow % NGUV) * (RG * UROW);
        uint32_t ubar = ubar_of(grow % NGUV), upar = (uint32_t)(grow / NGUV) & 1u;
        int cslot = grow % NGC;
        uint32_t cbar = cbar_of(cslot), cpar = (uint32_t)(grow / NGC) & 1u;
        // refills (lane 0): next u/v group -> the slot just drained; next coefficient group -> the slot retired
        int uy = (g0v + grow + NGUV) * RG;                     // first row of the next u/v group to request
        int cy = (g0v + grow - DRET + NGC) * RG;               // ... of the next coefficient group
        int fslot = (grow - DRET) % NGC;                       // slot that retires at the end of this trip
        const int ylast = glast * RG;
        const int cymin = (g0 + NGC) * RG;                     // earlier groups were requested by the prologue
        float* uo_row = uo + (size_t)(r - T) * A.row_pitch;    // advanced by one row per tick
        float* vo_row = vo + (size_t)(r - T) * A.row_pitch;
        unsigned orow = (unsigned)(r - T - R0);                // output row of the tick relative to the chunk
#pragma unroll kTripUnroll
        for (; r + RG - 1 <= r_end; r += RG) {
            mbar_wait(ubar, upar);
            mbar_wait(cbar, cpar);
            auto tick = [&](auto j_tag) {
                constexpr int J = decltype(j_tag)::value;
                float cu[4], cv[4];
                const float4 tu = lds128<J * UROW>(up);
                const float4 tv = lds128<J * UROW + ROWB>(up);
                cu[0] = tu.x; cu[1] = tu.y; cu[2] = tu.z; cu[3] = tu.w;
                cv[0] = tv.x; cv[1] = tv.y; cv[2] = tv.z; cv[3] = tv.w;
                // stage S consumes row (r + J - S) and needs the coefficients of row (r + J - S - 1): row
                // J - S - 1 relative to the first row of the trip's group, i.e. row REL + K*RG of group -K
                auto stage = [&](auto s_tag) {
                    constexpr int S = decltype(s_tag)::value;
                    constexpr int REL = J - S - 1;
                    constexpr int K = REL >= 0 ? 0 : (-REL + RG - 1) / RG;
                    constexpr int OFF = (REL + K * RG) * CROW;
                    if constexpr (TRACK) {
                        if (S >= lim) { pass_through(S, cu, cv); return; }
                    }
                    const float4 ka = lds128<OFF>(kg[K]), kb = lds128<OFF + ROWB>(kg[K]), kc = lds128<OFF + 2 * ROWB>(kg[K]);
                    // the row this stage puts out is (T - S - 1) rows below the row the tick stores
                    stage_row(edge_tag, s_tag, cu, cv, ka, kb, kc, TRACK && (orow + (unsigned)(J + T - S - 1) < out_rows));
                };
                [&]<int... S>(std::integer_sequence<int, S...>) {
                    (stage(std::integral_constant<int, S>{}), ...);
                }(std::make_integer_sequence<int, T>{});
                if (orow + J < out_rows) {
                    *reinterpret_cast<float4*>(uo_row) = make_float4(cu[0], cu[1], cu[2], cu[3]);
                    *reinterpret_cast<float4*>(vo_row) = make_float4(cv[0], cv[1], cv[2], cv[3]);
                    // no peer stores here: the ticks whose output rows a neighbour keeps as ghosts run in tick_gen (see the
                    // driver loop below), so that the steady state of a PEER launch is the steady state of any launch
                }
                uo_row += A.row_pitch; vo_row += A.row_pitch;
            };
            [&]<int... J>(std::integer_sequence<int, J...>) {
                (tick(std::integral_constant<int, J>{}), ...);
            }(std::make_integer_sequence<int, RG>{});
            orow += RG;
            __syncwarp();
            if (lane == 0) {
                if (uy <= ylast) {
                    mbar_expect_tx(ubar, (uint32_t)RG * UROW);
                    tma_load_4d(up, &tm_uv, x0, 0, uy, A.z_in0 + z, ubar);   // lane 0: up is the slot base
                }
                if ((unsigned)(cy - cymin) <= (unsigned)(ylast - cymin) && ylast >= cymin) {
                    const uint32_t fbar = cbar_of(fslot);
                    mbar_expect_tx(fbar, (uint32_t)RG * CROW);
                    tma_load_4d(kg[DRET], &tm_c, x0, 0, HS_COEF_Y(cy), HS_COEF_Z(A.z_c0 + z), fbar);
                }
            }
            uy += RG; cy += RG;
            if (++fslot == NGC) fslot = 0;
            // rotate: the u/v slot toggles (its phase flips when slot 1 -> 0), the coefficient groups age by one
            upar ^= (ubar >> 3) & 1u;                          // ubar_of(1) = bar0 + 8 (bar0 is 128-byte aligned)
            ubar ^= 8u;
            up = (up == su_l32) ? su_l32 + RG * UROW : su_l32;
#pragma unroll
            for (int k = DRET; k > 0; --k) kg[k] = kg[k - 1];
            if (++cslot == NGC) { cslot = 0; cpar ^= 1u; }
            cbar = cbar_of(cslot);
            kg[0] = sa_l32 + cslot * (RG * CROW);
        }
        return r;
    };

    int r = rs;
    int gen_end = min(fill_end, last_tick + 1);                  // pipeline fill (and tiny frames)
    int steady_end = min(H - 1, last_tick);                      // last tick with a real input row
    if constexpr (PEER) {
        // Output row ro leaves at tick ro + T.  Rows [up_lo, up_hi) / [dn_lo, dn_hi) are also stored into a neighbour:
        // those few ticks (T + 1 rows at the top of the first chunk, T at the bottom of the last one) stay in the
        // generic path, which carries the push code; the branch-free loop in between carries none.  (With the push
        // inside the steady loop every unit of a PEER launch paid 25 % more instructions per stage-row, and the seam
        // units -- the slowest of a single-wave launch -- set its duration.)
        if (A.peer_up != nullptr && R0 < A.up_hi && R1 > A.up_lo) gen_end = max(gen_end, min((A.up_hi + T + RG - 1) / RG * RG, last_tick + 1));
        if (A.peer_dn != nullptr && R0 < A.dn_hi && R1 > A.dn_lo) steady_end = min(steady_end, A.dn_lo + T - 1);
    }
    for (int pass = 0; pass < 2; ++pass) {
        for (; r < gen_end; ++r) tick_gen(r);
        if (pass == 1 || r > last_tick) break;
        // whole groups only, so that the generic ticks that follow see consistent ring bookkeeping
        const int st_end = r + ((steady_end - r + 1) / RG) * RG - 1;
        if (edge) {
            r = steady(std::true_type{}, r, st_end);
        } else {
            r = steady(std::false_type{}, r, st_end);
        }
        gen_end = last_tick + 1;                                 // bottom edge / remainder: generic ticks
    }

    if constexpr (TRACK) {
        if (A.trk_mode == 0) {                     // main launch: this unit's share of the per-stage max-norms
#pragma unroll
            for (int s = 0; s < TS; ++s) {
                float m = em[s];
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, d));
                if (lane == 0 && s < lim && m > 0.f) atomicMax(A.emax + zt * kMaxT + s, __float_as_uint(m));
            }
        }
    }

    // Halo-exchange signal: every counted unit (all of them, or just the two seam chunks -- StreamArgs::seam_first)
    // counts itself done after its stores (local and peer) are visible system-wide; the last one resets the counter
    // and publishes the epoch to both neighbours, whose streams wait on that word (cuStreamWaitValue32) before they
    // launch the next block.  No kernel ever spins on it.
    // (whether this unit counts is recomputed from blockIdx here instead of being kept alive through the pipeline)
    bool counted = PEER;
    if constexpr (PEER) {
        if (A.seam_first && A.ncy > 2) { const int q = (int)((blockIdx.x / (unsigned)A.nsx) % (unsigned)A.ncy); counted = q < 2; }
    }
    if (PEER && A.done_counter != nullptr && counted) {
        __syncwarp();
        if (lane == 0) {
            __threadfence_system();
            const unsigned prev = atomicAdd(A.done_counter, 1u);
            if (prev == (unsigned)(A.signal_units - 1)) {
                *A.done_counter = 0u;
                __threadfence_system();
                if (A.flag_up) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(A.flag_up), "r"(A.epoch) : "memory");
                if (A.flag_dn) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(A.flag_dn), "r"(A.epoch) : "memory");
            }
        }
    }
}


}  // namespace hs
