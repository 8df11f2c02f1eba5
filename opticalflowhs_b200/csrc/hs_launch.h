// hs_launch.h -- argument blocks and host-side launchers shared by the kernels and the C ABI.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hs {

struct DerivArgs {
    const uint8_t* f1;          // frame planes [pair][row][bytes]
    const uint8_t* f2;
    long long f_row_pitch;      // bytes
    long long f_pair_pitch;     // bytes
    float* c0;                  // Ex|a, Ey|b, Et|c planes [pair][row][col]
    float* c1;
    float* c2;
    long long c_row_pitch;      // elements
    long long c_pair_pitch;     // elements
    int W, H;
    int normalise;              // 1: write a,b,c = (Ex,Ey,Et)/sqrt(rho+Ex^2+Ey^2); 0: raw derivatives
    int zero_b;                 // 1: store b = 0 (LITERAL mode on the streaming kernel, see literal_on_stream in hsflow_capi.cu)
    float rho;                  // alpha^2 (Kernels.cl:85) or 1/lambda
};

struct Jacobi1Args {
    const float* u_in;
    const float* v_in;
    float* u_out;
    float* v_out;
    const float* c0;
    const float* c1;
    const float* c2;
    long long row_pitch;        // elements, u/v planes (source and destination)
    long long in_pair_pitch, out_pair_pitch;
    long long c_row_pitch, c_pair_pitch;   // coefficient planes
    int W, H;                   // H = rows of the buffer (replicate beyond both ends)
    int out_lo, out_hi;         // rows to produce
    int chunk_rows;
    float rho;                  // EXACT only
    int cv_form;                // EXACT only: 1 = the operation order of cvCalcOpticalFlowHS (update_exact_cv), 0 = Kernels.cl:84-86
    // Convergence tracking (hsflow_set_epsilon; the CV_TERMCRIT_EPS half of cv.cpp:29), nullptr = off.
    // emax[pair]: max |new - old| of this sweep as float bits (atomicMax; non-negative floats order like uints).
    // stop[pair]: 0 = still iterating, else (sweeps executed << 1) | parity of the buffer that holds the field.
    // A stopped pair skips the sweep; the LAST sweep of a call copies it when it sits in the other buffer.
    unsigned* emax;
    int* stop;
    int last_sweep;             // 1: final sweep of this hsflow_iterate / hsflow_compute call
    int total_sweeps;           // sweeps since hsflow_prepare at the end of the call (parity of the final buffer)
};

struct StreamArgs {
    float* u_out;
    float* v_out;
    long long row_pitch;        // elements
    long long out_pair_pitch;
    int W, H;
    int out_lo, out_hi;
    int chunk_rows;
    int nsx, ncy;
    int z_in0, z_c0;            // first pair index inside the u/v source map and the coefficient maps
    long long total_units;
    // Row-strip mode with peer transport (hsflow_strip_connect): output rows [up_lo, up_hi) are ALSO stored into the
    // upper neighbour's destination buffer at row + up_delta (its bottom ghost rows), rows [dn_lo, dn_hi) into the
    // lower neighbour's at row + dn_delta (its top ghost rows), over NVLink peer memory.  The last work unit of the
    // launch to finish publishes `epoch` in both neighbours' flag words (fused compute + halo exchange + signal).
    float* peer_up;             // neighbour buffers (u plane; v = u + (v_out - u_out)), nullptr: no neighbour
    float* peer_dn;
    int up_lo, up_hi, up_delta;
    int dn_lo, dn_hi, dn_delta;
    unsigned* done_counter;     // device word, 0 between launches; nullptr: no signalling
    // seam_first = 1: the top and bottom row chunks are scheduled first and only they are counted (signal_units of
    // them): the neighbours hear about this launch as soon as its seam rows are out, long before it ends, so the
    // wait before the next launch is already satisfied.  Set by launch_jacobi_stream when all ghost-row reads and
    // all pushes fall into those two chunks; otherwise every unit counts (signal_units = total_units).
    int seam_first;
    long long signal_units;
    unsigned* flag_up;          // word in the upper / lower neighbour's memory that receives `epoch`
    unsigned* flag_dn;
    unsigned epoch;
    // EPS criterion on the temporally blocked kernel (TRACK instantiation; cvTermCriteria(ITER | EPS), cv.cpp:29).
    // A block of trk_t <= T sweeps runs as TWO launches over the same source and destination buffers:
    //   main   (trk_mode 0): pairs that are still iterating run all trk_t sweeps; every stage S reduces
    //          max |x(S+1) - x(S)| over the pixels the unit owns into emax[pair][S] (atomicMax on the float bits);
    //   replay (trk_mode 1): a pair whose FIRST stage s with emax[pair][s] < eps exists should have stopped after
    //          s + 1 sweeps: its units run again from the (still intact) source buffer with the stages >= s + 1
    //          passing their input through, overwrite the destination with the field after exactly s + 1 sweeps and
    //          record stop[pair] = ((trk_base + s + 1) << 1) | trk_dst_parity.  Units of every other pair exit at once.
    // No host round trip: the launch sequence stays asynchronous.  Bit-identical in sweep count and field to the
    // single-sweep kernel with its per-sweep check (k_jacobi1<.., TRACK>).
    int* stop;                  // [pairs] 0 = still iterating, else (sweeps executed << 1) | buffer that holds the field (0 = A planes, 1 = B planes)
    unsigned* emax;             // [pairs][kMaxT], this block's bank (zero on entry of the main launch)
    unsigned* emax_next;        // the other bank: zeroed by the replay launch for the next block
    int trk_mode, trk_t, trk_base, trk_dst_parity;
    int z_trk0;                 // first pair's index into stop / emax
    double eps;
};

struct StreamGeom {             // filled by stream_geometry()
    int halo, valid_w, smem_per_warp, max_T;
    int rows_per_box;           // rows of one TMA box (2, 3 or 4): selects the tensor maps the launch needs
};

cudaError_t launch_deriv(const DerivArgs& A, int fmt, int pairs, cudaStream_t s);
cudaError_t launch_box3(const uint8_t* src, uint8_t* dst, int W, int H, long long rp, long long pp, int pairs, cudaStream_t s);
cudaError_t launch_bgr2gray(const uint8_t* src, uint8_t* dst, int W, int H, long long rp, long long pp, int pairs, cudaStream_t s);
cudaError_t launch_deriv_cv(const DerivArgs& A, int pairs, cudaStream_t s);
cudaError_t launch_jacobi1(const Jacobi1Args& A, bool exact, int stencil, bool update_v, int pairs, cudaStream_t s);
// after sweep number `sweep` (1-based since prepare): stop[z] = (sweep << 1) | (sweep & 1) where emax[z] < eps; emax[z] = 0
cudaError_t launch_eps_check(unsigned* emax, int* stop, double eps, int sweep, int pairs, cudaStream_t s);
// end of a call that ran up to `total` sweeps: every stopped pair now sits in the buffer of parity total & 1
cudaError_t launch_eps_settle(int* stop, int total, int pairs, cudaStream_t s);
cudaError_t launch_synth(uint8_t* f1, uint8_t* f2, int W, int rows, int full_h, int row0, long long rp, long long pp,
                         uint32_t seed0, int pairs, cudaStream_t s);
// out[z][i/step][j/step] = plane[z][i][j] for both fields (dense [pairs][ceil(H/step)][ceil(W/step)] outputs)
cudaError_t launch_sample_uv(const float* u, const float* v, int W, int H, long long row_pitch, long long pair_pitch, int step,
                             float* us, float* vs, int pairs, cudaStream_t s);
cudaError_t launch_dot_mask(const float* u, const float* v, int W, int H, long long pitch, int step, float thr,
                            uint8_t* mask, int* count, cudaStream_t s);

// temporally blocked streaming kernel (hs_stream.cu)
constexpr int kMaxT = 8;
// temporal_block = 0 (auto), three regimes (profiles/README.md, round-2 T sweeps on the final build, sustained under the 1000 W cap):
constexpr int kBigT = 7;               // >= kBigPixels per launch.  256 4K pairs: T = 6 / 7 / 8 = 1 040 / 1 075 / 1 065 k, 64 pairs 1 013 / 1 045 / 1 023 k,
                                       // a 16384^2 frame 125.2 / 121.6 / 124.3 ms per 500 iterations: the deeper block moves less HBM traffic, draws
                                       // less power and lets the SM clock rise (1.55 -> 1.64 -> 1.75 GHz); T = 8 no longer fits 255 registers
                                       // with the scalar pipeline state (hs_stream.cuh) and pays 12 % register moves in its packed form
constexpr int kDefaultT = 6;           // one to three 4K pairs (one pair: 842 k at T = 6), and every strip of a sharded frame (2048-row strip: 7.51 ms
                                       // per 240 iterations at T = 6, 7.90 at T = 7)
constexpr int kSmallT = 4;             // the job cannot fill the GPU twice over: shorter warm-up, shorter units (1080p: 573 k against 496 k at T = 6)
constexpr long long kBigPixels = 24LL << 20;   // ~ three 4K pairs
StreamGeom stream_geometry(int T);
cudaError_t stream_prepare(int device);     // opt in to large dynamic shared memory for every instantiation
int stream_warps_per_sm(int T, int stencil);   // resident warps (= CTAs) per SM (occupancy API)
// maps: row-interleaved u/v source {W,2,H,pairs} and coefficients {W,3,H,pairs}, box = 128 columns x all planes x
// stream_geometry(T).rows_per_box rows.  One warp per CTA.
cudaError_t launch_jacobi_stream(int T, int stencil, const CUtensorMap& tm_uv, const CUtensorMap& tm_c,
                                 StreamArgs A, int pairs, cudaStream_t s);
// whether that launch would schedule and count the two seam chunks first (StreamArgs::seam_first); pure geometry
bool stream_seam_first(int T, const StreamArgs& A);
// EPS mode: the TRACK instantiation exists for one depth only; shorter (tail) blocks run on it with trk_t < kTrackT
constexpr int kTrackT = 4;
cudaError_t launch_jacobi_stream_track(int stencil, const CUtensorMap& tm_uv, const CUtensorMap& tm_c, StreamArgs A, int pairs,
                                       cudaStream_t s);
// end of a call in EPS mode: pairs that stopped in the other ping-pong buffer are copied into the final one
// (a, b: the A-plane and B-plane blocks of the n pairs; final_parity: which of them holds the result), then re-labelled
cudaError_t launch_copy_stopped(float* a, float* b, long long pair_floats, int* stop, int final_parity, int pairs, cudaStream_t s);

}  // namespace hs
