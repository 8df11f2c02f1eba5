// hs_kernels.cu -- single-pass kernels of the B200 Horn-Schunck engine (sm_100a):
//   k_deriv     fused grayscale conversion + Ix/Iy/It cube derivative (+ coefficient normalisation)
//               replaces cvCvtColor (HSOpticalFlowOpenCL.cpp:727-728), readInputImage (cpp:6-45)
//               and ComputeDerivativesKernel (Kernels.cl:13-39)
//   k_jacobi1   one fused u_v_avgKernel + u_v_updateKernel sweep (Kernels.cl:43-90), row-streaming,
//               FAST or EXACT arithmetic; the T = 1 member of the iteration family and the
//               cross-check for the temporally blocked kernel in hs_stream.cu
//   k_box3 / k_deriv_cv   OpenCV-mode pre-blur and Sobel estimator (OpticalFlowOpenCV.cpp:27-29)
//   k_synth     deterministic synthetic frame pairs (mirror of oracle hso_synth_pair)
//   k_dot_mask  drawing predicate of cpp:762-765
// All kernels are HBM-bound byte/float streaming: vectorised (float4 / uchar4) coalesced access,
// one warp per 128-column strip, no tensor cores by design.
#include <algorithm>

#include "hs_common.cuh"
#include "hs_launch.h"

namespace hs {

// ------------------------------------------------------------------------------------------
// k_deriv
// ------------------------------------------------------------------------------------------
template <int FMT> struct PxLoader;

template <> struct PxLoader<FMT_GRAY8> {
    // v[0..4] = columns x..x+4 of one row, clamped at W-1 (Tex2D, Kernels.cl:6)
    static __device__ __forceinline__ void load(const uint8_t* row, int x, int W, float (&v)[5]) {
        const uchar4 q = *reinterpret_cast<const uchar4*>(row + x);
        v[0] = (float)q.x; v[1] = (float)q.y; v[2] = (float)q.z; v[3] = (float)q.w;
        v[4] = (x + 4 <= W - 1) ? (float)row[x + 4] : 0.f;
    }
};
template <> struct PxLoader<FMT_BGR8> {
    static __device__ __forceinline__ float gray(uint32_t b, uint32_t g, uint32_t r) {
        // OpenCV 2.1 BGR2GRAY fixed point (cpp:727-728): (B*1868 + G*9617 + R*4899 + 8192) >> 14
        return (float)((b * 1868u + g * 9617u + r * 4899u + 8192u) >> 14);
    }
    static __device__ __forceinline__ void load(const uint8_t* row, int x, int W, float (&v)[5]) {
        const uint32_t* p = reinterpret_cast<const uint32_t*>(row + 3 * (size_t)x);   // 12-byte groups, 4-aligned
        const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
        v[0] = gray(w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u);
        v[1] = gray(w0 >> 24, w1 & 255u, (w1 >> 8) & 255u);
        v[2] = gray((w1 >> 16) & 255u, w1 >> 24, w2 & 255u);
        v[3] = gray((w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24);
        if (x + 4 <= W - 1) {
            const uint8_t* q = row + 3 * (size_t)(x + 4);
            v[4] = gray(q[0], q[1], q[2]);
        } else v[4] = 0.f;
    }
};
template <> struct PxLoader<FMT_F32> {
    static __device__ __forceinline__ void load(const uint8_t* row8, int x, int W, float (&v)[5]) {
        const float* row = reinterpret_cast<const float*>(row8);
        const float4 q = *reinterpret_cast<const float4*>(row + x);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        v[4] = (x + 4 <= W - 1) ? row[x + 4] : 0.f;
    }
};

__device__ __forceinline__ void clamp5(float (&v)[5], int x, int W) {
#pragma unroll
    for (int j = 1; j < 5; ++j)
        if (x + j > W - 1) v[j] = v[j - 1];
}

template <int FMT>
__global__ void __launch_bounds__(128) k_deriv(DerivArgs A) {
    const int x = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int j = blockIdx.y * 4 + threadIdx.y;
    const int z = blockIdx.z;
    if (x >= A.W || j >= A.H) return;
    const int j1 = min(j + 1, A.H - 1);                        // Tex2D: j >= h -> h-1
    const uint8_t* f1 = A.f1 + (size_t)z * A.f_pair_pitch;
    const uint8_t* f2 = A.f2 + (size_t)z * A.f_pair_pitch;
    float a0[5], a1[5], b0[5], b1[5];
    PxLoader<FMT>::load(f1 + (size_t)j * A.f_row_pitch, x, A.W, a0);
    PxLoader<FMT>::load(f1 + (size_t)j1 * A.f_row_pitch, x, A.W, a1);
    PxLoader<FMT>::load(f2 + (size_t)j * A.f_row_pitch, x, A.W, b0);
    PxLoader<FMT>::load(f2 + (size_t)j1 * A.f_row_pitch, x, A.W, b1);
    clamp5(a0, x, A.W); clamp5(a1, x, A.W); clamp5(b0, x, A.W); clamp5(b1, x, A.W);
    float o0[4], o1[4], o2[4];
    const float q = (float)(1.0 / 4);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        // Kernels.cl:25-38, left-associated as written
        float ex = __fsub_rn(a0[k + 1], a0[k]);
        ex = __fadd_rn(ex, a1[k + 1]); ex = __fsub_rn(ex, a1[k]);
        ex = __fadd_rn(ex, b0[k + 1]); ex = __fsub_rn(ex, b0[k]);
        ex = __fadd_rn(ex, b1[k + 1]); ex = __fsub_rn(ex, b1[k]);
        ex = __fmul_rn(q, ex);
        float ey = __fsub_rn(a1[k], a0[k]);
        ey = __fadd_rn(ey, a1[k + 1]); ey = __fsub_rn(ey, a0[k + 1]);
        ey = __fadd_rn(ey, b1[k]);     ey = __fsub_rn(ey, b0[k]);
        ey = __fadd_rn(ey, b1[k + 1]); ey = __fsub_rn(ey, b0[k + 1]);
        ey = __fmul_rn(q, ey);
        float et = __fsub_rn(b0[k], a0[k]);
        et = __fadd_rn(et, b0[k + 1]); et = __fsub_rn(et, a0[k + 1]);
        et = __fadd_rn(et, b1[k]);     et = __fsub_rn(et, a1[k]);
        et = __fadd_rn(et, b1[k + 1]); et = __fsub_rn(et, a1[k + 1]);
        et = __fmul_rn(q, et);
        if (A.normalise) normalise_coefs(ex, ey, et, A.rho, o0[k], o1[k], o2[k]);
        else { o0[k] = ex; o1[k] = ey; o2[k] = et; }
        if (A.zero_b) o1[k] = 0.f;
    }
    const size_t o = (size_t)z * A.c_pair_pitch + (size_t)j * A.c_row_pitch + x;
    *reinterpret_cast<float4*>(A.c0 + o) = make_float4(o0[0], o0[1], o0[2], o0[3]);
    *reinterpret_cast<float4*>(A.c1 + o) = make_float4(o1[0], o1[1], o1[2], o1[3]);
    *reinterpret_cast<float4*>(A.c2 + o) = make_float4(o2[0], o2[1], o2[2], o2[3]);
}

// ------------------------------------------------------------------------------------------
// OpenCV-mode estimator: 3x3 box blur (cvSmooth CV_BLUR, cv.cpp:27-28) then Sobel/8 on frame 1
// and It = frame2 - frame1 (icvCalcOpticalFlowHS_8u32fR).  One thread per pixel.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_box3(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                               int W, int H, long long row_pitch, long long pair_pitch) {
    const int x = blockIdx.x * 64 + threadIdx.x, y = blockIdx.y * 4 + threadIdx.y;
    if (x >= W || y >= H) return;
    const uint8_t* s = src + (size_t)blockIdx.z * pair_pitch;
    const int xl = max(x - 1, 0), xr = min(x + 1, W - 1);
    int sum = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const uint8_t* r = s + (size_t)min(max(y + dy, 0), H - 1) * row_pitch;
        sum += r[xl] + r[x] + r[xr];
    }
    // round-half-even of sum/9 as cvRound(sum * (1/9.)) does; sum/9 is never exactly .5
    dst[(size_t)blockIdx.z * pair_pitch + (size_t)y * row_pitch + x] = (uint8_t)__double2int_rn((double)sum * (1.0 / 9.0));
}

// cvCvtColor(CV_BGR2GRAY) of OpenCV 2.1 (cv.cpp:17, 20) for the OpenCV-mode path when the frames arrive as BGR
__global__ void __launch_bounds__(256) k_bgr2gray(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int W, int H,
                                                   long long row_pitch, long long pair_pitch) {
    const int x = blockIdx.x * 64 + threadIdx.x, y = blockIdx.y * 4 + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t r = (size_t)blockIdx.z * pair_pitch + (size_t)y * row_pitch;
    const uint8_t* p = src + r + 3 * (size_t)x;
    dst[r + x] = (uint8_t)((p[0] * 1868u + p[1] * 9617u + p[2] * 4899u + 8192u) >> 14);
}

__global__ void __launch_bounds__(256) k_deriv_cv(DerivArgs A) {
    const int x = blockIdx.x * 64 + threadIdx.x, y = blockIdx.y * 4 + threadIdx.y;
    if (x >= A.W || y >= A.H) return;
    const uint8_t* f1 = A.f1 + (size_t)blockIdx.z * A.f_pair_pitch;
    const uint8_t* f2 = A.f2 + (size_t)blockIdx.z * A.f_pair_pitch;
    const int xl = max(x - 1, 0), xr = min(x + 1, A.W - 1);
    const uint8_t* r0 = f1 + (size_t)max(y - 1, 0) * A.f_row_pitch;
    const uint8_t* r1 = f1 + (size_t)y * A.f_row_pitch;
    const uint8_t* r2 = f1 + (size_t)min(y + 1, A.H - 1) * A.f_row_pitch;
    const int gx = (r0[xr] + 2 * r1[xr] + r2[xr]) - (r0[xl] + 2 * r1[xl] + r2[xl]);
    const int gy = (r2[xl] + 2 * r2[x] + r2[xr]) - (r0[xl] + 2 * r0[x] + r0[xr]);
    float ex = __fmul_rn((float)gx, 0.125f), ey = __fmul_rn((float)gy, 0.125f);
    float et = (float)((int)f2[(size_t)y * A.f_row_pitch + x] - (int)r1[x]);
    float o0 = ex, o1 = ey, o2 = et;
    if (A.normalise) normalise_coefs(ex, ey, et, A.rho, o0, o1, o2);
    if (A.zero_b) o1 = 0.f;
    const size_t o = (size_t)blockIdx.z * A.c_pair_pitch + (size_t)y * A.c_row_pitch + x;
    A.c0[o] = o0; A.c1[o] = o1; A.c2[o] = o2;
}

// ------------------------------------------------------------------------------------------
// k_jacobi1
// ------------------------------------------------------------------------------------------
struct Row {
    float c[4];   // columns col0..col0+3
    float l, r;   // columns col0-1, col0+4 (clamped)
};

__device__ __forceinline__ Row load_row(const float* __restrict__ plane, long long pitch, int row, int H,
                                         int col0, int W, int lane) {
    row = min(max(row, 0), H - 1);                             // Tex2D: clamp j
    const float* rp = plane + (size_t)row * pitch;
    Row R;
    if (col0 < W) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(rp + col0));
        R.c[0] = t.x; R.c[1] = t.y; R.c[2] = t.z; R.c[3] = t.w;
    } else { R.c[0] = R.c[1] = R.c[2] = R.c[3] = 0.f; }
    if (W & 3) sanitize_right(R.c, col0, W);
    R.l = __shfl_up_sync(kFull, R.c[3], 1);
    R.r = __shfl_down_sync(kFull, R.c[0], 1);
    if (lane == 0) R.l = (col0 > 0) ? __ldg(rp + col0 - 1) : R.c[0];
    if (lane == 31) R.r = (col0 + 4 <= W - 1) ? __ldg(rp + col0 + 4) : R.c[3];
    clamp_lr(R.c, col0, W, R.l, R.r);
    return R;
}
__device__ __forceinline__ float nb(const Row& R, int j) { return j < 0 ? R.l : (j > 3 ? R.r : R.c[j]); }

template <bool EXACT, int ST, bool UPDATE_V, bool TRACK>
__global__ void __launch_bounds__(128) k_jacobi1(Jacobi1Args A) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col0 = ((blockIdx.x * 4 + warp) * 32 + lane) * 4;
    if (((blockIdx.x * 4 + warp) * kStripW) >= A.W) return;    // whole warp outside
    const int R0 = A.out_lo + blockIdx.y * A.chunk_rows;
    const int R1 = min(R0 + A.chunk_rows, A.out_hi);
    const int z = blockIdx.z;
    const float* ui = A.u_in + (size_t)z * A.in_pair_pitch;
    const float* vi = A.v_in + (size_t)z * A.in_pair_pitch;
    float* uo = A.u_out + (size_t)z * A.out_pair_pitch;
    float* vo = A.v_out + (size_t)z * A.out_pair_pitch;
    const float* c0 = A.c0 + (size_t)z * A.c_pair_pitch;
    const float* c1 = A.c1 + (size_t)z * A.c_pair_pitch;
    const float* c2 = A.c2 + (size_t)z * A.c_pair_pitch;
    const bool inb = col0 < A.W;
    bool copy_only = false;                                    // TRACK: a converged pair is carried over, not iterated
    if (TRACK) {
        const int st = A.stop[z];
        if (st) {
            if (!A.last_sweep || ((st ^ A.total_sweeps) & 1) == 0) return;
            copy_only = true;
        }
    }
    float emax = 0.f;

    Row um = load_row(ui, A.row_pitch, R0 - 1, A.H, col0, A.W, lane);
    Row vm = load_row(vi, A.row_pitch, R0 - 1, A.H, col0, A.W, lane);
    Row u0 = load_row(ui, A.row_pitch, R0, A.H, col0, A.W, lane);
    Row v0 = load_row(vi, A.row_pitch, R0, A.H, col0, A.W, lane);
    for (int rho = R0; rho < R1; ++rho) {
        const Row up = load_row(ui, A.row_pitch, rho + 1, A.H, col0, A.W, lane);
        const Row vp = load_row(vi, A.row_pitch, rho + 1, A.H, col0, A.W, lane);
        float4 k0 = make_float4(0, 0, 0, 0), k1 = k0, k2 = k0;
        if (inb) {
            const size_t o = (size_t)rho * A.c_row_pitch + col0;
            k0 = __ldg(reinterpret_cast<const float4*>(c0 + o));
            k1 = __ldg(reinterpret_cast<const float4*>(c1 + o));
            k2 = __ldg(reinterpret_cast<const float4*>(c2 + o));
        }
        const float ka[4] = {k0.x, k0.y, k0.z, k0.w}, kb[4] = {k1.x, k1.y, k1.z, k1.w}, kc[4] = {k2.x, k2.y, k2.z, k2.w};
        float un[4], vn[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float ub, vb;
            if (EXACT) {
                const float we = ST == ST_CL8 ? (float)(1.0 / 6) : 0.25f;
                const float wd = ST == ST_CL8 ? (float)(1.0 / 12) : 0.0f;
                ub = avg_exact(we, wd, nb(u0, j - 1), nb(u0, j + 1), um.c[j], up.c[j],
                               nb(um, j - 1), nb(um, j + 1), nb(up, j - 1), nb(up, j + 1));
                vb = avg_exact(we, wd, nb(v0, j - 1), nb(v0, j + 1), vm.c[j], vp.c[j],
                               nb(vm, j - 1), nb(vm, j + 1), nb(vp, j - 1), nb(vp, j + 1));
                if (A.cv_form) update_exact_cv(ub, vb, ka[j], kb[j], kc[j], A.rho, un[j], vn[j]);
                else update_exact(ub, vb, ka[j], kb[j], kc[j], A.rho, un[j], vn[j]);
            } else {
                const float hum = __fadd_rn(nb(um, j - 1), nb(um, j + 1));
                const float hu0 = __fadd_rn(nb(u0, j - 1), nb(u0, j + 1));
                const float hup = __fadd_rn(nb(up, j - 1), nb(up, j + 1));
                const float hvm = __fadd_rn(nb(vm, j - 1), nb(vm, j + 1));
                const float hv0 = __fadd_rn(nb(v0, j - 1), nb(v0, j + 1));
                const float hvp = __fadd_rn(nb(vp, j - 1), nb(vp, j + 1));
                ub = combine<ST>(pOf<ST>(rowG<ST>(um.c[j], hum), hu0), rowG<ST>(up.c[j], hup));
                vb = combine<ST>(pOf<ST>(rowG<ST>(vm.c[j], hvm), hv0), rowG<ST>(vp.c[j], hvp));
                update_fast(ub, vb, ka[j], kb[j], kc[j], un[j], vn[j]);
            }
            if (!UPDATE_V) vn[j] = v0.c[j];                    // Kernels.cl:87-89: v is never written
            if (TRACK) {
                if (copy_only) { un[j] = u0.c[j]; vn[j] = v0.c[j]; }
                else if (col0 + j < A.W)                       // Eps = max |new - old| (icvCalcOpticalFlowHS)
                    emax = fmaxf(emax, fmaxf(fabsf(__fsub_rn(un[j], u0.c[j])), fabsf(__fsub_rn(vn[j], v0.c[j]))));
            }
        }
        if (inb) {
            const size_t o = (size_t)rho * A.row_pitch + col0;
            *reinterpret_cast<float4*>(uo + o) = make_float4(un[0], un[1], un[2], un[3]);
            *reinterpret_cast<float4*>(vo + o) = make_float4(vn[0], vn[1], vn[2], vn[3]);
        }
        um = u0; vm = v0; u0 = up; v0 = vp;
    }
    if (TRACK && !copy_only) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) emax = fmaxf(emax, __shfl_xor_sync(kFull, emax, d));
        if (lane == 0 && emax > 0.f) atomicMax(A.emax + z, __float_as_uint(emax));
    }
}

__global__ void k_eps_check(unsigned* emax, int* stop, double eps, int sweep, int pairs) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z >= pairs) return;
    if (!stop[z] && (double)__uint_as_float(emax[z]) < eps) stop[z] = (sweep << 1) | (sweep & 1);
    emax[z] = 0u;
}
__global__ void k_eps_settle(int* stop, int total, int pairs) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z < pairs && stop[z]) stop[z] = (stop[z] & ~1) | (total & 1);
}

// ------------------------------------------------------------------------------------------
// k_synth: integer value-noise texture warped by a smooth flow; mirrors oracle/hs_oracle.c
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash32(uint32_t seed, uint32_t ix, uint32_t iy) {
    uint32_t h = seed * 0x9E3779B1u ^ (ix * 0x85EBCA77u) ^ (iy * 0xC2B2AE3Du);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}
__device__ __forceinline__ int32_t octave(uint32_t seed, int k, int32_t X, int32_t Y) {
    const int sh = 8 + k;
    const int32_t cx = X >> sh, cy = Y >> sh;
    const int32_t fx = (X & ((1 << sh) - 1)) >> k, fy = (Y & ((1 << sh) - 1)) >> k;
    const int32_t v00 = (int32_t)(hash32(seed, (uint32_t)cx, (uint32_t)cy) & 255u);
    const int32_t v10 = (int32_t)(hash32(seed, (uint32_t)(cx + 1), (uint32_t)cy) & 255u);
    const int32_t v01 = (int32_t)(hash32(seed, (uint32_t)cx, (uint32_t)(cy + 1)) & 255u);
    const int32_t v11 = (int32_t)(hash32(seed, (uint32_t)(cx + 1), (uint32_t)(cy + 1)) & 255u);
    const int32_t top = v00 * (256 - fx) + v10 * fx, bot = v01 * (256 - fx) + v11 * fx;
    return (top * (256 - fy) + bot * fy) >> 8;
}
__device__ __forceinline__ uint8_t texture(uint32_t seed, int32_t X, int32_t Y) {
    const int32_t s = 3 * octave(seed, 5, X, Y) + 3 * octave(seed + 0x632BE5ABu, 3, X, Y) +
                      2 * octave(seed + 0xC6A4A793u, 2, X, Y);
    return (uint8_t)(s >> 11);
}
__device__ __forceinline__ int32_t sinlike(int32_t p) {
    p &= 1023;
    const int32_t q = p & 511, val = (q * (512 - q)) >> 6;
    return p < 512 ? val : -val;
}
__global__ void __launch_bounds__(256) k_synth(uint8_t* f1, uint8_t* f2, int W, int rows, int full_h, int row0,
                                                long long row_pitch, long long pair_pitch, uint32_t seed0) {
    const int x = blockIdx.x * 64 + threadIdx.x, r = blockIdx.y * 4 + threadIdx.y;
    if (x >= W || r >= rows) return;
    const int y = row0 + r;
    const uint32_t seed = seed0 + blockIdx.z;
    const int32_t X = x << 8, Y = y << 8;
    const int32_t py = (int32_t)(((long long)y << 10) / full_h), px = (int32_t)(((long long)x << 10) / W);
    const int32_t dx = 384 + ((128 * sinlike(py)) >> 10);
    const int32_t dy = -192 + ((128 * sinlike(px + 256)) >> 10);
    const size_t o = (size_t)blockIdx.z * pair_pitch + (size_t)r * row_pitch + x;
    f1[o] = texture(seed, X, Y);
    f2[o] = texture(seed, X - dx, Y - dy);
}

// ------------------------------------------------------------------------------------------
// k_dot_mask: cpp:762-765 on the stride-`step` grid
// ------------------------------------------------------------------------------------------
__global__ void k_dot_mask(const float* __restrict__ u, const float* __restrict__ v, int W, int H, long long pitch,
                           int step, float thr, uint8_t* mask, int gw, int gh, int* count) {
    const int gj = blockIdx.x * blockDim.x + threadIdx.x, gi = blockIdx.y * blockDim.y + threadIdx.y;
    if (gj >= gw || gi >= gh) return;
    const size_t p = (size_t)(gi * step) * pitch + gj * step;
    const float a = u[p], b = v[p];
    const int on = (a > thr || b > thr || a < -thr || b < -thr);
    mask[(size_t)gi * gw + gj] = (uint8_t)on;
    if (on) atomicAdd(count, 1);
}

// ------------------------------------------------------------------------------------------
// k_sample_uv: the fields on the stride-`step` grid the reference's consumer reads (cpp:762-767 looks at u, v
// only where i % 4 == 0 and j % 4 == 0): out[z][i/step][j/step] = plane[z][i][j].  Results that cross PCIe shrink
// from 8 B to 8/step^2 B per pixel.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sample_uv(const float* __restrict__ u, const float* __restrict__ v, long long row_pitch,
                                                    long long pair_pitch, int step, float* __restrict__ us, float* __restrict__ vs,
                                                    int gw, int gh) {
    const int gj = blockIdx.x * 64 + threadIdx.x, gi = blockIdx.y * 4 + threadIdx.y;
    if (gj >= gw || gi >= gh) return;
    const size_t p = (size_t)blockIdx.z * pair_pitch + (size_t)(gi * step) * row_pitch + (size_t)gj * step;
    const size_t o = ((size_t)blockIdx.z * gh + gi) * gw + gj;
    us[o] = u[p];
    vs[o] = v[p];
}

// k_copy_stopped / k_relabel_stopped: EPS mode on the temporally blocked kernel.  A pair that met the criterion keeps
// its field in the ping-pong buffer it stopped in (low bit of its stop word); at the end of a call the pairs whose
// buffer is not the final one are copied over, then their words are re-labelled (a second launch: every block of the
// copy reads the old label).
__global__ void __launch_bounds__(256) k_copy_stopped(float4* __restrict__ a, float4* __restrict__ b, long long pair_f4,
                                                       const int* __restrict__ stop, int final_parity) {
    const int z = blockIdx.y;
    const int st = stop[z];
    if (!st || ((st ^ final_parity) & 1) == 0) return;          // still iterating (its field is where the blocks left it) or in place
    const float4* src = (final_parity ? a : b) + (size_t)z * pair_f4;
    float4* dst = (final_parity ? b : a) + (size_t)z * pair_f4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pair_f4; i += (long long)gridDim.x * blockDim.x)
        dst[i] = src[i];
}
__global__ void k_relabel_stopped(int* stop, int final_parity, int pairs) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z < pairs && stop[z]) stop[z] = (stop[z] & ~1) | (final_parity & 1);
}

// ------------------------------------------------------------------------------------------
// host-side launchers (called from hsflow_capi.cu)
// ------------------------------------------------------------------------------------------
cudaError_t launch_deriv(const DerivArgs& A, int fmt, int pairs, cudaStream_t s) {
    dim3 blk(32, 4), grd((A.W + 127) / 128, (A.H + 3) / 4, pairs);
    switch (fmt) {
        case FMT_GRAY8: k_deriv<FMT_GRAY8><<<grd, blk, 0, s>>>(A); break;
        case FMT_BGR8: k_deriv<FMT_BGR8><<<grd, blk, 0, s>>>(A); break;
        case FMT_F32: k_deriv<FMT_F32><<<grd, blk, 0, s>>>(A); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
cudaError_t launch_box3(const uint8_t* src, uint8_t* dst, int W, int H, long long rp, long long pp, int pairs, cudaStream_t s) {
    dim3 blk(64, 4), grd((W + 63) / 64, (H + 3) / 4, pairs);
    k_box3<<<grd, blk, 0, s>>>(src, dst, W, H, rp, pp);
    return cudaGetLastError();
}
cudaError_t launch_bgr2gray(const uint8_t* src, uint8_t* dst, int W, int H, long long rp, long long pp, int pairs, cudaStream_t s) {
    dim3 blk(64, 4), grd((W + 63) / 64, (H + 3) / 4, pairs);
    k_bgr2gray<<<grd, blk, 0, s>>>(src, dst, W, H, rp, pp);
    return cudaGetLastError();
}
cudaError_t launch_deriv_cv(const DerivArgs& A, int pairs, cudaStream_t s) {
    dim3 blk(64, 4), grd((A.W + 63) / 64, (A.H + 3) / 4, pairs);
    k_deriv_cv<<<grd, blk, 0, s>>>(A);
    return cudaGetLastError();
}

template <bool EXACT, int ST> static void launch_j1(const Jacobi1Args& A, bool upd, dim3 g, cudaStream_t s) {
    const bool track = A.stop != nullptr;
    if (upd) { if (track) k_jacobi1<EXACT, ST, true, true><<<g, 128, 0, s>>>(A); else k_jacobi1<EXACT, ST, true, false><<<g, 128, 0, s>>>(A); }
    else     { if (track) k_jacobi1<EXACT, ST, false, true><<<g, 128, 0, s>>>(A); else k_jacobi1<EXACT, ST, false, false><<<g, 128, 0, s>>>(A); }
}
cudaError_t launch_eps_check(unsigned* emax, int* stop, double eps, int sweep, int pairs, cudaStream_t s) {
    k_eps_check<<<(pairs + 127) / 128, 128, 0, s>>>(emax, stop, eps, sweep, pairs);
    return cudaGetLastError();
}
cudaError_t launch_eps_settle(int* stop, int total, int pairs, cudaStream_t s) {
    k_eps_settle<<<(pairs + 127) / 128, 128, 0, s>>>(stop, total, pairs);
    return cudaGetLastError();
}
cudaError_t launch_jacobi1(const Jacobi1Args& A, bool exact, int stencil, bool update_v, int pairs, cudaStream_t s) {
    const int rows = A.out_hi - A.out_lo;
    if (rows <= 0) return cudaSuccess;
    dim3 g((A.W + 511) / 512, (rows + A.chunk_rows - 1) / A.chunk_rows, pairs);
    if (exact) { if (stencil == ST_CL8) launch_j1<true, ST_CL8>(A, update_v, g, s); else launch_j1<true, ST_CV4>(A, update_v, g, s); }
    else       { if (stencil == ST_CL8) launch_j1<false, ST_CL8>(A, update_v, g, s); else launch_j1<false, ST_CV4>(A, update_v, g, s); }
    return cudaGetLastError();
}
cudaError_t launch_synth(uint8_t* f1, uint8_t* f2, int W, int rows, int full_h, int row0, long long rp, long long pp,
                         uint32_t seed0, int pairs, cudaStream_t s) {
    dim3 blk(64, 4), grd((W + 63) / 64, (rows + 3) / 4, pairs);
    k_synth<<<grd, blk, 0, s>>>(f1, f2, W, rows, full_h, row0, rp, pp, seed0);
    return cudaGetLastError();
}
cudaError_t launch_sample_uv(const float* u, const float* v, int W, int H, long long row_pitch, long long pair_pitch, int step,
                             float* us, float* vs, int pairs, cudaStream_t s) {
    const int gw = (W + step - 1) / step, gh = (H + step - 1) / step;
    dim3 blk(64, 4), grd((gw + 63) / 64, (gh + 3) / 4, pairs);
    k_sample_uv<<<grd, blk, 0, s>>>(u, v, row_pitch, pair_pitch, step, us, vs, gw, gh);
    return cudaGetLastError();
}
cudaError_t launch_copy_stopped(float* a, float* b, long long pair_floats, int* stop, int final_parity, int pairs, cudaStream_t s) {
    if (pairs <= 0) return cudaSuccess;
    const long long f4 = pair_floats / 4;                       // plane pitches are multiples of 32 floats
    dim3 grd((unsigned)std::min<long long>(1024, (f4 + 255) / 256), pairs);
    k_copy_stopped<<<grd, 256, 0, s>>>(reinterpret_cast<float4*>(a), reinterpret_cast<float4*>(b), f4, stop, final_parity);
    k_relabel_stopped<<<(pairs + 127) / 128, 128, 0, s>>>(stop, final_parity, pairs);
    return cudaGetLastError();
}
cudaError_t launch_dot_mask(const float* u, const float* v, int W, int H, long long pitch, int step, float thr,
                            uint8_t* mask, int* count, cudaStream_t s) {
    const int gw = (W + step - 1) / step, gh = (H + step - 1) / step;
    dim3 blk(32, 8), grd((gw + 31) / 32, (gh + 7) / 8);
    k_dot_mask<<<grd, blk, 0, s>>>(u, v, W, H, pitch, step, thr, mask, gw, gh, count);
    return cudaGetLastError();
}

}  // namespace hs
