// hs_common.cuh -- shared device helpers of the B200 Horn-Schunck kernels (sm_100a).
//
// Arithmetic is written with explicit round-to-nearest intrinsics (__fadd_rn, __fmul_rn,
// __fmaf_rn, __fdiv_rn) so that nvcc never re-associates or contracts: the single-sweep kernel
// (hs_kernels.cu: k_jacobi1) and the temporally blocked streaming kernel (hs_stream.cu)
// evaluate the SAME operation sequence per pixel and are bit-identical to each other, and
// the EXACT variant is operand-for-operand Kernels.cl:55-63, 84-86 (reference paths relative
// to /root/reference/OpticalFlowHS/).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hs {

constexpr int kStripW = 128;    // columns one warp streams (32 lanes x 4 px)
constexpr unsigned kFull = 0xffffffffu;

enum Stencil { ST_CL8 = 0, ST_CV4 = 1 };
enum FrameFmt { FMT_GRAY8 = 0, FMT_BGR8 = 1, FMT_F32 = 2 };

// ---- FAST formulation -------------------------------------------------------------------
// 8-neighbour (Kernels.cl:55-58 with c6 == 2*c12 exactly in fp32):
//   ubar(r) = c12 * ( (G(r-1) + 2 h(r)) + G(r+1) ),  h = W + E,  G = 2 c + h
// 4-neighbour (cvCalcOpticalFlowHS):
//   ubar(r) = 0.25 * ( (c(r-1) + h(r)) + c(r+1) )
// update with normalised coefficients a,b,c = (Ex,Ey,Et)/sqrt(rho+Ex^2+Ey^2) (Kernels.cl:84-86):
//   t = a*ubar + b*vbar + c ; u = ubar - a t ; v = vbar - b t
template <int ST> __device__ __forceinline__ float rowG(float c, float h) {
    return ST == ST_CL8 ? __fmaf_rn(2.0f, c, h) : c;
}
template <int ST> __device__ __forceinline__ float pOf(float gprev, float h) {
    return ST == ST_CL8 ? __fmaf_rn(2.0f, h, gprev) : __fadd_rn(gprev, h);
}
template <int ST> __device__ __forceinline__ float combine(float p, float G) {
    return __fmul_rn(ST == ST_CL8 ? (float)(1.0 / 12) : 0.25f, __fadd_rn(p, G));
}
__device__ __forceinline__ void update_fast(float ub, float vb, float a, float b, float c, float& un, float& vn) {
    const float t = __fmaf_rn(a, ub, __fmaf_rn(b, vb, c));
    un = __fmaf_rn(-a, t, ub);
    vn = __fmaf_rn(-b, t, vb);
}
__device__ __forceinline__ void normalise_coefs(float ex, float ey, float et, float rho, float& a, float& b, float& c) {
    const float den = __fmaf_rn(ey, ey, __fmaf_rn(ex, ex, rho));
    const float r = __frsqrt_rn(den);
    a = __fmul_rn(ex, r);
    b = __fmul_rn(ey, r);
    c = __fmul_rn(et, r);
}

// ---- packed FP32 (Blackwell FFMA2 / FADD2 / FMUL2: fma.rn.f32x2 etc., sm_100+) --------------------
// One instruction works on TWO adjacent pixels held in an aligned register pair.  Each half is the
// IEEE round-to-nearest result of the scalar operation, so the packed FAST formulation is bit-identical
// to the scalar one above (k_jacobi1) while issuing half the arithmetic instructions: the streaming
// kernel is bound by instruction issue, not by the FMA pipe (tools/microbench/ffma2_bench.cu: FFMA2
// has the flop rate of FFMA at half the issue slots).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void unpk2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// h = W + E for the four pixels of a lane (l, r: the neighbouring lanes' edge pixels), as two packed pairs
__device__ __forceinline__ void hsum4(const float (&c)[4], float l, float r, f32x2 (&h)[2]) {
    h[0] = pk2(__fadd_rn(l, c[1]), __fadd_rn(c[0], c[2]));
    h[1] = pk2(__fadd_rn(c[1], c[3]), __fadd_rn(c[2], r));
}
template <int ST> __device__ __forceinline__ f32x2 rowG2(f32x2 c, f32x2 h) {
    return ST == ST_CL8 ? fma2(pk2(2.0f, 2.0f), c, h) : c;
}
template <int ST> __device__ __forceinline__ f32x2 pOf2(f32x2 gprev, f32x2 h) {
    return ST == ST_CL8 ? fma2(pk2(2.0f, 2.0f), h, gprev) : add2(gprev, h);
}
template <int ST> __device__ __forceinline__ f32x2 combine2(f32x2 p, f32x2 G) {
    const float k = ST == ST_CL8 ? (float)(1.0 / 12) : 0.25f;
    return mul2(pk2(k, k), add2(p, G));
}
// t = a*ub + (b*vb + c); u = ub - a t; v = vb - b t for two pixels (na, nb = -a, -b: the negation folds into FFMA2)
__device__ __forceinline__ void update_fast2(f32x2 ub, f32x2 vb, f32x2 a, f32x2 b, f32x2 c, f32x2 na, f32x2 nb, f32x2& un, f32x2& vn) {
    const f32x2 t = fma2(a, ub, fma2(b, vb, c));
    un = fma2(na, t, ub);
    vn = fma2(nb, t, vb);
}

// ---- EXACT formulation: Kernels.cl:55-58 and 84-86, one rounding per operator -------------
__device__ __forceinline__ float avg_exact(float we, float wd, float W, float E, float N, float S,
                                           float NW, float NE, float SW, float SE) {
    const float e = __fadd_rn(__fadd_rn(__fadd_rn(W, E), N), S);
    const float d = __fadd_rn(__fadd_rn(__fadd_rn(NW, NE), SW), SE);
    return __fadd_rn(__fmul_rn(we, e), __fmul_rn(wd, d));
}
__device__ __forceinline__ void update_exact(float ub, float vb, float ex, float ey, float et, float rho,
                                             float& un, float& vn) {
    float t = __fadd_rn(__fadd_rn(__fmul_rn(ex, ub), __fmul_rn(ey, vb)), et);
    const float den = __fadd_rn(__fadd_rn(rho, __fmul_rn(ex, ex)), __fmul_rn(ey, ey));
    t = __fdiv_rn(t, den);
    un = __fsub_rn(ub, __fmul_rn(ex, t));
    vn = __fsub_rn(vb, __fmul_rn(ey, t));
}

// cvCalcOpticalFlowHS (OpenCV 2.1, restated in oracle/hs_oracle.c hso_cvhs from the disassembly), operand for
// operand: products xx, xy, yy, xt, yt and a = 1 / (rho + xx + yy) as the routine stores them per pixel, then
//   u' = ubar - (xx ubar + xy vbar + xt) a ,  v' = vbar - (xy ubar + yy vbar + yt) a
// Mathematically the update above; numerically another rounding sequence.  The restated routine reproduces the shipped
// *_cv_out.jpg pixel for pixel, so this is the form the EXACT OpenCV-mode path uses.
__device__ __forceinline__ void update_exact_cv(float ub, float vb, float ix, float iy, float it, float rho, float& un, float& vn) {
    const float xx = __fmul_rn(ix, ix), xy = __fmul_rn(ix, iy), yy = __fmul_rn(iy, iy);
    const float xt = __fmul_rn(ix, it), yt = __fmul_rn(iy, it);
    const float a = __fdiv_rn(1.0f, __fadd_rn(__fadd_rn(rho, xx), yy));
    un = __fsub_rn(ub, __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(xx, ub), __fmul_rn(xy, vb)), xt), a));
    vn = __fsub_rn(vb, __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(xy, ub), __fmul_rn(yy, vb)), yt), a));
}

// ---- clamp-to-edge (Tex2D, Kernels.cl:2-9) inside a lane-of-4 layout ------------------------
// c[] holds columns col0..col0+3.  Make out-of-image columns replicate the edge column so
// that the W/E taps read clamped values; l/r are the neighbours fetched by shuffle.
__device__ __forceinline__ void sanitize_right(float (&c)[4], int col0, int W) {
#pragma unroll
    for (int j = 1; j < 4; ++j)
        if (col0 + j > W - 1) c[j] = c[j - 1];
}
__device__ __forceinline__ void clamp_lr(const float (&c)[4], int col0, int W, float& l, float& r) {
    if (col0 == 0) l = c[0];
    if (col0 + 3 == W - 1) r = c[3];
}

}  // namespace hs
