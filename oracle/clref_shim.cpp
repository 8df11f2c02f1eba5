// oracle/clref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Compiles the reference's OWN device code, /root/reference/OpticalFlowHS/Kernels.cl,
// verbatim for the host: the file is #include'd below (path given at build time with
// -DKERNELS_CL_PATH=...), nothing from it is copied into this repository.  The ~40 lines
// above the #include are the minimum OpenCL-C vocabulary Kernels.cl uses:
//   float4 + elementwise operators         (Kernels.cl:25-38, 55-63, 84-86)
//   __kernel / __global / uint             (Kernels.cl:13, 17)
//   get_global_id / get_global_size        (Kernels.cl:17-21)
//   convert_float4                         (Kernels.cl:8)
//   double-literal * float4  -> the literal narrows to float, as an OpenCL 1.0 device
//   without cl_khr_fp64 does               (Kernels.cl:25 "(1.0/4) * (...)")
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fopenmp (see oracle/Makefile); outputs go
// to oracle/_ref/ only.  -ffp-contract=off pins one evaluation order; the OpenCL
// compiler was free to contract (FP_CONTRACT ON), so the reference itself is only
// tolerance-defined -- this oracle is ONE valid evaluation of it.
//
// The driver functions at the bottom replay HSOpticalFlowOpenCL.cpp:748-751
// (runDerivatives once, then runCLKernels x iterations) without the PCIe copies.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

struct float4 { float s[4]; };
typedef unsigned int uint;
#define __kernel
#define __global
#define DEF_VV(op) static inline float4 operator op(const float4& a, const float4& b) { \
    float4 r; for (int k = 0; k < 4; ++k) r.s[k] = a.s[k] op b.s[k]; return r; }
DEF_VV(+) DEF_VV(-) DEF_VV(*) DEF_VV(/)
#undef DEF_VV
static inline float4 operator*(double lit, const float4& b) {   // (1.0/4) * float4
    const float f = (float)lit; float4 r; for (int k = 0; k < 4; ++k) r.s[k] = f * b.s[k]; return r; }
static inline float4 operator+(float a, const float4& b) {      // alpha*alpha + float4
    float4 r; for (int k = 0; k < 4; ++k) r.s[k] = a + b.s[k]; return r; }
static inline float4& operator/=(float4& a, const float4& b) { a = a / b; return a; }
static inline float4 convert_float4(const float4& a) { return a; }

static thread_local unsigned g_id[2], g_size[2];
static inline uint get_global_id(int d) { return g_id[d]; }
static inline uint get_global_size(int d) { return g_size[d]; }

#ifndef KERNELS_CL_PATH
#error "build with -DKERNELS_CL_PATH=\"/root/reference/OpticalFlowHS/Kernels.cl\""
#endif
#include KERNELS_CL_PATH   // <- the reference's kernels, verbatim

// The line the shipped kernel does not have (Kernels.cl:87-89 are blank): FULL mode
// (BASELINE.json north_star "and likewise for v").  Same arithmetic as Kernels.cl:84-85.
static void v_updateExt(float4* v, float4* u_avg, float4* v_avg, float4* Ex, float4* Ey,
                        float4* Et, const float alpha) {
    uint i = get_global_id(0), j = get_global_id(1), w = get_global_size(0);
    int pos = j * w + i;
    float4 t = Ex[pos] * u_avg[pos] + Ey[pos] * v_avg[pos] + Et[pos];
    t /= alpha * alpha + Ex[pos] * Ex[pos] + Ey[pos] * Ey[pos];
    v[pos] = v_avg[pos] - Ey[pos] * t;
}

template <class F> static void ndrange(int w, int h, F f) {   // clEnqueueNDRangeKernel {w,h}
#pragma omp parallel for schedule(static)
    for (int j = 0; j < h; ++j) {
        g_size[0] = (unsigned)w; g_size[1] = (unsigned)h; g_id[1] = (unsigned)j;
        for (int i = 0; i < w; ++i) { g_id[0] = (unsigned)i; f(); }
    }
}

static void widen(const float* src, float4* dst, size_t n) {    // cpp:15-22: lane 0 = value
    for (size_t k = 0; k < n; ++k) { dst[k].s[0] = src[k]; dst[k].s[1] = dst[k].s[2] = dst[k].s[3] = 0.f; }
}
static void narrow(const float4* src, float* dst, size_t n) {   // cpp:765: only .s[0] is read
    for (size_t k = 0; k < n; ++k) dst[k] = src[k].s[0];
}

extern "C" {

// ComputeDerivativesKernel over scalar planes (lane 0).  cpp:321-474.
int clref_derivatives(const float* I1, const float* I2, int w, int h, float* Ex, float* Ey, float* Et) {
    size_t n = (size_t)w * h;
    std::vector<float4> a(n), b(n), ex(n), ey(n), et(n);
    widen(I1, a.data(), n); widen(I2, b.data(), n);
    ndrange(w, h, [&] { ComputeDerivativesKernel(a.data(), b.data(), ex.data(), ey.data(), et.data()); });
    narrow(ex.data(), Ex, n); narrow(ey.data(), Ey, n); narrow(et.data(), Et, n);
    return 0;
}

// iterations x (u_v_avgKernel ; u_v_updateKernel [; v extension]).  cpp:476-679, 750-751.
// u, v are in/out (the reference zero-fills them in runDerivatives, cpp:331-332).
int clref_iterate(float* u, float* v, const float* Ex, const float* Ey, const float* Et,
                  int w, int h, float alpha, int iterations, int update_v) {
    size_t n = (size_t)w * h;
    std::vector<float4> U(n), V(n), UA(n), VA(n), ex(n), ey(n), et(n);
    widen(u, U.data(), n); widen(v, V.data(), n);
    widen(Ex, ex.data(), n); widen(Ey, ey.data(), n); widen(Et, et.data(), n);
    memset(UA.data(), 0, n * sizeof(float4)); memset(VA.data(), 0, n * sizeof(float4));
    for (int it = 0; it < iterations; ++it) {
        ndrange(w, h, [&] { u_v_avgKernel(U.data(), V.data(), UA.data(), VA.data()); });
        ndrange(w, h, [&] {
            u_v_updateKernel(U.data(), V.data(), UA.data(), VA.data(), ex.data(), ey.data(), et.data(), alpha);
            if (update_v) v_updateExt(V.data(), UA.data(), VA.data(), ex.data(), ey.data(), et.data(), alpha);
        });
    }
    narrow(U.data(), u, n); narrow(V.data(), v, n);
    return 0;
}

// Whole timed region of cpp:748-751 on float4 planes, for the CPU baseline: gray values
// in, u/v out; planes stay float4 (16 B/px) exactly as the reference keeps them.
int clref_run(const float* I1, const float* I2, int w, int h, float alpha, int iterations,
              int update_v, float* u, float* v) {
    size_t n = (size_t)w * h;
    std::vector<float4> a(n), b(n), ex(n), ey(n), et(n), U(n), V(n), UA(n), VA(n);
    widen(I1, a.data(), n); widen(I2, b.data(), n);
    memset(U.data(), 0, n * sizeof(float4)); memset(V.data(), 0, n * sizeof(float4));
    ndrange(w, h, [&] { ComputeDerivativesKernel(a.data(), b.data(), ex.data(), ey.data(), et.data()); });
    for (int it = 0; it < iterations; ++it) {
        ndrange(w, h, [&] { u_v_avgKernel(U.data(), V.data(), UA.data(), VA.data()); });
        ndrange(w, h, [&] {
            u_v_updateKernel(U.data(), V.data(), UA.data(), VA.data(), ex.data(), ey.data(), et.data(), alpha);
            if (update_v) v_updateExt(V.data(), UA.data(), VA.data(), ex.data(), ey.data(), et.data(), alpha);
        });
    }
    narrow(U.data(), u, n); narrow(V.data(), v, n);
    return 0;
}

int clref_max_threads(void);
}
#ifdef _OPENMP
#include <omp.h>
extern "C" int clref_max_threads(void) { return omp_get_max_threads(); }
extern "C" void clref_set_threads(int n) { omp_set_num_threads(n); }
#else
extern "C" int clref_max_threads(void) { return 1; }
extern "C" void clref_set_threads(int) {}
#endif
