/* oracle/hs_oracle.c -- TEST INFRASTRUCTURE ONLY (see hs_oracle.h for the contract).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fopenmp -fPIC -shared (oracle/Makefile).
 * -ffp-contract=off matters: every expression below is written in the reference's operand
 * order and must be evaluated with one rounding per operator, like oracle/clref_shim.cpp.
 * All citations are relative to /root/reference/OpticalFlowHS/.
 */
#include "hs_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- ingest ------------------------------------------------------------------------- */

/* OpenCV 2.1 cvCvtColor(BGR2GRAY) fixed point, 14 fractional bits (cpp:727-728):
 * gray = (B*1868 + G*9617 + R*4899 + 8192) >> 14.  Equal to cv2 4.13 on every fixture pixel. */
void hso_bgr2gray(const uint8_t* bgr, int w, int h, size_t pitch, uint8_t* gray) {
    for (int j = 0; j < h; ++j) {
        const uint8_t* row = bgr + (size_t)j * pitch;
        for (int i = 0; i < w; ++i) {
            int b = row[3 * i], g = row[3 * i + 1], r = row[3 * i + 2];
            gray[(size_t)j * w + i] = (uint8_t)((b * 1868 + g * 9617 + r * 4899 + 8192) >> 14);
        }
    }
}

/* cpp:15-22: pixelData[..].s[0] = (cl_float)s.val[0] */
void hso_u8_to_f32(const uint8_t* g, size_t n, float* out) {
    for (size_t k = 0; k < n; ++k) out[k] = (float)g[k];
}

/* ---- Kernels.cl restated ------------------------------------------------------------- */

/* Tex2D, Kernels.cl:2-9 (clamp-to-edge / Neumann). */
static inline float tex(const float* x, int w, int h, int i, int j) {
    if (i < 0) i = 0;
    if (j < 0) j = 0;
    if (i >= w) i = w - 1;
    if (j >= h) j = h - 1;
    return x[(size_t)j * w + i];
}

/* ComputeDerivativesKernel, Kernels.cl:25-38.  (1.0/4) is a double token that narrows to
 * float 0.25f; sums are left-associated exactly as written. */
void hso_derivatives(const float* I1, const float* I2, int w, int h, float* Ex, float* Ey, float* Et) {
    const float q = (float)(1.0 / 4);
#pragma omp parallel for schedule(static)
    for (int j = 0; j < h; ++j)
        for (int i = 0; i < w; ++i) {
            size_t pos = (size_t)j * w + i;
            float a00 = tex(I1, w, h, i, j), a10 = tex(I1, w, h, i + 1, j);
            float a01 = tex(I1, w, h, i, j + 1), a11 = tex(I1, w, h, i + 1, j + 1);
            float b00 = tex(I2, w, h, i, j), b10 = tex(I2, w, h, i + 1, j);
            float b01 = tex(I2, w, h, i, j + 1), b11 = tex(I2, w, h, i + 1, j + 1);
            Ex[pos] = q * (a10 - a00 + a11 - a01 + b10 - b00 + b11 - b01); /* cl:25-28 */
            Ey[pos] = q * (a01 - a00 + a11 - a10 + b01 - b00 + b11 - b10); /* cl:30-33 */
            Et[pos] = q * (b00 - a00 + b10 - a10 + b01 - a01 + b11 - a11); /* cl:35-38 */
        }
}

/* One field of u_v_avgKernel, Kernels.cl:55-58 with general weights. */
static inline float avg8(const float* x, int w, int h, int i, int j, float we, float wd) {
    return we * (tex(x, w, h, i - 1, j) + tex(x, w, h, i + 1, j) + tex(x, w, h, i, j - 1) + tex(x, w, h, i, j + 1)) +
           wd * (tex(x, w, h, i - 1, j - 1) + tex(x, w, h, i + 1, j - 1) + tex(x, w, h, i - 1, j + 1) +
                 tex(x, w, h, i + 1, j + 1));
}

/* The iteration with the termination rule of cvCalcOpticalFlowHS (cv.cpp:29: CV_TERMCRIT_ITER | CV_TERMCRIT_EPS):
 * stop after max_iter sweeps, or after the first sweep whose max |new - old| over u and v is < eps.
 * eps <= 0: exactly max_iter sweeps (the Kernels.cl path, cpp:750-751).  Returns the sweeps executed. */
int hso_jacobi_general_eps(float* u, float* v, const float* Ex, const float* Ey, const float* Et,
                           int w, int h, float we, float wd, float rho, int iterations, double eps, int update_v) {
    size_t n = (size_t)w * h;
    float* ua = (float*)malloc(n * sizeof(float));
    float* va = (float*)malloc(n * sizeof(float));
    if (!ua || !va) { free(ua); free(va); return -1; }
    int it = 0;
    for (; it < iterations;) {
        float emax = 0.f;
        /* u_v_avgKernel over the whole frame first (cpp:537-551) ... */
#pragma omp parallel for schedule(static)
        for (int j = 0; j < h; ++j)
            for (int i = 0; i < w; ++i) {
                size_t pos = (size_t)j * w + i;
                ua[pos] = avg8(u, w, h, i, j, we, wd);
                va[pos] = avg8(v, w, h, i, j, we, wd);
            }
        /* ... then u_v_updateKernel (cpp:623-637), Kernels.cl:84-86. */
#pragma omp parallel for schedule(static) reduction(max : emax)
        for (int j = 0; j < h; ++j)
            for (int i = 0; i < w; ++i) {
                size_t pos = (size_t)j * w + i;
                float t = Ex[pos] * ua[pos] + Ey[pos] * va[pos] + Et[pos];
                t /= rho + Ex[pos] * Ex[pos] + Ey[pos] * Ey[pos];
                float un = ua[pos] - Ex[pos] * t;
                float du = fabsf(un - u[pos]);
                if (du > emax) emax = du;
                u[pos] = un;
                if (update_v) { /* absent at Kernels.cl:87-89 */
                    float vn = va[pos] - Ey[pos] * t;
                    float dv = fabsf(vn - v[pos]);
                    if (dv > emax) emax = dv;
                    v[pos] = vn;
                }
            }
        ++it;
        if (eps > 0 && (double)emax < eps) break;
    }
    free(ua); free(va);
    return it;
}

int hso_jacobi_general(float* u, float* v, const float* Ex, const float* Ey, const float* Et,
                       int w, int h, float we, float wd, float rho, int iterations, int update_v) {
    return hso_jacobi_general_eps(u, v, Ex, Ey, Et, w, h, we, wd, rho, iterations, 0.0, update_v) < 0 ? -1 : 0;
}

int hso_jacobi(float* u, float* v, const float* Ex, const float* Ey, const float* Et,
               int w, int h, float alpha, int iterations, int update_v) {
    /* (1.0/6), (1.0/12): double tokens narrowing to float; alpha*alpha in float (cl:85). */
    return hso_jacobi_general(u, v, Ex, Ey, Et, w, h, (float)(1.0 / 6), (float)(1.0 / 12),
                              alpha * alpha, iterations, update_v);
}

int hso_run_cl(const uint8_t* g1, const uint8_t* g2, int w, int h, float alpha, int iterations,
               int update_v, float* u, float* v) {
    size_t n = (size_t)w * h;
    float* buf = (float*)malloc(5 * n * sizeof(float));
    if (!buf) return -1;
    float *I1 = buf, *I2 = buf + n, *Ex = buf + 2 * n, *Ey = buf + 3 * n, *Et = buf + 4 * n;
    hso_u8_to_f32(g1, n, I1);
    hso_u8_to_f32(g2, n, I2);
    memset(u, 0, n * sizeof(float)); /* cpp:331-332 */
    memset(v, 0, n * sizeof(float));
    hso_derivatives(I1, I2, w, h, Ex, Ey, Et);
    int rc = hso_jacobi(u, v, Ex, Ey, Et, w, h, alpha, iterations, update_v);
    free(buf);
    return rc;
}

/* ---- OpenCV 2.1 CPU path restated (third-party cv210.dll) ----------------- */

static inline int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* cvSmooth(CV_BLUR,3,3): normalised box, BORDER_REPLICATE, result rounded to nearest u8
 * (cv.cpp:27-28).  src and dst must not alias. */
void hso_box3_u8(const uint8_t* src, int w, int h, uint8_t* dst) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int s = 0;
            for (int dy = -1; dy <= 1; ++dy) {
                const uint8_t* r = src + (size_t)clampi(y + dy, 0, h - 1) * w;
                s += r[clampi(x - 1, 0, w - 1)] + r[x] + r[clampi(x + 1, 0, w - 1)];
            }
            dst[(size_t)y * w + x] = (uint8_t)lrint((double)s * (1.0 / 9.0));
        }
}

void hso_cv_derivatives(const uint8_t* A, const uint8_t* B, int w, int h, float* Ix, float* Iy, float* It) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) {
        const uint8_t* r0 = A + (size_t)clampi(y - 1, 0, h - 1) * w;
        const uint8_t* r1 = A + (size_t)y * w;
        const uint8_t* r2 = A + (size_t)clampi(y + 1, 0, h - 1) * w;
        for (int x = 0; x < w; ++x) {
            int xl = clampi(x - 1, 0, w - 1), xr = clampi(x + 1, 0, w - 1);
            int gx = (r0[xr] + 2 * r1[xr] + r2[xr]) - (r0[xl] + 2 * r1[xl] + r2[xl]);
            int gy = (r2[xl] + 2 * r2[x] + r2[xr]) - (r0[xl] + 2 * r0[x] + r0[xr]);
            size_t pos = (size_t)y * w + x;
            Ix[pos] = (float)gx * 0.125f;
            Iy[pos] = (float)gy * 0.125f;
            It[pos] = (float)((int)B[pos] - (int)r1[x]);
        }
    }
}

int hso_cvhs(const uint8_t* A, const uint8_t* B, int w, int h, float lambda, int max_iter,
             double eps, int use_previous, float* velx, float* vely) {
    size_t n = (size_t)w * h;
    float* buf = (float*)malloc(11 * n * sizeof(float));
    if (!buf) return -1;
    float *Ix = buf, *Iy = buf + n, *It = buf + 2 * n;
    float *xx = buf + 3 * n, *xy = buf + 4 * n, *yy = buf + 5 * n, *xt = buf + 6 * n, *yt = buf + 7 * n,
          *ai = buf + 8 * n, *nu = buf + 9 * n, *nv = buf + 10 * n;
    const float rho = 1.0f / lambda;
    hso_cv_derivatives(A, B, w, h, Ix, Iy, It);
    for (size_t k = 0; k < n; ++k) {
        xx[k] = Ix[k] * Ix[k]; xy[k] = Ix[k] * Iy[k]; yy[k] = Iy[k] * Iy[k];
        xt[k] = Ix[k] * It[k]; yt[k] = Iy[k] * It[k];
        ai[k] = 1.0f / (rho + xx[k] + yy[k]);
    }
    if (!use_previous) { memset(velx, 0, n * sizeof(float)); memset(vely, 0, n * sizeof(float)); }
    int iter = 0;
    for (;;) {
        float emax = 0.f;
#pragma omp parallel for schedule(static) reduction(max : emax)
        for (int y = 0; y < h; ++y) {
            int yu = clampi(y - 1, 0, h - 1), yd = clampi(y + 1, 0, h - 1);
            for (int x = 0; x < w; ++x) {
                int xl = clampi(x - 1, 0, w - 1), xr = clampi(x + 1, 0, w - 1);
                size_t pos = (size_t)y * w + x;
                float ub = (velx[(size_t)y * w + xl] + velx[(size_t)y * w + xr] + velx[(size_t)yu * w + x] +
                            velx[(size_t)yd * w + x]) * 0.25f;
                float vb = (vely[(size_t)y * w + xl] + vely[(size_t)y * w + xr] + vely[(size_t)yu * w + x] +
                            vely[(size_t)yd * w + x]) * 0.25f;
                float un = ub - (xx[pos] * ub + xy[pos] * vb + xt[pos]) * ai[pos];
                float vn = vb - (xy[pos] * ub + yy[pos] * vb + yt[pos]) * ai[pos];
                float du = fabsf(un - velx[pos]), dv = fabsf(vn - vely[pos]);
                if (du > emax) emax = du;
                if (dv > emax) emax = dv;
                nu[pos] = un; nv[pos] = vn;
            }
        }
        memcpy(velx, nu, n * sizeof(float));
        memcpy(vely, nv, n * sizeof(float));
        ++iter;
        if (max_iter > 0 && iter >= max_iter) break;
        if (eps > 0 && (double)emax < eps) break;
        if (max_iter <= 0 && eps <= 0) break;
    }
    free(buf);
    return iter;
}

int hso_run_cv(const uint8_t* g1, const uint8_t* g2, int w, int h, float lambda, int max_iter,
               double eps, float* velx, float* vely) {
    size_t n = (size_t)w * h;
    uint8_t* b = (uint8_t*)malloc(2 * n);
    if (!b) return -1;
    hso_box3_u8(g1, w, h, b);      /* cv.cpp:27 */
    hso_box3_u8(g2, w, h, b + n);  /* cv.cpp:28 */
    int it = hso_cvhs(b, b + n, w, h, lambda, max_iter, eps, 0, velx, vely); /* cv.cpp:29 */
    free(b);
    return it;
}

/* ---- drawing predicate ------------------------------------------------------------- */

int hso_dot_mask(const float* u, const float* v, int w, int h, int step, float thr, uint8_t* mask) {
    int gw = (w + step - 1) / step, cnt = 0;
    for (int i = 0; i < h; i += step)
        for (int j = 0; j < w; j += step) {
            size_t p = (size_t)i * w + j;
            int on = (u[p] > thr || v[p] > thr || u[p] < -thr || v[p] < -thr); /* cpp:765 */
            mask[(size_t)(i / step) * gw + j / step] = (uint8_t)on;
            cnt += on;
        }
    return cnt;
}

/* ---- synthetic frames (integer only; mirrored bit-for-bit by the CUDA generator) ------- */

static inline uint32_t hash32(uint32_t seed, uint32_t ix, uint32_t iy) {
    uint32_t h = seed * 0x9E3779B1u ^ (ix * 0x85EBCA77u) ^ (iy * 0xC2B2AE3Du);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}
/* bilinear value noise, cell = 2^k px; X,Y in 1/256 px (may be negative); result 0..65280 */
static inline int32_t octave(uint32_t seed, int k, int32_t X, int32_t Y) {
    int sh = 8 + k;
    int32_t cx = X >> sh, cy = Y >> sh;
    int32_t fx = (X & ((1 << sh) - 1)) >> k, fy = (Y & ((1 << sh) - 1)) >> k; /* 0..255 */
    int32_t v00 = (int32_t)(hash32(seed, (uint32_t)cx, (uint32_t)cy) & 255u);
    int32_t v10 = (int32_t)(hash32(seed, (uint32_t)(cx + 1), (uint32_t)cy) & 255u);
    int32_t v01 = (int32_t)(hash32(seed, (uint32_t)cx, (uint32_t)(cy + 1)) & 255u);
    int32_t v11 = (int32_t)(hash32(seed, (uint32_t)(cx + 1), (uint32_t)(cy + 1)) & 255u);
    int32_t top = v00 * (256 - fx) + v10 * fx, bot = v01 * (256 - fx) + v11 * fx;
    return (top * (256 - fy) + bot * fy) >> 8;
}
static inline uint8_t texture(uint32_t seed, int32_t X, int32_t Y) {
    int32_t s = 3 * octave(seed, 5, X, Y) + 3 * octave(seed + 0x632BE5ABu, 3, X, Y) +
                2 * octave(seed + 0xC6A4A793u, 2, X, Y);
    return (uint8_t)(s >> 11); /* /8 weights, /256 scale */
}
/* parabolic sine, phase p in [0,1024) -> [-1024,1024] */
static inline int32_t sinlike(int32_t p) {
    p &= 1023;
    int32_t q = p & 511, val = (q * (512 - q)) >> 6;
    return p < 512 ? val : -val;
}

void hso_synth_pair(int W, int H, int row0, int rows, uint32_t seed, uint8_t* f1, uint8_t* f2) {
#pragma omp parallel for schedule(static)
    for (int r = 0; r < rows; ++r) {
        int y = row0 + r;
        for (int x = 0; x < W; ++x) {
            int32_t X = x << 8, Y = y << 8;
            int32_t py = (int32_t)(((int64_t)y << 10) / H), px = (int32_t)(((int64_t)x << 10) / W);
            int32_t dx = 384 + ((128 * sinlike(py)) >> 10);        /* 1.5 +- 0.5 px  */
            int32_t dy = -192 + ((128 * sinlike(px + 256)) >> 10); /* -0.75 +- 0.5 px */
            f1[(size_t)r * W + x] = texture(seed, X, Y);
            f2[(size_t)r * W + x] = texture(seed, X - dx, Y - dy);
        }
    }
}

int hso_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void hso_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : 1);
#else
    (void)n;
#endif
}
