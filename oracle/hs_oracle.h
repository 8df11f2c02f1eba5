/* oracle/hs_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, scalar fp32, -ffp-contract=off) of the Horn-Schunck hot path
 * of miczi/OpticalFlowHS.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load this library; the product
 * (opticalflowhs_b200/, libhsflow.so) never does and has no CPU fallback.
 *
 * Parity status: PINNED.
 *   - hso_derivatives / hso_jacobi (Kernels.cl semantics) are checked bit-for-bit against
 *     the reference's own Kernels.cl compiled for the host (oracle/clref_shim.cpp ->
 *     oracle/_ref/libclref.so) and against the shipped *_cl_out.jpg pictures: drawn like
 *     HSOpticalFlowOpenCL.cpp:758-770 and saved like cvSaveImage, the fields reproduce EVERY PIXEL of
 *     OpticalFlowHS/{city,bunny}_cl_out.jpg and Release/bunny_cl_out.jpg (tests/golden/pictures.npz).
 *   - hso_cvhs (OpenCV 2.1 cvCalcOpticalFlowHS, third-party cv210.dll, source NOT under
 *     /root/reference) is a restatement of the disassembled algorithm.  It is pinned by the
 *     reference's own outputs as well: with both 3x3 blurs, lambda = 0.1, 10 iterations its fields, drawn like
 *     OpticalFlowOpenCV.cpp:32-46 and saved like cvSaveImage, reproduce EVERY PIXEL of
 *     OpticalFlowHS/{city,bunny}_cv_out.jpg (threshold decisions of all grid points, end points
 *     trunc(x + u/2) of the 950 / 1955 drawn lines); N +- 1 or no blur change thousands of pixels
 *     (tests/test_oracle.py).  Not pinned: the EPS termination rule (the shipped pictures ran into the
 *     iteration cap) and use_previous.
 */
#ifndef HS_ORACLE_H_
#define HS_ORACLE_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* cvCvtColor(CV_BGR2GRAY) of OpenCV 2.1 (HSOpticalFlowOpenCL.cpp:727-728, 738-739). */
void hso_bgr2gray(const uint8_t* bgr, int w, int h, size_t pitch_bytes, uint8_t* gray);

/* readInputImage (HSOpticalFlowOpenCL.cpp:15-22): u8 gray -> float lane 0. */
void hso_u8_to_f32(const uint8_t* g, size_t n, float* out);

/* ComputeDerivativesKernel (Kernels.cl:13-39). */
void hso_derivatives(const float* I1, const float* I2, int w, int h, float* Ex, float* Ey, float* Et);

/* iterations x (u_v_avgKernel Kernels.cl:43-68 ; u_v_updateKernel Kernels.cl:71-90).
 * update_v = 0 : LITERAL (shipped kernel, v never written); 1 : FULL (north-star).
 * u, v in/out.  Returns 0. */
int hso_jacobi(float* u, float* v, const float* Ex, const float* Ey, const float* Et,
               int w, int h, float alpha, int iterations, int update_v);

/* Same iteration with general weights: ubar = w_edge*(W+E+N+S) + w_diag*(NW+NE+SW+SE),
 * t = (Ex*ubar + Ey*vbar + Et) / (rho + Ex^2 + Ey^2).  (w_edge,w_diag,rho) = (1/6,1/12,alpha^2)
 * reproduces hso_jacobi; (1/4, 0, 1/lambda) is the OpenCV-style 4-neighbour iteration. */
int hso_jacobi_general(float* u, float* v, const float* Ex, const float* Ey, const float* Et,
                       int w, int h, float w_edge, float w_diag, float rho, int iterations, int update_v);

/* hso_jacobi_general with the termination rule of cvCalcOpticalFlowHS (OpticalFlowOpenCV.cpp:29,
 * CV_TERMCRIT_ITER | CV_TERMCRIT_EPS): at most `iterations` sweeps, stop after the first sweep whose
 * max |new - old| over u and v is < eps (eps <= 0: never).  Returns the sweeps executed. */
int hso_jacobi_general_eps(float* u, float* v, const float* Ex, const float* Ey, const float* Et,
                           int w, int h, float w_edge, float w_diag, float rho, int iterations, double eps, int update_v);

/* Whole CL path: gray u8 pair -> u, v (runDerivatives + iterations x runCLKernels, cpp:748-751). */
int hso_run_cl(const uint8_t* g1, const uint8_t* g2, int w, int h, float alpha, int iterations,
               int update_v, float* u, float* v);

/* cvSmooth(CV_BLUR, 3, 3) in place semantics (OpticalFlowOpenCV.cpp:27-28): normalised 3x3 box,
 * replicate border, rounded back to u8. */
void hso_box3_u8(const uint8_t* src, int w, int h, uint8_t* dst);

/* OpenCV 2.1 cvCalcOpticalFlowHS restated (OpticalFlowOpenCV.cpp:29).
 * Returns the number of iterations executed.  eps <= 0 disables the EPS criterion. */
int hso_cvhs(const uint8_t* A, const uint8_t* B, int w, int h, float lambda, int max_iter,
             double eps, int use_previous, float* velx, float* vely);

/* OpenCV-mode derivative estimator alone (Sobel/8 on A, It = B - A). */
void hso_cv_derivatives(const uint8_t* A, const uint8_t* B, int w, int h, float* Ix, float* Iy, float* It);

/* Timed region of OpticalFlowOpenCV.cpp:26-30: two blurs + cvCalcOpticalFlowHS. */
int hso_run_cv(const uint8_t* g1, const uint8_t* g2, int w, int h, float lambda, int max_iter,
               double eps, float* velx, float* vely);

/* Drawing predicate of HSOpticalFlowOpenCL.cpp:762-770 / OpticalFlowOpenCV.cpp:34-46:
 * mask[(i/step)*(ceil(w/step)) + j/step] = 1 where |u|>thr or |v|>thr on the stride-`step` grid.
 * Returns the number of dots. */
int hso_dot_mask(const float* u, const float* v, int w, int h, int step, float thr, uint8_t* mask);

/* Deterministic synthetic frame pair (integer arithmetic only, so the CUDA generator in
 * opticalflowhs_b200/csrc is bit-identical).  Writes rows [row0, row0+rows) of a W x H frame. */
void hso_synth_pair(int W, int H, int row0, int rows, uint32_t seed, uint8_t* f1, uint8_t* f2);

int hso_max_threads(void);
void hso_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
