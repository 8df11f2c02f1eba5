/* include/hsflow.h -- C ABI of libhsflow.so, the B200-native Horn-Schunck engine.
 *
 * This is the drop-in boundary for the hot path of miczi/OpticalFlowHS.  The reference has
 * no FFI of its own: its boundary is the C++ class HSOpticalFlowOpenCL (HSOpticalFlowOpenCL.hpp:26-264)
 * sitting on the OpenCL C API.  include/HSOpticalFlowOpenCL.hpp re-declares that class for the
 * unchanged main.cpp and implements it on the entry points below; every entry point cites the
 * reference code it replaces (paths relative to /root/reference/OpticalFlowHS/).
 *
 * Conventions: plain C types only; every function returns 0 (HSFLOW_OK) or a negative
 * HSFLOW_E* code, with a thread-local message in hsflow_last_error().  There is NO CPU
 * fallback: without a CUDA device hsflow_create() fails with HSFLOW_ENODEV.
 * All work of one handle is ordered on that handle's CUDA stream; calls taking host
 * pointers marked "sync" return when the data is usable, everything else is asynchronous
 * until hsflow_sync().
 */
#ifndef HSFLOW_H_
#define HSFLOW_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library itself is built with -fvisibility=hidden */
#endif

typedef struct hsflow hsflow_t;

enum {
    HSFLOW_OK = 0,
    HSFLOW_EINVAL = -1,  /* bad argument / call order */
    HSFLOW_ENODEV = -2,  /* no usable CUDA device (sm_100) */
    HSFLOW_ECUDA = -3,   /* CUDA runtime / driver error, see hsflow_last_error() */
    HSFLOW_ENOMEM = -4
};

/* neighbourhood of the smoothness average (BASELINE.json north_star: "4- or 8-neighbour") */
enum {
    HSFLOW_STENCIL_CL8 = 0, /* Kernels.cl:55-63  1/6 edge + 1/12 diagonal, rho = alpha^2        */
    HSFLOW_STENCIL_CV4 = 1  /* cvCalcOpticalFlowHS: 1/4 edge, rho = 1/lambda     */
};

/* arithmetic contract of the iteration kernels */
enum {
    HSFLOW_MATH_FAST = 0,  /* normalised coefficients + FMA; |du|,|dv| <= 1e-3 px vs the oracle   */
    HSFLOW_MATH_EXACT = 1  /* operand-for-operand Kernels.cl:55-63,84-86 without contraction and
                              with IEEE division: bit-identical to the host oracle.  T = 1 only.
                              With HSFLOW_DERIV_CV + HSFLOW_STENCIL_CV4 the update follows the rounding
                              sequence of cvCalcOpticalFlowHS instead (products xx, xy, yy, xt, yt and
                              1 / (rho + xx + yy) per pixel): bit-identical to the restated routine that
                              reproduces the shipped *_cv_out.jpg pixel for pixel.                    */
};

/* which derivative estimator feeds the iteration */
enum {
    HSFLOW_DERIV_CL = 0,   /* ComputeDerivativesKernel, Kernels.cl:13-39 (2x2x2 cube)            */
    HSFLOW_DERIV_CV = 1    /* OpenCV 2.1: 3x3 box blur of both frames, Sobel/8 on frame 1,
                              It = frame2 - frame1 (OpticalFlowOpenCV.cpp:27-29)  */
};

enum { HSFLOW_PHASE_LOAD = 0, HSFLOW_PHASE_DERIV = 1, HSFLOW_PHASE_ITER = 2, HSFLOW_PHASE_READ = 3 };

/* ---- library ---------------------------------------------------------------------- */
const char* hsflow_last_error(void);
int hsflow_version(void);
int hsflow_device_count(void);

/* ---- lifecycle: replaces setupCL() (cpp:67-319) and cleanup() (cpp:849-892) ---------- */
int hsflow_create(int device, hsflow_t** out);
int hsflow_destroy(hsflow_t* h);
/* Run on a caller-owned CUDA stream (a cudaStream_t passed as void*); NULL restores the own stream. */
int hsflow_set_stream(hsflow_t* h, void* cuda_stream);

/* ---- parameters: replaces the ctor arguments alp/it/gs (hpp:130-161) and kernel arg 7
 *      (cpp:617-621).  temporal_block = Jacobi iterations fused per launch (0 = auto). ---- */
int hsflow_set_params(hsflow_t* h, float alpha, int iterations, int stencil, int update_v, int temporal_block);
int hsflow_set_lambda(hsflow_t* h, float lambda);         /* rho = 1/lambda (cv.cpp:29)          */
int hsflow_set_math(hsflow_t* h, int math_mode);          /* HSFLOW_MATH_*                      */
int hsflow_set_deriv(hsflow_t* h, int deriv_mode);        /* HSFLOW_DERIV_*                     */
/* 0 = auto.  warps_per_cta is accepted and ignored: the streaming kernel always runs one autonomous warp per CTA.
 * sub_batch (pairs per launch / scratch size) takes effect at the next hsflow_configure. */
int hsflow_set_tuning(hsflow_t* h, int chunk_rows, int warps_per_cta, int sub_batch);
int hsflow_set_warm_start(hsflow_t* h, int keep_uv);      /* use_previous (cv.h:481-483)        */
/* The EPS half of cvTermCriteria(CV_TERMCRIT_ITER | CV_TERMCRIT_EPS, it, 1e-6) (OpticalFlowOpenCV.cpp:29,
 * 94): every pair stops after the first sweep whose max |new - old| over u and v is < eps, or after
 * `iterations` sweeps.  eps <= 0 (default) = ITER only, as runCLKernels (cpp:750-751).  The decision is taken
 * on the device (max-norm reduction inside the iteration kernel, a stop word per pair): no host round trip, the
 * launch sequence stays asynchronous.  FAST math keeps the temporally blocked kernel: blocks of up to 4 sweeps that
 * track the max-norm of every sweep, and a pair that met the criterion inside a block is replayed from the block's
 * input for exactly that many sweeps -- sweep count and field are bit-identical to checking after every sweep.
 * EXACT math runs one sweep per launch.  Not available in strip mode. */
int hsflow_set_epsilon(hsflow_t* h, double eps);
/* 0 = auto; 1 = single-sweep kernel only (one launch per iteration); 2 = streaming kernel even for T = 1 */
int hsflow_set_kernel(hsflow_t* h, int which);

/* CUDA graphs: when hsflow_compute is asked for the same computation again (geometry, parameters, frame planes), the
 * second call captures the launch sequence and every later one replays it with a single cudaGraphLaunch -- the win
 * for launch-bound jobs (a 600 x 480 pair x 100 iterations is 26 launches).  0 = auto (jobs up to 4 Mi pixels),
 * 1 = never, 2 = always. */
int hsflow_set_graph(hsflow_t* h, int mode);

/* ---- geometry: W x H frames, `pairs` independent frame pairs per handle ------------------ */
int hsflow_configure(hsflow_t* h, int width, int height, int pairs);
/* Row-strip of a taller frame: the handle's H rows are [own rows + ghost rows]; edges that are
 * not true image edges (is_top/is_bottom = 0) are fed by the caller's halo exchange, see
 * hsflow_iterate().  Default after configure: both are true edges. */
int hsflow_set_strip(hsflow_t* h, int is_top_edge, int is_bottom_edge);

/* ---- ingest: replaces readInputImage/readInputFrame (cpp:6-64), cvCvtColor (cpp:727-728) and the
 *      clEnqueueWriteBuffer of both frames (cpp:339-357).  Host pointers; pitch in bytes. ------- */
int hsflow_set_frames_gray8(hsflow_t* h, int pair, const uint8_t* f1, const uint8_t* f2, size_t pitch);
int hsflow_set_frames_bgr8(hsflow_t* h, int pair, const uint8_t* f1, const uint8_t* f2, size_t pitch);
int hsflow_set_frames_f32(hsflow_t* h, int pair, const float* f1, const float* f2, size_t pitch);
/* Same with device pointers (frames already in HBM). */
int hsflow_set_frames_gray8_dev(hsflow_t* h, int pair, const uint8_t* d_f1, const uint8_t* d_f2, size_t pitch);
int hsflow_set_frames_bgr8_dev(hsflow_t* h, int pair, const uint8_t* d_f1, const uint8_t* d_f2, size_t pitch);
/* Zero-copy ingest (cvLoadImage + cvCvtColor, cpp:721-728, without a host bounce): allocates the handle's frame
 * planes in `frame_format` and returns them, so that an on-GPU decoder (nvJPEG: see hsflow_host_load_pair_jpeg in
 * libhsflow_host.so) writes its output straight into them; BGR planes are converted to gray inside the derivative
 * kernel.  Plane layout: [pair][row][row_pitch bytes].  Call again (or any set_frames) after rewriting the planes. */
enum { HSFLOW_FRAMES_GRAY8 = 0, HSFLOW_FRAMES_BGR8 = 1 };
int hsflow_map_frames(hsflow_t* h, int frame_format, uint8_t** d_f1, uint8_t** d_f2, size_t* row_pitch, size_t* pair_pitch);
/* The second-frame planes become the first-frame planes and vice versa (cpp:834 memcpy I2 -> I1 as a pointer swap). */
int hsflow_swap_frames(hsflow_t* h);
/* Device-side synthetic frames for benches and large-frame tests (bit-identical to
 * oracle hso_synth_pair): rows [row0, row0+H) of a W x full_height frame, seed0 + pair. */
int hsflow_synth_frames(hsflow_t* h, int full_height, int row0, uint32_t seed0);
/* One-call form used by the class: configure(w,h,1) + set_frames(0). */
int hsflow_load_pair_gray8(hsflow_t* h, const uint8_t* f1, const uint8_t* f2, int w, int hgt, size_t pitch);
int hsflow_load_pair_bgr8(hsflow_t* h, const uint8_t* f1, const uint8_t* f2, int w, int hgt, size_t pitch);
int hsflow_load_pair_f32(hsflow_t* h, const float* f1, const float* f2, int w, int hgt, size_t pitch);

/* ---- compute ---------------------------------------------------------------------------
 * hsflow_compute = the timed region of run() (cpp:748-751): runDerivatives() once, then
 * runCLKernels() x iterations, for every pair of the handle, without the per-iteration PCIe
 * round trip.  It is hsflow_prepare() followed by hsflow_iterate(iterations). */
int hsflow_compute(hsflow_t* h);
/* The same for pairs [p0, p0 + n) only (n <= hsflow_sub_batch()), asynchronously; the other pairs' frames and fields are
 * not touched (fields of pairs no range has computed are unspecified).  Lets a caller overlap ingest of one half of
 * the pair slots with compute on the other half. */
int hsflow_compute_range(hsflow_t* h, int p0, int n);
int hsflow_prepare(hsflow_t* h);            /* runDerivatives (cpp:321-474): coefficients; u = v = 0 */
int hsflow_iterate(hsflow_t* h, int n);     /* n x runCLKernels (cpp:476-679)                         */
/* Strip mode: after the caller refreshed the ghost rows of the current u/v planes. */
int hsflow_halo_refreshed(hsflow_t* h);
int hsflow_sync(hsflow_t* h);

/* ---- row strips with PEER transport (BASELINE.json north_star: "halo exchange over NVLink via ... P2P stores").
 * The reference keeps the whole frame on one device (cpp:158 devices[0]); this is the multi-GPU form of
 * runCLKernels (cpp:476-679) for one very large frame.  Each strip publishes an opaque handle
 * (CUDA IPC handles of its two u/v buffers and of its signal words; plain pointers for strips living in
 * the same process); the caller carries the handles to the neighbours (any transport: torch.distributed
 * object all-gather, a pipe, shared memory) and connects.  From then on hsflow_iterate() needs no halo
 * call at all: the iteration kernel stores rows [lo, hi) of its output ALSO into the neighbour's buffer at
 * row + delta over NVLink, and the last thread block to finish publishes an epoch number in the
 * neighbours' signal words; each strip's stream waits on its own words (cuStreamWaitValue32) between
 * launches.  Ghost rows per seam must be >= the temporal block.  All strips must issue the same sequence
 * of prepare/iterate calls.  NULL handle = no neighbour on that side (true image edge). */
typedef struct hsflow_strip_handle { unsigned char opaque[320]; } hsflow_strip_handle_t;
int hsflow_strip_export(hsflow_t* h, hsflow_strip_handle_t* out);
int hsflow_strip_connect(hsflow_t* h,
                         const hsflow_strip_handle_t* up, int up_row_lo, int up_row_hi, int up_row_delta,
                         const hsflow_strip_handle_t* down, int down_row_lo, int down_row_hi, int down_row_delta);
int hsflow_strip_disconnect(hsflow_t* h);   /* call on every strip before any of them is reconfigured or destroyed */

/* ---- results: replaces clEnqueueReadBuffer of u,v (cpp:655-675) / Ex,Ey,Et (cpp:437-468) - */
int hsflow_read_uv(hsflow_t* h, int pair, float* u, float* v, size_t pitch);              /* sync */
int hsflow_read_derivatives(hsflow_t* h, int pair, float* Ex, float* Ey, float* Et, size_t pitch); /* sync */
int hsflow_write_uv(hsflow_t* h, int pair, const float* u, const float* v, size_t pitch); /* warm start */
/* Current device planes: element pitch of a row and of a pair (floats). */
int hsflow_get_device_uv(hsflow_t* h, float** d_u, float** d_v, size_t* row_pitch, size_t* pair_pitch);
int hsflow_get_device_frames(hsflow_t* h, uint8_t** d_f1, uint8_t** d_f2, size_t* row_pitch, size_t* pair_pitch);
/* Drawing predicate of cpp:762-765 on the device: mask[(i/step)*ceil(W/step) + j/step]. */
int hsflow_dot_mask(hsflow_t* h, int pair, int step, float threshold, uint8_t* mask, int* count); /* sync */

/* Consumer-shaped read-back: u, v of one pair on the stride-`step` grid only -- the reference's only consumer
 * looks at u[i*w+j], v[i*w+j] for i % 4 == 0, j % 4 == 0 (cpp:762-767).  u_s, v_s: ceil(H/step) x ceil(W/step)
 * floats each, dense.  8/step^2 bytes per pixel cross PCIe instead of 8. */
int hsflow_sample_uv(hsflow_t* h, int pair, int step, float* u_s, float* v_s);                    /* sync */

/* ---- pipelined host-to-host batch: H2D, compute and D2H of consecutive pairs overlap --------
 * frames: n_pairs x 2 gray8 images (f1 then f2, densely packed W*H each) in host memory;
 * u_out/v_out: n_pairs x W*H floats.  Pinned host memory gives full PCIe rate.
 * These calls RECONFIGURE the handle (geometry w x hgt, an internal number of pair slots) and leave no current field
 * on the device: hsflow_read_uv / hsflow_dot_mask / hsflow_get_device_uv fail with HSFLOW_EINVAL until the next
 * hsflow_configure + hsflow_compute.  They return only after every copy from / into the caller's buffers has
 * finished, on success and on error alike. */
int hsflow_run_batch_host(hsflow_t* h, const uint8_t* frames, int n_pairs, int w, int hgt, float* u_out, float* v_out);

/* ---- frame sequences: replaces the camera loop (cpp:800-842), where every grabbed frame is paired with
 *      the previous one and then becomes the previous one itself (cpp:834 memcpy I2 -> I1) ---------------
 * Batch form: frames = n_frames consecutive gray8 images (W*H each); pair k = (frame k, frame k+1), so
 * u_out/v_out receive n_frames - 1 fields.  Same three-stage pipeline as hsflow_run_batch_host, but a
 * frame crosses PCIe once and is never copied on the device: the second-frame plane of a sub-batch is its
 * first-frame plane shifted by one frame. */
int hsflow_run_sequence_host(hsflow_t* h, const uint8_t* frames, int n_frames, int w, int hgt, float* u_out, float* v_out);
/* General form of the two calls above.  frame_format: HSFLOW_FRAMES_GRAY8 (1 byte per pixel) or HSFLOW_FRAMES_BGR8
 * (3 bytes, cvLoadImage order; gray conversion of cpp:727-728 fused into the derivative kernel).  flags:
 * HSFLOW_PIPE_SEQUENCE = `frames` holds n_pairs + 1 consecutive frames.  sample_step > 0: u_out / v_out receive the
 * stride-`step` samples only, n_pairs x ceil(hgt/step) x ceil(w/step) floats each (see hsflow_sample_uv). */
enum { HSFLOW_PIPE_SEQUENCE = 1 };
int hsflow_run_pipeline_host(hsflow_t* h, const uint8_t* frames, int n_pairs, int w, int hgt, int frame_format, int flags,
                             int sample_step, float* u_out, float* v_out);
/* The same over several handles at once -- usually one per GPU of the box: pair k goes to handle k / ceil(n_pairs /
 * n_handles) (contiguous blocks), every handle runs its block on its own host thread, no data crosses
 * between GPUs.  The reference owns exactly one device (cpp:158 devices[0]); this is its batch loop over all of them.
 * The caller creates the handles (hsflow_create(device_k, ..)) and sets the same parameters on each. */
int hsflow_run_pipeline_host_multi(hsflow_t* const* handles, int n_handles, const uint8_t* frames, int n_pairs, int w, int hgt,
                                   int frame_format, int flags, int sample_step, float* u_out, float* v_out);
/* Streaming form for a handle configured with one pair: the current second frame becomes the first
 * (pointer swap in HBM), `frame` is uploaded as the new second frame; then hsflow_compute as usual.
 * The very first frame pushed fills both planes (zero flow). */
int hsflow_push_frame_gray8(hsflow_t* h, const uint8_t* frame, size_t pitch);

/* ---- instrumentation ------------------------------------------------------------------ */
float hsflow_last_ms(hsflow_t* h, int phase);         /* CUDA-event time of the last call's phase  */
long long hsflow_kernel_launches(hsflow_t* h);        /* kernels launched by this handle so far    */
int hsflow_iterations_done(hsflow_t* h, int pair, int* done);  /* sweeps a pair ran since prepare (< iterations
                                                                   when hsflow_set_epsilon stopped it).  sync */
int hsflow_effective_temporal_block(hsflow_t* h);
int hsflow_sub_batch(hsflow_t* h);                    /* pairs per launch chosen by hsflow_configure          */
int hsflow_device(hsflow_t* h);                       /* CUDA device the handle lives on                      */
void* hsflow_alloc_pinned(size_t bytes);              /* cudaMallocHost / cudaFreeHost helpers     */
void hsflow_free_pinned(void* p);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* HSFLOW_H_ */
