// hsflow_capi.cu -- the C ABI of libhsflow.so (declared in include/hsflow.h).
//
// Host orchestration of the Horn-Schunck path, replacing HSOpticalFlowOpenCL::setupCL /
// runDerivatives / runCLKernels / cleanup (HSOpticalFlowOpenCL.cpp:67-679, 849-892): device
// planes live in HBM for the whole computation (no per-iteration PCIe round trip, cpp:483-501,
// 655-675), scalar fp32 planes instead of float4 (4 B/px instead of 16 B/px), one CUDA stream
// per handle, TMA descriptors built once per geometry.  No CPU fallback of any kind.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <unistd.h>

#include "../../include/hsflow.h"
#include "hs_common.cuh"
#include "hs_launch.h"

using namespace hs;

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(HSFLOW_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
// Every entry point makes the handle's device current first: one process may drive handles on several GPUs
// (LocalStripSolver: row strips of one frame over the GPUs of a box from a single host thread).
#define NEED(h)                                                                                                   \
    do {                                                                                                          \
        if (!(h)) return fail(HSFLOW_EINVAL, "null handle");                                                      \
        cudaError_t e_ = cudaSetDevice((h)->device);                                                              \
        if (e_ != cudaSuccess) return fail(HSFLOW_ECUDA, "cudaSetDevice(%d): %s", (h)->device, cudaGetErrorString(e_)); \
    } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

typedef CUresult (*StreamWaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

// What one strip publishes to its neighbours (hsflow_strip_export): geometry for validation, CUDA IPC handles of the
// two u/v ping-pong buffers and of the signal words, and -- for neighbours living in the same process -- the raw
// pointers.  Must fit hsflow_strip_handle_t (include/hsflow.h).
struct StripBlob {
    uint32_t magic;
    int32_t W, H, device;
    int64_t pitch;
    int64_t pid;
    void* raw[3];                 // uvA, uvB, sig
    cudaIpcMemHandle_t mem[3];
};
static_assert(sizeof(StripBlob) <= sizeof(hsflow_strip_handle_t), "strip handle blob too large");
constexpr uint32_t kStripMagic = 0x48534631u;   // "HSF1"

// CUDA graphs for launch-bound jobs (small frames: 600x480 x 100 iterations is 1 + 25 launches of a few microseconds
// each).  hsflow_compute replays a captured graph when the same computation -- same geometry, parameters and frame
// planes -- comes again; everything that decides (the EPS criterion included) lives on the device, so the launch
// sequence is the same every time.
struct GraphKey {
    int W, H, P, S, iterations, stencil, update_v, tblock, math, deriv, kernel_sel, chunk_rows, fmt, uv_dirty;
    float rho;
    double eps;
    const void *f1, *f2;
    bool operator==(const GraphKey& o) const { return memcmp(this, &o, sizeof *this) == 0; }
};
struct GraphEntry {
    GraphKey key;
    cudaGraphExec_t exec = nullptr;
    int final_cur = 0;
    long long launches = 0;
    unsigned long long stamp = 0;
};

struct hsflow {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr, s_in = nullptr, s_out = nullptr;
    EncodeTiledFn encode = nullptr;
    int sm_count = 148;
    // parameters (defaults = main.cpp:4-8: ALPHA "15", ITERATIONS "100", LAMBDA ".1")
    float alpha = 15.f, lambda = 0.1f, rho = 225.f;
    int iterations = 100, stencil = HSFLOW_STENCIL_CL8, update_v = 1, tblock = 0;
    int math = HSFLOW_MATH_FAST, deriv = HSFLOW_DERIV_CL, kernel_sel = 0;
    int chunk_rows = 0, wpc = 0, sub_batch = 0, warm = 0;
    // geometry
    int W = 0, H = 0, P = 0, S = 0;
    int fmt = -1;
    long long f_row_pitch = 0, f_pair_pitch = 0;   // bytes
    // fp32 planes are ROW-INTERLEAVED: u and v rows alternate in one buffer ([pair][row][2][pitch]), the three
    // coefficient planes likewise ([pair][row][3][pitch]).  One TMA box then fetches a row group of all planes.
    long long pitch = 0;                           // elements of one plane row
    long long uv_rp = 0, uv_pp = 0;                // u/v: row pitch (2*pitch) and pair pitch, elements
    long long c_rp = 0, c_pp = 0;                  // coefficients: row pitch (3*pitch) and pair pitch, elements
    int top_edge = 1, bottom_edge = 1;
    // device memory
    uint8_t *f1 = nullptr, *f2 = nullptr, *fb1 = nullptr, *fb2 = nullptr, *fg1 = nullptr, *fg2 = nullptr;   // frames, blurred, gray-of-BGR
    float *uA = nullptr, *vA = nullptr, *uB = nullptr, *vB = nullptr, *c0 = nullptr, *c1 = nullptr, *c2 = nullptr;
    float* dtmp = nullptr;                         // 3 planes of one pair, for hsflow_read_derivatives
    uint8_t* d_mask = nullptr;
    float* d_sample = nullptr;                     // hsflow_sample_uv staging
    size_t sample_cap = 0;
    int* d_count = nullptr;
    size_t mask_cap = 0;
    CUtensorMap tm_uvA[3], tm_uvB[3], tm_c[3];     // [m]: TMA boxes of m + 2 rows (stream_geometry(T).rows_per_box = 2, 3, 4)
    // state
    int cur = 0;                                   // 0: uA/vA hold the current field, 1: uB/vB
    int valid_lo = 0, valid_hi = 0;
    int prepared = 0, coef_norm = -1;
    cudaEvent_t ev0[4] = {}, ev1[4] = {};
    int ev_set[4] = {};
    long long launches = 0;
    // peer transport of the row-strip mode: [0] = upper neighbour, [1] = lower neighbour
    StreamWaitValue32Fn wait_value = nullptr;
    int can_flush = 0;                             // CU_STREAM_WAIT_VALUE_FLUSH supported (cudaDevAttrCanFlushRemoteWrites)
    unsigned* sig = nullptr;                       // device words: [0] epoch from the upper neighbour, [1] from the lower, [2] done counter
    void* peer[2][3] = {};                         // neighbour's uvA, uvB, sig as seen from this device
    int peer_ipc[2] = {};                          // mapped with cudaIpcOpenMemHandle (else same-process raw pointers)
    int has_peer[2] = {};
    int push_lo[2] = {}, push_hi[2] = {}, push_delta[2] = {};
    int connected = 0;
    unsigned epoch = 0;
    unsigned waited_epoch = 0;                     // neighbours' epoch the stream has already waited for (cuStreamWaitValue32)
    // EPS termination (hsflow_set_epsilon): per-pair device words, see Jacobi1Args
    double eps = 0.0;
    unsigned* d_emax = nullptr;
    int* d_stop = nullptr;
    int eps_cap = 0;                               // pairs the two arrays hold
    int sweeps = 0;                                // sweeps since hsflow_prepare (hsflow_iterate path)
    int ec_on = 0, ec_sweep = 0, ec_total = 0, ec_off = 0;   // tracking context of the sweep run_block launches next
    int ec_blk = 0;                                // EPS on the streaming kernel: blocks launched since the words were reset (emax bank = parity)
    // host pipeline: the last launch of a sub-batch stores PLANAR, densely packed fields ([pair][H][W] of u, then of v)
    // into a staging slot, so that the read-back is two contiguous copies (56.9 GB/s over PCIe against 51.8 GB/s for
    // the row-by-row de-interleaving 2-D copy out of the u|v buffer, tools/microbench/d2h_bench.cu)
    float* stage = nullptr;
    size_t stage_bytes = 0;
    float *ov_u = nullptr, *ov_v = nullptr;        // output override of the launch run_block issues next
    int uv_dirty = 0;                              // hsflow_write_uv stored a v field since the last zeroing
    // u = v = 0 at the start of a computation (cpp:331-332) costs no memory traffic on the streaming kernel: its first
    // launch requests u/v boxes of a pair index beyond the tensor map, which the TMA unit answers with zero fill
    // without touching HBM (kZeroPair).  zero_pending = "the current buffer is logically zero but was not written":
    // everything else that looks at the buffer (read-back, single-sweep kernel, halo exchange by the caller)
    // materialises the zeros first.
    int graph_mode = 0;                            // hsflow_set_graph: 0 auto (small jobs), 1 never, 2 always
    int capturing = 0;                             // inside cudaStreamBeginCapture: no event timing
    std::vector<GraphEntry> graphs;                // at most kMaxGraphs, least recently used goes first
    GraphKey eager_key;                            // the computation that last ran eagerly (a repeat gets captured)
    int have_eager_key = 0;
    unsigned long long graph_stamp = 0;
    int zero_pending = 0;
    int uv_valid = 0;                              // the A/B planes hold a field a caller may read (not after pipelined calls)
    int coef_zero_b = 0;                           // the coefficient planes in memory were written with b = 0
};

static void strip_disconnect(hsflow* h) {
    for (int d = 0; d < 2; ++d) {
        if (h->has_peer[d] && h->peer_ipc[d])
            for (int k = 0; k < 3; ++k) if (h->peer[d][k]) cudaIpcCloseMemHandle(h->peer[d][k]);
        for (int k = 0; k < 3; ++k) h->peer[d][k] = nullptr;
        h->has_peer[d] = h->peer_ipc[d] = 0;
    }
    h->connected = 0;
    cudaGetLastError();
}

constexpr size_t kMaxGraphs = 4;
static void drop_graphs(hsflow* h) {               // device pointers baked into the graphs are about to die
    for (GraphEntry& g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    h->graphs.clear();
    h->have_eager_key = 0;
}

static void free_planes(hsflow* h) {
    drop_graphs(h);
    if (h->connected) strip_disconnect(h);         // the neighbours' mappings of OUR buffers die with the buffers: reconnect
    cudaFree(h->f1); cudaFree(h->f2); cudaFree(h->fb1); cudaFree(h->fb2); cudaFree(h->fg1); cudaFree(h->fg2);
    h->fg1 = h->fg2 = nullptr;
    cudaFree(h->uA); cudaFree(h->uB); cudaFree(h->c0); cudaFree(h->dtmp);   // vA, vB, c1, c2 point into these
    cudaFree(h->stage); h->stage = nullptr; h->stage_bytes = 0;
    h->f1 = h->f2 = h->fb1 = h->fb2 = nullptr;
    h->uA = h->vA = h->uB = h->vB = h->c0 = h->c1 = h->c2 = h->dtmp = nullptr;
    h->fmt = -1;
    h->prepared = 0;
}

// temporal_block = 0.  Throughput regime (enough work units for two full waves of resident warps even at the
// smallest chunk height): the deepest block that still runs without spills.  Latency regime (single frames up to
// about 4K): a launch is one partly filled wave whose length is (chunk + 2T) x T stage-rows, so a shallower block wins
// (tools/small_frame_probe.py: 600x480 x 100 iterations 0.31 ms at T = 4 against 0.48 ms at T = 6).  Strips keep the
// default: every strip of a frame has to issue the same launch sequence.
static int auto_T(const hsflow* h) {
    if (h->W <= 0 || h->connected || !h->top_edge || !h->bottom_edge) return kDefaultT;
    const StreamGeom G = stream_geometry(kDefaultT);
    const long long nsx = (h->W + G.valid_w - 1) / G.valid_w;
    const long long units = nsx * ((h->H + 4 * kDefaultT - 1) / (4 * kDefaultT)) * std::max(1, std::min(h->P, h->S));
    if (units < 2LL * h->sm_count * 8) return kSmallT;
    const long long px = (long long)h->W * h->H * std::max(1, std::min(h->P, h->S));
    return px >= kBigPixels ? kBigT : kDefaultT;
}
// Iterations of the next launch when `left` remain and blocks hold at most T: the ceil(left / T) launches are made as
// equal as possible (100 iterations at T = 8: 9 x 8 + 4 x 7 instead of 12 x 8 + 4, whose last block would run at the
// speed of the shallow T = 4 instantiation).  Any partition gives the same bits: streaming == direct sweeps.
static int next_block(int left, int T) {
    const int blocks = (left + T - 1) / T;
    return (left + blocks - 1) / blocks;
}
// LITERAL mode (update_v = 0: the shipped u_v_updateKernel never writes v, Kernels.cl:87-89) on the FULL streaming
// kernel.  While v is identically zero, "v stays what it was" and "v' = vbar - b t with b = 0" are the same thing, and
// t = a ubar + b vbar + c is a ubar + c either way: the derivative pass stores b = 0 (the normalisation still uses
// Ex^2 + Ey^2) and the temporally blocked kernel runs unchanged -- bit-identical to k_jacobi1<.., UPDATE_V = false>
// (fma(b, +0, c) == c == fma(0, +0, c) since c is never -0), 17 launches per 100 iterations instead of 100.
// Needs v == 0: no warm start, no hsflow_write_uv of a v field.
static bool literal_on_stream(const hsflow* h) {
    return !h->update_v && h->math == HSFLOW_MATH_FAST && !h->warm && !h->uv_dirty;
}
static bool use_stream_kernel(const hsflow* h, int t) {
    if (h->math == HSFLOW_MATH_EXACT || (!h->update_v && !literal_on_stream(h)) || h->kernel_sel == 1) return false;
    (void)t;                                       // also for a single iteration: at T = 1 the TMA-fed streaming kernel
    return true;                                   // moves 5.3 TB/s where k_jacobi1 moves 4.6-4.8 (tools/gpu_perf_probe.py)
}
// EPS criterion (hsflow_set_epsilon) on the temporally blocked kernel: the TRACK instantiation (StreamArgs::stop ...),
// blocks of up to kTrackT sweeps, each a main launch + a replay launch.  EXACT math keeps the single-sweep kernel.
static bool eps_on_stream(const hsflow* h) { return h->eps > 0.0 && use_stream_kernel(h, 1); }
static int effective_T(const hsflow* h) {
    if (!use_stream_kernel(h, 1)) return 1;
    if (h->eps > 0.0) return h->tblock > 0 ? std::min(h->tblock, kTrackT) : kTrackT;
    int T = h->tblock > 0 ? h->tblock : auto_T(h);
    return std::min(T, kMaxT);
}

// 4-D map {W, planes, H, pairs} over a row-interleaved buffer; box = 128 columns x all planes x box_rows rows
static int make_map(hsflow* h, CUtensorMap* tm, float* base, int planes, int pairs, int box_rows) {
    cuuint64_t dims[4] = {(cuuint64_t)h->W, (cuuint64_t)planes, (cuuint64_t)h->H, (cuuint64_t)pairs};
    cuuint64_t strides[3] = {(cuuint64_t)h->pitch * 4, (cuuint64_t)h->pitch * planes * 4,
                             (cuuint64_t)h->pitch * planes * h->H * 4};
    cuuint32_t box[4] = {(cuuint32_t)kStripW, (cuuint32_t)planes, (cuuint32_t)box_rows, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = h->encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HSFLOW_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return HSFLOW_OK;
}

constexpr int kZeroPair = 1 << 28;                 // pair coordinate outside every tensor map: the box is all zero fill

static int materialize_zero(hsflow* h) {
    if (!h->zero_pending) return HSFLOW_OK;
    h->zero_pending = 0;
    float* uv = h->cur == 0 ? h->uA : h->uB;       // u and v rows share the buffer
    CK(cudaMemsetAsync(uv, 0, (size_t)h->uv_pp * std::min(h->P, h->cur == 0 ? h->P : h->S) * sizeof(float), h->stream));
    return HSFLOW_OK;
}

static void phase_begin(hsflow* h, int ph) { if (!h->capturing) cudaEventRecord(h->ev0[ph], h->stream); }
static void phase_end(hsflow* h, int ph) { if (!h->capturing) { cudaEventRecord(h->ev1[ph], h->stream); h->ev_set[ph] = 1; } }

extern "C" {

const char* hsflow_last_error(void) { return g_err; }
int hsflow_version(void) { return 100; }
int hsflow_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int hsflow_create(int device, hsflow_t** out) {
    if (!out) return fail(HSFLOW_EINVAL, "out is null");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(HSFLOW_ENODEV, "no CUDA device: libhsflow has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(HSFLOW_EINVAL, "device %d out of range (0..%d)", device, n - 1);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(HSFLOW_ENODEV, "device %d is sm_%d%d; libhsflow is built for sm_100a only", device, prop.major, prop.minor);
    hsflow* h = new (std::nothrow) hsflow();
    if (!h) return fail(HSFLOW_ENOMEM, "out of host memory");
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
        delete h;
        return fail(HSFLOW_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
    }
    h->encode = (EncodeTiledFn)fn;
    fn = nullptr;
    e = cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) h->wait_value = (StreamWaitValue32Fn)fn;
    else cudaGetLastError();
    int flush = 0;
    if (cudaDeviceGetAttribute(&flush, cudaDevAttrCanFlushRemoteWrites, device) == cudaSuccess) h->can_flush = flush;
    else cudaGetLastError();
    e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete h; return fail(HSFLOW_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    h->stream = h->own_stream;
    for (int i = 0; i < 4; ++i) { cudaEventCreate(&h->ev0[i]); cudaEventCreate(&h->ev1[i]); }
    e = stream_prepare(device);
    if (e != cudaSuccess) { hsflow_destroy(h); return fail(HSFLOW_ECUDA, "kernel attribute setup: %s", cudaGetErrorString(e)); }
    if (cudaMalloc(&h->d_count, sizeof(int)) != cudaSuccess) { hsflow_destroy(h); return fail(HSFLOW_ENOMEM, "cudaMalloc"); }
    *out = h;
    return HSFLOW_OK;
}

int hsflow_destroy(hsflow_t* h) {
    if (!h) return HSFLOW_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    free_planes(h);
    cudaFree(h->sig);
    cudaFree(h->d_emax); cudaFree(h->d_stop);
    cudaFree(h->d_mask); cudaFree(h->d_count); cudaFree(h->d_sample);
    for (int i = 0; i < 4; ++i) { if (h->ev0[i]) cudaEventDestroy(h->ev0[i]); if (h->ev1[i]) cudaEventDestroy(h->ev1[i]); }
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    delete h;
    return HSFLOW_OK;
}

int hsflow_set_stream(hsflow_t* h, void* s) {
    NEED(h);
    h->stream = s ? (cudaStream_t)s : h->own_stream;
    return HSFLOW_OK;
}

int hsflow_set_params(hsflow_t* h, float alpha, int iterations, int stencil, int update_v, int temporal_block) {
    NEED(h);
    if (iterations < 0) return fail(HSFLOW_EINVAL, "iterations must be >= 0");
    if (stencil != HSFLOW_STENCIL_CL8 && stencil != HSFLOW_STENCIL_CV4) return fail(HSFLOW_EINVAL, "unknown stencil %d", stencil);
    if (temporal_block < 0 || temporal_block > kMaxT) return fail(HSFLOW_EINVAL, "temporal_block must be 0..%d", kMaxT);
    h->alpha = alpha;
    h->rho = alpha * alpha;                        // Kernels.cl:85 alpha*alpha, in float
    h->iterations = iterations;
    h->stencil = stencil;
    h->update_v = update_v ? 1 : 0;
    h->tblock = temporal_block;
    h->prepared = 0;
    return HSFLOW_OK;
}
int hsflow_set_lambda(hsflow_t* h, float lambda) {
    NEED(h);
    if (!(lambda > 0.f)) return fail(HSFLOW_EINVAL, "lambda must be > 0");
    h->lambda = lambda;
    h->rho = 1.0f / lambda;                        // cvCalcOpticalFlowHS: rho = 1/lambda
    h->prepared = 0;
    return HSFLOW_OK;
}
int hsflow_set_math(hsflow_t* h, int m) {
    NEED(h);
    if (m != HSFLOW_MATH_FAST && m != HSFLOW_MATH_EXACT) return fail(HSFLOW_EINVAL, "unknown math mode %d", m);
    h->math = m; h->prepared = 0;
    return HSFLOW_OK;
}
int hsflow_set_deriv(hsflow_t* h, int d) {
    NEED(h);
    if (d != HSFLOW_DERIV_CL && d != HSFLOW_DERIV_CV) return fail(HSFLOW_EINVAL, "unknown derivative mode %d", d);
    h->deriv = d; h->prepared = 0;
    return HSFLOW_OK;
}
int hsflow_set_tuning(hsflow_t* h, int chunk_rows, int warps_per_cta, int sub_batch) {
    NEED(h);
    if (chunk_rows < 0 || warps_per_cta < 0 || warps_per_cta > 8 || sub_batch < 0) return fail(HSFLOW_EINVAL, "bad tuning value");
    h->chunk_rows = chunk_rows; h->wpc = warps_per_cta;
    h->sub_batch = sub_batch;                      // takes effect at the next hsflow_configure (which re-allocates when it changes S)
    return HSFLOW_OK;
}
int hsflow_set_kernel(hsflow_t* h, int which) {   /* 0 auto, 1 single-sweep kernel only, 2 streaming kernel even for T = 1 */
    NEED(h);
    if (which < 0 || which > 2) return fail(HSFLOW_EINVAL, "kernel selector 0..2");
    h->kernel_sel = which;
    return HSFLOW_OK;
}
int hsflow_set_graph(hsflow_t* h, int mode) {
    NEED(h);
    if (mode < 0 || mode > 2) return fail(HSFLOW_EINVAL, "graph mode 0 (auto), 1 (never) or 2 (always)");
    h->graph_mode = mode;
    return HSFLOW_OK;
}
int hsflow_set_warm_start(hsflow_t* h, int keep) { NEED(h); h->warm = keep ? 1 : 0; return HSFLOW_OK; }
int hsflow_set_epsilon(hsflow_t* h, double eps) {
    NEED(h);
    if (eps != eps) return fail(HSFLOW_EINVAL, "epsilon is NaN");
    h->eps = eps > 0.0 ? eps : 0.0;
    h->prepared = 0;
    return HSFLOW_OK;
}

// EPS mode: the per-pair convergence words, zeroed ("every pair still iterating") for pairs [off, off + n)
static int eps_reset(hsflow* h, int off, int n) {
    if (h->eps_cap < h->P) {
        cudaFree(h->d_emax); cudaFree(h->d_stop);
        h->d_emax = nullptr; h->d_stop = nullptr; h->eps_cap = 0;
        if (cudaMalloc(&h->d_emax, 2 * (size_t)h->P * kMaxT * sizeof(unsigned)) != cudaSuccess ||   // two banks x pairs x stages
            cudaMalloc(&h->d_stop, (size_t)h->P * sizeof(int)) != cudaSuccess) {
            cudaGetLastError();
            return fail(HSFLOW_ENOMEM, "cudaMalloc of the convergence words failed");
        }
        h->eps_cap = h->P;
    }
    // layouts: single-sweep kernel emax[pair]; streaming kernel emax[bank][pair][kMaxT] with bank stride eps_cap * kMaxT
    CK(cudaMemsetAsync(h->d_emax + off, 0, (size_t)n * sizeof(unsigned), h->stream));
    for (int bank = 0; bank < 2; ++bank)
        CK(cudaMemsetAsync(h->d_emax + ((size_t)bank * h->eps_cap + off) * kMaxT, 0, (size_t)n * kMaxT * sizeof(unsigned), h->stream));
    CK(cudaMemsetAsync(h->d_stop + off, 0, (size_t)n * sizeof(int), h->stream));
    h->ec_blk = 0;
    return HSFLOW_OK;
}
static int eps_guard(const hsflow* h) {
    if (h->eps > 0.0 && (!h->top_edge || !h->bottom_edge || h->connected))
        return fail(HSFLOW_EINVAL, "the EPS criterion needs the whole frame on one device (not available in strip mode)");
    return HSFLOW_OK;
}

int hsflow_configure(hsflow_t* h, int W, int H, int P) {
    NEED(h);
    if (W <= 0 || H <= 0 || P <= 0) return fail(HSFLOW_EINVAL, "width, height, pairs must be positive");
    if (W > (1 << 24) || H > (1 << 24)) return fail(HSFLOW_EINVAL, "frame too large");
    CK(cudaSetDevice(h->device));
    if (h->W == W && h->H == H && h->P == P && h->uA && (h->sub_batch == 0 || h->S == std::min(h->sub_batch, P))) {
        h->prepared = 0;
        if (!h->connected) h->top_edge = h->bottom_edge = 1;   // a connected strip keeps its seams
        return HSFLOW_OK;
    }
    CK(cudaStreamSynchronize(h->stream));
    free_planes(h);
    h->W = W; h->H = H; h->P = P;
    h->pitch = ((long long)W + 31) / 32 * 32;
    h->uv_rp = 2 * h->pitch; h->uv_pp = h->uv_rp * H;
    h->c_rp = 3 * h->pitch; h->c_pp = h->c_rp * H;
    h->top_edge = h->bottom_edge = 1;
    // Pairs per launch (the scratch sub-batch: ping-pong partner + coefficient planes, 20 B/px per pair).  As many as
    // fit: one launch over 256 4K pairs runs 8 waves of work units that drift apart and keep HBM and the SMs busy
    // through each other's fill and drain phases -- 1 014 k Mpixel-iterations/s against 916 k for eight launches of 32
    // pairs (tools/subbatch_probe.py).  Auto: everything, unless results + frames + scratch would take more than
    // half of the free device memory (256 4K pairs: 64 GB of 180).
    if (h->sub_batch > 0) h->S = std::min(h->sub_batch, P);
    else {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = 0; }
        const double per_pair = (double)(h->uv_pp + h->c_pp) * sizeof(float);
        const double fixed = ((double)h->uv_pp * sizeof(float) + 8.0 * W * H) * P;       // results + two frames of <= 4 B/px
        const double room = 0.5 * (double)free_b - fixed;
        h->S = (int)std::max(1.0, std::min((double)P, room / per_pair));
        if (free_b == 0) h->S = std::min(P, 32);
    }
    const size_t uvP = (size_t)h->uv_pp * P * sizeof(float), uvS = (size_t)h->uv_pp * h->S * sizeof(float);
    const size_t cS = (size_t)h->c_pp * h->S * sizeof(float);
    if (cudaMalloc(&h->uA, uvP) != cudaSuccess || cudaMalloc(&h->uB, uvS) != cudaSuccess ||
        cudaMalloc(&h->c0, cS) != cudaSuccess) {
        cudaGetLastError();
        free_planes(h);
        h->W = h->H = h->P = 0;
        return fail(HSFLOW_ENOMEM, "cudaMalloc of %d x %d x %d planes failed", W, H, P);
    }
    h->vA = h->uA + h->pitch; h->vB = h->uB + h->pitch;
    h->c1 = h->c0 + h->pitch; h->c2 = h->c0 + 2 * h->pitch;
    for (int m = 0; m < 3; ++m) {
        const int box_rows = m + 2;
        int rc;
        if ((rc = make_map(h, &h->tm_uvA[m], h->uA, 2, P, box_rows)) || (rc = make_map(h, &h->tm_uvB[m], h->uB, 2, h->S, box_rows)) ||
            (rc = make_map(h, &h->tm_c[m], h->c0, 3, h->S, box_rows)))
            return rc;
    }
    CK(cudaMemsetAsync(h->uA, 0, uvP, h->stream));
    h->cur = 0; h->valid_lo = 0; h->valid_hi = H;
    h->zero_pending = 0; h->uv_valid = 1;
    return HSFLOW_OK;
}

int hsflow_set_strip(hsflow_t* h, int top, int bottom) {
    NEED(h);
    h->top_edge = top ? 1 : 0; h->bottom_edge = bottom ? 1 : 0;
    return HSFLOW_OK;
}

static int ensure_frames(hsflow* h, int fmt) {
    if (h->P <= 0) return fail(HSFLOW_EINVAL, "hsflow_configure first");
    if (h->fmt == fmt && h->f1) return HSFLOW_OK;
    CK(cudaStreamSynchronize(h->stream));
    drop_graphs(h);
    cudaFree(h->f1); cudaFree(h->f2); cudaFree(h->fb1); cudaFree(h->fb2); cudaFree(h->fg1); cudaFree(h->fg2);
    h->f1 = h->f2 = h->fb1 = h->fb2 = h->fg1 = h->fg2 = nullptr;
    h->fmt = -1;                                   // no frame planes until both allocations succeeded
    h->prepared = 0;
    const int bpp = fmt == FMT_GRAY8 ? 1 : (fmt == FMT_BGR8 ? 3 : 4);
    h->f_row_pitch = ((long long)h->W * bpp + 16 + 127) / 128 * 128;   // +16: vector loads may overrun the last pixel
    h->f_pair_pitch = h->f_row_pitch * h->H;
    const size_t bytes = (size_t)h->f_pair_pitch * h->P;
    if (cudaMalloc(&h->f1, bytes) != cudaSuccess || cudaMalloc(&h->f2, bytes) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(h->f1); cudaFree(h->f2);
        h->f1 = h->f2 = nullptr;
        h->f_row_pitch = h->f_pair_pitch = 0;
        return fail(HSFLOW_ENOMEM, "cudaMalloc of frame planes failed");
    }
    CK(cudaMemsetAsync(h->f1, 0, bytes, h->stream));
    CK(cudaMemsetAsync(h->f2, 0, bytes, h->stream));
    h->fmt = fmt;
    return HSFLOW_OK;
}

static int set_frames(hsflow* h, int fmt, int pair, const void* a, const void* b, size_t pitch, cudaMemcpyKind kind) {
    NEED(h);
    if (!a || !b) return fail(HSFLOW_EINVAL, "null frame pointer");
    if (pair < 0 || pair >= h->P) return fail(HSFLOW_EINVAL, "pair %d out of range (configured %d)", pair, h->P);
    int rc = ensure_frames(h, fmt);
    if (rc) return rc;
    const int bpp = fmt == FMT_GRAY8 ? 1 : (fmt == FMT_BGR8 ? 3 : 4);
    const size_t wbytes = (size_t)h->W * bpp;
    if (pitch == 0) pitch = wbytes;
    if (pitch < wbytes) return fail(HSFLOW_EINVAL, "pitch %zu smaller than a row (%zu bytes)", pitch, wbytes);
    phase_begin(h, HSFLOW_PHASE_LOAD);
    CK(cudaMemcpy2DAsync(h->f1 + (size_t)pair * h->f_pair_pitch, h->f_row_pitch, a, pitch, wbytes, h->H, kind, h->stream));
    CK(cudaMemcpy2DAsync(h->f2 + (size_t)pair * h->f_pair_pitch, h->f_row_pitch, b, pitch, wbytes, h->H, kind, h->stream));
    phase_end(h, HSFLOW_PHASE_LOAD);
    h->prepared = 0;
    return HSFLOW_OK;
}
int hsflow_set_frames_gray8(hsflow_t* h, int pair, const uint8_t* f1, const uint8_t* f2, size_t pitch) {
    return set_frames(h, FMT_GRAY8, pair, f1, f2, pitch, cudaMemcpyHostToDevice);
}
int hsflow_set_frames_bgr8(hsflow_t* h, int pair, const uint8_t* f1, const uint8_t* f2, size_t pitch) {
    return set_frames(h, FMT_BGR8, pair, f1, f2, pitch, cudaMemcpyHostToDevice);
}
int hsflow_set_frames_f32(hsflow_t* h, int pair, const float* f1, const float* f2, size_t pitch) {
    return set_frames(h, FMT_F32, pair, f1, f2, pitch, cudaMemcpyHostToDevice);
}
int hsflow_set_frames_gray8_dev(hsflow_t* h, int pair, const uint8_t* f1, const uint8_t* f2, size_t pitch) {
    return set_frames(h, FMT_GRAY8, pair, f1, f2, pitch, cudaMemcpyDeviceToDevice);
}
int hsflow_set_frames_bgr8_dev(hsflow_t* h, int pair, const uint8_t* f1, const uint8_t* f2, size_t pitch) {
    return set_frames(h, FMT_BGR8, pair, f1, f2, pitch, cudaMemcpyDeviceToDevice);
}
// Zero-copy ingest: the caller (an on-GPU decoder such as nvJPEG) writes frames straight into the handle's planes.
int hsflow_map_frames(hsflow_t* h, int frame_format, uint8_t** f1, uint8_t** f2, size_t* row_pitch, size_t* pair_pitch) {
    NEED(h);
    const int fmt = frame_format == HSFLOW_FRAMES_GRAY8 ? FMT_GRAY8 : (frame_format == HSFLOW_FRAMES_BGR8 ? FMT_BGR8 : -1);
    if (fmt < 0) return fail(HSFLOW_EINVAL, "frame format must be HSFLOW_FRAMES_GRAY8 or HSFLOW_FRAMES_BGR8");
    int rc = ensure_frames(h, fmt);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));          // the planes are idle and (when new) zeroed before a foreign stream writes them
    if (f1) *f1 = h->f1;
    if (f2) *f2 = h->f2;
    if (row_pitch) *row_pitch = (size_t)h->f_row_pitch;
    if (pair_pitch) *pair_pitch = (size_t)h->f_pair_pitch;
    h->prepared = 0;
    return HSFLOW_OK;
}
int hsflow_swap_frames(hsflow_t* h) {
    NEED(h);
    if (!h->f1 || !h->f2) return fail(HSFLOW_EINVAL, "no frame planes");
    std::swap(h->f1, h->f2);
    h->prepared = 0;
    return HSFLOW_OK;
}
int hsflow_synth_frames(hsflow_t* h, int full_height, int row0, uint32_t seed0) {
    NEED(h);
    int rc = ensure_frames(h, FMT_GRAY8);
    if (rc) return rc;
    if (full_height <= 0) full_height = h->H;
    CK(launch_synth(h->f1, h->f2, h->W, h->H, full_height, row0, h->f_row_pitch, h->f_pair_pitch, seed0, h->P, h->stream));
    h->launches++;
    h->prepared = 0;
    return HSFLOW_OK;
}
int hsflow_load_pair_gray8(hsflow_t* h, const uint8_t* f1, const uint8_t* f2, int w, int hgt, size_t pitch) {
    int rc = hsflow_configure(h, w, hgt, 1);
    return rc ? rc : hsflow_set_frames_gray8(h, 0, f1, f2, pitch);
}
int hsflow_load_pair_bgr8(hsflow_t* h, const uint8_t* f1, const uint8_t* f2, int w, int hgt, size_t pitch) {
    int rc = hsflow_configure(h, w, hgt, 1);
    return rc ? rc : hsflow_set_frames_bgr8(h, 0, f1, f2, pitch);
}
int hsflow_load_pair_f32(hsflow_t* h, const float* f1, const float* f2, int w, int hgt, size_t pitch) {
    int rc = hsflow_configure(h, w, hgt, 1);
    return rc ? rc : hsflow_set_frames_f32(h, 0, f1, f2, pitch);
}

// derivatives of pairs [p0, p0+n) into coefficient slots [0, n)
static int run_deriv(hsflow* h, int p0, int n, int normalise, float* o0, float* o1, float* o2, long long c_rp, long long c_pp,
                     const uint8_t* f1base = nullptr, const uint8_t* f2base = nullptr) {
    if (!f1base) f1base = h->f1;                   // sequence mode passes f2base = f1base + one frame: pair k = frames k, k+1
    if (!f2base) f2base = h->f2;
    DerivArgs A;
    A.f_row_pitch = h->f_row_pitch; A.f_pair_pitch = h->f_pair_pitch;
    A.c0 = o0; A.c1 = o1; A.c2 = o2;
    A.c_row_pitch = c_rp; A.c_pair_pitch = c_pp;
    A.W = h->W; A.H = h->H; A.normalise = normalise; A.rho = h->rho;
    A.zero_b = h->coef_zero_b;
    if (h->deriv == HSFLOW_DERIV_CL) {
        A.f1 = f1base + (size_t)p0 * h->f_pair_pitch;
        A.f2 = f2base + (size_t)p0 * h->f_pair_pitch;
        CK(launch_deriv(A, h->fmt, n, h->stream));
        h->launches++;
    } else {
        if (h->fmt != FMT_GRAY8 && h->fmt != FMT_BGR8) return fail(HSFLOW_EINVAL, "HSFLOW_DERIV_CV needs 8-bit frames (gray or BGR)");
        if (!h->top_edge || !h->bottom_edge) return fail(HSFLOW_EINVAL, "HSFLOW_DERIV_CV is not available in strip mode");
        const size_t bytes = (size_t)h->f_pair_pitch * h->S;
        if (!h->fb1) {
            if (cudaMalloc(&h->fb1, bytes) != cudaSuccess || cudaMalloc(&h->fb2, bytes) != cudaSuccess) {
                cudaGetLastError();
                cudaFree(h->fb1); cudaFree(h->fb2); h->fb1 = h->fb2 = nullptr;
                return fail(HSFLOW_ENOMEM, "cudaMalloc of blur planes failed");
            }
        }
        const uint8_t* g1 = f1base + (size_t)p0 * h->f_pair_pitch;
        const uint8_t* g2 = f2base + (size_t)p0 * h->f_pair_pitch;
        if (h->fmt == FMT_BGR8) {                  // cvCvtColor(CV_BGR2GRAY) (cv.cpp:17, 20) into gray planes of the same pitches
            if (!h->fg1) {
                if (cudaMalloc(&h->fg1, bytes) != cudaSuccess || cudaMalloc(&h->fg2, bytes) != cudaSuccess) {
                    cudaGetLastError();
                    cudaFree(h->fg1); cudaFree(h->fg2); h->fg1 = h->fg2 = nullptr;
                    return fail(HSFLOW_ENOMEM, "cudaMalloc of gray planes failed");
                }
            }
            CK(launch_bgr2gray(g1, h->fg1, h->W, h->H, h->f_row_pitch, h->f_pair_pitch, n, h->stream));
            CK(launch_bgr2gray(g2, h->fg2, h->W, h->H, h->f_row_pitch, h->f_pair_pitch, n, h->stream));
            h->launches += 2;
            g1 = h->fg1; g2 = h->fg2;
        }
        CK(launch_box3(g1, h->fb1, h->W, h->H, h->f_row_pitch, h->f_pair_pitch, n, h->stream));
        CK(launch_box3(g2, h->fb2, h->W, h->H, h->f_row_pitch, h->f_pair_pitch, n, h->stream));
        A.f1 = h->fb1; A.f2 = h->fb2;
        CK(launch_deriv_cv(A, n, h->stream));
        h->launches += 3;
    }
    return HSFLOW_OK;
}

// Rows per work unit.  A unit streams chunk + 2T rows (T warm-up rows above and below; those ticks run in the
// generic, predicated path at about half the steady-state speed) plus a fixed prologue worth about 4 rows, and a
// launch runs ceil(units / resident warps) waves of units one after the other: pick the chunk height that minimises
//   waves x (chunk + 4T + 4) x (1 + 0.2 / waves).
// The last factor is the lockstep penalty of short launches: in a single wave every warp runs its pipeline fill, its
// steady state and its drain at the same time as all the others, with several waves the units drift apart and keep
// HBM and the SMs busy through each other's phases (one 16384^2 frame, T = 6: 14.15 ms per 60 iterations with 1 wave
// of 2049-row chunks, 13.28 ms with 4 waves of 513-row chunks; 32 4K pairs: 14.75 -> 13.8 ms).  With many units the
// rule is "least redundant work, fullest last wave"; with few (one small frame) it makes the single wave as short as
// possible.  Heights are multiples of the TMA box rows so that every chunk enters the steady state without extra
// generic ticks (1080p, T = 4: 464 us per 100 iterations with 16-row chunks, 544 us with 15).
static int chunk_rows_for(const hsflow* h, int rows, int nsx, int pairs, int T) {
    if (h->chunk_rows > 0) return std::min(h->chunk_rows, rows);
    const long long slots = (long long)h->sm_count * std::max(1, stream_warps_per_sm(T, h->stencil));
    double best = -1.0;
    int best_ch = rows;
    const int rg = stream_geometry(T).rows_per_box;
    for (int ncy = 1; ncy <= rows; ++ncy) {
        const int ch = ((rows + ncy - 1) / ncy + rg - 1) / rg * rg;
        if (ch < 8 && ncy > 1) break;
        const long long units = (long long)nsx * ((rows + ch - 1) / ch) * pairs;
        const long long waves = (units + slots - 1) / slots;
        const double cost = (double)waves * (ch + 4.0 * T + 4.0) * (1.0 + 0.2 / (double)waves);
        if (best < 0 || cost < best - 1e-9) { best = cost; best_ch = ch; }
        if (ncy > 4096) break;
    }
    return best_ch;
}

// Peer transport: make the handle's stream wait until both neighbours published epoch `e` (their seam rows of that
// launch are in our buffer, and they finished reading the rows our later launches push into).  No host involvement,
// no spinning kernel.
static int strip_stream_wait(hsflow* h, unsigned e) {
    if ((int)(h->waited_epoch - e) >= 0) return HSFLOW_OK;
    for (int d = 0; d < 2; ++d)
        if (h->has_peer[d]) {
            // FLUSH: the neighbour's seam rows were written over NVLink BEFORE the flag; the next kernel on this stream
            // must see them, which the driver only guarantees with the flush flag (where supported -- otherwise the
            // ordering rests on the writer's __threadfence_system + st.release.sys alone).
            CUresult r = h->wait_value((CUstream)h->stream, (CUdeviceptr)(h->sig + d), e,
                                       CU_STREAM_WAIT_VALUE_GEQ | (h->can_flush ? CU_STREAM_WAIT_VALUE_FLUSH : 0));
            if (r != CUDA_SUCCESS) return fail(HSFLOW_ECUDA, "cuStreamWaitValue32 failed with CUresult %d", (int)r);
        }
    h->waited_epoch = e;
    return HSFLOW_OK;
}

// one launch advancing t iterations for n pairs.  src/dst: 0 = A planes (pair offset pA), 1 = B planes (offset 0)
static int run_block(hsflow* h, int t, int src, int pA, int n, int out_lo, int out_hi, bool zero_in = false) {
    float* uo = src == 0 ? h->uB : h->uA + (size_t)pA * h->uv_pp;
    float* vo = src == 0 ? h->vB : h->vA + (size_t)pA * h->uv_pp;
    if (use_stream_kernel(h, t)) {
        StreamArgs A;
        memset(&A, 0, sizeof A);
        A.u_out = uo; A.v_out = vo;
        A.row_pitch = h->uv_rp; A.out_pair_pitch = h->uv_pp;
        if (h->ov_u) {                             // planar, dense output (host pipeline, last launch of a sub-batch)
            A.u_out = h->ov_u; A.v_out = h->ov_v;
            A.row_pitch = h->W; A.out_pair_pitch = (long long)h->W * h->H;
        }
        A.W = h->W; A.H = h->H; A.out_lo = out_lo; A.out_hi = out_hi;
        const StreamGeom G = stream_geometry(t);
        const int nsx = (h->W + G.valid_w - 1) / G.valid_w;
        A.z_in0 = zero_in ? kZeroPair : (src == 0 ? pA : 0);
        A.z_c0 = 0;
        A.chunk_rows = chunk_rows_for(h, out_hi - out_lo, nsx, n, t);
        if (h->connected) {                        // fused halo exchange: seam rows go straight into the neighbours' buffers
            const int dst = src == 0 ? 1 : 0;      // 0 = A planes, 1 = B planes; every strip flips in step
            A.peer_up = h->has_peer[0] ? (float*)h->peer[0][dst] : nullptr;
            A.peer_dn = h->has_peer[1] ? (float*)h->peer[1][dst] : nullptr;
            A.up_lo = h->push_lo[0]; A.up_hi = h->push_hi[0]; A.up_delta = h->push_delta[0];
            A.dn_lo = h->push_lo[1]; A.dn_hi = h->push_hi[1]; A.dn_delta = h->push_delta[1];
            A.done_counter = h->sig + 2;
            A.flag_up = h->has_peer[0] ? (unsigned*)h->peer[0][2] + 1 : nullptr;   // we are its lower neighbour
            A.flag_dn = h->has_peer[1] ? (unsigned*)h->peer[1][2] + 0 : nullptr;   // we are its upper neighbour
            A.epoch = h->epoch;
            // This launch needs the neighbours' previous epoch: the stream waits for it (no kernel ever spins).
            // Tried in round 2 and removed: waiting inside the seam units of the kernel instead (ld.acquire.sys on the
            // epoch word) so that launches overlap through programmatic dependent launch.  Bit-identical, but no faster on
            // 2 x B200 (4.530 vs 4.544 ms per 120 iterations of 2048-row strips) and the spin loop cost the PEER
            // instantiation 19 registers (236 -> 255 with spills at T = 6), which slowed every unit of the launch.
            { int rc = strip_stream_wait(h, h->epoch - 1u); if (rc) return rc; }
        }
        if (h->ec_on) {                            // EPS criterion: main launch + replay launch on the TRACK instantiation
            if (h->connected || h->ov_u || t > kTrackT) return fail(HSFLOW_EINVAL, "internal: EPS block outside its envelope");
            const StreamGeom GT = stream_geometry(kTrackT);
            A.chunk_rows = chunk_rows_for(h, out_hi - out_lo, (h->W + GT.valid_w - 1) / GT.valid_w, n, kTrackT);
            const int mt = GT.rows_per_box - 2;
            const int bank = h->ec_blk & 1;
            A.stop = h->d_stop; A.z_trk0 = h->ec_off;
            A.emax = h->d_emax + (size_t)bank * h->eps_cap * kMaxT;
            A.emax_next = h->d_emax + (size_t)(bank ^ 1) * h->eps_cap * kMaxT;
            A.trk_t = t; A.trk_base = h->ec_sweep; A.trk_dst_parity = src == 0 ? 1 : 0; A.eps = h->eps;
            const CUtensorMap& tuv = src == 0 ? h->tm_uvA[mt] : h->tm_uvB[mt];
            A.trk_mode = 0;
            CK(launch_jacobi_stream_track(h->stencil, tuv, h->tm_c[mt], A, n, h->stream));
            A.trk_mode = 1;
            CK(launch_jacobi_stream_track(h->stencil, tuv, h->tm_c[mt], A, n, h->stream));
            h->launches += 2;
            h->ec_blk++;
            return HSFLOW_OK;
        }
        const int m = G.rows_per_box - 2;
        if (m < 0 || m > 2) return fail(HSFLOW_EINVAL, "internal: no tensor map for %d-row boxes", G.rows_per_box);
        CK(launch_jacobi_stream(t, h->stencil, src == 0 ? h->tm_uvA[m] : h->tm_uvB[m], h->tm_c[m], A, n, h->stream));
        h->launches++;
        return HSFLOW_OK;
    }
    if (t != 1 || zero_in) return fail(HSFLOW_EINVAL, "internal: single-sweep kernel advances one iteration per launch from a real buffer");
    Jacobi1Args A;
    memset(&A, 0, sizeof A);
    A.u_in = src == 0 ? h->uA + (size_t)pA * h->uv_pp : h->uB;
    A.v_in = src == 0 ? h->vA + (size_t)pA * h->uv_pp : h->vB;
    A.u_out = uo; A.v_out = vo;
    A.c0 = h->c0; A.c1 = h->c1; A.c2 = h->c2;
    A.row_pitch = h->uv_rp; A.in_pair_pitch = A.out_pair_pitch = h->uv_pp;
    A.c_row_pitch = h->c_rp; A.c_pair_pitch = h->c_pp;
    A.W = h->W; A.H = h->H; A.out_lo = out_lo; A.out_hi = out_hi;
    if (h->chunk_rows > 0) A.chunk_rows = h->chunk_rows;
    else {
        // 64-row chunks re-read 2 halo rows per 64 (3 %).  A single frame cannot fill the GPU that way (1080p: 4 x 17
        // CTAs): shrink the chunk until there are about eight waves of seven 4-warp CTAs per SM, but not below 8 rows
        // (25 % re-reads, absorbed by L2) -- or 4 rows when even 8-row chunks leave SMs idle.  100 sweeps with 64-row
        // chunks / with this rule: 600x480 5.9 / 0.8 ms, 1080p 6.9 / 1.4 ms (tools/small_frame_probe.py).
        const long long rows = out_hi - out_lo, nx = (h->W + 511) / 512;
        const long long fit = rows * nx * n / ((long long)h->sm_count * 56);
        const long long lo = nx * ((rows + 7) / 8) * n < 2LL * h->sm_count ? 4 : 8;
        A.chunk_rows = (int)std::max<long long>(1, std::min<long long>(rows, std::max<long long>(lo, std::min<long long>(64, fit))));
    }
    A.rho = h->rho;
    // EXACT OpenCV-mode path: the rounding sequence of cvCalcOpticalFlowHS, so that the field is bit-identical to the
    // restated routine the shipped *_cv_out.jpg pin
    A.cv_form = (h->math == HSFLOW_MATH_EXACT && h->deriv == HSFLOW_DERIV_CV && h->stencil == HSFLOW_STENCIL_CV4) ? 1 : 0;
    if (h->ec_on) {                                // EPS mode: track max |new - old|, carry converged pairs over
        A.emax = h->d_emax + h->ec_off; A.stop = h->d_stop + h->ec_off;
        A.last_sweep = h->ec_sweep == h->ec_total; A.total_sweeps = h->ec_total;
    }
    CK(launch_jacobi1(A, h->math == HSFLOW_MATH_EXACT, h->stencil, h->update_v != 0, n, h->stream));
    h->launches++;
    if (h->ec_on) {
        CK(launch_eps_check(A.emax, A.stop, h->eps, h->ec_sweep, n, h->stream));
        h->launches++;
        if (A.last_sweep) { CK(launch_eps_settle(A.stop, h->ec_total, n, h->stream)); h->launches++; }
    }
    return HSFLOW_OK;
}

int hsflow_prepare(hsflow_t* h) {
    NEED(h);
    if (h->P <= 0 || !h->f1 || !h->f2) return fail(HSFLOW_EINVAL, "configure and load frames first");
    if (h->P > h->S) return fail(HSFLOW_EINVAL, "hsflow_prepare/iterate need pairs <= sub_batch (%d > %d); use hsflow_compute", h->P, h->S);
    CK(cudaSetDevice(h->device));
    phase_begin(h, HSFLOW_PHASE_DERIV);
    const int norm = h->math == HSFLOW_MATH_FAST ? 1 : 0;
    if (!h->warm) h->uv_dirty = 0;                 // u, v are zeroed below
    h->coef_zero_b = (!h->update_v && use_stream_kernel(h, 1)) ? 1 : 0;
    int rc = run_deriv(h, 0, h->P, norm, h->c0, h->c1, h->c2, h->c_rp, h->c_pp);
    if (rc) return rc;
    h->coef_norm = norm;
    h->zero_pending = 0;
    if (!h->warm) {                                // cpp:331-332: u, v start at zero for every pair
        h->zero_pending = 1;
        if (!use_stream_kernel(h, 1)) { if ((rc = materialize_zero(h))) return rc; }
    }
    h->uv_valid = 1;
    phase_end(h, HSFLOW_PHASE_DERIV);
    h->valid_lo = 0; h->valid_hi = h->H;
    h->sweeps = 0;
    if (h->eps > 0.0) { if ((rc = eps_guard(h)) || (rc = eps_reset(h, 0, h->P))) return rc; }
    h->prepared = 1;
    return HSFLOW_OK;
}

int hsflow_iterate(hsflow_t* h, int n) {
    NEED(h);
    if (!h->prepared) return fail(HSFLOW_EINVAL, "hsflow_prepare first");
    if (n < 0) return fail(HSFLOW_EINVAL, "n must be >= 0");
    CK(cudaSetDevice(h->device));
    const int T = effective_T(h);
    phase_begin(h, HSFLOW_PHASE_ITER);
    if (h->connected) {
        // Peer transport: every launch refreshes the neighbours' ghost rows itself, so the valid range never shrinks
        // beyond one block.  After each launch the stream waits (cuStreamWaitValue32, no host involvement, no
        // spinning kernel) until both neighbours published the same epoch: their seam rows are in our buffer and
        // they finished reading the buffer our next launch stores into.
        if (!use_stream_kernel(h, 1)) return fail(HSFLOW_EINVAL, "peer transport needs the streaming kernel (FAST math, update_v = 1)");
        while (n > 0) {
            const int t = next_block(n, T);
            const int lo = h->top_edge ? 0 : t, hi = h->bottom_edge ? h->H : h->H - t;
            if ((h->has_peer[0] && h->push_lo[0] < t) || (h->has_peer[1] && h->H - h->push_hi[1] < t) || lo >= hi)
                return fail(HSFLOW_EINVAL, "strip has fewer ghost rows than the temporal block (%d)", t);
            ++h->epoch;
            int rc = run_block(h, t, h->cur, 0, h->P, lo, hi, h->zero_pending != 0);
            if (rc) return rc;
            h->zero_pending = 0;
            h->cur ^= 1;
            h->sweeps += t;
            n -= t;
        }
        // end of the call: the stream catches up with both neighbours, so that whatever follows on it (read-back, a
        // new hsflow_prepare) sees a quiescent seam
        { int rc = strip_stream_wait(h, h->epoch); if (rc) return rc; }
        h->valid_lo = 0; h->valid_hi = h->H;
        phase_end(h, HSFLOW_PHASE_ITER);
        return HSFLOW_OK;
    }
    const int sweeps_at_end = h->sweeps + n;
    { int rc = eps_guard(h); if (rc) return rc; }
    while (n > 0) {
        const int t = next_block(n, T);
        const int lo = h->top_edge ? 0 : h->valid_lo + t;
        const int hi = h->bottom_edge ? h->H : h->valid_hi - t;
        if (lo >= hi) return fail(HSFLOW_EINVAL, "ghost rows exhausted: refresh the halo (valid rows [%d,%d), block %d)", h->valid_lo, h->valid_hi, t);
        // the single-sweep path ping-pongs internally; keep the bookkeeping identical for both
        if (use_stream_kernel(h, t)) {
            h->ec_on = h->eps > 0.0; h->ec_sweep = h->sweeps; h->ec_off = 0;
            int rc = run_block(h, t, h->cur, 0, h->P, lo, hi, h->zero_pending != 0);
            h->ec_on = 0;
            if (rc) return rc;
            h->zero_pending = 0;
            h->cur ^= 1;
            h->sweeps += t;
        } else {
            { int rc = materialize_zero(h); if (rc) return rc; }
            for (int k = 0; k < t; ++k) {
                const int lo1 = h->top_edge ? 0 : h->valid_lo + k + 1;
                const int hi1 = h->bottom_edge ? h->H : h->valid_hi - k - 1;
                h->ec_on = h->eps > 0.0; h->ec_sweep = h->sweeps + 1; h->ec_total = sweeps_at_end; h->ec_off = 0;
                int rc = run_block(h, 1, h->cur, 0, h->P, lo1, hi1);
                h->ec_on = 0;
                if (rc) return rc;
                h->cur ^= 1;
                h->sweeps++;
            }
        }
        h->valid_lo = lo; h->valid_hi = hi;
        n -= t;
    }
    // EPS on the streaming kernel: pairs that stopped in the other ping-pong buffer move into the current one
    if (eps_on_stream(h) && h->d_stop) { CK(launch_copy_stopped(h->uA, h->uB, h->uv_pp, h->d_stop, h->cur, h->P, h->stream)); h->launches += 2; }
    phase_end(h, HSFLOW_PHASE_ITER);
    return HSFLOW_OK;
}

int hsflow_halo_refreshed(hsflow_t* h) { NEED(h); h->valid_lo = 0; h->valid_hi = h->H; return HSFLOW_OK; }

// derivative pass + all iterations for pairs [p0, p0+n) (n <= S); the result lands in the A planes
// fin_u / fin_v: where the final fields go instead of the A planes (planar [n][H][W]; needs the streaming kernel for the
// last block and W % 4 == 0, checked by the caller)
static int compute_subbatch(hsflow* h, int p0, int n, const uint8_t* f1base = nullptr, const uint8_t* f2base = nullptr,
                            float* fin_u = nullptr, float* fin_v = nullptr) {
    const int T = effective_T(h), N = h->iterations;
    int L = 0;                                     // ping-pong flips
    for (int left = N; left > 0;) { const int t = next_block(left, T); L += use_stream_kernel(h, t) ? 1 : t; left -= t; }
    const int norm = h->math == HSFLOW_MATH_FAST ? 1 : 0;
    h->coef_zero_b = (!h->update_v && use_stream_kernel(h, 1)) ? 1 : 0;     // u, v are zeroed below: only warm / dirty say no
    int rc = run_deriv(h, p0, n, norm, h->c0, h->c1, h->c2, h->c_rp, h->c_pp, f1base, f2base);
    if (rc) return rc;
    h->coef_norm = norm;
    int src = (L % 2 == 0) ? 0 : 1;                // so that the last flip lands in A
    // cpp:331-332 u = v = 0: zero fill by the TMA unit in the first streaming launch, a real memset otherwise
    bool zero_in = N > 0 && use_stream_kernel(h, std::min(N, T));
    if (!zero_in) {
        float* uv = src == 0 ? h->uA + (size_t)p0 * h->uv_pp : h->uB;
        CK(cudaMemsetAsync(uv, 0, (size_t)h->uv_pp * n * sizeof(float), h->stream));
    }
    int sweep = 0;
    if (h->eps > 0.0 && (rc = eps_reset(h, p0, n))) return rc;
    for (int left = N; left > 0;) {
        const int t = next_block(left, T);
        if (use_stream_kernel(h, t)) {
            if (left == t) { h->ov_u = fin_u; h->ov_v = fin_v; }
            h->ec_on = h->eps > 0.0; h->ec_sweep = N - left; h->ec_off = p0;
            rc = run_block(h, t, src, p0, n, 0, h->H, zero_in);
            h->ec_on = 0;
            zero_in = false;
            h->ov_u = h->ov_v = nullptr;
            if (rc) return rc;
            src ^= 1;
        } else {
            for (int k = 0; k < t; ++k) {
                h->ec_on = h->eps > 0.0; h->ec_sweep = ++sweep; h->ec_total = N; h->ec_off = p0;
                rc = run_block(h, 1, src, p0, n, 0, h->H);
                h->ec_on = 0;
                if (rc) return rc;
                src ^= 1;
            }
        }
        left -= t;
    }
    if (src != 0) return fail(HSFLOW_ECUDA, "internal: sub-batch result not in the output planes");
    if (eps_on_stream(h) && N > 0) {               // pairs that stopped in the B planes move into the A planes
        CK(launch_copy_stopped(h->uA + (size_t)p0 * h->uv_pp, h->uB, h->uv_pp, h->d_stop + p0, 0, n, h->stream));
        h->launches += 2;
    }
    return HSFLOW_OK;
}

int hsflow_compute(hsflow_t* h) {
    NEED(h);
    if (h->P <= 0 || !h->f1 || !h->f2) return fail(HSFLOW_EINVAL, "configure and load frames first");
    if (h->P <= h->S) {
        const bool small = (long long)h->W * h->H * h->P <= (4LL << 20);
        const bool eligible = h->graph_mode != 1 && (h->graph_mode == 2 || small) && !h->warm && !h->connected && h->top_edge &&
                              h->bottom_edge && h->iterations > 0;
        if (!eligible) {
            int rc = hsflow_prepare(h);
            return rc ? rc : hsflow_iterate(h, h->iterations);
        }
        CK(cudaSetDevice(h->device));
        GraphKey key;
        memset(&key, 0, sizeof key);               // padding bytes take part in the comparison
        key.W = h->W; key.H = h->H; key.P = h->P; key.S = h->S; key.iterations = h->iterations; key.stencil = h->stencil;
        key.update_v = h->update_v; key.tblock = h->tblock; key.math = h->math; key.deriv = h->deriv; key.kernel_sel = h->kernel_sel;
        key.chunk_rows = h->chunk_rows; key.fmt = h->fmt; key.uv_dirty = h->uv_dirty; key.rho = h->rho; key.eps = h->eps;
        key.f1 = h->f1; key.f2 = h->f2;
        GraphEntry* hit = nullptr;
        for (GraphEntry& g : h->graphs) if (g.key == key) hit = &g;
        if (!hit && !(h->have_eager_key && h->eager_key == key)) {
            // first time: run eagerly (one-off allocations and attribute calls happen here, never inside a capture)
            memcpy(&h->eager_key, &key, sizeof key); h->have_eager_key = 1;
            h->cur = 0;                            // same starting buffer as the graph of this computation will use
            int rc = hsflow_prepare(h);
            return rc ? rc : hsflow_iterate(h, h->iterations);
        }
        if (!hit) {                                // second time: capture the launch sequence
            h->cur = 0;
            const long long l0 = h->launches;
            cudaGraph_t graph = nullptr;
            CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            h->capturing = 1;
            int rc = hsflow_prepare(h);
            if (!rc) rc = hsflow_iterate(h, h->iterations);
            h->capturing = 0;
            const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
            const long long captured = h->launches - l0;
            h->launches = l0;                      // nothing ran yet: the replay below counts them
            if (rc || ce != cudaSuccess || !graph) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                h->graph_mode = 1;                 // do not try again on this handle
                h->prepared = 0;
                if (rc) return rc;
                rc = hsflow_prepare(h);
                return rc ? rc : hsflow_iterate(h, h->iterations);
            }
            GraphEntry g;
            memcpy(&g.key, &key, sizeof key); g.final_cur = h->cur; g.launches = captured;
            const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ie != cudaSuccess) {
                cudaGetLastError();
                h->graph_mode = 1;
                h->cur = 0; h->prepared = 0;
                rc = hsflow_prepare(h);
                return rc ? rc : hsflow_iterate(h, h->iterations);
            }
            if (h->graphs.size() >= kMaxGraphs) {  // evict the least recently used
                size_t lru = 0;
                for (size_t i = 1; i < h->graphs.size(); ++i) if (h->graphs[i].stamp < h->graphs[lru].stamp) lru = i;
                cudaGraphExecDestroy(h->graphs[lru].exec);
                h->graphs.erase(h->graphs.begin() + (long)lru);
            }
            h->graphs.push_back(g);
            hit = &h->graphs.back();
        }
        hit->stamp = ++h->graph_stamp;
        phase_begin(h, HSFLOW_PHASE_ITER);
        h->ev_set[HSFLOW_PHASE_DERIV] = 0;         // the graph runs as one piece: its time is reported as the iteration phase
        CK(cudaGraphLaunch(hit->exec, h->stream));
        phase_end(h, HSFLOW_PHASE_ITER);
        h->launches += hit->launches;
        h->cur = hit->final_cur;
        h->valid_lo = 0; h->valid_hi = h->H;
        h->sweeps = h->iterations;
        h->zero_pending = 0; h->prepared = 1; h->uv_valid = 1;
        h->coef_norm = h->math == HSFLOW_MATH_FAST ? 1 : 0;
        h->coef_zero_b = (!h->update_v && use_stream_kernel(h, 1)) ? 1 : 0;
        return HSFLOW_OK;
    }
    // batch larger than the scratch: sub-batches of S pairs; results always end in the A planes
    if (h->warm) return fail(HSFLOW_EINVAL, "warm start needs pairs <= sub_batch");
    CK(cudaSetDevice(h->device));
    { int rc = eps_guard(h); if (rc) return rc; }
    h->sweeps = h->iterations;
    phase_begin(h, HSFLOW_PHASE_ITER);
    for (int p0 = 0; p0 < h->P; p0 += h->S) {
        int rc = compute_subbatch(h, p0, std::min(h->S, h->P - p0));
        if (rc) return rc;
    }
    phase_end(h, HSFLOW_PHASE_ITER);
    h->cur = 0;
    h->prepared = 0;
    h->zero_pending = 0; h->uv_valid = 1;
    return HSFLOW_OK;
}

int hsflow_compute_range(hsflow_t* h, int p0, int n) {
    NEED(h);
    if (h->P <= 0 || !h->f1 || !h->f2) return fail(HSFLOW_EINVAL, "configure and load frames first");
    if (p0 < 0 || n < 1 || p0 + n > h->P) return fail(HSFLOW_EINVAL, "pairs [%d, %d) outside the configured %d", p0, p0 + n, h->P);
    if (n > h->S) return fail(HSFLOW_EINVAL, "a range holds at most sub_batch = %d pairs", h->S);
    if (h->warm) return fail(HSFLOW_EINVAL, "warm start is not available for pair ranges");
    if (!h->top_edge || !h->bottom_edge || h->connected) return fail(HSFLOW_EINVAL, "not available in strip mode");
    CK(cudaSetDevice(h->device));
    h->zero_pending = 0;
    h->cur = 0;                                    // ranges always deliver into the result planes; a prepare/iterate session ends here
    phase_begin(h, HSFLOW_PHASE_ITER);
    int rc = compute_subbatch(h, p0, n);
    if (rc) return rc;
    phase_end(h, HSFLOW_PHASE_ITER);
    h->sweeps = h->iterations;
    h->prepared = 0;
    h->uv_valid = 1;
    return HSFLOW_OK;
}

int hsflow_strip_export(hsflow_t* h, hsflow_strip_handle_t* out) {
    NEED(h);
    if (!out) return fail(HSFLOW_EINVAL, "out is null");
    if (!h->uA || h->P != 1) return fail(HSFLOW_EINVAL, "hsflow_strip_export needs a configured handle with one pair");
    CK(cudaSetDevice(h->device));
    if (!h->sig) {
        if (cudaMalloc(&h->sig, 256) != cudaSuccess) { cudaGetLastError(); return fail(HSFLOW_ENOMEM, "cudaMalloc"); }
    }
    CK(cudaMemsetAsync(h->sig, 0, 256, h->stream));
    CK(cudaStreamSynchronize(h->stream));          // the words are zero before any neighbour can see them
    h->epoch = 0;
    h->waited_epoch = 0;
    StripBlob b;
    memset(&b, 0, sizeof b);
    b.magic = kStripMagic; b.W = h->W; b.H = h->H; b.device = h->device; b.pitch = h->pitch; b.pid = (int64_t)getpid();
    b.raw[0] = h->uA; b.raw[1] = h->uB; b.raw[2] = h->sig;
    CK(cudaIpcGetMemHandle(&b.mem[0], h->uA));
    CK(cudaIpcGetMemHandle(&b.mem[1], h->uB));
    CK(cudaIpcGetMemHandle(&b.mem[2], h->sig));
    memset(out, 0, sizeof *out);
    memcpy(out, &b, sizeof b);
    return HSFLOW_OK;
}

int hsflow_strip_connect(hsflow_t* h, const hsflow_strip_handle_t* up, int up_lo, int up_hi, int up_delta,
                         const hsflow_strip_handle_t* down, int dn_lo, int dn_hi, int dn_delta) {
    NEED(h);
    if (!h->uA || h->P != 1 || !h->sig) return fail(HSFLOW_EINVAL, "hsflow_strip_export first");
    if (!h->wait_value) return fail(HSFLOW_ECUDA, "cuStreamWaitValue32 not available from the driver");
    CK(cudaSetDevice(h->device));
    if (h->connected) strip_disconnect(h);
    const hsflow_strip_handle_t* blobs[2] = {up, down};
    const int lo[2] = {up_lo, dn_lo}, hi[2] = {up_hi, dn_hi}, delta[2] = {up_delta, dn_delta};
    for (int d = 0; d < 2; ++d) {
        if (!blobs[d]) continue;
        StripBlob b;
        memcpy(&b, blobs[d], sizeof b);
        if (b.magic != kStripMagic) { strip_disconnect(h); return fail(HSFLOW_EINVAL, "not a strip handle"); }
        if (b.W != h->W || b.pitch != h->pitch) { strip_disconnect(h); return fail(HSFLOW_EINVAL, "neighbour strip has a different width (%d vs %d)", b.W, h->W); }
        if (lo[d] < 0 || hi[d] < lo[d] || hi[d] > h->H || lo[d] + delta[d] < 0 || hi[d] + delta[d] > b.H) {
            strip_disconnect(h);
            return fail(HSFLOW_EINVAL, "seam rows [%d,%d) + %d fall outside the strips (%d and %d rows)", lo[d], hi[d], delta[d], h->H, b.H);
        }
        if (b.pid == (int64_t)getpid()) {          // same process: plain peer access
            if (b.device != h->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    strip_disconnect(h);
                    return fail(HSFLOW_ECUDA, "cudaDeviceEnablePeerAccess(%d): %s", b.device, cudaGetErrorString(e));
                }
                cudaGetLastError();
            }
            for (int k = 0; k < 3; ++k) h->peer[d][k] = b.raw[k];
            h->peer_ipc[d] = 0;
        } else {
            h->peer_ipc[d] = 1;
            h->has_peer[d] = 1;                    // so that a partial failure closes what was opened
            for (int k = 0; k < 3; ++k) {
                cudaError_t e = cudaIpcOpenMemHandle(&h->peer[d][k], b.mem[k], cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) {
                    h->peer[d][k] = nullptr;
                    strip_disconnect(h);
                    return fail(HSFLOW_ECUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
                }
            }
        }
        h->has_peer[d] = 1;
        h->push_lo[d] = lo[d]; h->push_hi[d] = hi[d]; h->push_delta[d] = delta[d];
    }
    h->top_edge = h->has_peer[0] ? 0 : 1;
    h->bottom_edge = h->has_peer[1] ? 0 : 1;
    h->connected = 1;
    h->prepared = 0;
    return HSFLOW_OK;
}

int hsflow_strip_disconnect(hsflow_t* h) {
    NEED(h);
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    strip_disconnect(h);
    return HSFLOW_OK;
}

int hsflow_sync(hsflow_t* h) {
    NEED(h);
    CK(cudaStreamSynchronize(h->stream));
    return HSFLOW_OK;
}

// Readers of the current field: it must exist on the device (the pipelined host calls deliver their results to host
// memory and leave the planes in an unspecified state) and pending zeros must have been written.
static int field_ready(hsflow* h) {
    if (!h->uA) return fail(HSFLOW_EINVAL, "not configured");
    if (!h->uv_valid)
        return fail(HSFLOW_EINVAL, "no flow field on the device: hsflow_run_batch_host / hsflow_run_sequence_host deliver their "
                                   "results to host memory; run hsflow_compute to read fields from the handle");
    return materialize_zero(h);
}
static float* cur_u(hsflow* h) { return h->cur == 0 ? h->uA : h->uB; }
static float* cur_v(hsflow* h) { return h->cur == 0 ? h->vA : h->vB; }

int hsflow_read_uv(hsflow_t* h, int pair, float* u, float* v, size_t pitch) {
    NEED(h);
    if (pair < 0 || pair >= h->P) return fail(HSFLOW_EINVAL, "pair %d out of range", pair);
    if (h->cur == 1 && pair >= h->S) return fail(HSFLOW_EINVAL, "internal: pair outside scratch planes");
    { int rc = field_ready(h); if (rc) return rc; }
    const size_t wb = (size_t)h->W * sizeof(float);
    if (pitch == 0) pitch = wb;
    if (pitch < wb) return fail(HSFLOW_EINVAL, "pitch too small");
    phase_begin(h, HSFLOW_PHASE_READ);
    if (u) CK(cudaMemcpy2DAsync(u, pitch, cur_u(h) + (size_t)pair * h->uv_pp, h->uv_rp * sizeof(float), wb, h->H, cudaMemcpyDeviceToHost, h->stream));
    if (v) CK(cudaMemcpy2DAsync(v, pitch, cur_v(h) + (size_t)pair * h->uv_pp, h->uv_rp * sizeof(float), wb, h->H, cudaMemcpyDeviceToHost, h->stream));
    phase_end(h, HSFLOW_PHASE_READ);
    CK(cudaStreamSynchronize(h->stream));
    return HSFLOW_OK;
}

int hsflow_write_uv(hsflow_t* h, int pair, const float* u, const float* v, size_t pitch) {
    NEED(h);
    if (pair < 0 || pair >= h->P || (h->cur == 1 && pair >= h->S)) return fail(HSFLOW_EINVAL, "pair %d out of range", pair);
    if (!h->uA) return fail(HSFLOW_EINVAL, "not configured");
    { int rc = materialize_zero(h); if (rc) return rc; }
    h->uv_valid = 1;
    const size_t wb = (size_t)h->W * sizeof(float);
    if (pitch == 0) pitch = wb;
    if (u) CK(cudaMemcpy2DAsync(cur_u(h) + (size_t)pair * h->uv_pp, h->uv_rp * sizeof(float), u, pitch, wb, h->H, cudaMemcpyHostToDevice, h->stream));
    if (v) CK(cudaMemcpy2DAsync(cur_v(h) + (size_t)pair * h->uv_pp, h->uv_rp * sizeof(float), v, pitch, wb, h->H, cudaMemcpyHostToDevice, h->stream));
    if (v) {
        h->uv_dirty = 1;                           // LITERAL mode leaves the streaming kernel (literal_on_stream)
        if (h->coef_zero_b && h->prepared && h->P <= h->S) {   // ... and needs the true b plane again
            h->coef_zero_b = 0;
            int rc = run_deriv(h, 0, h->P, h->coef_norm, h->c0, h->c1, h->c2, h->c_rp, h->c_pp);
            if (rc) return rc;
        }
    }
    CK(cudaStreamSynchronize(h->stream));
    return HSFLOW_OK;
}

int hsflow_read_derivatives(hsflow_t* h, int pair, float* Ex, float* Ey, float* Et, size_t pitch) {
    NEED(h);
    if (pair < 0 || pair >= h->P || !h->f1) return fail(HSFLOW_EINVAL, "pair %d out of range or no frames", pair);
    const size_t wb = (size_t)h->W * sizeof(float);
    if (pitch == 0) pitch = wb;
    const long long plane = h->pitch * h->H;
    if (!h->dtmp && cudaMalloc(&h->dtmp, 3 * (size_t)plane * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        return fail(HSFLOW_ENOMEM, "cudaMalloc");
    }
    float* d[3] = {h->dtmp, h->dtmp + plane, h->dtmp + 2 * plane};
    const int keep_zero_b = h->coef_zero_b;
    h->coef_zero_b = 0;                            // the caller gets Ex, Ey, Et as ComputeDerivativesKernel writes them
    int rc = run_deriv(h, pair, 1, 0, d[0], d[1], d[2], h->pitch, plane);
    h->coef_zero_b = keep_zero_b;
    if (rc) return rc;
    float* o[3] = {Ex, Ey, Et};
    for (int k = 0; k < 3; ++k)
        if (o[k]) CK(cudaMemcpy2DAsync(o[k], pitch, d[k], h->pitch * sizeof(float), wb, h->H, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return HSFLOW_OK;
}

int hsflow_get_device_uv(hsflow_t* h, float** u, float** v, size_t* row_pitch, size_t* pair_pitch) {
    NEED(h);
    { int rc = field_ready(h); if (rc) return rc; }
    if (u) *u = cur_u(h);
    if (v) *v = cur_v(h);
    if (row_pitch) *row_pitch = (size_t)h->uv_rp;       // v == u + row_pitch/2: u and v rows interleave
    if (pair_pitch) *pair_pitch = (size_t)h->uv_pp;
    return HSFLOW_OK;
}
int hsflow_get_device_frames(hsflow_t* h, uint8_t** f1, uint8_t** f2, size_t* row_pitch, size_t* pair_pitch) {
    NEED(h);
    int rc = ensure_frames(h, h->fmt >= 0 ? h->fmt : FMT_GRAY8);
    if (rc) return rc;
    if (f1) *f1 = h->f1;
    if (f2) *f2 = h->f2;
    if (row_pitch) *row_pitch = (size_t)h->f_row_pitch;
    if (pair_pitch) *pair_pitch = (size_t)h->f_pair_pitch;
    return HSFLOW_OK;
}

int hsflow_dot_mask(hsflow_t* h, int pair, int step, float thr, uint8_t* mask, int* count) {
    NEED(h);
    if (pair < 0 || pair >= h->P || step <= 0 || !mask) return fail(HSFLOW_EINVAL, "bad argument");
    if (h->cur == 1 && pair >= h->S) return fail(HSFLOW_EINVAL, "internal: pair outside scratch planes");
    { int rc = field_ready(h); if (rc) return rc; }
    const int gw = (h->W + step - 1) / step, gh = (h->H + step - 1) / step;
    const size_t need = (size_t)gw * gh;
    if (need > h->mask_cap) {
        cudaFree(h->d_mask); h->d_mask = nullptr; h->mask_cap = 0;
        if (cudaMalloc(&h->d_mask, need) != cudaSuccess) { cudaGetLastError(); return fail(HSFLOW_ENOMEM, "cudaMalloc"); }
        h->mask_cap = need;
    }
    CK(cudaMemsetAsync(h->d_count, 0, sizeof(int), h->stream));
    CK(launch_dot_mask(cur_u(h) + (size_t)pair * h->uv_pp, cur_v(h) + (size_t)pair * h->uv_pp, h->W, h->H, h->uv_rp, step, thr,
                       h->d_mask, h->d_count, h->stream));
    h->launches++;
    CK(cudaMemcpyAsync(mask, h->d_mask, need, cudaMemcpyDeviceToHost, h->stream));
    int c = 0;
    CK(cudaMemcpyAsync(&c, h->d_count, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (count) *count = c;
    return HSFLOW_OK;
}

int hsflow_sample_uv(hsflow_t* h, int pair, int step, float* u_s, float* v_s) {
    NEED(h);
    if (pair < 0 || pair >= h->P || step <= 0 || !u_s || !v_s) return fail(HSFLOW_EINVAL, "bad argument");
    if (h->cur == 1 && pair >= h->S) return fail(HSFLOW_EINVAL, "internal: pair outside scratch planes");
    { int rc = field_ready(h); if (rc) return rc; }
    const int gw = (h->W + step - 1) / step, gh = (h->H + step - 1) / step;
    const size_t need = 2 * (size_t)gw * gh * sizeof(float);
    if (need > h->sample_cap) {
        cudaFree(h->d_sample); h->d_sample = nullptr; h->sample_cap = 0;
        if (cudaMalloc(&h->d_sample, need) != cudaSuccess) { cudaGetLastError(); return fail(HSFLOW_ENOMEM, "cudaMalloc"); }
        h->sample_cap = need;
    }
    float* ds_u = h->d_sample; float* ds_v = ds_u + (size_t)gw * gh;
    CK(launch_sample_uv(cur_u(h) + (size_t)pair * h->uv_pp, cur_v(h) + (size_t)pair * h->uv_pp, h->W, h->H, h->uv_rp, h->uv_pp, step,
                        ds_u, ds_v, 1, h->stream));
    h->launches++;
    CK(cudaMemcpyAsync(u_s, ds_u, need / 2, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(v_s, ds_v, need / 2, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return HSFLOW_OK;
}

// Three-stage pipeline over sub-batches of B pairs: H2D of sub-batch i+1, compute of i and D2H of i-1 overlap on three
// streams.  sequence = 0: `frames` holds n_pairs x (f1, f2).  sequence = 1: `frames` holds n_pairs + 1 consecutive
// frames and pair k = (frame k, frame k+1) -- the camera loop of cpp:800-842, where the second frame of one pair is
// the first of the next (cpp:834 memcpy I2 -> I1): here each slot holds B + 1 frames in ONE plane and the second-frame
// pointer is the first-frame pointer plus one frame, so a frame is uploaded once and never copied.
// fmt: gray8 or interleaved BGR8 (cvLoadImage order; the derivative kernel does the cvCvtColor of cpp:727-728).
// sample_step > 0: u_out / v_out receive the fields on the stride-`step` grid only (what cpp:762-767 reads).
struct PipeOpts { int fmt, sequence, sample_step; };

static int run_pipeline(hsflow* h, const uint8_t* frames, int n_pairs, int w, int hgt, float* u_out, float* v_out, PipeOpts o) {
    if (!frames || !u_out || !v_out || n_pairs <= 0 || w <= 0 || hgt <= 0) return fail(HSFLOW_EINVAL, "bad argument");
    if (o.fmt != FMT_GRAY8 && o.fmt != FMT_BGR8) return fail(HSFLOW_EINVAL, "frame format must be gray8 or bgr8");
    if (o.sample_step < 0) return fail(HSFLOW_EINVAL, "sample_step must be >= 0");
    if (!h->top_edge || !h->bottom_edge || h->connected) return fail(HSFLOW_EINVAL, "not available in strip mode");
    CK(cudaSetDevice(h->device));
    const long long px = (long long)w * hgt;
    const int bpp = o.fmt == FMT_BGR8 ? 3 : 1;
    const int K = 3;                               // sub-batches in flight: H2D | compute | D2H
    // Pairs per sub-batch: enough pixels that a compute launch runs several waves of work units (units drift apart and
    // keep HBM and the SMs busy through each other's fill and drain phases): 128 Mi pixels = 16 4K pairs,
    // at most 64 pairs, and never more than a third of the free device memory for the K slots.
    int B = (int)std::max<long long>(1, std::min<long long>(64, (128LL << 20) / px));
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = 0; }
        if (h->uA && h->W == w && h->H == hgt) free_b += (size_t)h->uv_pp * 4 * (h->P + h->S) + (size_t)h->c_pp * 4 * h->S + h->stage_bytes;
        const double per_pair = (double)px * (K * (8.0 + 2.0 * bpp + 8.0) + 20.0) * 1.1;
        if (free_b) B = (int)std::max(1.0, std::min((double)B, (double)free_b / 3.0 / per_pair));
    }
    if (h->sub_batch > 0) B = std::min(B, h->sub_batch);    // hsflow_set_tuning: tests force several ragged sub-batches
    B = std::min(B, n_pairs);
    const int slot_pairs = B + (o.sequence ? 1 : 0); // frame (and result) slots per sub-batch
    if (!(h->W == w && h->H == hgt && h->P == K * slot_pairs && h->S == B)) {
        const int keep = h->sub_batch;
        CK(cudaStreamSynchronize(h->stream));
        free_planes(h); h->W = h->H = h->P = 0;
        h->sub_batch = B;
        int rc = hsflow_configure(h, w, hgt, K * slot_pairs);
        h->sub_batch = keep;
        if (rc) return rc;
    }
    int rc = ensure_frames(h, o.fmt);
    if (rc) return rc;
    if (!h->s_in) { CK(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking)); }
    // Staging slots for the read-back.  Full fields: the last block of a sub-batch stores PLANAR, densely packed
    // fields straight into the slot when it runs in the streaming kernel (its output addressing is free) and rows of
    // W floats keep the float4 stores aligned.  Sampled fields: a gather kernel fills the slot from the A planes.
    const int T_eff = effective_T(h);
    const int T_last = h->iterations > 0 ? (h->iterations % T_eff ? h->iterations % T_eff : T_eff) : 0;
    const int step = o.sample_step;
    const int gw = step ? (w + step - 1) / step : w, gh = step ? (hgt + step - 1) / step : hgt;
    const size_t opx = (size_t)gw * gh;            // floats per field and pair that go home
    const bool planar = !step && (w % 4 == 0) && h->iterations > 0 && use_stream_kernel(h, T_last) && !(h->eps > 0.0);
    const bool staged = planar || step > 0;
    const size_t slot_floats = 2 * (size_t)B * opx;
    if (staged && h->stage_bytes < K * slot_floats * sizeof(float)) {
        CK(cudaStreamSynchronize(h->stream));
        cudaFree(h->stage); h->stage = nullptr; h->stage_bytes = 0;
        if (cudaMalloc(&h->stage, K * slot_floats * sizeof(float)) != cudaSuccess) { cudaGetLastError(); return fail(HSFLOW_ENOMEM, "cudaMalloc of the read-back staging slots failed"); }
        h->stage_bytes = K * slot_floats * sizeof(float);
    }
    struct Events {                                // destroyed on every exit path, error returns included
        cudaEvent_t e[3 * K] = {};
        ~Events() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); }
    } evs;
    cudaEvent_t *ev_in = evs.e, *ev_comp = evs.e + K, *ev_out = evs.e + 2 * K;
    for (int k = 0; k < 3 * K; ++k) CK(cudaEventCreateWithFlags(&evs.e[k], cudaEventDisableTiming));
    // order the side streams after whatever the handle's stream did so far (allocation memsets)
    CK(cudaEventRecord(ev_comp[0], h->stream));
    CK(cudaStreamWaitEvent(h->s_in, ev_comp[0], 0));
    CK(cudaStreamWaitEvent(h->s_out, ev_comp[0], 0));
    const size_t fbytes = (size_t)px * bpp, rowb = (size_t)w * bpp, wb = (size_t)w * sizeof(float);
    // Sub-batch sizes ramp up from one pair: the first result can only leave after upload + compute of the FIRST
    // sub-batch.  With full fields the read-back is the bottleneck (8 bytes of u, v per pixel against 1-3 bytes of
    // frames), so the sizes also ramp down to one pair again: what is left after the last compute is the read-back of
    // the LAST sub-batch.  Sampled fields are small: no ramp down.
    std::vector<int> sizes;
    {
        std::vector<int> head, tail;
        int left = n_pairs;
        const int ends = step ? 1 : 2;
        for (int s = 1; s < B && left > ends * B; s *= 2) {
            head.push_back(s); left -= s;
            if (!step) { tail.push_back(s); left -= s; }
        }
        sizes = head;
        for (; left > 0; left -= B) sizes.push_back(std::min(B, left));
        sizes.insert(sizes.end(), tail.rbegin(), tail.rend());
    }
    // From here on copies into the caller's buffers may be in flight: every failure leaves through the common exit
    // below, which drains all three streams before the caller gets its buffers back.
    int status = HSFLOW_OK, first = 0;
#define PK(call)                                                                                                        \
    do {                                                                                                               \
        cudaError_t e_ = (call);                                                                                       \
        if (e_ != cudaSuccess) {                                                                                       \
            status = fail(HSFLOW_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);  \
            goto drain;                                                                                                \
        }                                                                                                              \
    } while (0)
    for (size_t i = 0; i < sizes.size(); first += sizes[i], ++i) {
        const int slot = (int)(i % K), p0 = slot * slot_pairs, n = sizes[i];
        if (i >= (size_t)K) PK(cudaStreamWaitEvent(h->s_in, ev_comp[slot], 0));      // frames of the slot were consumed
        if (o.sequence) {
            for (int k = 0; k <= n; ++k)
                PK(cudaMemcpy2DAsync(h->f1 + (size_t)(p0 + k) * h->f_pair_pitch, h->f_row_pitch, frames + (size_t)(first + k) * fbytes,
                                     rowb, rowb, hgt, cudaMemcpyHostToDevice, h->s_in));
        } else {
            for (int k = 0; k < n; ++k) {
                const uint8_t* src = frames + (size_t)(first + k) * 2 * fbytes;
                PK(cudaMemcpy2DAsync(h->f1 + (size_t)(p0 + k) * h->f_pair_pitch, h->f_row_pitch, src, rowb, rowb, hgt, cudaMemcpyHostToDevice, h->s_in));
                PK(cudaMemcpy2DAsync(h->f2 + (size_t)(p0 + k) * h->f_pair_pitch, h->f_row_pitch, src + fbytes, rowb, rowb, hgt, cudaMemcpyHostToDevice, h->s_in));
            }
        }
        PK(cudaEventRecord(ev_in[slot], h->s_in));
        PK(cudaStreamWaitEvent(h->stream, ev_in[slot], 0));
        if (i >= (size_t)K) PK(cudaStreamWaitEvent(h->stream, ev_out[slot], 0));     // u/v of the slot were read back
        float* su = staged ? h->stage + slot * slot_floats : nullptr;                  // U block [n][gh][gw], then V block
        float* sv = staged ? su + (size_t)n * opx : nullptr;
        status = o.sequence ? compute_subbatch(h, p0, n, h->f1, h->f1 + h->f_pair_pitch, planar ? su : nullptr, planar ? sv : nullptr)
                            : compute_subbatch(h, p0, n, nullptr, nullptr, planar ? su : nullptr, planar ? sv : nullptr);
        if (status) goto drain;
        if (step) {
            PK(launch_sample_uv(h->uA + (size_t)p0 * h->uv_pp, h->vA + (size_t)p0 * h->uv_pp, w, hgt, h->uv_rp, h->uv_pp, step, su, sv, n, h->stream));
            h->launches++;
        }
        PK(cudaEventRecord(ev_comp[slot], h->stream));
        PK(cudaStreamWaitEvent(h->s_out, ev_comp[slot], 0));
        if (staged) {
            PK(cudaMemcpyAsync(u_out + (size_t)first * opx, su, (size_t)n * opx * sizeof(float), cudaMemcpyDeviceToHost, h->s_out));
            PK(cudaMemcpyAsync(v_out + (size_t)first * opx, sv, (size_t)n * opx * sizeof(float), cudaMemcpyDeviceToHost, h->s_out));
        } else for (int k = 0; k < n; ++k) {
            PK(cudaMemcpy2DAsync(u_out + (size_t)(first + k) * px, wb, h->uA + (size_t)(p0 + k) * h->uv_pp, h->uv_rp * sizeof(float), wb, hgt, cudaMemcpyDeviceToHost, h->s_out));
            PK(cudaMemcpy2DAsync(v_out + (size_t)(first + k) * px, wb, h->vA + (size_t)(p0 + k) * h->uv_pp, h->uv_rp * sizeof(float), wb, hgt, cudaMemcpyDeviceToHost, h->s_out));
        }
        PK(cudaEventRecord(ev_out[slot], h->s_out));
    }
#undef PK
drain:
    {
        // the caller's frames / u_out / v_out are ours until all three streams are idle, error or not
        const cudaError_t e1 = cudaStreamSynchronize(h->s_in), e2 = cudaStreamSynchronize(h->stream), e3 = cudaStreamSynchronize(h->s_out);
        // The slots hold the fields of the last K sub-batches (and, with planar staging, not even those): there is no
        // "current field" a later hsflow_read_uv / hsflow_dot_mask could mean.
        h->cur = 0; h->prepared = 0; h->zero_pending = 0; h->uv_valid = 0;
        if (status) return status;
        const cudaError_t e = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
        if (e != cudaSuccess) return fail(HSFLOW_ECUDA, "pipeline drain: %s", cudaGetErrorString(e));
    }
    CK(cudaGetLastError());
    return HSFLOW_OK;
}

int hsflow_run_batch_host(hsflow_t* h, const uint8_t* frames, int n_pairs, int w, int hgt, float* u_out, float* v_out) {
    NEED(h);
    return run_pipeline(h, frames, n_pairs, w, hgt, u_out, v_out, PipeOpts{FMT_GRAY8, 0, 0});
}
int hsflow_run_sequence_host(hsflow_t* h, const uint8_t* frames, int n_frames, int w, int hgt, float* u_out, float* v_out) {
    NEED(h);
    if (n_frames < 2) return fail(HSFLOW_EINVAL, "a sequence needs at least two frames");
    return run_pipeline(h, frames, n_frames - 1, w, hgt, u_out, v_out, PipeOpts{FMT_GRAY8, 1, 0});
}
int hsflow_run_pipeline_host(hsflow_t* h, const uint8_t* frames, int n_pairs, int w, int hgt, int frame_format, int flags,
                             int sample_step, float* u_out, float* v_out) {
    NEED(h);
    if (flags & ~HSFLOW_PIPE_SEQUENCE) return fail(HSFLOW_EINVAL, "unknown pipeline flags 0x%x", flags);
    const int fmt = frame_format == HSFLOW_FRAMES_GRAY8 ? FMT_GRAY8 : (frame_format == HSFLOW_FRAMES_BGR8 ? FMT_BGR8 : -1);
    return run_pipeline(h, frames, n_pairs, w, hgt, u_out, v_out, PipeOpts{fmt, (flags & HSFLOW_PIPE_SEQUENCE) ? 1 : 0, sample_step});
}

// Pair sharding inside one process (independent frame pairs, contiguous blocks of ceil(P / n) pairs per GPU, no
// data-path collective): one host thread per handle runs the pipelined call on its block.  The handles usually live on
// different GPUs; the caller created them and set their parameters.
int hsflow_run_pipeline_host_multi(hsflow_t* const* handles, int n_handles, const uint8_t* frames, int n_pairs, int w, int hgt,
                                   int frame_format, int flags, int sample_step, float* u_out, float* v_out) {
    if (!handles || n_handles <= 0) return fail(HSFLOW_EINVAL, "no handles");
    for (int k = 0; k < n_handles; ++k) if (!handles[k]) return fail(HSFLOW_EINVAL, "null handle");
    if (!frames || !u_out || !v_out || n_pairs <= 0 || w <= 0 || hgt <= 0) return fail(HSFLOW_EINVAL, "bad argument");
    if (flags & ~HSFLOW_PIPE_SEQUENCE) return fail(HSFLOW_EINVAL, "unknown pipeline flags 0x%x", flags);
    if (sample_step < 0) return fail(HSFLOW_EINVAL, "sample_step must be >= 0");
    const int fmt = frame_format == HSFLOW_FRAMES_GRAY8 ? FMT_GRAY8 : (frame_format == HSFLOW_FRAMES_BGR8 ? FMT_BGR8 : -1);
    if (fmt < 0) return fail(HSFLOW_EINVAL, "frame format must be gray8 or bgr8");
    const int seq = (flags & HSFLOW_PIPE_SEQUENCE) ? 1 : 0;
    const size_t fbytes = (size_t)w * hgt * (fmt == FMT_BGR8 ? 3 : 1);
    const size_t opx = sample_step ? (size_t)((w + sample_step - 1) / sample_step) * ((hgt + sample_step - 1) / sample_step) : (size_t)w * hgt;
    const int per = (n_pairs + n_handles - 1) / n_handles;
    std::vector<int> rc((size_t)n_handles, HSFLOW_OK);
    std::vector<std::string> msg((size_t)n_handles);
    std::vector<std::thread> th;
    for (int k = 0; k < n_handles; ++k) {
        const int lo = std::min(k * per, n_pairs), hi = std::min(lo + per, n_pairs);
        if (lo >= hi) continue;
        th.emplace_back([&, k, lo, hi] {
            const uint8_t* f = frames + (size_t)lo * (seq ? 1 : 2) * fbytes;      // a sequence block starts at its first frame
            rc[(size_t)k] = run_pipeline(handles[k], f, hi - lo, w, hgt, u_out + (size_t)lo * opx, v_out + (size_t)lo * opx,
                                         PipeOpts{fmt, seq, sample_step});
            if (rc[(size_t)k]) msg[(size_t)k] = g_err;                              // g_err is thread-local: carry it out
        });
    }
    for (std::thread& t : th) t.join();
    for (int k = 0; k < n_handles; ++k)
        if (rc[(size_t)k]) return fail(rc[(size_t)k], "handle %d: %s", k, msg[(size_t)k].c_str());
    return HSFLOW_OK;
}

// Camera-loop step on the device (cpp:800-842): the second frame of the handle's pair becomes the first one (cpp:834
// copies I2 over I1; here the two plane pointers swap) and `frame` is uploaded as the new second frame.
int hsflow_push_frame_gray8(hsflow_t* h, const uint8_t* frame, size_t pitch) {
    NEED(h);
    if (!frame) return fail(HSFLOW_EINVAL, "null frame pointer");
    if (h->P != 1) return fail(HSFLOW_EINVAL, "hsflow_push_frame_gray8 needs a handle configured for one pair");
    const bool first = h->fmt != FMT_GRAY8 || !h->f1;
    int rc = ensure_frames(h, FMT_GRAY8);
    if (rc) return rc;
    const size_t wbytes = (size_t)h->W;
    if (pitch == 0) pitch = wbytes;
    if (pitch < wbytes) return fail(HSFLOW_EINVAL, "pitch %zu smaller than a row (%zu bytes)", pitch, wbytes);
    phase_begin(h, HSFLOW_PHASE_LOAD);
    std::swap(h->f1, h->f2);
    CK(cudaMemcpy2DAsync(h->f2, h->f_row_pitch, frame, pitch, wbytes, h->H, cudaMemcpyHostToDevice, h->stream));
    if (first)                                     // very first frame: both planes hold it (zero flow until the next push)
        CK(cudaMemcpy2DAsync(h->f1, h->f_row_pitch, frame, pitch, wbytes, h->H, cudaMemcpyHostToDevice, h->stream));
    phase_end(h, HSFLOW_PHASE_LOAD);
    h->prepared = 0;
    return HSFLOW_OK;
}

float hsflow_last_ms(hsflow_t* h, int phase) {
    if (!h || phase < 0 || phase > 3 || !h->ev_set[phase]) return -1.f;
    cudaSetDevice(h->device);
    if (cudaEventSynchronize(h->ev1[phase]) != cudaSuccess) return -1.f;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, h->ev0[phase], h->ev1[phase]) != cudaSuccess) { cudaGetLastError(); return -1.f; }
    return ms;
}
long long hsflow_kernel_launches(hsflow_t* h) { return h ? h->launches : 0; }
int hsflow_sub_batch(hsflow_t* h) { return h ? h->S : 0; }
int hsflow_iterations_done(hsflow_t* h, int pair, int* done) {
    NEED(h);
    if (pair < 0 || pair >= h->P || !done) return fail(HSFLOW_EINVAL, "bad argument");
    *done = h->sweeps;
    if (h->eps > 0.0 && h->d_stop && pair < h->eps_cap) {
        int st = 0;
        CK(cudaMemcpyAsync(&st, h->d_stop + pair, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (st) *done = st >> 1;
    }
    return HSFLOW_OK;
}
int hsflow_effective_temporal_block(hsflow_t* h) { return h ? effective_T(h) : 0; }
int hsflow_device(hsflow_t* h) { return h ? h->device : -1; }

void* hsflow_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void hsflow_free_pinned(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
