// hs_ingest.cpp -- JPEG ingest on the GPU (include/hsflow_ingest.h).
//
// Replaces cvLoadImage + cvCvtColor + readInputImage (HSOpticalFlowOpenCL.cpp:721-740, 6-45): the reference decodes on
// the CPU, converts to gray on the CPU, widens every pixel to a float4 through per-pixel cvGet2D calls and uploads
// 16 B/px.  Here nvJPEG writes interleaved BGR bytes straight into the engine's frame planes in HBM
// (hsflow_map_frames) and the derivative kernel converts to gray on the fly: 3 B/px written once, nothing on the host.
#include "../../../include/hsflow_ingest.h"

#include <cuda_runtime.h>
#include <nvjpeg.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "hs_image.h"

namespace {

thread_local char g_err[512] = "";
int fail(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
    return code;
}
#define CKC(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) return fail(HSFLOW_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_));  \
    } while (0)
#define CKH(call)                                                                                          \
    do {                                                                                                   \
        int rc_ = (call);                                                                                  \
        if (rc_ != HSFLOW_OK) return fail(rc_, "%s", hsflow_last_error());                                 \
    } while (0)

// One decoder per process: a plain handle for single images and a batched state on the backend HSFLOW_NVJPEG_BACKEND
// names ("default" | "hybrid" | "gpu" | "hardware"; default = "default": Huffman stage on the host, IDCT + colour
// conversion on the GPU).  Batches first try the parallel decoupled path below (decode_parallel) and fall back to
// nvjpegDecodeBatched.  Measured on the B200 boxes with 4K frames (tools/ingest_probe.py), nvjpegDecodeBatched: default
// 56-58 images/s with 1 or 16 library threads, GPU-assisted Huffman ("gpu") 25 images/s; the hardware engine
// ("hardware") is refused by nvjpegCreateEx on this driver and falls back to default.
struct Decoder {
    nvjpegHandle_t single = nullptr, batched = nullptr;
    nvjpegJpegState_t st_single = nullptr, st_batched = nullptr;
    cudaStream_t stream = nullptr;
    int backend = 0, batch_size = 0, threads = 1, last_path = 0;   // last_path 1: parallel decoupled decode, 0: nvjpegDecodeBatched
    bool ok = false;
    Decoder() {
        if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) return;
        if (nvjpegCreateSimple(&single) != NVJPEG_STATUS_SUCCESS) return;
        if (nvjpegJpegStateCreate(single, &st_single) != NVJPEG_STATUS_SUCCESS) return;
        const char* want = getenv("HSFLOW_NVJPEG_BACKEND");
        nvjpegBackend_t order[2] = {NVJPEG_BACKEND_DEFAULT, NVJPEG_BACKEND_DEFAULT};
        if (want && !strcmp(want, "hardware")) order[0] = NVJPEG_BACKEND_HARDWARE;
        else if (want && !strcmp(want, "hybrid")) order[0] = NVJPEG_BACKEND_HYBRID;
        else if (want && !strcmp(want, "gpu")) order[0] = NVJPEG_BACKEND_GPU_HYBRID;
        const char* th = getenv("HSFLOW_NVJPEG_THREADS");
        threads = th ? atoi(th) : 0;
        if (threads <= 0) {
            cpu_set_t set;
            CPU_ZERO(&set);
            threads = 1;                           // measured: more threads do not speed nvjpegDecodeBatched up on this backend
            (void)set;
        }
        for (nvjpegBackend_t b : order) {
            if (nvjpegCreateEx(b, nullptr, nullptr, NVJPEG_FLAGS_DEFAULT, &batched) == NVJPEG_STATUS_SUCCESS &&
                nvjpegJpegStateCreate(batched, &st_batched) == NVJPEG_STATUS_SUCCESS) {
                backend = (int)b;
                break;
            }
            if (batched) { nvjpegDestroy(batched); batched = nullptr; }
            st_batched = nullptr;
        }
        ok = batched != nullptr;
    }
};
// one decoder per device, created on first use with that device current (every entry point that takes a handle makes
// the handle's device current first: use_device)
constexpr int kMaxDevices = 64;
Decoder& dec() {
    static Decoder* d[kMaxDevices] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) { cudaGetLastError(); dev = 0; }
    if (!d[dev]) d[dev] = new Decoder();
    return *d[dev];
}
int use_device(hsflow_t* h) {
    const int dev = hsflow_device(h);
    if (dev < 0) return fail(HSFLOW_EINVAL, "null handle");
    CKC(cudaSetDevice(dev));
    return HSFLOW_OK;
}

bool is_jpeg(const uint8_t* p, size_t n) { return n >= 2 && p[0] == 0xFF && p[1] == 0xD8; }

bool read_file(const char* path, std::vector<uint8_t>& buf) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    if (n < 0) { fclose(f); return false; }
    buf.resize((size_t)n);
    const bool ok = n == 0 || fread(buf.data(), 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    return ok;
}

int jpeg_info(const uint8_t* jpeg, size_t len, int* w, int* h, int* comps) {
    Decoder& D = dec();
    if (!D.ok) return fail(HSFLOW_ENODEV, "nvJPEG initialisation failed (no CUDA device?)");
    int nc = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
    nvjpegChromaSubsampling_t ss;
    if (!jpeg || nvjpegGetImageInfo(D.single, jpeg, len, &nc, &ss, ws, hs) != NVJPEG_STATUS_SUCCESS)
        return fail(HSFLOW_EINVAL, "not a decodable JPEG stream");
    if (w) *w = ws[0];
    if (h) *h = hs[0];
    if (comps) *comps = nc;
    return HSFLOW_OK;
}

int decode_one(const uint8_t* jpeg, size_t len, uint8_t* d_bgr, size_t pitch, int w, int h, cudaStream_t s) {
    Decoder& D = dec();
    int iw = 0, ih = 0;
    int rc = jpeg_info(jpeg, len, &iw, &ih, nullptr);
    if (rc) return rc;
    if (iw != w || ih != h) return fail(HSFLOW_EINVAL, "JPEG is %d x %d, expected %d x %d", iw, ih, w, h);
    if (!d_bgr || pitch < (size_t)w * 3) return fail(HSFLOW_EINVAL, "bad destination");
    nvjpegImage_t img;
    memset(&img, 0, sizeof img);
    img.channel[0] = d_bgr; img.pitch[0] = pitch;
    const nvjpegStatus_t st = nvjpegDecode(D.single, D.st_single, jpeg, len, NVJPEG_OUTPUT_BGRI, &img, s);
    if (st != NVJPEG_STATUS_SUCCESS) return fail(HSFLOW_ECUDA, "nvjpegDecode failed (%d)", (int)st);
    return HSFLOW_OK;
}

// ---- parallel decode: nvJPEG's decoupled API, one worker context per host thread -------------------------------------
// The Huffman stage of the hybrid decoder runs on the host and is what bounds a 4K clip (17 ms per frame on one core,
// against 1 ms of Horn-Schunck compute per pair): HSFLOW_NVJPEG_THREADS workers (default: the cores this process may
// use, at most 16) each own a decoder state with its pinned and device buffers and a CUDA stream, take images off a
// shared counter, run the host stage (nvjpegDecodeJpegHost) and queue the device stages (transfer, IDCT, colour
// conversion straight into the destination plane) on their stream.
struct Worker {
    nvjpegHandle_t h = nullptr;                    // a library handle of its own: workers sharing one handle serialise inside nvJPEG
    nvjpegJpegDecoder_t dec = nullptr;
    nvjpegJpegState_t st = nullptr;
    nvjpegBufferPinned_t pin = nullptr;
    nvjpegBufferDevice_t dev = nullptr;
    nvjpegJpegStream_t js = nullptr;
    nvjpegDecodeParams_t par = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    bool ok = false;
    bool init() {
        if (nvjpegCreateSimple(&h) != NVJPEG_STATUS_SUCCESS) return false;
        if (nvjpegDecoderCreate(h, NVJPEG_BACKEND_HYBRID, &dec) != NVJPEG_STATUS_SUCCESS) return false;
        if (nvjpegDecoderStateCreate(h, dec, &st) != NVJPEG_STATUS_SUCCESS) return false;
        if (nvjpegBufferPinnedCreate(h, nullptr, &pin) != NVJPEG_STATUS_SUCCESS) return false;
        if (nvjpegBufferDeviceCreate(h, nullptr, &dev) != NVJPEG_STATUS_SUCCESS) return false;
        if (nvjpegJpegStreamCreate(h, &js) != NVJPEG_STATUS_SUCCESS) return false;
        if (nvjpegDecodeParamsCreate(h, &par) != NVJPEG_STATUS_SUCCESS) return false;
        if (nvjpegDecodeParamsSetOutputFormat(par, NVJPEG_OUTPUT_BGRI) != NVJPEG_STATUS_SUCCESS) return false;
        if (nvjpegStateAttachPinnedBuffer(st, pin) != NVJPEG_STATUS_SUCCESS) return false;
        if (nvjpegStateAttachDeviceBuffer(st, dev) != NVJPEG_STATUS_SUCCESS) return false;
        if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) return false;
        if (cudaEventCreateWithFlags(&done, cudaEventDisableTiming) != cudaSuccess) return false;
        return ok = true;
    }
};
struct Pool {
    std::vector<Worker> w;
    int device = 0;
    bool tried = false;
};
Pool& pool() {                                     // per device, like the decoder
    static Pool* p[kMaxDevices] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) { cudaGetLastError(); dev = 0; }
    if (!p[dev]) p[dev] = new Pool();
    return *p[dev];
}

int pool_threads() {
    const char* th = getenv("HSFLOW_NVJPEG_THREADS");
    int n = th ? atoi(th) : 0;
    if (n <= 0) {
        cpu_set_t set;
        CPU_ZERO(&set);
        n = sched_getaffinity(0, sizeof set, &set) == 0 ? CPU_COUNT(&set) : 1;
        n = std::min(n, 16);
    }
    return std::max(1, n);
}

// returns HSFLOW_OK, or a negative code when the parallel path cannot take this batch (the caller falls back)
int decode_parallel(const uint8_t* const* jpegs, const size_t* sizes, uint8_t* const* dsts, int n, size_t pitch) {
    Pool& P = pool();
    const int want = std::min(pool_threads(), n);
    if (want < 2) return HSFLOW_EINVAL;
    if (!P.tried) {
        P.tried = true;
        cudaGetDevice(&P.device);
        P.w.resize((size_t)pool_threads());
        for (Worker& wk : P.w) if (!wk.init()) { P.w.clear(); break; }
    }
    if (P.w.empty()) return HSFLOW_EINVAL;
    std::atomic<int> next(0), failed(0);
    auto body = [&](Worker* wk) {
        cudaSetDevice(P.device);
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n || failed.load()) break;
            nvjpegImage_t img;
            memset(&img, 0, sizeof img);
            img.channel[0] = dsts[i]; img.pitch[0] = pitch;
            bool ok = nvjpegJpegStreamParse(wk->h, jpegs[i], sizes[i], 0, 0, wk->js) == NVJPEG_STATUS_SUCCESS &&
                      nvjpegDecodeJpegHost(wk->h, wk->dec, wk->st, wk->par, wk->js) == NVJPEG_STATUS_SUCCESS &&
                      nvjpegDecodeJpegTransferToDevice(wk->h, wk->dec, wk->st, wk->js, wk->stream) == NVJPEG_STATUS_SUCCESS &&
                      nvjpegDecodeJpegDevice(wk->h, wk->dec, wk->st, &img, wk->stream) == NVJPEG_STATUS_SUCCESS;
            // the state's pinned buffer is reused by the next image of this worker: its transfer has to be over
            if (ok) ok = cudaStreamSynchronize(wk->stream) == cudaSuccess;
            if (!ok) { failed.store(1); break; }
        }
    };
    std::vector<std::thread> th;
    for (int k = 1; k < want; ++k) th.emplace_back(body, &P.w[(size_t)k]);
    body(&P.w[0]);
    for (std::thread& t : th) t.join();
    if (failed.load()) { cudaGetLastError(); return HSFLOW_ECUDA; }
    return HSFLOW_OK;                              // every worker stream is idle: the planes are complete
}

// Batched decode of `n` bitstreams into `n` device destinations (all w x h, pitch bytes per row), on D.stream.
int decode_batch(const uint8_t* const* jpegs, const size_t* sizes, uint8_t* const* dsts, int n, size_t pitch, int w, int h) {
    Decoder& D = dec();
    if (!D.ok) return fail(HSFLOW_ENODEV, "nvJPEG initialisation failed (no CUDA device?)");
    for (int k = 0; k < n; ++k) {
        int iw = 0, ih = 0;
        int rc = jpeg_info(jpegs[k], sizes[k], &iw, &ih, nullptr);
        if (rc) return rc;
        if (iw != w || ih != h) return fail(HSFLOW_EINVAL, "image %d is %d x %d, expected %d x %d", k, iw, ih, w, h);
    }
    if (decode_parallel(jpegs, sizes, dsts, n, pitch) == HSFLOW_OK) { D.last_path = 1; return HSFLOW_OK; }
    D.last_path = 0;
    if (D.batch_size != n) {
        const nvjpegStatus_t st = nvjpegDecodeBatchedInitialize(D.batched, D.st_batched, n, D.threads, NVJPEG_OUTPUT_BGRI);
        if (st != NVJPEG_STATUS_SUCCESS) return fail(HSFLOW_ECUDA, "nvjpegDecodeBatchedInitialize(%d) failed (%d)", n, (int)st);
        D.batch_size = n;
    }
    std::vector<nvjpegImage_t> out((size_t)n);
    memset(out.data(), 0, sizeof(nvjpegImage_t) * (size_t)n);
    for (int k = 0; k < n; ++k) { out[k].channel[0] = dsts[k]; out[k].pitch[0] = pitch; }
    nvjpegStatus_t st = nvjpegDecodeBatched(D.batched, D.st_batched, jpegs, sizes, out.data(), D.stream);
    if (st != NVJPEG_STATUS_SUCCESS) {
        // e.g. a progressive stream on a backend that only takes baseline: one image at a time on the plain decoder
        D.batch_size = 0;
        for (int k = 0; k < n; ++k) {
            int rc = decode_one(jpegs[k], sizes[k], dsts[k], pitch, w, h, D.stream);
            if (rc) return rc;
        }
    }
    return HSFLOW_OK;
}

}  // namespace

extern "C" {

const char* hsingest_last_error(void) { return g_err; }

int hsingest_jpeg_info(const uint8_t* jpeg, size_t len, int* w, int* h, int* comps) { return jpeg_info(jpeg, len, w, h, comps); }

int hsingest_decode_to_device(const uint8_t* jpeg, size_t len, uint8_t* d_bgr, size_t pitch, int w, int h, void* cuda_stream) {
    Decoder& D = dec();
    if (!D.ok) return fail(HSFLOW_ENODEV, "nvJPEG initialisation failed (no CUDA device?)");
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : D.stream;
    int rc = decode_one(jpeg, len, d_bgr, pitch, w, h, s);
    if (rc) return rc;
    if (!cuda_stream) CKC(cudaStreamSynchronize(s));
    return HSFLOW_OK;
}

int hsingest_load_pair_jpeg(hsflow_t* h, const uint8_t* j1, size_t n1, const uint8_t* j2, size_t n2, int* width, int* height) {
    if (!h) return fail(HSFLOW_EINVAL, "null handle");
    { int rc0 = use_device(h); if (rc0) return rc0; }
    int w = 0, ht = 0, w2 = 0, h2 = 0;
    int rc;
    if ((rc = jpeg_info(j1, n1, &w, &ht, nullptr)) || (rc = jpeg_info(j2, n2, &w2, &h2, nullptr))) return rc;
    if (w != w2 || ht != h2) return fail(HSFLOW_EINVAL, "the two frames differ in size (%d x %d vs %d x %d)", w, ht, w2, h2);
    CKH(hsflow_configure(h, w, ht, 1));
    uint8_t *d1 = nullptr, *d2 = nullptr;
    size_t rp = 0, pp = 0;
    CKH(hsflow_map_frames(h, HSFLOW_FRAMES_BGR8, &d1, &d2, &rp, &pp));
    Decoder& D = dec();
    if ((rc = decode_one(j1, n1, d1, rp, w, ht, D.stream)) || (rc = decode_one(j2, n2, d2, rp, w, ht, D.stream))) return rc;
    CKC(cudaStreamSynchronize(D.stream));
    if (width) *width = w;
    if (height) *height = ht;
    return HSFLOW_OK;
}

int hsingest_load_pair_files(hsflow_t* h, const char* p1, const char* p2, int* width, int* height) {
    if (!h || !p1 || !p2) return fail(HSFLOW_EINVAL, "null argument");
    { int rc0 = use_device(h); if (rc0) return rc0; }
    std::vector<uint8_t> b1, b2;
    if (!read_file(p1, b1)) return fail(HSFLOW_EINVAL, "cannot read %s", p1);
    if (!read_file(p2, b2)) return fail(HSFLOW_EINVAL, "cannot read %s", p2);
    if (is_jpeg(b1.data(), b1.size()) && is_jpeg(b2.data(), b2.size()))
        return hsingest_load_pair_jpeg(h, b1.data(), b1.size(), b2.data(), b2.size(), width, height);
    // PGM / PPM (or a mix): decoded on the host by hsimg_read, uploaded as they are; the kernel does the gray conversion
    int w[2], ht[2], ch[2];
    uint8_t* px[2] = {nullptr, nullptr};
    const char* paths[2] = {p1, p2};
    for (int k = 0; k < 2; ++k)
        if (hsimg_read(paths[k], &w[k], &ht[k], &ch[k], &px[k]) != 0) {
            if (k) hsimg_free(px[0]);
            return fail(HSFLOW_EINVAL, "%s: %s", paths[k], hsimg_last_error());
        }
    int rc = HSFLOW_OK;
    if (w[0] != w[1] || ht[0] != ht[1]) rc = fail(HSFLOW_EINVAL, "the two frames differ in size");
    else if (ch[0] == 1 && ch[1] == 1) { rc = hsflow_load_pair_gray8(h, px[0], px[1], w[0], ht[0], 0); if (rc) fail(rc, "%s", hsflow_last_error()); }
    else {
        std::vector<uint8_t> tmp[2];
        for (int k = 0; k < 2; ++k)
            if (ch[k] == 1) {                      // mixed pair: replicate the gray frame to BGR (its gray conversion is the identity)
                tmp[k].resize((size_t)w[k] * ht[k] * 3);
                for (size_t i = 0; i < (size_t)w[k] * ht[k]; ++i) tmp[k][3 * i] = tmp[k][3 * i + 1] = tmp[k][3 * i + 2] = px[k][i];
            }
        rc = hsflow_load_pair_bgr8(h, tmp[0].empty() ? px[0] : tmp[0].data(), tmp[1].empty() ? px[1] : tmp[1].data(), w[0], ht[0], 0);
        if (rc) fail(rc, "%s", hsflow_last_error());
    }
    if (rc == HSFLOW_OK) { rc = hsflow_sync(h); if (rc) fail(rc, "%s", hsflow_last_error()); }   // the host buffers go away below
    hsimg_free(px[0]); hsimg_free(px[1]);
    if (rc == HSFLOW_OK) { if (width) *width = w[0]; if (height) *height = ht[0]; }
    return rc;
}

int hsingest_push_frame_file(hsflow_t* h, const char* path) {
    if (!h || !path) return fail(HSFLOW_EINVAL, "null argument");
    { int rc0 = use_device(h); if (rc0) return rc0; }
    std::vector<uint8_t> b;
    if (!read_file(path, b)) return fail(HSFLOW_EINVAL, "cannot read %s", path);
    uint8_t *d1 = nullptr, *d2 = nullptr;
    size_t rp = 0, pp = 0;
    if (is_jpeg(b.data(), b.size())) {
        int w = 0, ht = 0;
        int rc = jpeg_info(b.data(), b.size(), &w, &ht, nullptr);
        if (rc) return rc;
        CKH(hsflow_map_frames(h, HSFLOW_FRAMES_BGR8, &d1, &d2, &rp, &pp));
        CKH(hsflow_swap_frames(h));                // cpp:834: the previous second frame is now the first one
        if ((rc = decode_one(b.data(), b.size(), d1, rp, w, ht, dec().stream))) return rc;   // d1 = the plane that is now "second"
        CKC(cudaStreamSynchronize(dec().stream));
        return HSFLOW_OK;
    }
    int w = 0, ht = 0, ch = 0;
    uint8_t* px = nullptr;
    if (hsimg_read(path, &w, &ht, &ch, &px) != 0) return fail(HSFLOW_EINVAL, "%s: %s", path, hsimg_last_error());
    int rc = hsflow_map_frames(h, ch == 3 ? HSFLOW_FRAMES_BGR8 : HSFLOW_FRAMES_GRAY8, &d1, &d2, &rp, &pp);
    if (rc == HSFLOW_OK) rc = hsflow_swap_frames(h);
    if (rc != HSFLOW_OK) { hsimg_free(px); return fail(rc, "%s", hsflow_last_error()); }
    const cudaError_t e = cudaMemcpy2D(d1, rp, px, (size_t)w * ch, (size_t)w * ch, ht, cudaMemcpyHostToDevice);
    hsimg_free(px);
    if (e != cudaSuccess) return fail(HSFLOW_ECUDA, "cudaMemcpy2D: %s", cudaGetErrorString(e));
    return HSFLOW_OK;
}

int hsingest_run_jpeg_batch(hsflow_t* h, const uint8_t* const* jpegs, const size_t* sizes, int n_images, int sequence,
                            int sample_step, float* u_out, float* v_out, double stats[4]) {
    if (!h || !jpegs || !sizes || !u_out || !v_out) return fail(HSFLOW_EINVAL, "null argument");
    const int n_pairs = sequence ? n_images - 1 : n_images / 2;
    if (n_pairs < 1 || (!sequence && (n_images & 1))) return fail(HSFLOW_EINVAL, "need an even number of images (pairs) or >= 2 consecutive frames");
    if (sample_step < 0) return fail(HSFLOW_EINVAL, "sample_step must be >= 0");
    { int rc0 = use_device(h); if (rc0) return rc0; }
    Decoder& D = dec();
    if (!D.ok) return fail(HSFLOW_ENODEV, "nvJPEG initialisation failed (no CUDA device?)");
    int w = 0, ht = 0;
    int rc = jpeg_info(jpegs[0], sizes[0], &w, &ht, nullptr);
    if (rc) return rc;
    const long long px = (long long)w * ht;
    // pairs per chunk: 64 Mi pixels (8 4K pairs), at most 32; the handle holds two chunks (ping-pong halves)
    const int B = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(32, n_pairs), (64LL << 20) / px));
    CKH(hsflow_set_tuning(h, 0, 0, B));
    rc = hsflow_configure(h, w, ht, 2 * B);
    hsflow_set_tuning(h, 0, 0, 0);
    if (rc) return fail(rc, "%s", hsflow_last_error());
    uint8_t *f1 = nullptr, *f2 = nullptr;
    size_t rp = 0, pp = 0;
    CKH(hsflow_map_frames(h, HSFLOW_FRAMES_BGR8, &f1, &f2, &rp, &pp));
    cudaStream_t cs = nullptr;                     // compute stream we can order against the decoder's stream
    CKC(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    cudaEvent_t ev = nullptr;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { cudaStreamDestroy(cs); return fail(HSFLOW_ECUDA, "cudaEventCreate"); }
    hsflow_set_stream(h, cs);
    const int gw = sample_step ? (w + sample_step - 1) / sample_step : w, gh = sample_step ? (ht + sample_step - 1) / sample_step : ht;
    const size_t opx = (size_t)gw * gh;
    double decode_ms = 0.0;
    long long decoded = 0;
    const int n_chunks = (n_pairs + B - 1) / B;
    int status = HSFLOW_OK;

    // decode chunk c into half (c & 1) of the pair slots (host-side Huffman stage + GPU stage on the decoder's stream)
    auto decode = [&](int c) -> int {
        const int half = c & 1, p0 = c * B, n = std::min(B, n_pairs - p0), s0 = half * B;
        std::vector<const uint8_t*> src; std::vector<size_t> len; std::vector<uint8_t*> dst;
        if (!sequence) {
            for (int j = 0; j < n; ++j) {
                src.push_back(jpegs[2 * (p0 + j)]); len.push_back(sizes[2 * (p0 + j)]); dst.push_back(f1 + (size_t)(s0 + j) * pp);
                src.push_back(jpegs[2 * (p0 + j) + 1]); len.push_back(sizes[2 * (p0 + j) + 1]); dst.push_back(f2 + (size_t)(s0 + j) * pp);
            }
        } else {
            // frame p0 + j is the first frame of pair j and the second frame of pair j - 1: decoded once (into the
            // first-frame slot, the very last one into the second-frame slot), copied device-to-device to its other place
            for (int j = (c == 0 ? 0 : 1); j < n; ++j) { src.push_back(jpegs[p0 + j]); len.push_back(sizes[p0 + j]); dst.push_back(f1 + (size_t)(s0 + j) * pp); }
            src.push_back(jpegs[p0 + n]); len.push_back(sizes[p0 + n]); dst.push_back(f2 + (size_t)(s0 + n - 1) * pp);
        }
        const auto t0 = std::chrono::steady_clock::now();
        int r = decode_batch(src.data(), len.data(), dst.data(), (int)src.size(), rp, w, ht);
        decode_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (r) return r;
        decoded += (long long)src.size();
        if (sequence) {
            if (c > 0) {                           // first frame of this chunk = last frame of the previous chunk (other half)
                const int ph = (c - 1) & 1;
                CKC(cudaMemcpyAsync(f1 + (size_t)s0 * pp, f2 + (size_t)(ph * B + B - 1) * pp, pp, cudaMemcpyDeviceToDevice, D.stream));
            }
            for (int j = 0; j + 1 < n; ++j)
                CKC(cudaMemcpyAsync(f2 + (size_t)(s0 + j) * pp, f1 + (size_t)(s0 + j + 1) * pp, pp, cudaMemcpyDeviceToDevice, D.stream));
        }
        CKC(cudaEventRecord(ev, D.stream));
        return HSFLOW_OK;
    };
    // queue the compute of chunk c behind its decode (asynchronous)
    auto launch = [&](int c) -> int {
        const int p0 = c * B, n = std::min(B, n_pairs - p0);
        CKC(cudaStreamWaitEvent(cs, ev, 0));
        CKH(hsflow_compute_range(h, (c & 1) * B, n));
        return HSFLOW_OK;
    };

    status = decode(0);
    if (status == HSFLOW_OK) status = launch(0);
    for (int c = 0; c < n_chunks && status == HSFLOW_OK; ++c) {
        // while the GPU computes chunk c, the host decodes chunk c + 1 into the other half; the read-back of chunk c
        // goes first on the compute stream, then the compute of chunk c + 1
        if (c + 1 < n_chunks) status = decode(c + 1);
        if (status) break;
        const int half = c & 1, p0 = c * B, n = std::min(B, n_pairs - p0);
        for (int j = 0; j < n && status == HSFLOW_OK; ++j) {
            float* uo = u_out + (size_t)(p0 + j) * opx; float* vo = v_out + (size_t)(p0 + j) * opx;
            status = sample_step ? hsflow_sample_uv(h, half * B + j, sample_step, uo, vo) : hsflow_read_uv(h, half * B + j, uo, vo, 0);
            if (status) fail(status, "%s", hsflow_last_error());
        }
        if (status == HSFLOW_OK && c + 1 < n_chunks) status = launch(c + 1);
    }
    cudaStreamSynchronize(D.stream);
    cudaStreamSynchronize(cs);
    hsflow_set_stream(h, nullptr);
    cudaEventDestroy(ev);
    cudaStreamDestroy(cs);
    if (stats) { stats[0] = decode_ms; stats[1] = (double)decoded; stats[2] = (double)B; stats[3] = (double)D.backend + 0.01 * (D.last_path ? std::min(pool_threads(), 99) : 1); }
    return status;
}

}  // extern "C"
