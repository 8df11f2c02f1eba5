// hs_image.cpp -- see hs_image.h.  nvJPEG (CUDA 12.9) decodes to interleaved BGR directly in
// device memory and encodes from it; no CPU JPEG codec is linked.
#include "hs_image.h"

#include <cuda_runtime.h>
#include <nvjpeg.h>

#include <cctype>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static thread_local char g_err[256] = "";
static int fail(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
    return -1;
}
static bool has_ext(const char* path, const char* ext) {
    size_t n = strlen(path), m = strlen(ext);
    if (n < m) return false;
    for (size_t i = 0; i < m; ++i)
        if (tolower((unsigned char)path[n - m + i]) != ext[i]) return false;
    return true;
}
static bool read_file(const char* path, std::vector<unsigned char>& buf) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    if (n < 0) { fclose(f); return false; }
    buf.resize((size_t)n);
    bool ok = n == 0 || fread(buf.data(), 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    return ok;
}

// ---- PNM -----------------------------------------------------------------------------------
static int pnm_token(const std::vector<unsigned char>& b, size_t& pos) {
    for (;;) {
        while (pos < b.size() && isspace(b[pos])) ++pos;
        if (pos < b.size() && b[pos] == '#') { while (pos < b.size() && b[pos] != '\n') ++pos; continue; }
        break;
    }
    int v = -1;
    while (pos < b.size() && isdigit(b[pos])) { v = (v < 0 ? 0 : v) * 10 + (b[pos] - '0'); ++pos; }
    return v;
}
static int read_pnm(const std::vector<unsigned char>& b, int* w, int* h, int* ch, uint8_t** data) {
    if (b.size() < 3 || b[0] != 'P' || (b[1] != '5' && b[1] != '6')) return fail("not a binary PGM/PPM file");
    const int c = b[1] == '5' ? 1 : 3;
    size_t pos = 2;
    int W = pnm_token(b, pos), H = pnm_token(b, pos), M = pnm_token(b, pos);
    if (W <= 0 || H <= 0 || M != 255) return fail("unsupported PNM header (%d x %d, max %d)", W, H, M);
    ++pos;   // single whitespace after maxval
    const size_t n = (size_t)W * H * c;
    if (b.size() < pos + n) return fail("truncated PNM data");
    uint8_t* out = (uint8_t*)malloc(n);
    if (!out) return fail("out of memory");
    if (c == 1) memcpy(out, b.data() + pos, n);
    else for (size_t k = 0; k < (size_t)W * H; ++k) {   // RGB on disk -> BGR in memory (cvLoadImage order)
        out[3 * k] = b[pos + 3 * k + 2]; out[3 * k + 1] = b[pos + 3 * k + 1]; out[3 * k + 2] = b[pos + 3 * k];
    }
    *w = W; *h = H; *ch = c; *data = out;
    return 0;
}
static int write_pnm(const char* path, const uint8_t* d, int w, int h, int c) {
    FILE* f = fopen(path, "wb");
    if (!f) return fail("cannot open %s for writing", path);
    fprintf(f, "P%d\n%d %d\n255\n", c == 1 ? 5 : 6, w, h);
    if (c == 1) fwrite(d, 1, (size_t)w * h, f);
    else {
        std::vector<uint8_t> row((size_t)w * 3);
        for (int y = 0; y < h; ++y) {
            const uint8_t* s = d + (size_t)y * w * 3;
            for (int x = 0; x < w; ++x) { row[3 * x] = s[3 * x + 2]; row[3 * x + 1] = s[3 * x + 1]; row[3 * x + 2] = s[3 * x]; }
            fwrite(row.data(), 1, row.size(), f);
        }
    }
    fclose(f);
    return 0;
}

// ---- JPEG through nvJPEG -------------------------------------------------------------------------
struct NvJpeg {
    nvjpegHandle_t h = nullptr;
    nvjpegJpegState_t st = nullptr;
    bool ok = false;
    NvJpeg() {
        if (nvjpegCreateSimple(&h) != NVJPEG_STATUS_SUCCESS) return;
        if (nvjpegJpegStateCreate(h, &st) != NVJPEG_STATUS_SUCCESS) return;
        ok = true;
    }
};
static NvJpeg& nvj() { static NvJpeg j; return j; }

static int read_jpeg(const std::vector<unsigned char>& b, int* w, int* h, int* ch, uint8_t** data) {
    NvJpeg& J = nvj();
    if (!J.ok) return fail("nvJPEG initialisation failed (no CUDA device?)");
    int ncomp = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
    nvjpegChromaSubsampling_t ss;
    if (nvjpegGetImageInfo(J.h, b.data(), b.size(), &ncomp, &ss, ws, hs) != NVJPEG_STATUS_SUCCESS)
        return fail("not a decodable JPEG stream");
    const int W = ws[0], H = hs[0];
    const size_t n = (size_t)W * H * 3;
    unsigned char* dbuf = nullptr;
    if (cudaMalloc(&dbuf, n) != cudaSuccess) return fail("cudaMalloc failed");
    nvjpegImage_t img;
    memset(&img, 0, sizeof img);
    img.channel[0] = dbuf; img.pitch[0] = (size_t)W * 3;
    nvjpegStatus_t s = nvjpegDecode(J.h, J.st, b.data(), b.size(), NVJPEG_OUTPUT_BGRI, &img, 0);
    if (s != NVJPEG_STATUS_SUCCESS) { cudaFree(dbuf); return fail("nvjpegDecode failed (%d)", (int)s); }
    uint8_t* out = (uint8_t*)malloc(n);
    if (!out) { cudaFree(dbuf); return fail("out of memory"); }
    cudaError_t e = cudaMemcpy(out, dbuf, n, cudaMemcpyDeviceToHost);
    cudaFree(dbuf);
    if (e != cudaSuccess) { free(out); return fail("cudaMemcpy: %s", cudaGetErrorString(e)); }
    *w = W; *h = H; *ch = 3; *data = out;   // cvLoadImage(path, 1) always returns 3-channel BGR (cpp:721)
    return 0;
}

static int write_jpeg(const char* path, const uint8_t* d, int w, int h, int c) {
    NvJpeg& J = nvj();
    if (!J.ok) return fail("nvJPEG initialisation failed (no CUDA device?)");
    std::vector<uint8_t> bgr;
    if (c == 1) { bgr.resize((size_t)w * h * 3); for (size_t k = 0; k < (size_t)w * h; ++k) bgr[3 * k] = bgr[3 * k + 1] = bgr[3 * k + 2] = d[k]; d = bgr.data(); }
    nvjpegEncoderState_t es = nullptr;
    nvjpegEncoderParams_t ep = nullptr;
    unsigned char* dbuf = nullptr;
    int rc = -1;
    const size_t n = (size_t)w * h * 3;
    do {
        if (nvjpegEncoderStateCreate(J.h, &es, 0) != NVJPEG_STATUS_SUCCESS) { fail("nvjpegEncoderStateCreate"); break; }
        if (nvjpegEncoderParamsCreate(J.h, &ep, 0) != NVJPEG_STATUS_SUCCESS) { fail("nvjpegEncoderParamsCreate"); break; }
        nvjpegEncoderParamsSetQuality(ep, 95, 0);                       // cvSaveImage default quality
        nvjpegEncoderParamsSetSamplingFactors(ep, NVJPEG_CSS_420, 0);   // libjpeg default for colour
        if (cudaMalloc(&dbuf, n) != cudaSuccess || cudaMemcpy(dbuf, d, n, cudaMemcpyHostToDevice) != cudaSuccess) { fail("cuda upload failed"); break; }
        nvjpegImage_t img; memset(&img, 0, sizeof img);
        img.channel[0] = dbuf; img.pitch[0] = (size_t)w * 3;
        if (nvjpegEncodeImage(J.h, es, ep, &img, NVJPEG_INPUT_BGRI, w, h, 0) != NVJPEG_STATUS_SUCCESS) { fail("nvjpegEncodeImage failed"); break; }
        size_t len = 0;
        if (nvjpegEncodeRetrieveBitstream(J.h, es, nullptr, &len, 0) != NVJPEG_STATUS_SUCCESS) { fail("nvjpeg bitstream size"); break; }
        cudaStreamSynchronize(0);
        std::vector<unsigned char> out(len);
        if (nvjpegEncodeRetrieveBitstream(J.h, es, out.data(), &len, 0) != NVJPEG_STATUS_SUCCESS) { fail("nvjpeg bitstream"); break; }
        cudaStreamSynchronize(0);
        FILE* f = fopen(path, "wb");
        if (!f) { fail("cannot open %s for writing", path); break; }
        fwrite(out.data(), 1, len, f); fclose(f);
        rc = 0;
    } while (0);
    if (dbuf) cudaFree(dbuf);
    if (ep) nvjpegEncoderParamsDestroy(ep);
    if (es) nvjpegEncoderStateDestroy(es);
    return rc;
}

extern "C" {
const char* hsimg_last_error(void) { return g_err; }
void hsimg_free(uint8_t* p) { free(p); }
int hsimg_read(const char* path, int* w, int* h, int* ch, uint8_t** data) {
    if (!path || !w || !h || !ch || !data) return fail("null argument");
    std::vector<unsigned char> buf;
    if (!read_file(path, buf)) return fail("cannot read %s", path);
    if (buf.size() >= 2 && buf[0] == 'P' && (buf[1] == '5' || buf[1] == '6')) return read_pnm(buf, w, h, ch, data);
    if (buf.size() >= 2 && buf[0] == 0xFF && buf[1] == 0xD8) return read_jpeg(buf, w, h, ch, data);
    return fail("%s: unsupported image format (JPEG, binary PGM/PPM)", path);
}
int hsimg_write(const char* path, const uint8_t* d, int w, int h, int c) {
    if (!path || !d || w <= 0 || h <= 0 || (c != 1 && c != 3)) return fail("bad argument");
    if (has_ext(path, ".jpg") || has_ext(path, ".jpeg")) return write_jpeg(path, d, w, h, c);
    return write_pnm(path, d, w, h, c);
}
}
