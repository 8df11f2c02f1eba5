// HSOpticalFlowOpenCL.cpp -- the drop-in host class on top of the C ABI (include/hsflow.h).
//
// Mirrors /root/reference/OpticalFlowHS/HSOpticalFlowOpenCL.cpp method by method; citations
// "cpp:" refer to that file.  What changed underneath: no OpenCL runtime, frames go to the GPU
// once as 8-bit pixels, grayscale conversion + derivatives are one fused kernel, the
// `iterations` Jacobi sweeps run temporally blocked without any per-iteration host round trip
// (the reference moves u and v over PCIe twice per iteration, cpp:483-501, 655-675).
#include "../../../include/HSOpticalFlowOpenCL.hpp"

#include <algorithm>
#include <chrono>

#include "../../../include/hsflow_ingest.h"
#include "hs_image.h"

namespace {
const unsigned char kCircleBGR[3] = {255, 0, 0};   // CV_RGB(0,0,255), cpp:3
const unsigned char kLineBGR[3] = {0, 0, 255};     // CV_RGB(255,0,0), cpp:4

inline void put(std::vector<unsigned char>& img, int w, int h, int x, int y, const unsigned char* c) {
    if (x < 0 || y < 0 || x >= w || y >= h) return;
    unsigned char* p = &img[((size_t)y * w + x) * 3];
    p[0] = c[0]; p[1] = c[1]; p[2] = c[2];
}
// cvCircle(img, centre, 2, colour, -1): the filled circle of radius 2 as OpenCV rasterises it -- 13 pixels, rows of
// 1, 3, 5, 3, 1 (|dx| + |dy| <= 2), not the 21-pixel Euclidean disc.
void filledCircle2(std::vector<unsigned char>& img, int w, int h, int cx, int cy, const unsigned char* c) {
    for (int dy = -2; dy <= 2; ++dy)
        for (int dx = -2; dx <= 2; ++dx)
            if (std::abs(dx) + std::abs(dy) <= 2) put(img, w, h, cx + dx, cy + dy, c);
}

// cv::clipLine: Cohen-Sutherland against [0, w-1] x [0, h-1], intersection points computed in double and truncated
bool clipLine(int w, int h, long long& x1, long long& y1, long long& x2, long long& y2) {
    const long long right = w - 1, bottom = h - 1;
    if (w <= 0 || h <= 0) return false;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (long long)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (long long)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (long long)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (long long)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// cvLine(img, p1, p2, colour, 1, 8): OpenCV's LineIterator -- the line is clipped to the image first, always walked
// from its left end point to its right one, one pixel per step along the major axis, the minor axis steps when the
// integer error term (starting at dx - 2 dy) is negative.
void line8(std::vector<unsigned char>& img, int w, int h, int ax, int ay, int bx, int by, const unsigned char* c) {
    long long x1 = ax, y1 = ay, x2 = bx, y2 = by;
    if ((unsigned long long)x1 >= (unsigned long long)w || (unsigned long long)x2 >= (unsigned long long)w ||
        (unsigned long long)y1 >= (unsigned long long)h || (unsigned long long)y2 >= (unsigned long long)h) {
        if (!clipLine(w, h, x1, y1, x2, y2)) return;
    }
    long long dx = x2 - x1, dy = y2 - y1;
    if (dx < 0) { dx = -dx; dy = -dy; x1 = x2; y1 = y2; }     // left to right
    int sx = 1, sy = 1;
    if (dy < 0) { dy = -dy; sy = -1; }
    const bool steep = dy > dx;                                 // the major axis is y
    const long long major = steep ? dy : dx, minor = steep ? dx : dy;
    long long err = major - (minor + minor);
    const long long plusDelta = major + major, minusDelta = -(minor + minor);
    long long x = x1, y = y1;
    for (long long k = 0; k <= major; ++k) {
        put(img, w, h, (int)x, (int)y, c);
        const bool neg = err < 0;
        err += minusDelta + (neg ? plusDelta : 0);
        if (steep) { y += sy; if (neg) x += sx; }
        else { x += sx; if (neg) y += sy; }
    }
}
}  // namespace

// Shared with OpticalFlowOpenCV.cpp: the drawing loop of cpp:758-770 / cv.cpp:32-46.
void hsflow_host_draw(std::vector<unsigned char>& img, int w, int h, const float* u, const float* v,
                      float thr, float lineScale) {
    img.assign((size_t)w * h * 3, 0);             // cvZero(imgFlow)
    const int step = 4;
    for (int i = 0; i < h; i += step)
        for (int j = 0; j < w; j += step) {
            const float a = u[(size_t)i * w + j], b = v[(size_t)i * w + j];
            if (a > thr || b > thr || a < -thr || b < -thr) {
                filledCircle2(img, w, h, j, i, kCircleBGR);
                // cvPoint(x + u, y + v): int + float evaluated by the reference's x87 build in extended precision, then
                // truncated; double keeps that (float would round a fraction just below 1 up to the next integer)
                const double ex = std::min(std::max((double)j + (double)a * (double)lineScale, -1e9), 1e9);
                const double ey = std::min(std::max((double)i + (double)b * (double)lineScale, -1e9), 1e9);
                line8(img, w, h, j, i, (int)ex, (int)ey, kLineBGR);
            }
        }
}

// C entry for tests and other bindings: bgr_out receives w * h * 3 bytes
extern "C" __attribute__((visibility("default"))) int hsimg_draw_flow(const float* u, const float* v, int w, int h, float thr,
                                                                      float line_scale, unsigned char* bgr_out) {
    if (!u || !v || !bgr_out || w <= 0 || h <= 0) return -1;
    std::vector<unsigned char> img;
    hsflow_host_draw(img, w, h, u, v, thr, line_scale);
    memcpy(bgr_out, img.data(), img.size());
    return 0;
}

HSOpticalFlowOpenCL::HSOpticalFlowOpenCL(const char* nm, char* src_, char* in1, char* in2, char* out,
                                         float alp, int it, int gs, char* dType)
    : name(nm ? nm : ""), pixelData(NULL), inputImageData1(NULL), inputImageData2(NULL),
      alpha(alp), engine(NULL), width(0), height(0), blockSizeX((size_t)gs), blockSizeY(1),
      src(src_), input1(in1), input2(in2), output(out), iterations(it),
      useGpu(!(dType && strcmp(dType, "CPU") == 0)), totalTime(0.0) {}

HSOpticalFlowOpenCL::HSOpticalFlowOpenCL(const char* nm, char* src_, float alp, int it, int gs, char* dType)
    : name(nm ? nm : ""), pixelData(NULL), inputImageData1(NULL), inputImageData2(NULL),
      alpha(alp), engine(NULL), width(0), height(0), blockSizeX((size_t)gs), blockSizeY(1),
      src(src_), input1(NULL), input2(NULL), output(NULL), iterations(it),
      useGpu(!(dType && strcmp(dType, "CPU") == 0)), totalTime(0.0) {}

HSOpticalFlowOpenCL::~HSOpticalFlowOpenCL() {}

int HSOpticalFlowOpenCL::initialize() { return SDK_SUCCESS; }   // cpp:681-704: only registered a dead "-i" option
int HSOpticalFlowOpenCL::setup() { return SDK_SUCCESS; }        // cpp:895
int HSOpticalFlowOpenCL::verifyResults() { return SDK_SUCCESS; }   // cpp:894 (stub in the reference too)
void HSOpticalFlowOpenCL::printStats() {
    std::cout << name << ": " << width << "x" << height << ", alpha " << alpha << ", " << iterations
              << " iterations, " << totalTime << " ms" << std::endl;
}

// cvLoadImage + cvCvtColor(BGR2GRAY) (cpp:721-728): the image is decoded to BGR (or gray) bytes.
int HSOpticalFlowOpenCL::loadGray(const char* path, std::vector<unsigned char>& out, int& w, int& h) {
    int ch = 0;
    uint8_t* data = NULL;
    if (hsimg_read(path, &w, &h, &ch, &data) != 0) return -1;
    out.resize((size_t)w * h);
    if (ch == 1) memcpy(out.data(), data, out.size());
    else for (size_t k = 0; k < out.size(); ++k)      // OpenCV 2.1 fixed-point BGR2GRAY
        out[k] = (unsigned char)((data[3 * k] * 1868 + data[3 * k + 1] * 9617 + data[3 * k + 2] * 4899 + 8192) >> 14);
    hsimg_free(data);
    return 0;
}

// cpp:6-45: stage the current gray frame as float4 (lane 0 = value) into a fresh plane.
int HSOpticalFlowOpenCL::readInputImage(cl_float4** inputImageData) {
    // run() no longer keeps a host copy of the gray frame (it is decoded straight into HBM): callers of this legacy
    // helper get it loaded here, from the second input (the frame run() stages last, cpp:732-740)
    if (gray.empty() && input2) { int w = 0, h = 0; if (loadGray(input2, gray, w, h) == 0) { width = (cl_uint)w; height = (cl_uint)h; } }
    if (!inputImageData || gray.empty()) return SDK_FAILURE;
    const size_t n = (size_t)width * height;
    free(pixelData);
    pixelData = (cl_float4*)calloc(n, sizeof(cl_float4));
    *inputImageData = (cl_float4*)malloc(n * sizeof(cl_float4));
    if (!pixelData || !*inputImageData) return SDK_FAILURE;
    for (size_t k = 0; k < n; ++k) pixelData[k].s[0] = (cl_float)gray[k];
    memcpy(*inputImageData, pixelData, n * sizeof(cl_float4));
    return SDK_SUCCESS;
}
// cpp:47-64: same into an existing plane.
int HSOpticalFlowOpenCL::readInputFrame(cl_float4** inputImageData) {
    if (!inputImageData || !*inputImageData || !pixelData || gray.empty()) return SDK_FAILURE;
    const size_t n = (size_t)width * height;
    memset(pixelData, 0, n * sizeof(cl_float4));
    for (size_t k = 0; k < n; ++k) pixelData[k].s[0] = (cl_float)gray[k];
    memcpy(*inputImageData, pixelData, n * sizeof(cl_float4));
    return 0;
}

// cpp:67-319: context, queue, 9 buffers, program build -> one engine handle.
int HSOpticalFlowOpenCL::setupCL() {
    if (engine) return SDK_SUCCESS;
    if (hsflow_create(0, &engine) != HSFLOW_OK) {
        std::cout << "hsflow: " << hsflow_last_error() << std::endl;
        engine = NULL;
        return SDK_FAILURE;
    }
    // Defaults reproduce the shipped kernel literally: u_v_updateKernel never writes v
    // (Kernels.cl:87-89).  HSFLOW_UPDATE_V=1 selects the full Horn-Schunck update.
    const char* uv = getenv("HSFLOW_UPDATE_V");
    const char* ex = getenv("HSFLOW_EXACT");
    hsflow_set_math(engine, (ex && atoi(ex)) ? HSFLOW_MATH_EXACT : HSFLOW_MATH_FAST);
    if (hsflow_set_params(engine, alpha, iterations, HSFLOW_STENCIL_CL8, (uv && atoi(uv)) ? 1 : 0, 0) != HSFLOW_OK) {
        std::cout << "hsflow: " << hsflow_last_error() << std::endl;
        return SDK_FAILURE;
    }
    return SDK_SUCCESS;
}

// cpp:321-474: zero u/v, upload both frames, derivative kernel.
int HSOpticalFlowOpenCL::runDerivatives() {
    if (!engine || !inputImageData1 || !inputImageData2) return SDK_FAILURE;
    const size_t n = (size_t)width * height;
    std::vector<float> a(n), b(n);
    for (size_t k = 0; k < n; ++k) { a[k] = inputImageData1[k].s[0]; b[k] = inputImageData2[k].s[0]; }
    if (hsflow_load_pair_f32(engine, a.data(), b.data(), (int)width, (int)height, 0) != HSFLOW_OK) return SDK_FAILURE;
    return hsflow_prepare(engine) == HSFLOW_OK ? SDK_SUCCESS : SDK_FAILURE;
}
// cpp:476-679: one iteration (no PCIe traffic here; results stay on the device).
int HSOpticalFlowOpenCL::runCLKernels() {
    if (!engine) return SDK_FAILURE;
    return hsflow_iterate(engine, 1) == HSFLOW_OK ? SDK_SUCCESS : SDK_FAILURE;
}

int HSOpticalFlowOpenCL::drawAndSave(const char* path) {
    std::vector<unsigned char> img;
    hsflow_host_draw(img, (int)width, (int)height, uHost.data(), vHost.data(), 0.5f, 1.0f);   // cpp:762-770
    if (hsimg_write(path, img.data(), (int)width, (int)height, 3) != 0) {
        std::cout << "Output image error: " << hsimg_last_error() << std::endl;
        return -1;
    }
    return 0;
}

// cpp:706-847
int HSOpticalFlowOpenCL::run() {
    if (!src) return SDK_FAILURE;
    if (strcmp(src, "-hd") != 0) return runFrameSequence();
    if (!input1 || !input2) { std::cout << "Input image error.\n"; return -1; }
    std::cout << "przed setupCL\n";
    if (setupCL() != SDK_SUCCESS) return SDK_FAILURE;
    std::cout << "po setupCL\n";
    // cvLoadImage + cvCvtColor + readInputImage (cpp:721-740): JPEGs are decoded on the GPU straight into the engine's
    // BGR frame planes, PGM/PPM are uploaded as they are; the gray conversion runs inside the derivative kernel
    int w1 = 0, h1 = 0;
    if (hsingest_load_pair_files(engine, input1, input2, &w1, &h1) != HSFLOW_OK) {
        std::cout << "Input image error.\n";                                             // cpp:722-725
        return -1;
    }
    width = (cl_uint)w1; height = (cl_uint)h1;
    gray.clear();
    const size_t n = (size_t)width * height;
    uHost.assign(n, 0.f); vHost.assign(n, 0.f);
    // timed region of cpp:748-752: derivatives, all iterations, read-back
    const auto t0 = std::chrono::steady_clock::now();
    int rc = hsflow_compute(engine);
    if (rc == HSFLOW_OK) rc = hsflow_read_uv(engine, 0, uHost.data(), vHost.data(), 0);
    totalTime = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (rc != HSFLOW_OK) { std::cout << "hsflow: " << hsflow_last_error() << std::endl; return SDK_FAILURE; }
    std::cout << "Avg time: " << totalTime << " [ms]" << std::endl;   // cpp:755
    if (output && drawAndSave(output) != 0) return -1;
    return 0;
}

// Camera branch (cpp:775-845).  There is no capture device here: consecutive frames are read
// from files named by HSFLOW_FRAMES (printf pattern, e.g. "frames/%04d.pgm", starting at 0),
// each new frame paired with the previous one (cpp:834), until a file is missing.
int HSOpticalFlowOpenCL::runFrameSequence() {
    const char* pattern = getenv("HSFLOW_FRAMES");
    if (!pattern) { fprintf(stderr, "ERROR: capture is NULL \n"); return -1; }   // cpp:779-783
    char path[1024];
    int w = 0, h = 0, count = 0;
    snprintf(path, sizeof path, pattern, 0);
    if (setupCL() != SDK_SUCCESS) return SDK_FAILURE;
    // the first frame fills both planes (cpp:800-806 grabs one frame before the loop)
    if (hsingest_load_pair_files(engine, path, path, &w, &h) != HSFLOW_OK) { fprintf(stderr, "ERROR: frame is null...\n"); return -1; }
    width = (cl_uint)w; height = (cl_uint)h;
    uHost.assign((size_t)w * h, 0.f); vHost.assign((size_t)w * h, 0.f);
    double ms = 0;
    for (int k = 1;; ++k) {
        snprintf(path, sizeof path, pattern, k);
        FILE* probe = fopen(path, "rb");
        if (!probe) break;
        fclose(probe);
        const auto t0 = std::chrono::steady_clock::now();
        // cpp:834: the previous frame is already in HBM -- it becomes the first frame by a pointer swap, only the new
        // frame crosses PCIe (as a JPEG bitstream or as raw pixels); u, v are re-zeroed per pair (cpp:331-332)
        int rc = hsingest_push_frame_file(engine, path);
        if (rc != HSFLOW_OK) { std::cout << "hsflow: " << hsingest_last_error() << std::endl; break; }
        rc = hsflow_compute(engine);
        if (rc == HSFLOW_OK) rc = hsflow_read_uv(engine, 0, uHost.data(), vHost.data(), 0);
        ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (rc != HSFLOW_OK) { std::cout << "hsflow: " << hsflow_last_error() << std::endl; return SDK_FAILURE; }
        const char* outp = getenv("HSFLOW_FRAMES_OUT");
        if (outp) { char op[1024]; snprintf(op, sizeof op, outp, k); drawAndSave(op); }
        ++count;
    }
    gray.clear();
    totalTime = count ? ms / count : 0.0;
    std::cout << "Avg time: " << totalTime << " [ms]" << std::endl;                  // cpp:838
    return SDK_SUCCESS;
}

// cpp:849-892
int HSOpticalFlowOpenCL::cleanup() {
    if (engine) { hsflow_destroy(engine); engine = NULL; }
    free(pixelData); free(inputImageData1); free(inputImageData2);
    pixelData = inputImageData1 = inputImageData2 = NULL;
    return SDK_SUCCESS;
}
