// OpticalFlowOpenCV.cpp -- drop-in for /root/reference/OpticalFlowHS/OpticalFlowOpenCV.cpp
// ("cv.cpp" below) on the CUDA engine: cvSmooth(CV_BLUR 3x3) x2 + cvCalcOpticalFlowHS
// (cv.cpp:27-29) = HSFLOW_DERIV_CV + HSFLOW_STENCIL_CV4 with rho = 1/lambda and the same
// termination criterion, CV_TERMCRIT_ITER | CV_TERMCRIT_EPS with eps = 1e-6 (hsflow_set_epsilon).
#include "../../../include/OpticalFlowOpenCV.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../../include/hsflow.h"
#include "../../../include/hsflow_ingest.h"
#include "hs_image.h"

void hsflow_host_draw(std::vector<unsigned char>& img, int w, int h, const float* u, const float* v,
                      float thr, float lineScale);

static hsflow_t* cv_engine(float lambda, int it) {
    hsflow_t* e = NULL;
    if (hsflow_create(0, &e) != HSFLOW_OK) { std::cout << "hsflow: " << hsflow_last_error() << std::endl; return NULL; }
    const char* ex = getenv("HSFLOW_EXACT");                                // bit-exact arithmetic (one sweep per launch), as for the CL class
    hsflow_set_math(e, (ex && atoi(ex)) ? HSFLOW_MATH_EXACT : HSFLOW_MATH_FAST);
    hsflow_set_deriv(e, HSFLOW_DERIV_CV);                                   // cvSmooth x2 + Sobel estimator (cv.cpp:27-29)
    hsflow_set_params(e, 0.f, it, HSFLOW_STENCIL_CV4, 1, 0);
    hsflow_set_lambda(e, lambda);
    hsflow_set_epsilon(e, 1e-6);                                            // cv.cpp:29 cvTermCriteria(ITER | EPS, it, 1e-6)
    return e;
}

int OpticalFlowOpenCV::runFromImg(char* input1, char* input2, char* output, float lambda, int it) {
    hsflow_t* e = cv_engine(lambda, it);
    if (!e) return -1;
    int w = 0, h = 0;
    // cvLoadImage + cvCvtColor (cv.cpp:15-20): decoded on the GPU straight into the engine's frame planes
    if (!input1 || !input2 || hsingest_load_pair_files(e, input1, input2, &w, &h) != HSFLOW_OK) {
        std::cout << "Input image error.\n";
        hsflow_destroy(e);
        return -1;
    }
    std::vector<float> u((size_t)w * h), v((size_t)w * h);
    const auto t0 = std::chrono::steady_clock::now();                       // cv.cpp:26
    int rc = hsflow_compute(e);
    if (rc == HSFLOW_OK) rc = hsflow_read_uv(e, 0, u.data(), v.data(), 0);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (rc != HSFLOW_OK) { std::cout << "hsflow: " << hsflow_last_error() << std::endl; hsflow_destroy(e); return -1; }
    hsflow_destroy(e);
    std::vector<unsigned char> img;
    hsflow_host_draw(img, w, h, u.data(), v.data(), 1.0f, 0.5f);              // cv.cpp:34-46: |v|>1, line length v/2
    if (output && hsimg_write(output, img.data(), w, h, 3) != 0) { std::cout << "Output image error.\n"; return -1; }
    std::cout << "Avg time: " << ms << " [ms]" << std::endl;                 // cv.cpp:49
    return 0;
}

// Camera loop (cv.cpp:56-131).  There is no capture device here: consecutive frames are read from files named by
// HSFLOW_FRAMES (printf pattern starting at 0, as the CL class does), each new frame paired with the previous one
// (cv.cpp:118 cvCopy), pictures optionally written to HSFLOW_FRAMES_OUT.  The reference passes use_previous = 0
// (cv.cpp:94); HSFLOW_USE_PREVIOUS=1 keeps u, v of the previous pair as the starting point (cv.h:481-483).
int OpticalFlowOpenCV::runFromCamera(float lambda, int it) {
    const char* pattern = getenv("HSFLOW_FRAMES");
    if (!pattern) { std::cout << "ERROR: capture is NULL \n"; return -1; }  // cv.cpp:68-73
    hsflow_t* e = cv_engine(lambda, it);
    if (!e) return -1;
    const char* up = getenv("HSFLOW_USE_PREVIOUS");
    const bool warm = up && atoi(up);
    char path[1024];
    snprintf(path, sizeof path, pattern, 0);
    int w = 0, h = 0, count = 0;
    if (hsingest_load_pair_files(e, path, path, &w, &h) != HSFLOW_OK) { std::cout << "ERROR: frame is null...\n"; hsflow_destroy(e); return -1; }
    std::vector<float> u((size_t)w * h), v((size_t)w * h);
    std::vector<unsigned char> img;
    double ms = 0;
    for (int k = 1;; ++k) {
        snprintf(path, sizeof path, pattern, k);
        FILE* probe = fopen(path, "rb");
        if (!probe) break;
        fclose(probe);
        const auto t0 = std::chrono::steady_clock::now();                   // cv.cpp:91
        int rc = hsingest_push_frame_file(e, path);
        if (rc != HSFLOW_OK) { std::cout << "hsflow: " << hsingest_last_error() << std::endl; break; }
        if (warm && k == 2) hsflow_set_warm_start(e, 1);                      // from the second pair on: start from the last field
        rc = hsflow_compute(e);
        if (rc == HSFLOW_OK) rc = hsflow_read_uv(e, 0, u.data(), v.data(), 0);
        ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (rc != HSFLOW_OK) { std::cout << "hsflow: " << hsflow_last_error() << std::endl; hsflow_destroy(e); return -1; }
        const char* outp = getenv("HSFLOW_FRAMES_OUT");
        if (outp) {
            char op[1024];
            snprintf(op, sizeof op, outp, k);
            hsflow_host_draw(img, w, h, u.data(), v.data(), 1.0f, 0.5f);      // cv.cpp:97-112
            hsimg_write(op, img.data(), w, h, 3);
        }
        ++count;
    }
    hsflow_destroy(e);
    std::cout << "Avg time: " << (count ? ms / count : 0.0) << " [ms]" << std::endl;   // cv.cpp:122
    return 0;
}
