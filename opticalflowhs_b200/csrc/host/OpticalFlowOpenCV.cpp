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
#include "hs_image.h"

void hsflow_host_draw(std::vector<unsigned char>& img, int w, int h, const float* u, const float* v,
                      float thr, float lineScale);

static int load_gray(const char* path, std::vector<unsigned char>& out, int& w, int& h) {
    int ch = 0;
    uint8_t* data = NULL;
    if (!path || hsimg_read(path, &w, &h, &ch, &data) != 0) return -1;
    out.resize((size_t)w * h);
    if (ch == 1) memcpy(out.data(), data, out.size());
    else for (size_t k = 0; k < out.size(); ++k)      // cvCvtColor(CV_BGR2GRAY), cv.cpp:17, 20
        out[k] = (unsigned char)((data[3 * k] * 1868 + data[3 * k + 1] * 9617 + data[3 * k + 2] * 4899 + 8192) >> 14);
    hsimg_free(data);
    return 0;
}

int OpticalFlowOpenCV::runFromImg(char* input1, char* input2, char* output, float lambda, int it) {
    std::vector<unsigned char> a, b;
    int w = 0, h = 0, w2 = 0, h2 = 0;
    if (load_gray(input1, a, w, h) != 0 || load_gray(input2, b, w2, h2) != 0 || w != w2 || h != h2) {
        std::cout << "Input image error.\n";
        return -1;
    }
    hsflow_t* e = NULL;
    if (hsflow_create(0, &e) != HSFLOW_OK) { std::cout << "hsflow: " << hsflow_last_error() << std::endl; return -1; }
    std::vector<float> u((size_t)w * h), v((size_t)w * h);
    hsflow_set_deriv(e, HSFLOW_DERIV_CV);
    hsflow_set_params(e, 0.f, it, HSFLOW_STENCIL_CV4, 1, 0);
    hsflow_set_lambda(e, lambda);
    hsflow_set_epsilon(e, 1e-6);                                            // cv.cpp:29 cvTermCriteria(ITER | EPS, it, 1e-6)
    const auto t0 = std::chrono::steady_clock::now();                       // cv.cpp:26
    int rc = hsflow_load_pair_gray8(e, a.data(), b.data(), w, h, 0);
    if (rc == HSFLOW_OK) rc = hsflow_compute(e);
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

int OpticalFlowOpenCV::runFromCamera(float lambda, int it) {
    (void)lambda; (void)it;
    std::cout << "ERROR: capture is NULL \n";                                // cv.cpp:68-73: no capture device
    return -1;
}
