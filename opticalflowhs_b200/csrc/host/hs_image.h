// hs_image.h -- image file I/O for the drop-in classes (replaces cvLoadImage / cvSaveImage,
// HSOpticalFlowOpenCL.cpp:721, 732, 772).  JPEG goes through nvJPEG on the GPU, PGM/PPM are
// read and written natively.  Pixels are 8-bit, interleaved BGR (3 channels) or gray (1).
#pragma once
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif
/* Returns 0 on success.  *data is malloc'ed (free with hsimg_free); channels is 1 or 3 (BGR). */
int hsimg_read(const char* path, int* width, int* height, int* channels, uint8_t** data);
int hsimg_write(const char* path, const uint8_t* data, int width, int height, int channels);
void hsimg_free(uint8_t* data);
const char* hsimg_last_error(void);
/* The drawing loop of the reference (HSOpticalFlowOpenCL.cpp:758-770: thr 0.5, line_scale 1; OpticalFlowOpenCV.cpp:32-46:
 * thr 1, line_scale 0.5) into a w x h BGR image: stride-4 grid, filled blue circle of radius 2, red 8-connected line. */
int hsimg_draw_flow(const float* u, const float* v, int width, int height, float thr, float line_scale, unsigned char* bgr_out);
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
