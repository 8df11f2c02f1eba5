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
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
