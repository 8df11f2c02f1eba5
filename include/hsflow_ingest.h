/* include/hsflow_ingest.h -- JPEG ingest on the GPU for the Horn-Schunck engine (exported by libhsflow_host.so).
 *
 * Replaces cvLoadImage + cvCvtColor(BGR2GRAY) + readInputImage (HSOpticalFlowOpenCL.cpp:721-740, 6-45) and their
 * OpenCV-class twins (OpticalFlowOpenCV.cpp:15-20): nvJPEG decodes the bitstream on the GPU as interleaved BGR
 * STRAIGHT INTO the engine's frame planes (hsflow_map_frames, include/hsflow.h); the gray conversion happens inside the
 * derivative kernel.  No decoded pixel ever visits host memory.  PGM / PPM files are uploaded as they are (gray / BGR).
 *
 * Plain C; every call returns 0 or a negative HSFLOW_E* code, message in hsingest_last_error().
 */
#ifndef HSFLOW_INGEST_H_
#define HSFLOW_INGEST_H_
#include <stddef.h>
#include <stdint.h>

#include "hsflow.h"

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

const char* hsingest_last_error(void);

/* Width, height and component count of a JPEG bitstream in host memory. */
int hsingest_jpeg_info(const uint8_t* jpeg, size_t len, int* width, int* height, int* components);

/* Decode one JPEG as interleaved BGR (cvLoadImage(path, 1) order, gray streams replicated) into device memory
 * d_bgr[row * pitch + 3 * col], on `cuda_stream` (a cudaStream_t as void*, NULL = the decoder's own stream,
 * synchronised before return).  The image must be width x height. */
int hsingest_decode_to_device(const uint8_t* jpeg, size_t len, uint8_t* d_bgr, size_t pitch, int width, int height,
                              void* cuda_stream);

/* run()'s loading step for one pair (cpp:721-740): hsflow_configure(w, h, 1), then both images go into the handle's
 * frame planes -- JPEG decoded on the GPU into BGR planes, PPM uploaded as BGR, PGM as gray.  width/height may be NULL. */
int hsingest_load_pair_files(hsflow_t* h, const char* path1, const char* path2, int* width, int* height);
int hsingest_load_pair_jpeg(hsflow_t* h, const uint8_t* jpeg1, size_t len1, const uint8_t* jpeg2, size_t len2,
                            int* width, int* height);
/* Camera-loop step (cpp:800-842) for a handle configured with one pair: the second frame becomes the first
 * (hsflow_swap_frames), the image at `path` (or the JPEG in memory) becomes the new second frame. */
int hsingest_push_frame_file(hsflow_t* h, const char* path);

/* Video batches: n_images JPEG bitstreams in host memory, all of the same size.  sequence = 0: images (2k, 2k+1) form
 * pair k; sequence = 1: consecutive frames, pair k = (image k, image k+1), every image decoded once.  Batched nvJPEG
 * decode of the next chunk of pairs (host-side Huffman stage included) overlaps the engine's compute of the current
 * chunk: the handle's pair slots are used as two halves.  u_out / v_out: n_pairs fields of W*H floats, or with
 * sample_step > 0 the stride-`step` samples (ceil(H/step) x ceil(W/step) floats per pair, see hsflow_sample_uv).
 * Uses the handle's parameters (hsflow_set_params ...); reconfigures it and resets hsflow_set_tuning to automatic.  stats (may be NULL) receives
 * {decode_ms_total, images decoded, pairs per chunk, nvjpegBackend_t used + host threads / 100}. */
int hsingest_run_jpeg_batch(hsflow_t* h, const uint8_t* const* jpegs, const size_t* sizes, int n_images, int sequence,
                            int sample_step, float* u_out, float* v_out, double stats[4]);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* HSFLOW_INGEST_H_ */
