// include/OpticalFlowOpenCV.hpp -- drop-in declaration of the reference's second class
// (/root/reference/OpticalFlowHS/OpticalFlowOpenCV.hpp:6-11) for the unchanged main.cpp
// (main:136-137, 146-147).  The reference runs OpenCV 2.1 cvSmooth + cvCalcOpticalFlowHS on the
// CPU (OpticalFlowOpenCV.cpp:27-29); here the same algorithm (3x3 box blur, Sobel/8 on frame 1,
// It = frame2 - frame1, 4-neighbour mean, rho = 1/lambda) runs on the CUDA engine in
// HSFLOW_DERIV_CV + HSFLOW_STENCIL_CV4 mode.  No CPU implementation is shipped.
#ifndef OPTICALFLOWOPENCV_HPP_
#define OPTICALFLOWOPENCV_HPP_
#include <cmath>
#include <iostream>

#if defined(__GNUC__) && !defined(HSFLOW_CLASS)
#define HSFLOW_CLASS __attribute__((visibility("default")))
#elif !defined(HSFLOW_CLASS)
#define HSFLOW_CLASS
#endif

class HSFLOW_CLASS OpticalFlowOpenCV {
public:
    int runFromImg(char* input1, char* input2, char* output, float lambda, int it);
    int runFromCamera(float lambda, int it);
};
#endif
