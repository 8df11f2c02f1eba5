// include/HSOpticalFlowOpenCL.hpp -- drop-in declaration of the reference's host class.
//
// Replaces /root/reference/OpticalFlowHS/HSOpticalFlowOpenCL.hpp (hpp:26-264) for the UNCHANGED
// main.cpp (main:104-108, 120-124): same include guard, same class name, same two constructors
// (hpp:130, hpp:171), same lifecycle initialize/setup/run/cleanup (hpp:239-257) and the same
// public helpers (hpp:109-116, 218-233, 263).  Nothing of OpenCL, the AMD SDK sample framework,
// OpenCV or <windows.h> is needed: the class owns one hsflow_t (include/hsflow.h) and the CUDA
// engine behind it.  Return conventions follow SDKCommon.hpp:23-24 (SDK_SUCCESS 0, SDK_FAILURE 1);
// run() returns -1 when an input cannot be loaded (cpp:722-725) and 0 on success (cpp:773).
#ifndef FILTERS_H_
#define FILTERS_H_

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "hsflow.h"

#ifndef SDK_SUCCESS
#define SDK_SUCCESS 0
#define SDK_FAILURE 1
#endif
#define GROUP_SIZE 32   // hpp:18 (unused there as well)

// The reference stages every plane as cl_float4 with the value in lane 0 (cpp:15-22); the
// helpers below keep that signature for source compatibility.
typedef float cl_float;
typedef unsigned int cl_uint;
typedef struct { cl_float s[4]; } cl_float4;

#if defined(__GNUC__)
#define HSFLOW_CLASS __attribute__((visibility("default")))
#else
#define HSFLOW_CLASS
#endif

class HSFLOW_CLASS HSOpticalFlowOpenCL {
    std::string name;
    cl_float4* pixelData;        // staging plane of readInputImage (lane 0 = gray value)
    cl_float4* inputImageData1;
    cl_float4* inputImageData2;
    cl_float alpha;              // flow smoothness coefficient
    hsflow_t* engine;            // replaces cl_context / queue / 9 cl_mem / 3 cl_kernel (hpp:46-71)
    cl_uint width, height;
    size_t blockSizeX, blockSizeY;   // accepted (work-group hint of the reference), not needed on CUDA
    char *src, *input1, *input2, *output;
    int iterations;
    bool useGpu;                 // dType: "CPU" is accepted and ignored -- there is no CPU fallback
    std::vector<unsigned char> gray;         // current gray8 frame (cvCvtColor result, cpp:727-728)
    std::vector<float> uHost, vHost;         // scalar flow read back by run()
    double totalTime;

    int loadGray(const char* path, std::vector<unsigned char>& out, int& w, int& h);
    int drawAndSave(const char* path);
    int runFrameSequence();

public:
    int readInputImage(cl_float4** inputImageData);   // hpp:109
    int readInputFrame(cl_float4** inputImageData);   // hpp:116

    HSOpticalFlowOpenCL(const char* name, char* src, char* input1, char* input2, char* output,
                        float alp, int it, int gs, char* dType);                       // hpp:130
    HSOpticalFlowOpenCL(const char* name, char* src, float alp, int it, int gs, char* dType);   // hpp:171
    ~HSOpticalFlowOpenCL();

    int setupCL();          // hpp:218  creates the engine (was: context, queue, buffers, program)
    int runDerivatives();   // hpp:221  derivative pass for the loaded pair; u = v = 0
    int runCLKernels();     // hpp:228  ONE Jacobi iteration (the reference calls it `iterations` times)
    void printStats();      // hpp:233  declared but never defined in the reference
    int initialize();       // hpp:239
    int setup();            // hpp:245
    int run();              // hpp:251
    int cleanup();          // hpp:257
    int verifyResults();    // hpp:263

    // extensions (not in the reference): results and mode for programmatic callers
    const float* flowU() const { return uHost.data(); }
    const float* flowV() const { return vHost.data(); }
    int flowWidth() const { return (int)width; }
    int flowHeight() const { return (int)height; }
    double lastMilliseconds() const { return totalTime; }
};

#endif  // FILTERS_H_
