"""opticalflowhs_b200 -- B200-native Horn-Schunck dense optical flow (hot path of miczi/OpticalFlowHS).

The product is libhsflow.so (hand-written CUDA for sm_100a behind the C ABI of include/hsflow.h)
plus the C++ drop-in class of include/HSOpticalFlowOpenCL.hpp.  This Python package is the thin
ctypes binding used by the tests, the bench and the multi-GPU sharding layer.  It never imports
oracle/ and has no CPU fallback: without the built library or without a GPU it raises.
"""
from .hsflow import (  # noqa: F401
    HSFlow, HSFlowError, lib, library_path,
    STENCIL_CL8, STENCIL_CV4, MATH_FAST, MATH_EXACT, DERIV_CL, DERIV_CV, FRAMES_GRAY8, FRAMES_BGR8, PIPE_SEQUENCE,
)
