"""Builds the C++ host side above the C ABI:
  opticalflowhs_b200/libhsflow_host.so : the drop-in classes HSOpticalFlowOpenCL / OpticalFlowOpenCV
                                         (include/*.hpp) + image I/O (nvJPEG, PGM/PPM)
  opticalflowhs_b200/bin/OpticalFlowHS : the reference's UNCHANGED main.cpp compiled against our
                                         headers -- only where /root/reference exists (this
                                         container); the source is read in place, never copied.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HOST = os.path.join(HERE, "csrc", "host")
LIB = os.path.join(HERE, "libhsflow_host.so")
EXE = os.path.join(HERE, "bin", "OpticalFlowHS")
REF_MAIN = "/root/reference/OpticalFlowHS/main.cpp"
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
SRCS = ["HSOpticalFlowOpenCL.cpp", "OpticalFlowOpenCV.cpp", "hs_image.cpp", "hs_ingest.cpp"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False):
    from . import build as core
    core.build()
    deps = [os.path.join(HOST, s) for s in SRCS + ["hs_image.h"]] + [
        os.path.join(ROOT, "include", f) for f in ("hsflow.h", "hsflow_ingest.h", "HSOpticalFlowOpenCL.hpp", "OpticalFlowOpenCV.hpp")]
    if force or _stale(LIB, deps):
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall",
               f"-I{CUDA}/include", "-o", LIB] + [os.path.join(HOST, s) for s in SRCS] + [
               f"-L{HERE}", "-lhsflow", f"-L{CUDA}/lib64", "-lnvjpeg", "-lcudart",
               "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{CUDA}/lib64"]
        subprocess.check_call(cmd)
    if os.path.exists(REF_MAIN) and (force or _stale(EXE, [LIB, REF_MAIN] + deps)):
        os.makedirs(os.path.dirname(EXE), exist_ok=True)
        # main.cpp is compiled from stdin so that `#include "HSOpticalFlowOpenCL.hpp"` resolves to
        # OUR include directory instead of the reference's own header next to it.
        with open(REF_MAIN, "rb") as f:
            subprocess.run(["g++", "-x", "c++", "-std=c++17", "-O2", "-Wno-write-strings", "-iquote",
                            os.path.join(ROOT, "include"), "-o", EXE, "-", f"-L{HERE}", "-lhsflow_host", "-lhsflow",
                            "-Wl,-rpath,$ORIGIN/..", f"-Wl,-rpath,{CUDA}/lib64"], stdin=f, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
