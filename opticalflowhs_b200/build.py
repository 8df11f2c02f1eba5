"""In-tree build of libhsflow.so (hand-written CUDA for sm_100a + the C ABI).

    python -m opticalflowhs_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with
the gpurun snapshot; nothing is JIT-compiled at run time.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhsflow.so")
STREAM_DEPTHS = range(1, 9)             # hs_stream_inst.cu is compiled once per temporal-block depth T
SOURCES = ["hs_kernels.cu", "hs_stream.cu", "hsflow_capi.cu", "hs_stream_inst.cu"]
HEADERS = ["hs_common.cuh", "hs_launch.h", "hs_stream.cuh", os.path.join("..", "..", "include", "hsflow.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-diag-suppress", "3288", "--fmad=false",           # arithmetic is written with explicit *_rn intrinsics; never contract anything else
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, variant=None, extra=()):
    """variant/extra: A/B kernel experiments -- libhsflow_<variant>.so built with additional nvcc flags
    (select it at run time with HSFLOW_LIBRARY=...); the product is the plain libhsflow.so."""
    if variant:
        return _build(os.path.join(HERE, f"libhsflow_{variant}.so"), os.path.join(HERE, "build", variant), list(extra), verbose)
    if not force and not needs_build():
        return LIB
    return _build(LIB, os.path.join(HERE, "build"), [], verbose)


def _build(lib, objdir, extra, verbose):
    objs = []
    procs = []
    os.makedirs(objdir, exist_ok=True)
    jobs = [(src, [], src.replace(".cu", ".o")) for src in SOURCES if src != "hs_stream_inst.cu"]
    jobs += [("hs_stream_inst.cu", [f"-DHS_STREAM_T={t}"], f"hs_stream_t{t}.o") for t in STREAM_DEPTHS]
    for src, defs, objname in jobs:
        obj = os.path.join(objdir, objname)
        cmd = [_nvcc()] + NVCC_FLAGS + extra + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src + " " + " ".join(defs), subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [_nvcc(), "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    # python -m opticalflowhs_b200.build [--force] [-v] [--variant NAME -- extra nvcc flags ...]
    args = sys.argv[1:]
    extra = args[args.index("--") + 1:] if "--" in args else []
    variant = args[args.index("--variant") + 1] if "--variant" in args else None
    print(build(force="--force" in args, verbose="-v" in args, variant=variant, extra=extra))
