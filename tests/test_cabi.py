"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/hsflow.h declares (and nothing is silently missing from the ctypes binding), fails
loudly without a GPU, and the product never touches oracle/."""
import os
import re
import subprocess

import pytest

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hsflow.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hsflow_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    import opticalflowhs_b200 as P
    from opticalflowhs_b200 import build, hsflow
    build.build()
    L = P.lib()
    names = declared_symbols()
    assert len(names) >= 38
    for n in names:
        assert hasattr(L, n), f"{n} declared in hsflow.h but not exported by libhsflow.so"
        assert n in hsflow.SIGNATURES, f"{n} missing from the ctypes binding"
    out = subprocess.check_output(["nm", "-D", "--defined-only", P.library_path()], text=True)
    exported = set(re.findall(r" T (hsflow_\w+)", out))
    assert set(names) <= exported
    assert L.hsflow_version() >= 100


def test_ingest_header_symbols_are_exported_and_bound(tmp_path):
    """include/hsflow_ingest.h (JPEG ingest on the GPU, libhsflow_host.so): every declared symbol is exported and bound,
    and the header compiles as C99.  Loading needs libnvjpeg + libcudart, not a GPU."""
    from opticalflowhs_b200 import build_host, ingest
    build_host.build()
    src = open(os.path.join(ROOT, "include", "hsflow_ingest.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(hsingest_[a-z0-9_]+)\s*\(", src)))
    assert len(names) >= 7
    L = ingest.lib()
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(ROOT, "opticalflowhs_b200", "libhsflow_host.so")], text=True)
    exported = set(re.findall(r" T (hsingest_\w+)", out))
    for n in names:
        assert n in exported and hasattr(L, n), f"{n} declared in hsflow_ingest.h but not exported by libhsflow_host.so"
        assert n in ingest.SIGNATURES, f"{n} missing from the ctypes binding"
    c = tmp_path / "hi.c"
    c.write_text('#include "%s"\nint main(void) { return hsingest_last_error() != 0 ? 0 : 1; }\n' % os.path.join(ROOT, "include", "hsflow_ingest.h"))
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-c", str(c), "-o", str(tmp_path / "hi.o")])


def test_header_is_plain_c(tmp_path):
    """include/hsflow.h is the C ABI: it has to compile as C99 (cgo / JNI / ctypes-generator consumers), not only as C++."""
    src = tmp_path / "hc.c"
    src.write_text('#include "%s"\nint main(void) { hsflow_t* h = 0; hsflow_strip_handle_t s; (void)s; '
                   'return hsflow_version() > 0 && h == 0 ? 0 : 1; }\n' % os.path.join(ROOT, "include", "hsflow.h"))
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-c", str(src), "-o", str(tmp_path / "hc.o")])


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import opticalflowhs_b200 as P
    assert P.lib().hsflow_device_count() == 0
    with pytest.raises(P.HSFlowError) as e:
        P.HSFlow(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "opticalflowhs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "libclref" not in text and "hs_oracle.h" not in text, f


def test_drop_in_headers_compile_the_reference_main_interface(tmp_path):
    """A caller written like main.cpp (main:104-108, 136-137) compiles against include/ alone."""
    src = tmp_path / "caller.cpp"
    src.write_text('#include "HSOpticalFlowOpenCL.hpp"\n#include "OpticalFlowOpenCV.hpp"\n'
                   'int f(char* a){ HSOpticalFlowOpenCL c("OpticalFlow", a, a, a, a, 15, 100, 1, a);\n'
                   ' c.initialize(); c.setup(); int r = c.run(); c.cleanup(); c.verifyResults();\n'
                   ' HSOpticalFlowOpenCL d("OpticalFlow", a, 15, 100, 1, a); cl_float4* p = 0; d.readInputImage(&p);\n'
                   ' OpticalFlowOpenCV* e = new OpticalFlowOpenCV(); e->runFromImg(a, a, a, .1f, 100); e->runFromCamera(.1f, 100);\n'
                   ' return r + SDK_SUCCESS + SDK_FAILURE + GROUP_SIZE; }\n')
    subprocess.check_call(["g++", "-std=c++17", "-Wno-write-strings", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                           "-o", str(tmp_path / "caller.o")])


def test_stream_model_is_self_consistent(oracle, frames):
    import numpy as np
    import stream_model as M
    g1, g2 = frames["bunny_1"][:40, :70], frames["bunny_2"][:40, :70]
    Ex, Ey, Et = oracle.derivatives(g1.astype(np.float32), g2.astype(np.float32))
    a, b, c = M.normalise(Ex, Ey, Et, 225.0)
    for st8 in (True, False):
        for T, chunk, vw, halo in [(1, 16, 32, 4), (4, 13, 24, 4), (6, 40, 48, 8)]:
            us, vs = M.stream_block_strips(np.zeros_like(Ex), np.zeros_like(Ex), a, b, c, T, chunk, vw, halo, st8)
            ud, vd = np.zeros_like(Ex), np.zeros_like(Ex)
            for _ in range(T):
                ud, vd = M.sweep_direct(ud, vd, a, b, c, st8)
            assert (us.view(np.uint32) == ud.view(np.uint32)).all() and (vs.view(np.uint32) == vd.view(np.uint32)).all()
    u, v = np.zeros_like(Ex), np.zeros_like(Ex)
    for _ in range(30):
        u, v = M.sweep_direct(u, v, a, b, c, True)
    uo, vo = oracle.jacobi(Ex, Ey, Et, 15.0, 30, True)
    assert np.abs(u - uo).max() < 1e-5 and np.abs(v - vo).max() < 1e-5     # contract is 1e-3


def test_host_rasteriser_equals_opencv_drawing(oracle, frames, pictures):
    """hsimg_draw_flow (libhsflow_host.so: the drawing loop the drop-in classes run, cpp:758-770 / cv.cpp:32-46) against
    cv2.circle + cv2.line: every pixel equal on the reference's own pictures and on random fields with flows of hundreds
    of pixels (lines clipped at all four borders).  Host code only: no GPU needed."""
    pytest.importorskip("cv2")
    import ctypes as C
    import numpy as np
    from opticalflowhs_b200 import build_host
    build_host.build()
    L = C.CDLL(os.path.join(ROOT, "opticalflowhs_b200", "libhsflow_host.so"))
    L.hsimg_draw_flow.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]

    def ours(u, v, thr, scale):
        u, v = np.ascontiguousarray(u, np.float32), np.ascontiguousarray(v, np.float32)
        out = np.zeros(u.shape + (3,), np.uint8)
        assert L.hsimg_draw_flow(u.ctypes.data, v.ctypes.data, u.shape[1], u.shape[0], thr, scale, out.ctypes.data) == 0
        return out

    g1, g2 = frames["bunny_1"], frames["bunny_2"]
    u, v = oracle.run_cl(g1, g2, 15.0, 10, False)
    assert (oracle.jpeg_roundtrip(ours(u, v, 0.5, 1.0)) == pictures["bunny_cl_a15_n10"]).all()
    u, v, _ = oracle.run_cv(g1, g2, 0.1, 10, eps=1e-6)
    assert (oracle.jpeg_roundtrip(ours(u, v, 1.0, 0.5)) == pictures["bunny_cv_l0.1_n10"]).all()
    rng = np.random.default_rng(0)
    for _ in range(120):
        h, w = int(rng.integers(5, 60)), int(rng.integers(5, 80))
        keep = rng.random((h, w)) < 0.15
        u = (rng.standard_normal((h, w)) * rng.choice([1, 5, 30, 200]) * keep).astype(np.float32)
        v = (rng.standard_normal((h, w)) * rng.choice([1, 5, 30, 200]) * keep).astype(np.float32)
        for thr, scale in ((0.5, 1.0), (1.0, 0.5)):
            assert (ours(u, v, thr, scale) == oracle.render_flow(u, v, thr, scale)).all(), (h, w)


def test_stream_model_stage_limit_and_max_norm_tracking(oracle):
    """The EPS criterion on the temporally blocked kernel, on the numpy model of its bookkeeping: a block whose stages
    >= lim pass their input through equals exactly `lim` direct sweeps (tail blocks, replay launch), and the per-stage
    max-norm over the rows a unit owns equals max |sweep s+1 - sweep s| there -- for chunks at the top, in the middle and
    at the bottom of the frame (first-row replicate, virtual bottom row)."""
    import numpy as np
    import stream_model as M
    f1, f2 = oracle.synth_pair(96, 40, seed=9)
    Ex, Ey, Et = oracle.derivatives(f1.astype(np.float32), f2.astype(np.float32))
    a, b, c = M.normalise(Ex, Ey, Et, 225.0)
    rng = np.random.default_rng(2)
    u0 = rng.standard_normal((40, 96)).astype(np.float32)
    v0 = rng.standard_normal((40, 96)).astype(np.float32)
    T = 4
    direct = [(u0, v0)]
    for _ in range(T):
        direct.append(M.sweep_direct(direct[-1][0], direct[-1][1], a, b, c))
    for (R0, R1) in ((0, 13), (13, 29), (29, 40), (0, 40)):
        for lim in range(0, T + 1):
            emax = [0.0] * T
            su, sv = M.stream_block(u0, v0, a, b, c, T, R0, R1, lim=lim, emax=emax)
            wu, wv = direct[lim]
            assert (su[R0:R1].view(np.uint32) == wu[R0:R1].view(np.uint32)).all(), (R0, R1, lim)
            assert (sv[R0:R1].view(np.uint32) == wv[R0:R1].view(np.uint32)).all(), (R0, R1, lim)
            for s_ in range(T):
                if s_ < lim:
                    want = max(float(np.abs(direct[s_ + 1][0][R0:R1] - direct[s_][0][R0:R1]).max()),
                               float(np.abs(direct[s_ + 1][1][R0:R1] - direct[s_][1][R0:R1]).max()))
                    assert emax[s_] == want, (R0, R1, lim, s_)
                else:
                    assert emax[s_] == 0.0
