"""ctypes front-end of oracle/liboracle.so (plain-C restatement) and oracle/_ref/libclref.so
(the reference's own Kernels.cl compiled for the host).  TEST INFRASTRUCTURE ONLY.

Function-by-function citations live in hs_oracle.h / hs_oracle.c / clref_shim.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_REF = os.path.join(_HERE, "_ref", "libclref.so")

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile liboracle.so (always possible: gcc) and, when /root/reference exists, _ref/libclref.so."""
    src_new = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("hs_oracle.c", "hs_oracle.h"))
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < src_new:
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/OpticalFlowHS/Kernels.cl"):
        shim = os.path.join(_HERE, "clref_shim.cpp")
        if force or not os.path.exists(_REF) or os.path.getmtime(_REF) < os.path.getmtime(shim):
            subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.hso_bgr2gray.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, _u8p]
        L.hso_derivatives.argtypes = [_f32p, _f32p, C.c_int, C.c_int, _f32p, _f32p, _f32p]
        L.hso_jacobi.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int]
        L.hso_jacobi_general.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int,
                                         C.c_float, C.c_float, C.c_float, C.c_int, C.c_int]
        L.hso_jacobi_general_eps.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int,
                                             C.c_float, C.c_float, C.c_float, C.c_int, C.c_double, C.c_int]
        L.hso_run_cl.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, _f32p, _f32p]
        L.hso_box3_u8.argtypes = [_u8p, C.c_int, C.c_int, _u8p]
        L.hso_cvhs.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_double, C.c_int, _f32p, _f32p]
        L.hso_cv_derivatives.argtypes = [_u8p, _u8p, C.c_int, C.c_int, _f32p, _f32p, _f32p]
        L.hso_run_cv.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_double, _f32p, _f32p]
        L.hso_dot_mask.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_float, _u8p]
        L.hso_synth_pair.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, _u8p, _u8p]
        L.hso_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def have_ref():
    if not os.path.exists(_REF):
        try:
            build()
        except Exception:
            pass
    return os.path.exists(_REF)


def ref():
    """The reference's own Kernels.cl on the host (None-safe: raises if the .so is absent)."""
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libclref.so not built (needs /root/reference at build time)")
        R = C.CDLL(_REF)
        R.clref_derivatives.argtypes = [_f32p, _f32p, C.c_int, C.c_int, _f32p, _f32p, _f32p]
        R.clref_iterate.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int]
        R.clref_run.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, _f32p, _f32p]
        R.clref_set_threads.argtypes = [C.c_int]
        _ref = R
    return _ref


# ---- plain-C restatement ---------------------------------------------------------------

def bgr2gray(bgr):
    bgr = np.ascontiguousarray(bgr, np.uint8)
    h, w, _ = bgr.shape
    out = np.empty((h, w), np.uint8)
    lib().hso_bgr2gray(bgr.reshape(-1), w, h, 3 * w, out.reshape(-1))
    return out


def derivatives(I1, I2):
    I1 = np.ascontiguousarray(I1, np.float32)
    I2 = np.ascontiguousarray(I2, np.float32)
    h, w = I1.shape
    Ex, Ey, Et = (np.empty((h, w), np.float32) for _ in range(3))
    lib().hso_derivatives(I1, I2, w, h, Ex, Ey, Et)
    return Ex, Ey, Et


def jacobi(Ex, Ey, Et, alpha, iterations, update_v=True, u0=None, v0=None):
    h, w = Ex.shape
    u = np.zeros((h, w), np.float32) if u0 is None else np.array(u0, np.float32, order="C")
    v = np.zeros((h, w), np.float32) if v0 is None else np.array(v0, np.float32, order="C")
    lib().hso_jacobi(u, v, np.ascontiguousarray(Ex), np.ascontiguousarray(Ey), np.ascontiguousarray(Et),
                     w, h, alpha, iterations, int(update_v))
    return u, v


def jacobi_general(Ex, Ey, Et, w_edge, w_diag, rho, iterations, update_v=True, u0=None, v0=None):
    h, w = Ex.shape
    u = np.zeros((h, w), np.float32) if u0 is None else np.array(u0, np.float32, order="C")
    v = np.zeros((h, w), np.float32) if v0 is None else np.array(v0, np.float32, order="C")
    lib().hso_jacobi_general(u, v, np.ascontiguousarray(Ex), np.ascontiguousarray(Ey), np.ascontiguousarray(Et),
                             w, h, w_edge, w_diag, rho, iterations, int(update_v))
    return u, v


def jacobi_general_eps(Ex, Ey, Et, w_edge, w_diag, rho, max_iter, eps, update_v=True, u0=None, v0=None):
    """jacobi_general with the ITER | EPS termination of cvCalcOpticalFlowHS (cv.cpp:29); returns u, v, sweeps executed."""
    h, w = Ex.shape
    u = np.zeros((h, w), np.float32) if u0 is None else np.array(u0, np.float32, order="C")
    v = np.zeros((h, w), np.float32) if v0 is None else np.array(v0, np.float32, order="C")
    it = lib().hso_jacobi_general_eps(u, v, np.ascontiguousarray(Ex), np.ascontiguousarray(Ey), np.ascontiguousarray(Et),
                                      w, h, w_edge, w_diag, rho, max_iter, float(eps), int(update_v))
    return u, v, it


def run_cl(g1, g2, alpha, iterations, update_v=True):
    g1 = np.ascontiguousarray(g1, np.uint8)
    g2 = np.ascontiguousarray(g2, np.uint8)
    h, w = g1.shape
    u, v = np.empty((h, w), np.float32), np.empty((h, w), np.float32)
    lib().hso_run_cl(g1, g2, w, h, alpha, iterations, int(update_v), u, v)
    return u, v


def box3(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.empty_like(img)
    lib().hso_box3_u8(img, w, h, out)
    return out


def cv_derivatives(A, B):
    A = np.ascontiguousarray(A, np.uint8)
    B = np.ascontiguousarray(B, np.uint8)
    h, w = A.shape
    Ix, Iy, It = (np.empty((h, w), np.float32) for _ in range(3))
    lib().hso_cv_derivatives(A, B, w, h, Ix, Iy, It)
    return Ix, Iy, It


def cvhs(A, B, lam, max_iter, eps=1e-6, u0=None, v0=None):
    A = np.ascontiguousarray(A, np.uint8)
    B = np.ascontiguousarray(B, np.uint8)
    h, w = A.shape
    use_prev = u0 is not None
    u = np.zeros((h, w), np.float32) if u0 is None else np.array(u0, np.float32, order="C")
    v = np.zeros((h, w), np.float32) if v0 is None else np.array(v0, np.float32, order="C")
    it = lib().hso_cvhs(A, B, w, h, lam, max_iter, eps, int(use_prev), u, v)
    return u, v, it


def run_cv(g1, g2, lam, max_iter, eps=1e-6):
    g1 = np.ascontiguousarray(g1, np.uint8)
    g2 = np.ascontiguousarray(g2, np.uint8)
    h, w = g1.shape
    u, v = np.empty((h, w), np.float32), np.empty((h, w), np.float32)
    it = lib().hso_run_cv(g1, g2, w, h, lam, max_iter, eps, u, v)
    return u, v, it


def dot_mask(u, v, step=4, thr=0.5):
    u = np.ascontiguousarray(u, np.float32)
    v = np.ascontiguousarray(v, np.float32)
    h, w = u.shape
    m = np.zeros(((h + step - 1) // step, (w + step - 1) // step), np.uint8)
    lib().hso_dot_mask(u, v, w, h, step, thr, m.reshape(-1))
    return m.astype(bool)


def synth_pair(W, H, seed=1234, row0=0, rows=None):
    rows = H - row0 if rows is None else rows
    f1, f2 = np.empty((rows, W), np.uint8), np.empty((rows, W), np.uint8)
    lib().hso_synth_pair(W, H, row0, rows, seed, f1, f2)
    return f1, f2


def set_threads(n):
    lib().hso_set_threads(n)
    if have_ref():
        ref().clref_set_threads(n)


def max_threads():
    return lib().hso_max_threads()


# ---- the reference's own kernels on the host ------------------------------------------

def ref_derivatives(I1, I2):
    I1 = np.ascontiguousarray(I1, np.float32)
    I2 = np.ascontiguousarray(I2, np.float32)
    h, w = I1.shape
    Ex, Ey, Et = (np.empty((h, w), np.float32) for _ in range(3))
    ref().clref_derivatives(I1, I2, w, h, Ex, Ey, Et)
    return Ex, Ey, Et


def ref_iterate(Ex, Ey, Et, alpha, iterations, update_v=True, u0=None, v0=None):
    h, w = Ex.shape
    u = np.zeros((h, w), np.float32) if u0 is None else np.array(u0, np.float32, order="C")
    v = np.zeros((h, w), np.float32) if v0 is None else np.array(v0, np.float32, order="C")
    ref().clref_iterate(u, v, np.ascontiguousarray(Ex), np.ascontiguousarray(Ey), np.ascontiguousarray(Et),
                        w, h, alpha, iterations, int(update_v))
    return u, v


def ref_run(g1, g2, alpha, iterations, update_v=True):
    I1 = np.ascontiguousarray(g1, np.float32)
    I2 = np.ascontiguousarray(g2, np.float32)
    h, w = I1.shape
    u, v = np.empty((h, w), np.float32), np.empty((h, w), np.float32)
    ref().clref_run(I1, I2, w, h, alpha, iterations, int(update_v), u, v)
    return u, v


# ---- large-frame checks without a full CPU run -------------------------------------------

def window_reference(W, H, N, seed, y, x, K, alpha=15.0, update_v=True, frames=None):
    """u, v of the K x K window at (row y, column x) of a W x H frame after N Jacobi sweeps, computed on the window's
    domain of dependence only: N sweeps reach N pixels, so a (K + 2N)^2 crop -- clamped where it touches a true image
    edge, which is then also an edge of the crop -- gives exactly the full-frame values inside the window.
    frames: (f1, f2) full frames; default = the synthetic pair hso_synth_pair(W, H, seed)."""
    y0, y1 = max(y - N - 1, 0), min(y + K + N + 1, H)      # one more row/column for the j+1 / i+1 derivative taps
    x0, x1 = max(x - N - 1, 0), min(x + K + N + 1, W)
    if frames is None:
        f1, f2 = synth_pair(W, H, seed=seed, row0=y0, rows=y1 - y0)
    else:
        f1, f2 = frames[0][y0:y1], frames[1][y0:y1]
    uo, vo = run_cl(np.ascontiguousarray(f1[:, x0:x1]), np.ascontiguousarray(f2[:, x0:x1]), alpha, N, update_v)
    yy, xx = y - y0, x - x0
    return uo[yy:yy + K, xx:xx + K], vo[yy:yy + K, xx:xx + K]


def window_error(u_win, v_win, W, H, N, seed, y, x, alpha=15.0, update_v=True, frames=None):
    """max |du|, max |dv| of a computed K x K window against window_reference."""
    K = u_win.shape[0]
    uo, vo = window_reference(W, H, N, seed, y, x, K, alpha, update_v, frames)
    return float(np.abs(u_win - uo).max()), float(np.abs(v_win - vo).max())


# ---- the drawing loop of the reference, for picture-level known-answer tests ---------------------------------------

def render_flow(u, v, thr, line_scale, step=4):
    """The output picture of the reference: HSOpticalFlowOpenCL.cpp:758-770 (thr 0.5, line to (j + u, i + v)) and
    OpticalFlowOpenCV.cpp:32-46 (thr 1, line to (x + u/2, y + v/2)): black image; on the stride-4 grid, where
    |u| > thr or |v| > thr, a filled blue circle of radius 2 and a red 8-connected line whose end point is truncated to
    int like cvPoint(float, float) does.  Drawn with cv2.circle / cv2.line (same rasterisers as the OpenCV 2.1 C API)."""
    import cv2
    h, w = u.shape
    img = np.zeros((h, w, 3), np.uint8)
    for y in range(0, h, step):
        for x in range(0, w, step):
            a, b = float(u[y, x]), float(v[y, x])
            if a > thr or b > thr or a < -thr or b < -thr:
                cv2.circle(img, (x, y), 2, (255, 0, 0), -1)                       # CV_RGB(0,0,255), cpp:3
                # int + float, evaluated in extended precision by the reference's x87 build, truncated by cvPoint
                cv2.line(img, (x, y), (int(x + a * float(line_scale)), int(y + b * float(line_scale))), (0, 0, 255), 1, 8)   # CV_RGB(255,0,0), cpp:4
    return img


def jpeg_roundtrip(img_bgr, quality=95):
    """cvSaveImage(path.jpg) with its default quality, then cvLoadImage: what a shipped *_out.jpg holds."""
    import cv2
    ok, enc = cv2.imencode(".jpg", img_bgr, [cv2.IMWRITE_JPEG_QUALITY, quality])
    assert ok
    return cv2.imdecode(enc, cv2.IMREAD_COLOR)
