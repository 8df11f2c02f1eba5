"""ctypes binding of the ingest entry points of libhsflow_host.so (include/hsflow_ingest.h): JPEG bitstreams are decoded
by nvJPEG on the GPU straight into an HSFlow handle's frame planes -- the replacement of cvLoadImage + cvCvtColor +
readInputImage (HSOpticalFlowOpenCL.cpp:721-740, 6-45) without a host bounce of the decoded pixels."""
import ctypes as C
import os

import numpy as np

from .hsflow import HSFlowError, lib as _core_lib

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libhsflow_host.so")
_lib = None
_P = C.c_void_p

SIGNATURES = {
    "hsingest_last_error": (C.c_char_p, []),
    "hsingest_jpeg_info": (C.c_int, [_P, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "hsingest_decode_to_device": (C.c_int, [_P, C.c_size_t, _P, C.c_size_t, C.c_int, C.c_int, _P]),
    "hsingest_load_pair_files": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "hsingest_load_pair_jpeg": (C.c_int, [_P, _P, C.c_size_t, _P, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "hsingest_push_frame_file": (C.c_int, [_P, C.c_char_p]),
    "hsingest_run_jpeg_batch": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_size_t), C.c_int, C.c_int, C.c_int, _P, _P, C.POINTER(C.c_double)]),
}


def lib():
    global _lib
    if _lib is None:
        _core_lib()                                  # libhsflow.so first (same instance the handles come from)
        if not os.path.exists(_LIBPATH):
            raise HSFlowError(-2, f"{_LIBPATH} is missing: build it with `python -m opticalflowhs_b200.build_host`")
        L = C.CDLL(_LIBPATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _ck(rc):
    if rc != 0:
        raise HSFlowError(rc, lib().hsingest_last_error().decode())


def _buf(b):
    a = np.frombuffer(bytes(b), np.uint8)
    return a, a.ctypes.data_as(_P), a.size


def jpeg_info(data):
    a, p, n = _buf(data)
    w, h, c = C.c_int(), C.c_int(), C.c_int()
    _ck(lib().hsingest_jpeg_info(p, n, C.byref(w), C.byref(h), C.byref(c)))
    return w.value, h.value, c.value


def load_pair_jpeg(engine, jpeg1, jpeg2):
    """configure(w, h, 1) + decode both bitstreams on the GPU into the handle's BGR frame planes."""
    a1, p1, n1 = _buf(jpeg1)
    a2, p2, n2 = _buf(jpeg2)
    w, h = C.c_int(), C.c_int()
    _ck(lib().hsingest_load_pair_jpeg(engine._h, p1, n1, p2, n2, C.byref(w), C.byref(h)))
    engine.W, engine.H, engine.P = w.value, h.value, 1
    return engine


def load_pair_files(engine, path1, path2):
    w, h = C.c_int(), C.c_int()
    _ck(lib().hsingest_load_pair_files(engine._h, os.fsencode(path1), os.fsencode(path2), C.byref(w), C.byref(h)))
    engine.W, engine.H, engine.P = w.value, h.value, 1
    return engine


def push_frame_file(engine, path):
    _ck(lib().hsingest_push_frame_file(engine._h, os.fsencode(path)))
    return engine


def run_jpeg_batch(engine, jpegs, sequence=False, sample_step=0):
    """jpegs: list of JPEG bitstreams (bytes) of one size.  Returns (u, v, stats): fields of every pair, or their
    stride-`sample_step` samples; stats = dict(decode_ms, images, pairs_per_chunk, backend)."""
    bufs = [_buf(j) for j in jpegs]
    n = len(bufs)
    ptrs = (_P * n)(*[b[1] for b in bufs])
    sizes = (C.c_size_t * n)(*[b[2] for b in bufs])
    w, h, _ = jpeg_info(jpegs[0])
    n_pairs = n - 1 if sequence else n // 2
    shape = (n_pairs, -(-h // sample_step), -(-w // sample_step)) if sample_step else (n_pairs, h, w)
    u, v = np.empty(shape, np.float32), np.empty(shape, np.float32)
    stats = (C.c_double * 4)()
    rc = lib().hsingest_run_jpeg_batch(engine._h, ptrs, sizes, n, int(sequence), int(sample_step), u.ctypes.data_as(_P),
                                       v.ctypes.data_as(_P), stats)
    engine.W, engine.H, engine.P = w, h, 0
    _ck(rc)
    return u, v, {"decode_ms": stats[0], "images": int(stats[1]), "pairs_per_chunk": int(stats[2]), "backend": int(stats[3]), "threads": int(round((stats[3] - int(stats[3])) * 100))}
