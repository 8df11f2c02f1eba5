"""ctypes binding of libhsflow.so (include/hsflow.h).  One HSFlow object = one hsflow_t handle =
one GPU + one CUDA stream.  Method names follow the C ABI, which in turn cites the reference
code each call replaces (HSOpticalFlowOpenCL.cpp / Kernels.cl)."""
import ctypes as C
import os

import numpy as np

STENCIL_CL8, STENCIL_CV4 = 0, 1
FRAMES_GRAY8, FRAMES_BGR8 = 0, 1
PIPE_SEQUENCE = 1
MATH_FAST, MATH_EXACT = 0, 1
DERIV_CL, DERIV_CV = 0, 1
PHASE_LOAD, PHASE_DERIV, PHASE_ITER, PHASE_READ = 0, 1, 2, 3

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.environ.get("HSFLOW_LIBRARY") or os.path.join(_HERE, "libhsflow.so")   # override: A/B kernel experiments
_lib = None


class HSFlowError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"hsflow error {code}: {msg}")
        self.code = code


def library_path():
    return _LIBPATH


# name -> (restype, argtypes); also the list tests use to check that every symbol of hsflow.h is exported
_P = C.c_void_p
SIGNATURES = {
    "hsflow_last_error": (C.c_char_p, []),
    "hsflow_version": (C.c_int, []),
    "hsflow_device_count": (C.c_int, []),
    "hsflow_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "hsflow_destroy": (C.c_int, [_P]),
    "hsflow_set_stream": (C.c_int, [_P, _P]),
    "hsflow_set_params": (C.c_int, [_P, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]),
    "hsflow_set_lambda": (C.c_int, [_P, C.c_float]),
    "hsflow_set_math": (C.c_int, [_P, C.c_int]),
    "hsflow_set_deriv": (C.c_int, [_P, C.c_int]),
    "hsflow_set_tuning": (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    "hsflow_set_warm_start": (C.c_int, [_P, C.c_int]),
    "hsflow_set_epsilon": (C.c_int, [_P, C.c_double]),
    "hsflow_set_kernel": (C.c_int, [_P, C.c_int]),
    "hsflow_set_graph": (C.c_int, [_P, C.c_int]),
    "hsflow_configure": (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    "hsflow_set_strip": (C.c_int, [_P, C.c_int, C.c_int]),
    "hsflow_set_frames_gray8": (C.c_int, [_P, C.c_int, _P, _P, C.c_size_t]),
    "hsflow_set_frames_bgr8": (C.c_int, [_P, C.c_int, _P, _P, C.c_size_t]),
    "hsflow_set_frames_f32": (C.c_int, [_P, C.c_int, _P, _P, C.c_size_t]),
    "hsflow_set_frames_gray8_dev": (C.c_int, [_P, C.c_int, _P, _P, C.c_size_t]),
    "hsflow_set_frames_bgr8_dev": (C.c_int, [_P, C.c_int, _P, _P, C.c_size_t]),
    "hsflow_map_frames": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "hsflow_swap_frames": (C.c_int, [_P]),
    "hsflow_synth_frames": (C.c_int, [_P, C.c_int, C.c_int, C.c_uint32]),
    "hsflow_load_pair_gray8": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_size_t]),
    "hsflow_load_pair_bgr8": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_size_t]),
    "hsflow_load_pair_f32": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_size_t]),
    "hsflow_compute": (C.c_int, [_P]),
    "hsflow_compute_range": (C.c_int, [_P, C.c_int, C.c_int]),
    "hsflow_prepare": (C.c_int, [_P]),
    "hsflow_iterate": (C.c_int, [_P, C.c_int]),
    "hsflow_halo_refreshed": (C.c_int, [_P]),
    "hsflow_sync": (C.c_int, [_P]),
    "hsflow_strip_export": (C.c_int, [_P, _P]),
    "hsflow_strip_connect": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int]),
    "hsflow_strip_disconnect": (C.c_int, [_P]),
    "hsflow_read_uv": (C.c_int, [_P, C.c_int, _P, _P, C.c_size_t]),
    "hsflow_read_derivatives": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_size_t]),
    "hsflow_write_uv": (C.c_int, [_P, C.c_int, _P, _P, C.c_size_t]),
    "hsflow_get_device_uv": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "hsflow_get_device_frames": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "hsflow_dot_mask": (C.c_int, [_P, C.c_int, C.c_int, C.c_float, _P, C.POINTER(C.c_int)]),
    "hsflow_sample_uv": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "hsflow_run_batch_host": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "hsflow_run_pipeline_host": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "hsflow_run_sequence_host": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "hsflow_run_pipeline_host_multi": (C.c_int, [C.POINTER(_P), C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "hsflow_push_frame_gray8": (C.c_int, [_P, _P, C.c_size_t]),
    "hsflow_last_ms": (C.c_float, [_P, C.c_int]),
    "hsflow_kernel_launches": (C.c_longlong, [_P]),
    "hsflow_iterations_done": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int)]),
    "hsflow_effective_temporal_block": (C.c_int, [_P]),
    "hsflow_sub_batch": (C.c_int, [_P]),
    "hsflow_device": (C.c_int, [_P]),
    "hsflow_alloc_pinned": (_P, [C.c_size_t]),
    "hsflow_free_pinned": (None, [_P]),
}


def lib():
    """Load libhsflow.so.  Fails loudly when it was not built (python -m opticalflowhs_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIBPATH):
            raise HSFlowError(-2, f"{_LIBPATH} is missing: build it with `python -m opticalflowhs_b200.build` "
                                  "(there is no CPU fallback)")
        L = C.CDLL(_LIBPATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class HSFlow:
    """Load a frame pair (or a batch), set alpha/lambda and the iteration count, compute, read back u/v."""

    def __init__(self, device=0):
        self._L = lib()
        self._h = C.c_void_p()
        self._ck(self._L.hsflow_create(device, C.byref(self._h)))
        self.device = device
        self.W = self.H = self.P = 0

    def _ck(self, rc):
        if rc != 0:
            raise HSFlowError(rc, self._L.hsflow_last_error().decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.hsflow_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- parameters
    def set_params(self, alpha=15.0, iterations=100, stencil=STENCIL_CL8, update_v=True, temporal_block=0):
        self._ck(self._L.hsflow_set_params(self._h, alpha, iterations, stencil, int(update_v), temporal_block))
        return self

    def set_lambda(self, lam):
        self._ck(self._L.hsflow_set_lambda(self._h, lam)); return self

    def set_math(self, mode):
        self._ck(self._L.hsflow_set_math(self._h, mode)); return self

    def set_deriv(self, mode):
        self._ck(self._L.hsflow_set_deriv(self._h, mode)); return self

    def set_tuning(self, chunk_rows=0, warps_per_cta=0, sub_batch=0):
        self._ck(self._L.hsflow_set_tuning(self._h, chunk_rows, warps_per_cta, sub_batch)); return self

    def set_kernel(self, which):
        self._ck(self._L.hsflow_set_kernel(self._h, which)); return self

    def set_graph(self, mode):
        """CUDA graph replay of repeated computes: 0 auto (small jobs), 1 never, 2 always."""
        self._ck(self._L.hsflow_set_graph(self._h, mode)); return self

    def set_warm_start(self, keep):
        self._ck(self._L.hsflow_set_warm_start(self._h, int(keep))); return self

    def set_epsilon(self, eps):
        """EPS termination of cvCalcOpticalFlowHS (cv.cpp:29): stop a pair once max |new - old| < eps; 0 = off."""
        self._ck(self._L.hsflow_set_epsilon(self._h, float(eps))); return self

    def iterations_done(self, pair=0):
        n = C.c_int(0)
        self._ck(self._L.hsflow_iterations_done(self._h, pair, C.byref(n)))
        return n.value

    def set_stream(self, cuda_stream):
        self._ck(self._L.hsflow_set_stream(self._h, C.c_void_p(cuda_stream or 0))); return self

    # ---- geometry / ingest
    def configure(self, width, height, pairs=1):
        self._ck(self._L.hsflow_configure(self._h, width, height, pairs))
        self.W, self.H, self.P = width, height, pairs
        return self

    def set_strip(self, is_top_edge, is_bottom_edge):
        self._ck(self._L.hsflow_set_strip(self._h, int(is_top_edge), int(is_bottom_edge))); return self

    def set_frames(self, f1, f2, pair=0):
        """numpy frames: uint8 (H,W) gray, uint8 (H,W,3) BGR or float32 (H,W)."""
        f1, f2 = np.ascontiguousarray(f1), np.ascontiguousarray(f2)
        if f1.shape != f2.shape or f1.dtype != f2.dtype:
            raise ValueError("frames differ in shape/dtype")
        if f1.dtype == np.uint8 and f1.ndim == 2:
            fn = self._L.hsflow_set_frames_gray8
        elif f1.dtype == np.uint8 and f1.ndim == 3 and f1.shape[2] == 3:
            fn = self._L.hsflow_set_frames_bgr8
        elif f1.dtype == np.float32 and f1.ndim == 2:
            fn = self._L.hsflow_set_frames_f32
        else:
            raise ValueError(f"unsupported frame array {f1.dtype} {f1.shape}")
        if (f1.shape[0], f1.shape[1]) != (self.H, self.W):
            raise ValueError("frame size differs from configure()")
        self._ck(fn(self._h, pair, _ptr(f1), _ptr(f2), 0))
        self._L.hsflow_sync(self._h)      # the numpy buffers may go away
        return self

    def load_pair(self, f1, f2):
        f1 = np.asarray(f1)
        self.configure(f1.shape[1], f1.shape[0], 1)
        return self.set_frames(f1, f2, 0)

    def set_frames_dev(self, d_f1, d_f2, pitch, pair=0):
        self._ck(self._L.hsflow_set_frames_gray8_dev(self._h, pair, C.c_void_p(d_f1), C.c_void_p(d_f2), pitch)); return self

    def synth_frames(self, full_height=0, row0=0, seed0=1234):
        self._ck(self._L.hsflow_synth_frames(self._h, full_height, row0, seed0)); return self

    # ---- compute
    def compute(self):
        self._ck(self._L.hsflow_compute(self._h)); return self

    def compute_range(self, p0, n):
        self._ck(self._L.hsflow_compute_range(self._h, p0, n)); return self

    def prepare(self):
        self._ck(self._L.hsflow_prepare(self._h)); return self

    def iterate(self, n):
        self._ck(self._L.hsflow_iterate(self._h, n)); return self

    def halo_refreshed(self):
        self._ck(self._L.hsflow_halo_refreshed(self._h)); return self

    def sync(self):
        self._ck(self._L.hsflow_sync(self._h)); return self

    # ---- row strips with peer transport (seam rows stored into the neighbours' buffers by the iteration kernel)
    STRIP_HANDLE_BYTES = 320

    def strip_export(self):
        buf = C.create_string_buffer(self.STRIP_HANDLE_BYTES)
        self._ck(self._L.hsflow_strip_export(self._h, buf))
        return buf.raw

    def strip_connect(self, up=None, up_rows=(0, 0, 0), down=None, down_rows=(0, 0, 0)):
        """up/down: handle bytes of the neighbours (None at a true image edge); *_rows = (lo, hi, delta): output rows
        [lo, hi) of this strip are also stored at row + delta of that neighbour's buffer."""
        ub = C.create_string_buffer(up, self.STRIP_HANDLE_BYTES) if up is not None else None
        db = C.create_string_buffer(down, self.STRIP_HANDLE_BYTES) if down is not None else None
        self._ck(self._L.hsflow_strip_connect(self._h, ub, *[int(x) for x in up_rows], db, *[int(x) for x in down_rows]))
        return self

    def strip_disconnect(self):
        self._ck(self._L.hsflow_strip_disconnect(self._h)); return self

    # ---- results
    def read_uv(self, pair=0):
        if self.P <= 0:
            raise HSFlowError(-1, "no flow field on the device (configure + compute first; the pipelined host calls "
                                  "deliver their results to host memory)")
        u = np.empty((self.H, self.W), np.float32)
        v = np.empty((self.H, self.W), np.float32)
        self._ck(self._L.hsflow_read_uv(self._h, pair, _ptr(u), _ptr(v), 0))
        return u, v

    def write_uv(self, u, v, pair=0):
        u, v = np.ascontiguousarray(u, np.float32), np.ascontiguousarray(v, np.float32)
        self._ck(self._L.hsflow_write_uv(self._h, pair, _ptr(u), _ptr(v), 0)); return self

    def read_derivatives(self, pair=0):
        out = [np.empty((self.H, self.W), np.float32) for _ in range(3)]
        self._ck(self._L.hsflow_read_derivatives(self._h, pair, _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), 0))
        return tuple(out)

    def device_uv(self):
        u, v, rp, pp = C.c_void_p(), C.c_void_p(), C.c_size_t(), C.c_size_t()
        self._ck(self._L.hsflow_get_device_uv(self._h, C.byref(u), C.byref(v), C.byref(rp), C.byref(pp)))
        return u.value, v.value, rp.value, pp.value

    def device_frames(self):
        a, b, rp, pp = C.c_void_p(), C.c_void_p(), C.c_size_t(), C.c_size_t()
        self._ck(self._L.hsflow_get_device_frames(self._h, C.byref(a), C.byref(b), C.byref(rp), C.byref(pp)))
        return a.value, b.value, rp.value, pp.value

    def dot_mask(self, pair=0, step=4, threshold=0.5):
        m = np.zeros(((self.H + step - 1) // step, (self.W + step - 1) // step), np.uint8)
        cnt = C.c_int()
        self._ck(self._L.hsflow_dot_mask(self._h, pair, step, threshold, _ptr(m), C.byref(cnt)))
        return m.astype(bool), cnt.value

    @staticmethod
    def _host_array(a, dtype, what):
        """The pipelined calls hand raw pointers to the C side: wrong dtype or a strided view would be read / written
        as if it were dense, so refuse instead of guessing."""
        if not isinstance(a, np.ndarray) or a.dtype != np.dtype(dtype) or not a.flags.c_contiguous:
            raise ValueError(f"{what} must be a C-contiguous numpy array of {np.dtype(dtype).name}")
        return a

    def _after_pipeline(self, W, H):
        # the call reconfigured the handle (internal pair slots) and left no current field on the device
        self.W, self.H, self.P = W, H, 0

    def run_batch_host(self, frames, u_out, v_out):
        """frames: uint8 (n,2,H,W); u_out/v_out: float32 (n,H,W); all C-contiguous (pinned for full rate).
        Reconfigures the handle; afterwards configure() + compute() are needed before read_uv()."""
        return self.run_pipeline_host(frames, u_out, v_out)

    def run_sequence_host(self, frames, u_out, v_out):
        """frames: uint8 (n+1,H,W) consecutive frames; pair k = (frame k, frame k+1); u_out/v_out: float32 (n,H,W)."""
        return self.run_pipeline_host(frames, u_out, v_out, sequence=True)

    def run_pipeline_host(self, frames, u_out, v_out, sequence=False, sample_step=0):
        """General pipelined call.  frames: uint8 (n,2,H,W[,3]) pairs or, with sequence=True, (n+1,H,W[,3]) consecutive
        frames; a trailing axis of 3 = interleaved BGR.  sample_step > 0: u_out/v_out are (n, ceil(H/step), ceil(W/step))."""
        frames = self._host_array(frames, np.uint8, "frames")
        lead = 1 if sequence else 2
        if frames.ndim not in (lead + 2, lead + 3) or (frames.ndim == lead + 3 and frames.shape[-1] != 3) or \
                (not sequence and frames.shape[1] != 2):
            raise ValueError(f"unsupported frames array {frames.shape}")
        bgr = frames.ndim == lead + 3
        n = frames.shape[0] - 1 if sequence else frames.shape[0]
        H, W = frames.shape[lead], frames.shape[lead + 1]
        if n < 1:
            raise ValueError("need at least one frame pair")
        oshape = (n, -(-H // sample_step), -(-W // sample_step)) if sample_step else (n, H, W)
        for a, what in ((u_out, "u_out"), (v_out, "v_out")):
            self._host_array(a, np.float32, what)
            if a.shape != oshape:
                raise ValueError(f"{what} must have shape {oshape}, got {a.shape}")
        rc = self._L.hsflow_run_pipeline_host(self._h, _ptr(frames), n, W, H, FRAMES_BGR8 if bgr else FRAMES_GRAY8,
                                              PIPE_SEQUENCE if sequence else 0, int(sample_step), _ptr(u_out), _ptr(v_out))
        self._after_pipeline(W, H)
        self._ck(rc)
        return self

    @staticmethod
    def run_pipeline_host_multi(engines, frames, u_out, v_out, sequence=False, sample_step=0):
        """run_pipeline_host over several handles (one per GPU): contiguous blocks of pairs, one host thread per handle."""
        e0 = engines[0]
        frames = e0._host_array(frames, np.uint8, "frames")
        lead = 1 if sequence else 2
        bgr = frames.ndim == lead + 3
        n = frames.shape[0] - 1 if sequence else frames.shape[0]
        H, W = frames.shape[lead], frames.shape[lead + 1]
        oshape = (n, -(-H // sample_step), -(-W // sample_step)) if sample_step else (n, H, W)
        for a, what in ((u_out, "u_out"), (v_out, "v_out")):
            e0._host_array(a, np.float32, what)
            if a.shape != oshape:
                raise ValueError(f"{what} must have shape {oshape}, got {a.shape}")
        hs = (_P * len(engines))(*[e._h for e in engines])
        rc = e0._L.hsflow_run_pipeline_host_multi(hs, len(engines), _ptr(frames), n, W, H, FRAMES_BGR8 if bgr else FRAMES_GRAY8,
                                                  PIPE_SEQUENCE if sequence else 0, int(sample_step), _ptr(u_out), _ptr(v_out))
        for e in engines:
            e._after_pipeline(W, H)
        e0._ck(rc)

    def sample_uv(self, pair=0, step=4):
        """u, v on the stride-`step` grid (what cpp:762-767 reads)."""
        shape = (-(-self.H // step), -(-self.W // step))
        u, v = np.empty(shape, np.float32), np.empty(shape, np.float32)
        self._ck(self._L.hsflow_sample_uv(self._h, pair, step, _ptr(u), _ptr(v)))
        return u, v

    def map_frames(self, bgr=False):
        """Device pointers of the handle's frame planes (f1, f2, row_pitch, pair_pitch) for on-GPU decoders."""
        a, b, rp, pp = C.c_void_p(), C.c_void_p(), C.c_size_t(), C.c_size_t()
        self._ck(self._L.hsflow_map_frames(self._h, FRAMES_BGR8 if bgr else FRAMES_GRAY8, C.byref(a), C.byref(b), C.byref(rp), C.byref(pp)))
        return a.value, b.value, rp.value, pp.value

    def set_frames_bgr_dev(self, d_f1, d_f2, pitch, pair=0):
        self._ck(self._L.hsflow_set_frames_bgr8_dev(self._h, pair, C.c_void_p(d_f1), C.c_void_p(d_f2), pitch)); return self

    def push_frame(self, frame):
        """Camera-loop step: the second frame becomes the first, `frame` (uint8 (H,W)) the new second one."""
        frame = np.ascontiguousarray(frame, np.uint8)
        assert frame.shape == (self.H, self.W)
        self._ck(self._L.hsflow_push_frame_gray8(self._h, _ptr(frame), 0)); return self

    # ---- instrumentation
    def last_ms(self, phase):
        return self._L.hsflow_last_ms(self._h, phase)

    @property
    def kernel_launches(self):
        return self._L.hsflow_kernel_launches(self._h)

    @property
    def sub_batch(self):
        return self._L.hsflow_sub_batch(self._h)

    @property
    def temporal_block(self):
        return self._L.hsflow_effective_temporal_block(self._h)


class _PinnedBuffer:
    """cudaMallocHost block exposing the buffer protocol through ctypes; freed with the last numpy view."""

    def __init__(self, nbytes):
        self.ptr = lib().hsflow_alloc_pinned(max(nbytes, 1))
        if not self.ptr:
            raise HSFlowError(-4, "cudaMallocHost failed")
        self.nbytes = nbytes

    def __del__(self):
        try:
            if self.ptr:
                lib().hsflow_free_pinned(self.ptr)
                self.ptr = None
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """numpy array over pinned (page-locked) host memory."""
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    owner = _PinnedBuffer(count * dtype.itemsize)
    raw = (C.c_uint8 * max(owner.nbytes, 1)).from_address(owner.ptr)
    raw._owner = owner                       # ctypes arrays accept attributes; numpy keeps `raw` as .base
    return np.frombuffer(raw, dtype=dtype, count=count).reshape(shape)
