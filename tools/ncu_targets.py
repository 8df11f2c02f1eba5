#!/usr/bin/env python
"""Small, fixed workloads for ncu captures of each kernel on the hot path (not a test):

    MODE=stream  32 x 4K pairs, two hsflow_compute (k_deriv<GRAY8>, then 17 launches of k_jacobi_stream<6> each)
    MODE=single  8 x 4K pairs, 6 sweeps on the single-sweep kernel k_jacobi1 (FAST math)
    MODE=track   8 x 4K pairs, EPS criterion on the blocked kernel: k_jacobi_stream<4, .., TRACK> main + replay launches
    MODE=bgr     8 x 4K BGR pairs: k_deriv<BGR8> (fused gray conversion)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import opticalflowhs_b200 as P  # noqa: E402

MODE = os.environ.get("MODE", "stream")
W, H = 3840, 2160
with P.HSFlow(0) as e:
    if MODE == "stream":
        e.set_params(15.0, 100, P.STENCIL_CL8, True, 0).configure(W, H, 32).synth_frames(0, 0, 1234)
        e.compute().compute().sync()
        print("stream: T =", e.temporal_block, "ms", e.last_ms(2))
    elif MODE == "single":
        e.set_kernel(1).set_params(15.0, 6, P.STENCIL_CL8, True, 1).configure(W, H, 8).synth_frames(0, 0, 1234)
        e.compute().sync()
        print("single: ms", e.last_ms(2))
    elif MODE == "track":
        e.set_params(15.0, 12, P.STENCIL_CL8, True, 0).set_epsilon(1e-6).configure(W, H, 8).synth_frames(0, 0, 1234)
        e.compute().sync()
        print("track: T =", e.temporal_block, "ms", e.last_ms(2), "sweeps", e.iterations_done(0))
    elif MODE == "bgr":
        rng = np.random.default_rng(0)
        f = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        e.set_params(15.0, 6, P.STENCIL_CL8, True, 0).configure(W, H, 8)
        for k in range(8):
            e.set_frames(f, np.roll(f, 2, axis=1), pair=k)
        e.compute().sync()
        print("bgr: ms", e.last_ms(1))
