#!/usr/bin/env python
"""EPS-criterion probe (not a test): one 4K pair in the OpenCV configuration of the reference (OpticalFlowOpenCV.cpp:27-29:
3x3 blurs, Sobel estimator, 4-neighbour stencil, lambda = 0.1, cvTermCriteria(ITER | EPS, 100, 1e-6)) and in the CL
configuration, on the single-sweep kernel (one sweep + one check launch per iteration -- round 1) and on the temporally
blocked TRACK kernel (blocks of 4 tracked sweeps + replay launch).  Prints ms per pair, sweeps done and launches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import opticalflowhs_b200 as P  # noqa: E402

W, H, N = int(os.environ.get("W", 3840)), int(os.environ.get("H", 2160)), int(os.environ.get("N", 100))
PAIRS = int(os.environ.get("NP", 1))
for mode in ("cv", "cl"):
    rows = {}
    for kernel, name in ((1, "single-sweep + check per iteration"), (0, "blocked TRACK kernel + replay")):
        with P.HSFlow(0) as e:
            e.set_kernel(kernel)
            if mode == "cv":
                e.set_deriv(P.DERIV_CV).set_params(0.0, N, P.STENCIL_CV4, True, 0).set_lambda(0.1)
            else:
                e.set_params(15.0, N, P.STENCIL_CL8, True, 0)
            e.set_epsilon(1e-6)
            e.configure(W, H, PAIRS).synth_frames(0, 0, 1234)
            ms = []
            for rep in range(6):
                l0 = e.kernel_launches
                e.compute(); e.sync()
                ms.append(max(e.last_ms(1), 0.0) + e.last_ms(2))
                launches = e.kernel_launches - l0
            u, v = e.read_uv(0)
            rows[kernel] = (min(ms[2:]), e.iterations_done(0), launches, u, v)
            print(f"{mode} {W}x{H} x {PAIRS} pair(s), eps 1e-6, cap {N}: {name}: {min(ms[2:]):8.3f} ms, {e.iterations_done(0)} sweeps, "
                  f"{launches} launches, {PAIRS * W * H * e.iterations_done(0) / min(ms[2:]) / 1e3:9.0f} Mpx-it/s", flush=True)
    same = (rows[0][3].view(np.uint32) == rows[1][3].view(np.uint32)).all() and (rows[0][4].view(np.uint32) == rows[1][4].view(np.uint32)).all()
    print(f"{mode}: speed-up {rows[1][0] / rows[0][0]:.2f}x, fields bit-identical: {bool(same)}, sweeps equal: {rows[0][1] == rows[1][1]}", flush=True)
