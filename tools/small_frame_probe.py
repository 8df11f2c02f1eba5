#!/usr/bin/env python
"""Latency probe (not a test): one frame pair of a small size, N iterations, per kernel / temporal block / chunk height.
Small frames cannot fill 148 SMs with 128-column strips, so what matters is launch count and the length of the
longest work unit, not bytes.

    SIZES=600x480,424x240,1920x1080 N=100 python tools/small_frame_probe.py
"""
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opticalflowhs_b200 as P  # noqa: E402

N = int(os.environ.get("N", 100))
sizes = [tuple(int(x) for x in s.split("x")) for s in os.environ.get("SIZES", "600x480,424x240,1920x1080").split(",")]
e = P.HSFlow(0)
for W, H in sizes:
    e.configure(W, H, 1).synth_frames(0, 0, 1234)
    print(f"# {W}x{H}, {N} iterations, iteration phase only (CUDA events), median of 15")
    for kern, T, chunk in [(0, 0, 0), (1, 1, 0), (1, 1, 16), (1, 1, 8)] + [(2, t, c) for t in (1, 2, 3, 4, 6) for c in (0, 8, 16, 32)]:
        try:
            e.set_kernel(kern).set_tuning(chunk, 0, 0).set_params(15.0, N, 0, True, T)
            ms = []
            for _ in range(18):
                e.prepare(); e.iterate(N); e.sync()
                ms.append(e.last_ms(2))
            m = statistics.median(ms[3:])
            print(f"kernel={kern} T={T} chunk={chunk:3d}: {1e3 * m:8.1f} us  ({W * H * N / m / 1e3:9.0f} Mpx-it/s)", flush=True)
        except Exception as ex:
            print(f"kernel={kern} T={T} chunk={chunk}: {ex}")
    # whole computes (derivative pass + all iterations), host wall clock over 50 back-to-back calls: eager launches
    # against CUDA-graph replay (hsflow_set_graph); HSFLOW_NO_PDL=1 in the environment switches programmatic dependent
    # launch off for the same comparison
    e.set_kernel(0).set_tuning(0, 0, 0).set_params(15.0, N, 0, True, 0)
    for graph, name in ((1, "eager launches"), (2, "graph replay")):
        e.set_graph(graph)
        for _ in range(3):
            e.compute()
        e.sync()
        t0 = time.perf_counter()
        for _ in range(50):
            e.compute()
        e.sync()
        us = (time.perf_counter() - t0) / 50 * 1e6
        print(f"compute() {name:15s} PDL {'off' if os.environ.get('HSFLOW_NO_PDL') else 'on '}: {us:8.1f} us per pair  ({W * H * N / us:9.0f} Mpx-it/s)", flush=True)
    e.set_graph(0)
e.close()
