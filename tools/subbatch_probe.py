#!/usr/bin/env python
"""Sub-batch size probe (not a test): 256 synthetic 4K pairs x 100 iterations per step, hsflow_compute, for several
sizes of the scratch sub-batch (pairs per launch).   SUBS=32,64,128,256 python tools/subbatch_probe.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opticalflowhs_b200 as P  # noqa: E402

W, H, NP, N = 3840, 2160, int(os.environ.get("NP", 256)), 100
for sub in [int(x) for x in os.environ.get("SUBS", "32,64,128,256").split(",")]:
    e = P.HSFlow(0)
    e.set_tuning(0, 0, sub).set_params(15.0, N, 0, True, int(os.environ.get("T", 0)))
    e.configure(W, H, NP).synth_frames(0, 0, 1234)
    for _ in range(2):
        e.compute(); e.sync()
    ms = []
    for _ in range(4):
        e.compute(); e.sync()
        ms.append(e.last_ms(2))
    m = sorted(ms)[len(ms) // 2]
    print(f"sub_batch={sub:4d}: {m:8.2f} ms per step  {NP * W * H * N / m / 1e3:9.0f} Mpx-it/s", flush=True)
    e.close()
