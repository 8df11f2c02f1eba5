#!/usr/bin/env python
"""Sustained-throughput probe (not a test): runs the stream kernel back to back for a few seconds per
configuration and reports throughput next to board power and SM clock (NVML), so that kernel variants can be
compared in the power-capped regime the 256-pair bench step runs in.

    T=2,4,8 SECS=3 NP=32 [HSFLOW_LIBRARY=...] python tools/sustained_probe.py
"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opticalflowhs_b200 as P  # noqa: E402

try:
    import pynvml
    pynvml.nvmlInit()
    NV = pynvml.nvmlDeviceGetHandleByIndex(0)
except Exception:
    NV = None


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.stop_flag, self.p, self.c = False, [], []

    def run(self):
        while not self.stop_flag and NV is not None:
            self.p.append(pynvml.nvmlDeviceGetPowerUsage(NV) / 1e3)
            self.c.append(pynvml.nvmlDeviceGetClockInfo(NV, pynvml.NVML_CLOCK_SM))
            time.sleep(0.05)


W, H, NP = int(os.environ.get("W", 3840)), int(os.environ.get("H", 2160)), int(os.environ.get("NP", 32))
SECS = float(os.environ.get("SECS", 3))
Ts = [int(x) for x in os.environ.get("T", "4").split(",")]
CH, WPC = int(os.environ.get("CHUNK", 0)), int(os.environ.get("WPC", 0))
e = P.HSFlow(0)
e.configure(W, H, NP).synth_frames(0, 0, 1234)
print(f"# {W}x{H} x {NP} pairs, {SECS:g} s per configuration, lib {os.path.basename(P.hsflow.library_path())}")
for T in Ts:
    N = 25 * T                                   # 25 launches per measurement slice
    e.set_kernel(2).set_tuning(CH, WPC, 0).set_params(15.0, N, 0, True, T)
    e.prepare(); e.iterate(N); e.sync()
    s = Sampler(); s.start()
    t0 = time.time(); ms = []
    while time.time() - t0 < SECS:
        e.iterate(N); e.sync()
        ms.append(e.last_ms(2))
    s.stop_flag = True; s.join()
    half = ms[len(ms) // 2:]                     # second half: thermally / power settled
    rate = NP * W * H * N / (sum(half) / len(half)) / 1e3
    burst = NP * W * H * N / min(ms) / 1e3
    k = len(s.p) // 2
    pw = sum(s.p[k:]) / max(1, len(s.p[k:])); ck = sum(s.c[k:]) / max(1, len(s.c[k:]))
    print(f"T={T} sustained {rate:9.0f} Mpx-it/s  best slice {burst:9.0f}  power {pw:6.0f} W  sm {ck:5.0f} MHz  slices {len(ms)}", flush=True)
e.close()
