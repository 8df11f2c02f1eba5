#!/usr/bin/env python
"""JPEG video-ingest probe (not a test): N consecutive 4K frames as JPEG bitstreams -> hsingest_run_jpeg_batch (batched nvJPEG
decode into the engine's frame planes, overlapped with compute; stride-4 samples back).  One nvJPEG backend per process:

    HSFLOW_NVJPEG_BACKEND=gpu|default|hybrid|hardware python tools/ingest_probe.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2  # noqa: E402
import numpy as np  # noqa: E402
import opticalflowhs_b200 as P  # noqa: E402
from opticalflowhs_b200 import ingest  # noqa: E402

W, H, N, IT = int(os.environ.get("W", 3840)), int(os.environ.get("H", 2160)), int(os.environ.get("FRAMES", 33)), int(os.environ.get("N", 100))
rng = np.random.default_rng(1)
base = cv2.GaussianBlur(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), (0, 0), 3.0)
streams = []
for k in range(N):
    ok, enc = cv2.imencode(".jpg", np.roll(base, (k, 2 * k), axis=(0, 1)), [cv2.IMWRITE_JPEG_QUALITY, 90])
    streams.append(enc.tobytes())
mb = sum(len(s) for s in streams) / 1e6
with P.HSFlow(0) as e:
    e.set_params(15.0, IT, P.STENCIL_CL8, True, 0)
    ingest.run_jpeg_batch(e, streams[:5], sequence=True, sample_step=4)         # warm-up: decoder state, allocations
    t0 = time.perf_counter()
    u, v, st = ingest.run_jpeg_batch(e, streams, sequence=True, sample_step=4)
    dt = time.perf_counter() - t0
print(f"backend {os.environ.get('HSFLOW_NVJPEG_BACKEND', 'default')} -> nvjpegBackend_t {st['backend']}, {st['threads']} host thread(s): {N} frames {W}x{H} "
      f"({mb:.1f} MB of JPEG), {IT} iterations: {dt * 1e3:.1f} ms total = {(N - 1) / dt:.1f} pairs/s, decode calls {st['decode_ms']:.1f} ms "
      f"({st['images'] / st['decode_ms'] * 1e3:.0f} images/s), {st['pairs_per_chunk']} pairs per chunk; finite: {bool(np.isfinite(u).all())}")
