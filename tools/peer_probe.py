#!/usr/bin/env python
"""Row strips of one frame over the GPUs of a box, driven by ONE process (LocalStripSolver: same-process branch of
hsflow_strip_connect, plain peer access) -- the form ncu can profile ("never wrap a multi-rank command in ncu").

    GPUS=2 W=16384 ROWS=2048 N=60 python tools/peer_probe.py            # timing + bit-exact check against one GPU
    ncu --set full -k regex:k_jacobi_stream -s 4 -c 2 ... python tools/peer_probe.py      # SLICE=1 is the default

Each GPU holds ROWS (+ ghost) rows x W columns -- with ROWS = 2048 and W = 16384 that is exactly one GPU's share of
BASELINE.json configs[4] at 8 GPUs.  SLICE = temporal blocks issued per strip before moving to the next strip; 1 keeps
every launch's dependencies issued before it, which a profiler that serialises launches needs."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import opticalflowhs_b200 as P  # noqa: E402
from opticalflowhs_b200.sharding import LocalStripSolver  # noqa: E402

G = int(os.environ.get("GPUS", P.lib().hsflow_device_count()))
W, ROWS, N, T = int(os.environ.get("W", 16384)), int(os.environ.get("ROWS", 2048)), int(os.environ.get("N", 60)), int(os.environ.get("T", 0))
SLICE, CHECK = int(os.environ.get("SLICE", 1)), int(os.environ.get("CHECK", 1))
H = ROWS * G
engs = [P.HSFlow(d) for d in range(G)]
CHUNK = int(os.environ.get("CHUNK", 0))            # rows per work unit (0 = the engine's rule)
for e in engs:
    e.set_params(15.0, N, P.STENCIL_CL8, True, T).set_tuning(chunk_rows=CHUNK)
Teff = engs[0].temporal_block if T else 6
s = LocalStripSolver(engs, W, H, Teff)
s.load_synth(1234)
s.run(N, slice_blocks=SLICE).sync()                       # warm-up (and the run a profiler looks at)
t0 = time.perf_counter()
reps = int(os.environ.get("REPS", 3))
for _ in range(reps):
    s.run(N, slice_blocks=SLICE)
s.sync()
dt = (time.perf_counter() - t0) / reps
out = {"gpus": G, "width": W, "height": H, "iterations": N, "temporal_block": engs[0].temporal_block, "ms": dt * 1e3,
       "mpx_it_per_s": W * H * N / dt / 1e6, "slice_blocks": SLICE, "chunk_rows": CHUNK,
       "seam_bytes_per_launch_per_direction": 2 * engs[0].temporal_block * W * 4,
       "launches_per_gpu": -(-N // engs[0].temporal_block)}
if CHECK and H <= 8192:
    u, v = s.gather_uv()
    with P.HSFlow(0) as e:
        e.set_params(15.0, N, P.STENCIL_CL8, True, T)
        e.configure(W, H, 1).synth_frames(0, 0, 1234).compute()
        uw, vw = e.read_uv()
    out["bit_identical_to_one_gpu"] = bool((u.view(np.uint32) == uw.view(np.uint32)).all() and (v.view(np.uint32) == vw.view(np.uint32)).all())
s.close()
for e in engs:
    e.close()
print(json.dumps(out))
