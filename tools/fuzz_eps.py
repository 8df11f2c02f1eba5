#!/usr/bin/env python
"""Randomised cross-check (not a test): EPS criterion on the temporally blocked TRACK kernel against the single-sweep kernel
with its per-sweep check, over odd shapes (W % 4 != 0, frames thinner than the block, one-pixel frames), both stencils,
LITERAL / FULL, random eps and iteration caps, batches with per-pair stops.  Prints mismatches; exit code 1 if any."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import opticalflowhs_b200 as P  # noqa: E402

rng = np.random.default_rng(int(os.environ.get("SEED", 1)))
SHAPES = [(1, 1), (1, 5), (7, 1), (5, 7), (17, 33), (9, 130), (66, 257), (40, 116), (23, 240), (31, 124), (3, 300), (130, 9), (64, 512), (50, 1000)]
bad = runs = 0
for trial in range(int(os.environ.get("TRIALS", 120))):
    H, W = SHAPES[rng.integers(len(SHAPES))]
    pairs = int(rng.choice([1, 1, 2, 5]))
    stencil = int(rng.integers(2))
    upd = bool(rng.integers(4))                      # LITERAL one time in four
    N = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 13, 40, 101]))
    eps = float(rng.choice([0.0, 1e-12, 1e-4, 3e-3, 2e-2, 0.2, 5.0, 1e9]))
    frames = rng.integers(0, 256, (pairs, 2, H, W), dtype=np.uint8)
    if pairs > 1:
        frames[1, 1] = frames[1, 0]                  # a pair without motion stops after one sweep
    res = []
    for kernel in (1, 0):
        with P.HSFlow(0) as e:
            e.set_kernel(kernel).set_params(15.0, N, stencil, upd, 0).set_epsilon(eps)
            if stencil == P.STENCIL_CV4:
                e.set_lambda(0.1)
            e.configure(W, H, pairs)
            for k in range(pairs):
                e.set_frames(frames[k, 0], frames[k, 1], pair=k)
            e.compute()
            res.append([(e.iterations_done(k),) + e.read_uv(k) for k in range(pairs)])
    runs += 1
    for k in range(pairs):
        (i1, u1, v1), (i4, u4, v4) = res[0][k], res[1][k]
        if i1 != i4 or not (u1.view(np.uint32) == u4.view(np.uint32)).all() or not (v1.view(np.uint32) == v4.view(np.uint32)).all():
            bad += 1
            print("MISMATCH", dict(H=H, W=W, pairs=pairs, pair=k, stencil=stencil, update_v=upd, N=N, eps=eps, sweeps=(i1, i4),
                                   du=float(np.abs(u1 - u4).max()), dv=float(np.abs(v1 - v4).max())), flush=True)
print(f"{runs} configurations, {bad} mismatching pairs")
sys.exit(1 if bad else 0)
