#!/usr/bin/env python
"""Randomised cross-checks (not a test): (1) hsflow_run_pipeline_host over random shapes, formats, layouts, sub-batch sizes,
sampling steps, block depths and modes against per-pair computes; (2) row strips with the fused peer transport (several handles
on one GPU, LocalStripSolver) over random geometries against the whole frame.  Exit code 1 on any mismatch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import opticalflowhs_b200 as P  # noqa: E402
from opticalflowhs_b200.sharding import LocalStripSolver  # noqa: E402

rng = np.random.default_rng(int(os.environ.get("SEED", 1)))
bad = 0


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def gray(f):                                         # OpenCV 2.1 fixed-point BGR2GRAY (cpp:727-728)
    f = f.astype(np.uint32)
    return ((f[..., 0] * 1868 + f[..., 1] * 9617 + f[..., 2] * 4899 + 8192) >> 14).astype(np.uint8)


for trial in range(int(os.environ.get("TRIALS", 40))):
    W = int(rng.choice([5, 33, 64, 116, 130, 232, 250, 257, 400]))
    H = int(rng.choice([1, 7, 16, 40, 90]))
    n = int(rng.integers(1, 20))
    seq, bgr = bool(rng.integers(2)), bool(rng.integers(2))
    step = int(rng.choice([0, 0, 3, 4]))
    N, T = int(rng.choice([1, 5, 12, 17])), int(rng.choice([0, 1, 3, 4, 6, 8]))
    mode = rng.choice(["fast", "fast", "literal", "exact", "eps"])
    sub = int(rng.choice([0, 1, 3, 4, 16]))
    nf = n + 1 if seq else 2 * n
    fr = rng.integers(0, 256, (nf, H, W, 3) if bgr else (nf, H, W), dtype=np.uint8)
    g = gray(fr) if bgr else fr

    def setup(e):
        e.set_params(15.0, N, P.STENCIL_CL8, mode != "literal", T)
        if mode == "exact":
            e.set_math(P.MATH_EXACT)
        if mode == "eps":
            e.set_epsilon(2e-2)
        return e

    want = []
    with setup(P.HSFlow(0)) as e:
        for k in range(n):
            a, b = (k, k + 1) if seq else (2 * k, 2 * k + 1)
            e.load_pair(g[a], g[b]).compute()
            want.append(e.read_uv())
    shape = (n, -(-H // step), -(-W // step)) if step else (n, H, W)
    u, v = np.empty(shape, np.float32), np.empty(shape, np.float32)
    frames = fr if seq else fr.reshape((n, 2) + fr.shape[1:])
    with setup(P.HSFlow(0)) as e:
        e.set_tuning(sub_batch=sub)
        e.run_pipeline_host(np.ascontiguousarray(frames), u, v, sequence=seq, sample_step=step)
    for k in range(n):
        wu, wv = (want[k][0][::step, ::step], want[k][1][::step, ::step]) if step else want[k]
        if not ((bits(u[k]) == bits(wu)).all() and (bits(v[k]) == bits(wv)).all()):
            bad += 1
            print("PIPELINE MISMATCH", dict(W=W, H=H, n=n, pair=k, seq=seq, bgr=bgr, step=step, N=N, T=T, mode=mode, sub=sub), flush=True)
            break
print("pipeline trials done, mismatches so far:", bad, flush=True)

for trial in range(int(os.environ.get("STRIP_TRIALS", 25))):
    world = int(rng.integers(2, 5))
    T = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 8]))
    ghost = T + int(rng.integers(0, 3))
    rows = int(rng.choice([ghost + 2, 2 * ghost + 5, 40, 97, 300]))
    W = int(rng.choice([64, 130, 500, 1500]))
    H = rows * world + int(rng.integers(0, 7))
    N = int(rng.choice([T, 2 * T + 1, 23, 40]))
    chunk = int(rng.choice([0, 0, 8, 16, 64]))
    with P.HSFlow(0) as e:
        e.set_params(15.0, N, P.STENCIL_CL8, True, T)
        e.configure(W, H, 1).synth_frames(0, 0, 99 + trial).compute()
        whole = e.read_uv()
    engs = [P.HSFlow(0) for _ in range(world)]
    try:
        for e in engs:
            e.set_params(15.0, N, P.STENCIL_CL8, True, T).set_tuning(chunk_rows=chunk)
        s = LocalStripSolver(engs, W, H, ghost)
        s.load_synth(99 + trial)
        for rep in range(2):
            s.run(N, slice_blocks=int(rng.integers(1, 5))).sync()
            uu, vv = s.gather_uv()
            if not ((bits(uu) == bits(whole[0])).all() and (bits(vv) == bits(whole[1])).all()):
                bad += 1
                print("STRIP MISMATCH", dict(world=world, T=T, ghost=ghost, rows=rows, W=W, H=H, N=N, chunk=chunk, rep=rep), flush=True)
                break
        s.close()
    except P.HSFlowError as ex:
        print("strip config refused:", dict(world=world, T=T, ghost=ghost, rows=rows, W=W, H=H, N=N, chunk=chunk), str(ex)[:100], flush=True)
    finally:
        for e in engs:
            e.close()
print("total mismatches:", bad)
sys.exit(1 if bad else 0)
