"""CPU stand-in for the HSFlow engine (test infrastructure): same method set, FAST formulation
from tests/stream_model.py.  Lets the gloo tests drive opticalflowhs_b200.sharding unchanged."""
import numpy as np
import torch

import oracle as O
import stream_model as M


class NumpyEngine:
    device = -1

    def __init__(self, alpha=15.0):
        self.rho = np.float32(alpha) * np.float32(alpha)
        self.top = self.bottom = True

    def configure(self, W, H, pairs=1):
        assert pairs == 1
        self.W, self.H = W, H
        self.u = np.zeros((H, W), np.float32)
        self.v = np.zeros((H, W), np.float32)
        return self

    def set_strip(self, top, bottom):
        self.top, self.bottom = bool(top), bool(bottom)
        return self

    def set_frames(self, f1, f2, pair=0):
        self.f1, self.f2 = np.ascontiguousarray(f1), np.ascontiguousarray(f2)
        return self

    def synth_frames(self, full_height=0, row0=0, seed0=1234):
        self.f1, self.f2 = O.synth_pair(self.W, full_height or self.H, seed=seed0, row0=row0, rows=self.H)
        return self

    def prepare(self):
        Ex, Ey, Et = O.derivatives(self.f1.astype(np.float32), self.f2.astype(np.float32))
        self.a, self.b, self.c = M.normalise(Ex, Ey, Et, self.rho)
        self.u[:] = 0
        self.v[:] = 0
        self.lo, self.hi = 0, self.H
        return self

    def iterate(self, n):
        for _ in range(n):
            lo = 0 if self.top else self.lo + 1
            hi = self.H if self.bottom else self.hi - 1
            if lo >= hi:
                raise RuntimeError("ghost rows exhausted")
            un, vn = M.sweep_direct(self.u, self.v, self.a, self.b, self.c, True)
            self.u[lo:hi], self.v[lo:hi] = un[lo:hi], vn[lo:hi]      # rows outside stay stale, like the GPU engine
            self.lo, self.hi = lo, hi
        return self

    def halo_refreshed(self):
        self.lo, self.hi = 0, self.H
        return self

    def uv_tensors(self):
        return [torch.from_numpy(self.u), torch.from_numpy(self.v)]
