"""CPU stand-in for the HSFlow engine (test infrastructure): same method set, FAST formulation
from tests/stream_model.py.  Lets the gloo tests drive opticalflowhs_b200.sharding unchanged.

It also models the PEER transport of the row-strip mode (hsflow_strip_export / hsflow_strip_connect):
the two u/v ping-pong buffers and the signal words of every strip live in POSIX shared memory, a block
of T iterations stores the seam rows into the neighbours' destination buffers and then publishes its
epoch in their signal words, and the next block starts when both neighbours published the same epoch --
the protocol of k_jacobi_stream + cuStreamWaitValue32, run by real concurrent processes."""
import pickle
import time
from multiprocessing import shared_memory

import numpy as np
import torch

import oracle as O
import stream_model as M


class NumpyEngine:
    device = -1

    def __init__(self, alpha=15.0, temporal_block=4):
        self.rho = np.float32(alpha) * np.float32(alpha)
        self.top = self.bottom = True
        self.temporal_block = temporal_block
        self.connected = False
        self._shm = []

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
        if self.connected:
            return self._iterate_connected(n)
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

    # ---- peer transport model -------------------------------------------------------------------
    def sync(self):
        return self

    def _alloc_shared(self, shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        m = shared_memory.SharedMemory(create=True, size=max(n, 8))
        self._shm.append(m)
        a = np.ndarray(shape, dtype, buffer=m.buf)
        a[...] = 0
        return m.name, a

    def _attach(self, name, shape, dtype):
        m = shared_memory.SharedMemory(name=name)
        self._shm.append(m)
        return np.ndarray(shape, dtype, buffer=m.buf)

    def strip_export(self):
        """Move u/v into two shared ping-pong buffers [2 planes, H, W] and publish their names."""
        names = []
        self.bufs = []
        for _ in range(2):
            name, a = self._alloc_shared((2, self.H, self.W), np.float32)
            names.append(name)
            self.bufs.append(a)
        name, self.sig = self._alloc_shared((2,), np.int64)
        names.append(name)
        self.cur, self.epoch = 0, 0
        self.u, self.v = self.bufs[0][0], self.bufs[0][1]
        return pickle.dumps({"names": names, "H": self.H, "W": self.W})

    def strip_connect(self, up, up_rows, down, down_rows):
        self.peer = [None, None]
        self.push = [tuple(up_rows), tuple(down_rows)]
        for d, blob in enumerate((up, down)):
            if blob is None:
                continue
            info = pickle.loads(blob)
            assert info["W"] == self.W
            lo, hi, delta = self.push[d]
            assert 0 <= lo <= hi <= self.H and lo + delta >= 0 and hi + delta <= info["H"]
            bufs = [self._attach(n, (2, info["H"], info["W"]), np.float32) for n in info["names"][:2]]
            self.peer[d] = (bufs, self._attach(info["names"][2], (2,), np.int64))
        self.top, self.bottom = self.peer[0] is None, self.peer[1] is None
        self.connected = True
        return self

    def strip_disconnect(self):
        self.connected = False
        self.peer = [None, None]
        self.u, self.v = self.u.copy(), self.v.copy()
        self.bufs = self.sig = None
        for m in self._shm:
            m.close()
            try:
                m.unlink()
            except FileNotFoundError:
                pass
        self._shm = []
        return self

    def _iterate_connected(self, n):
        T = self.temporal_block
        while n > 0:
            t = min(n, T)
            src, dst = self.bufs[self.cur], self.bufs[self.cur ^ 1]
            lo = 0 if self.top else t
            hi = self.H if self.bottom else self.H - t
            assert (self.top or self.push[0][0] >= t) and (self.bottom or self.H - self.push[1][1] >= t) and lo < hi
            # Seam-first schedule of the kernel (StreamArgs::seam_first): the row chunks that read ghost rows or push seam
            # rows are computed and stored first and the epoch goes out right after them; the interior follows LATER
            # and from a fresh read of the source rows it depends on -- by then a neighbour that ran ahead may already
            # have stored its next block's seam rows into this source buffer's ghost rows.
            first_end = lo if self.peer[0] is None else max(self.push[0][1], self.push[0][0] + t)
            last_start = hi if self.peer[1] is None else min(self.push[1][0], self.push[1][1] - t)
            seam_first = first_end <= last_start

            def block(r0, r1):                                                # rows [r0, r1) of time step +t from src
                a0, a1 = max(r0 - t, 0), min(r1 + t, self.H)
                uu, vv = src[0][a0:a1].copy(), src[1][a0:a1].copy()
                for _ in range(t):                                            # rows within t of a crop edge go stale, unused
                    uu, vv = M.sweep_direct(uu, vv, self.a[a0:a1], self.b[a0:a1], self.c[a0:a1], True)
                return uu[r0 - a0:r1 - a0], vv[r0 - a0:r1 - a0]

            self.epoch += 1
            parts = [(lo, first_end), (last_start, hi)] if seam_first else [(lo, hi)]
            for r0, r1 in parts:
                if r0 < r1:
                    uu, vv = block(r0, r1)
                    dst[0][r0:r1], dst[1][r0:r1] = uu, vv
                    for d in range(2):                                        # the kernel's second store: seam rows
                        if self.peer[d] is not None:
                            plo, phi, delta = self.push[d]
                            q0, q1 = max(plo, r0), min(phi, r1)
                            if q0 < q1:
                                pb = self.peer[d][0][self.cur ^ 1]
                                pb[0][q0 + delta:q1 + delta] = uu[q0 - r0:q1 - r0]
                                pb[1][q0 + delta:q1 + delta] = vv[q0 - r0:q1 - r0]
            for d in range(2):                                                # ... then the epoch word
                if self.peer[d] is not None:
                    self.peer[d][1][1 - d] = self.epoch                      # upper neighbour's word [1], lower's word [0]
            if seam_first and first_end < last_start:
                time.sleep(0.002 * (1 + self.epoch % 3))                      # let the neighbours run ahead
                uu, vv = block(first_end, last_start)
                dst[0][first_end:last_start], dst[1][first_end:last_start] = uu, vv
            deadline = time.time() + 60
            for d in range(2):                                                # cuStreamWaitValue32(sig[d] >= epoch)
                while self.peer[d] is not None and self.sig[d] < self.epoch:
                    if time.time() > deadline:
                        raise RuntimeError("peer signal never arrived")
                    time.sleep(0.0005)
            self.cur ^= 1
            self.u, self.v = self.bufs[self.cur][0], self.bufs[self.cur][1]
            n -= t
        return self
