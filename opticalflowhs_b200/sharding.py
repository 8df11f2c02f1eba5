"""Sharding layer of the Horn-Schunck engine (the reference has no multi-device code).

One process per GPU (torchrun), `torch.distributed` only for plumbing:

  * pair sharding  -- independent frame pairs, contiguous blocks of ceil(P/n) pairs per rank, no
                      data-path collective (`pair_block`);
  * row strips     -- one very large frame cut into `world` horizontal strips.  Every rank keeps
                      `ghost` extra rows of u, v (and of the frames / coefficients) on each inner
                      seam.  Two transports for the seam rows:
                        "p2p"  (default on GPUs) the iteration kernel itself stores the rows its
                               neighbours keep as ghosts into THEIR buffers over NVLink peer memory
                               and publishes an epoch word; streams wait on the word between
                               launches (hsflow_strip_connect).  No exchange step, no collective:
                               torch.distributed only carries the IPC handles once.
                        "nccl" advance `ghost` iterations, then swap seam rows with both
                               neighbours (`StripSolver.exchange`: NCCL send/recv, non-periodic).
                      Either way the result is bit-identical to the single-GPU run because every
                      pixel sees the same operation sequence.

The solver talks to an engine object with the HSFlow method set (configure, set_strip,
set_frames / synth_frames, prepare, iterate, halo_refreshed) plus a callable returning the flow
planes as a list of [rows, n] torch tensors; the CPU test-suite
plugs a numpy stand-in into the same exchange code under the gloo backend.
"""
import math


def pair_block(n_pairs, world, rank):
    """Contiguous block [lo, hi) of frame pairs owned by `rank` (ceil(P/n) per rank)."""
    per = (n_pairs + world - 1) // world
    lo = min(rank * per, n_pairs)
    return lo, min(lo + per, n_pairs)


class StripPlan:
    """Rows of a W x H frame owned by one rank, and the buffer rows (owned + ghosts) it holds.

    own  = [lo, hi)          rows this rank is responsible for
    buf  = [a, b)            rows resident on the rank: `ghost` rows above lo, `ghost + 1` rows
                             below hi (the extra row feeds the j+1 tap of the derivative
                             stencil, Kernels.cl:26), clipped at the true image edges
    """

    def __init__(self, height, world, rank, ghost):
        if world < 1 or not (0 <= rank < world):
            raise ValueError("bad rank/world")
        base = height // world
        if base < ghost + 1:
            raise ValueError(f"strips of {base} rows are thinner than ghost+1 = {ghost + 1}")
        self.height, self.world, self.rank, self.ghost = height, world, rank, ghost
        self.lo = rank * base
        self.hi = height if rank == world - 1 else (rank + 1) * base
        self.a = max(self.lo - ghost, 0)
        self.b = min(self.hi + ghost + 1, height)
        self.is_top = rank == 0
        self.is_bottom = rank == world - 1

    @property
    def rows(self):
        return self.b - self.a

    @property
    def top_ghost(self):
        return self.lo - self.a

    @property
    def bottom_ghost(self):
        return self.b - self.hi

    def neighbour(self, rank):
        return StripPlan(self.height, self.world, rank, self.ghost)

    def push_rows(self, direction):
        """Peer transport: (lo, hi, delta) in THIS rank's buffer rows -- output rows [lo, hi) are also stored at
        row + delta of the neighbour's buffer (direction -1: upper neighbour, whose bottom ghost rows they are;
        +1: lower neighbour, its top ghost rows).  None at a true image edge."""
        if direction < 0:
            if self.is_top:
                return None
            n = self.neighbour(self.rank - 1)
            return self.lo - self.a, n.b - self.a, self.a - n.a
        if self.is_bottom:
            return None
        n = self.neighbour(self.rank + 1)
        return n.a - self.a, self.hi - self.a, self.a - n.a

    def halo_bytes_per_exchange(self, width):
        """fp32 bytes this rank SENDS per exchange (u and v)."""
        rows = (0 if self.is_top else self.ghost + 1) + (0 if self.is_bottom else self.ghost)
        return rows * width * 4 * 2


class DeviceView:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can wrap it."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def engine_uv_tensors(engine, device_index):
    """Current flow planes of an HSFlow handle as torch tensors over device memory (no copy).

    The engine keeps u and v row-interleaved in ONE buffer ([row][u|v][pitch]), so a block of rows of
    the returned [rows, 2*pitch] tensor carries both fields and one send per seam and direction is enough."""
    import torch
    u, _, rp, _ = engine.device_uv()
    dev = torch.device("cuda", device_index)
    return [torch.as_tensor(DeviceView(u, (engine.H, rp)), device=dev)]


class StripSolver:
    """Row-strip Horn-Schunck over `world` ranks.  transport "p2p": seam rows pushed by the iteration kernel
    into the neighbours' buffers (needs ghost >= temporal block; one exchange per launch, fused); "nccl": halo
    exchange every `ghost` iterations with send/recv."""

    def __init__(self, engine, width, height, rank, world, ghost, dist=None, uv_tensors=None, transport="nccl"):
        if transport not in ("nccl", "p2p"):
            raise ValueError("transport must be 'nccl' or 'p2p'")
        self.e, self.W, self.H = engine, width, height
        self.plan = StripPlan(height, world, rank, ghost)
        self.dist = dist
        self.transport = transport if world > 1 else "nccl"
        self._uv = uv_tensors or (lambda: engine_uv_tensors(engine, engine.device))
        self.exchanges = 0
        engine.configure(width, self.plan.rows, 1)
        engine.set_strip(self.plan.is_top, self.plan.is_bottom)
        if self.transport == "p2p":
            self._connect()

    def _connect(self):
        """Carry every strip's handle to its neighbours (one object all-gather, plumbing only) and connect."""
        p, dist = self.plan, self.dist
        mine = self.e.strip_export()
        handles = [None] * p.world
        dist.all_gather_object(handles, mine)
        up, dn = p.push_rows(-1), p.push_rows(+1)
        self.e.strip_connect(handles[p.rank - 1] if up else None, up or (0, 0, 0),
                             handles[p.rank + 1] if dn else None, dn or (0, 0, 0))
        dist.barrier()                 # every strip is connected before anyone launches

    def close(self):
        """Peer mappings must go before any strip frees its buffers."""
        if self.transport == "p2p":
            self.e.sync()
            self.dist.barrier()
            self.e.strip_disconnect()
            self.dist.barrier()
            self.transport = "nccl"

    # ---- frames -------------------------------------------------------------------------------
    def load_synth(self, seed):
        self.e.synth_frames(self.H, self.plan.a, seed)

    def load_frames(self, f1, f2):
        """f1, f2: the FULL frames (numpy); only this rank's rows are uploaded."""
        self.e.set_frames(f1[self.plan.a:self.plan.b], f2[self.plan.a:self.plan.b])

    # ---- halo exchange ---------------------------------------------------------------------------
    def exchange(self):
        """Swap seam rows of the current u/v planes with the upper and lower neighbour."""
        p, g = self.plan, self.plan.ghost
        if p.world > 1:
            dist = self.dist
            ops = []
            for t in self._uv():      # one [rows, 2*pitch] tensor on the GPU engine, [u, v] on the CPU stand-in
                if not p.is_top:      # upper neighbour's bottom ghost holds g+1 rows; my top ghost holds g rows
                    ops.append(dist.P2POp(dist.isend, t[p.top_ghost:p.top_ghost + g + 1], p.rank - 1))
                    ops.append(dist.P2POp(dist.irecv, t[0:p.top_ghost], p.rank - 1))
                if not p.is_bottom:
                    own_end = p.hi - p.a
                    ops.append(dist.P2POp(dist.isend, t[own_end - g:own_end], p.rank + 1))
                    ops.append(dist.P2POp(dist.irecv, t[own_end:own_end + p.bottom_ghost], p.rank + 1))
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            self.exchanges += 1
        self.e.halo_refreshed()

    # ---- driver ---------------------------------------------------------------------------------
    def run(self, iterations):
        self.e.prepare()
        if self.transport == "p2p":    # the engine exchanges inside every launch
            self.e.iterate(iterations)
            T = max(1, self.e.temporal_block)
            self.exchanges += -(-iterations // T)
            return self
        done = 0
        while done < iterations:
            step = min(self.plan.ghost, iterations - done)
            self.e.iterate(step)
            done += step
            if done < iterations:
                self.exchange()
        return self

    def owned_rows(self, plane):
        """Slice of a [rows, ...] buffer plane that this rank owns."""
        return plane[self.plan.lo - self.plan.a:self.plan.hi - self.plan.a]


class LocalStripSolver:
    """Row strips of ONE frame driven by ONE process: `engines[k]` owns strip k (any mix of devices -- eight GPUs of a
    box, or several handles on one GPU).  Always the fused peer transport: the strips are connected through the
    same-process branch of hsflow_strip_connect (plain peer access instead of CUDA IPC), every launch stores its seam
    rows into the neighbours' buffers and the streams wait on each other's epoch words.  No torch.distributed at all.

    All work is asynchronous, so one host thread can feed every strip; launches are issued round-robin in slices of
    a few temporal blocks so that no strip's stream runs dry (or its launch queue fills up) while the host is still
    busy with another strip."""

    def __init__(self, engines, width, height, ghost):
        world = len(engines)
        self.engines, self.W, self.H = list(engines), width, height
        self.plans = [StripPlan(height, world, r, ghost) for r in range(world)]
        for e, p in zip(self.engines, self.plans):
            e.configure(width, p.rows, 1)
            e.set_strip(p.is_top, p.is_bottom)
        self.connected = False
        if world > 1:
            handles = [e.strip_export() for e in self.engines]
            for r, (e, p) in enumerate(zip(self.engines, self.plans)):
                up, dn = p.push_rows(-1), p.push_rows(+1)
                e.strip_connect(handles[r - 1] if up else None, up or (0, 0, 0), handles[r + 1] if dn else None, dn or (0, 0, 0))
            self.connected = True

    def load_synth(self, seed):
        for e, p in zip(self.engines, self.plans):
            e.synth_frames(self.H, p.a, seed)

    def load_frames(self, f1, f2):
        for e, p in zip(self.engines, self.plans):
            e.set_frames(f1[p.a:p.b], f2[p.a:p.b])

    def run(self, iterations, slice_blocks=8):
        for e in self.engines:
            e.prepare()
        T = max(1, self.engines[0].temporal_block)
        done = 0
        while done < iterations:
            n = min(iterations - done, slice_blocks * T)
            for e in self.engines:
                e.iterate(n)
            done += n
        return self

    def sync(self):
        for e in self.engines:
            e.sync()
        return self

    def gather_uv(self):
        """The whole field, assembled on the host from every strip's owned rows (tests; small frames)."""
        import numpy as np
        us, vs = [], []
        for e, p in zip(self.engines, self.plans):
            u, v = e.read_uv()
            us.append(u[p.lo - p.a:p.hi - p.a]); vs.append(v[p.lo - p.a:p.hi - p.a])
        return np.concatenate(us), np.concatenate(vs)

    def close(self):
        """Peer mappings must go before any strip frees its buffers."""
        if self.connected:
            self.sync()
            for e in self.engines:
                e.strip_disconnect()
            self.connected = False


def ideal_strip_time_s(width, height, iterations, world, hbm_gbs, bytes_per_px_it=28.0):
    """Lower bound of SURVEY.md 8d: algorithmic bytes / (world x HBM bandwidth)."""
    return width * height * iterations * bytes_per_px_it / (hbm_gbs * 1e9) / world


def halo_fraction(width, rows_per_rank, ghost):
    """Share of redundant (ghost) rows a rank computes per block, averaged over the block."""
    return (ghost + 1) / float(rows_per_rank) if rows_per_rank else math.inf
