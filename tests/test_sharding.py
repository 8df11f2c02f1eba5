"""Host logic of the sharding layer on CPU: pair blocks, strip plans, and the halo exchange of
opticalflowhs_b200.sharding.StripSolver run for real over torch.distributed (gloo, world_size 2
and 3) with a numpy stand-in engine.  The sharded result must be bit-identical to the whole frame."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def test_pair_blocks_cover_the_batch_contiguously():
    from opticalflowhs_b200.sharding import pair_block
    for P in (1, 7, 256, 257):
        for n in (1, 2, 3, 4, 8):
            blocks = [pair_block(P, n, r) for r in range(n)]
            assert blocks[0][0] == 0 and blocks[-1][1] == P
            assert all(b[0] == a[1] for a, b in zip(blocks, blocks[1:]))
            assert max(b[1] - b[0] for b in blocks) == -(-P // n)


def test_strip_plan_rows_and_halo_sizes():
    from opticalflowhs_b200.sharding import StripPlan
    H, n, g = 16384, 8, 4
    plans = [StripPlan(H, n, r, g) for r in range(n)]
    assert plans[0].lo == 0 and plans[-1].hi == H
    assert all(p.hi - p.lo == 2048 for p in plans)
    assert plans[0].top_ghost == 0 and plans[0].bottom_ghost == g + 1
    assert plans[3].top_ghost == g and plans[3].bottom_ghost == g + 1 and plans[3].rows == 2048 + 2 * g + 1
    assert plans[-1].bottom_ghost == 0
    # one row of u and v per direction is 2 * W * 4 B = 131072 B at W = 16384
    assert plans[3].halo_bytes_per_exchange(16384) == (2 * g + 1) * 131072
    with pytest.raises(ValueError):
        StripPlan(16, 8, 0, 4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, N, ghost, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from numpy_engine import NumpyEngine
    from opticalflowhs_b200.sharding import StripSolver
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    eng = NumpyEngine(15.0)
    s = StripSolver(eng, W, H, rank, world, ghost, dist=dist, uv_tensors=eng.uv_tensors)
    s.load_synth(77)
    s.run(N)
    np.save(os.path.join(out_dir, f"u{rank}.npy"), s.owned_rows(eng.u))
    np.save(os.path.join(out_dir, f"v{rank}.npy"), s.owned_rows(eng.v))
    np.save(os.path.join(out_dir, f"x{rank}.npy"), np.array([s.exchanges]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,ghost,N", [(2, 3, 10), (3, 1, 4), (2, 4, 8)])
def test_strip_solver_halo_exchange_over_gloo_is_bit_identical(tmp_path, world, ghost, N):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from numpy_engine import NumpyEngine
    W, H = 96, 60
    whole = NumpyEngine(15.0).configure(W, H).synth_frames(H, 0, 77).prepare().iterate(N)
    mp.spawn(_worker, args=(world, _free_port(), W, H, N, ghost, str(tmp_path)), nprocs=world, join=True)
    u = np.concatenate([np.load(tmp_path / f"u{r}.npy") for r in range(world)])
    v = np.concatenate([np.load(tmp_path / f"v{r}.npy") for r in range(world)])
    assert u.shape == (H, W)
    assert (u.view(np.uint32) == whole.u.view(np.uint32)).all() and (v.view(np.uint32) == whole.v.view(np.uint32)).all()
    expected = -(-N // ghost) - 1
    assert int(np.load(tmp_path / "x0.npy")[0]) == expected


def test_push_rows_mirror_the_neighbours_ghost_rows():
    """Peer transport: the rows a strip pushes are exactly the neighbour's ghost rows, in both coordinate systems."""
    from opticalflowhs_b200.sharding import StripPlan
    for H, n, g in ((16384, 8, 4), (61, 3, 2), (137, 4, 8), (64, 2, 1)):
        plans = [StripPlan(H, n, r, g) for r in range(n)]
        assert plans[0].push_rows(-1) is None and plans[-1].push_rows(+1) is None
        for r in range(n):
            p = plans[r]
            if r > 0:
                lo, hi, d = p.push_rows(-1)
                up = plans[r - 1]
                assert (lo + p.a, hi + p.a) == (up.hi, up.b)              # global rows = its bottom ghost
                assert (lo + d, hi + d) == (up.hi - up.a, up.rows)        # its buffer rows
                assert lo == p.top_ghost and hi - lo == g + 1
            if r < n - 1:
                lo, hi, d = p.push_rows(+1)
                dn = plans[r + 1]
                assert (lo + p.a, hi + p.a) == (dn.a, dn.lo)              # global rows = its top ghost
                assert (lo + d, hi + d) == (0, dn.top_ghost)
                assert hi == p.hi - p.a and hi - lo == g


def _worker_p2p(rank, world, port, W, H, N, T, ghost, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from numpy_engine import NumpyEngine
    from opticalflowhs_b200.sharding import StripSolver
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    eng = NumpyEngine(15.0, temporal_block=T)
    s = StripSolver(eng, W, H, rank, world, ghost, dist=dist, uv_tensors=eng.uv_tensors, transport="p2p")
    s.load_synth(77)
    s.run(N)
    first = (s.owned_rows(eng.u).copy(), s.owned_rows(eng.v).copy())
    s.run(N)                      # a second run on the same connection: epochs keep counting, buffers are reused
    assert (s.owned_rows(eng.u) == first[0]).all() and (s.owned_rows(eng.v) == first[1]).all()
    np.save(os.path.join(out_dir, f"u{rank}.npy"), s.owned_rows(eng.u))
    np.save(os.path.join(out_dir, f"v{rank}.npy"), s.owned_rows(eng.v))
    np.save(os.path.join(out_dir, f"x{rank}.npy"), np.array([s.exchanges]))
    s.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,T,ghost,N", [(2, 4, 4, 10), (3, 2, 3, 7), (3, 1, 1, 3)])
def test_strip_solver_peer_transport_protocol_over_shared_memory(tmp_path, world, T, ghost, N):
    """StripSolver(transport="p2p") with the numpy model of the peer protocol (push seam rows into the neighbour's
    destination buffer, publish the epoch, wait for both neighbours): bit-identical to the whole frame."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from numpy_engine import NumpyEngine
    W, H = 96, 60
    whole = NumpyEngine(15.0).configure(W, H).synth_frames(H, 0, 77).prepare().iterate(N)
    mp.spawn(_worker_p2p, args=(world, _free_port(), W, H, N, T, ghost, str(tmp_path)), nprocs=world, join=True)
    u = np.concatenate([np.load(tmp_path / f"u{r}.npy") for r in range(world)])
    v = np.concatenate([np.load(tmp_path / f"v{r}.npy") for r in range(world)])
    assert (u.view(np.uint32) == whole.u.view(np.uint32)).all() and (v.view(np.uint32) == whole.v.view(np.uint32)).all()
    assert int(np.load(tmp_path / "x0.npy")[0]) == 2 * -(-N // T)
