"""Multi-GPU parity (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).
On a 1-GPU box these tests skip; the exchange logic itself is covered on CPU by tests/test_sharding.py."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, W, H, N, T, ghost, out_dir, transport="nccl"):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import opticalflowhs_b200 as P
    from opticalflowhs_b200.sharding import StripSolver
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    eng = P.HSFlow(rank)
    eng.set_stream(stream.cuda_stream).set_params(15.0, N, P.STENCIL_CL8, True, T)
    s = StripSolver(eng, W, H, rank, world, ghost, dist=dist, transport=transport)
    s.load_synth(4321)
    s.run(N)
    if transport == "p2p":
        s.run(N)                  # second run on the same connection (epochs keep counting, ghost rows are re-zeroed)
    u, v = eng.read_uv()
    np.save(os.path.join(out_dir, f"u{rank}.npy"), s.owned_rows(u))
    np.save(os.path.join(out_dir, f"v{rank}.npy"), s.owned_rows(v))
    s.close()
    dist.barrier(); dist.destroy_process_group(); eng.close()


@pytest.mark.parametrize("T,ghost,N", [(4, 4, 22), (4, 8, 40), (1, 2, 7)])
def test_row_strips_over_nccl_bit_identical_to_one_gpu(tmp_path, T, ghost, N):
    world = min(_ngpu(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    import opticalflowhs_b200 as P
    W, H = 1000, 64 * world + 9
    with P.HSFlow(0) as e:
        e.set_params(15.0, N, P.STENCIL_CL8, True, T)
        e.configure(W, H, 1).synth_frames(0, 0, 4321).compute()
        whole = e.read_uv()
    mp.spawn(_worker, args=(world, _free_port(), W, H, N, T, ghost, str(tmp_path)), nprocs=world, join=True)
    u = np.concatenate([np.load(tmp_path / f"u{r}.npy") for r in range(world)])
    v = np.concatenate([np.load(tmp_path / f"v{r}.npy") for r in range(world)])
    assert (u.view(np.uint32) == whole[0].view(np.uint32)).all() and (v.view(np.uint32) == whole[1].view(np.uint32)).all()


@pytest.mark.parametrize("T,ghost,N,rows", [(4, 4, 22, 64), (4, 6, 40, 64), (1, 1, 7, 64), (8, 8, 19, 64),
                                            (6, 6, 60, 1024), (0, 6, 45, 700)])
def test_row_strips_with_peer_stores_bit_identical_to_one_gpu(tmp_path, T, ghost, N, rows):
    """transport="p2p": the iteration kernel pushes the seam rows into the neighbours' buffers over NVLink and
    signals an epoch word; no exchange step.  Must equal the single-GPU field bit for bit.  The tall cases have
    many row chunks per strip: the seam chunks run first and the epoch is published while the interior still runs."""
    world = min(_ngpu(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    import opticalflowhs_b200 as P
    W, H = (1000 if rows == 64 else 3000), rows * world + 9
    with P.HSFlow(0) as e:
        e.set_params(15.0, N, P.STENCIL_CL8, True, T)
        e.configure(W, H, 1).synth_frames(0, 0, 4321).compute()
        whole = e.read_uv()
    mp.spawn(_worker, args=(world, _free_port(), W, H, N, T, ghost, str(tmp_path), "p2p"), nprocs=world, join=True)
    u = np.concatenate([np.load(tmp_path / f"u{r}.npy") for r in range(world)])
    v = np.concatenate([np.load(tmp_path / f"v{r}.npy") for r in range(world)])
    assert (u.view(np.uint32) == whole[0].view(np.uint32)).all() and (v.view(np.uint32) == whole[1].view(np.uint32)).all()


@pytest.mark.parametrize("T,N,rows", [(6, 60, 400), (4, 22, 130), (0, 45, 1030)])
def test_row_strips_driven_by_one_process_over_several_gpus(T, N, rows):
    """LocalStripSolver: ONE process, one handle per GPU, strips connected through the same-process branch of
    hsflow_strip_connect (cudaDeviceEnablePeerAccess instead of CUDA IPC).  The iteration kernel on GPU k stores its
    seam rows into the buffers of GPUs k - 1 and k + 1 over NVLink.  Bit-identical to the whole frame on one GPU."""
    world = min(_ngpu(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import opticalflowhs_b200 as P
    from opticalflowhs_b200.sharding import LocalStripSolver
    W, H = 3000, rows * world + 9
    with P.HSFlow(0) as e:
        e.set_params(15.0, N, P.STENCIL_CL8, True, T)
        e.configure(W, H, 1).synth_frames(0, 0, 4321).compute()
        whole = e.read_uv()
    engs = [P.HSFlow(d) for d in range(world)]
    try:
        for e in engs:
            e.set_params(15.0, N, P.STENCIL_CL8, True, T)
        s = LocalStripSolver(engs, W, H, 6 if T == 0 else T)
        s.load_synth(4321)
        for rep in range(2):
            s.run(N).sync()
            u, v = s.gather_uv()
            assert (u.view(np.uint32) == whole[0].view(np.uint32)).all() and (v.view(np.uint32) == whole[1].view(np.uint32)).all(), rep
        s.close()
    finally:
        for e in engs:
            e.close()
