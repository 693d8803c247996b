"""world_size-2 gloo tests (CPU) of the N>1 path's host logic: order-preserving shard / scatter / gather
(mirrors the order guarantee of transcribeWithOptions, WhisperKit.swift:801)."""
import os
import socket
from types import SimpleNamespace

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from whisperkit_b200 import distributed as D


def test_shard_bounds_cover_in_order():
    for n in (1, 2, 5, 64, 129):
        for world in (1, 2, 4, 8):
            spans = [D.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_windows, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stride = 64
    allp = None
    if rank == 0:
        allp = torch.arange(n_windows * stride, dtype=torch.float32).view(n_windows, stride)

    def fake_transcribe(shard):
        # "tokens" derived from the window content so order / content mix-ups are detected
        return [SimpleNamespace(tokens=[int(w[0].item()) // stride, int(w.sum().item()) % 1000, 7]) for w in shard]

    out = D.transcribe_sharded(allp, n_windows, stride, torch.device("cpu"), fake_transcribe)
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_windows", [4, 5])
def test_scatter_transcribe_gather_world2(n_windows):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_windows, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(out) == n_windows
    stride = 64
    for i, toks in enumerate(out):
        w = torch.arange(i * stride, (i + 1) * stride, dtype=torch.float32)
        assert toks == [i, int(w.sum().item()) % 1000, 7]
