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


def test_assign_streams_is_balanced_and_deterministic():
    lens = [480000 * 7, 1000, 480000 * 3, 480000 * 3, 0, 480000 * 2, 480000]
    for world in (1, 2, 3, 8):
        a = D.assign_streams(lens, world)
        assert sorted(sum(a, [])) == list(range(len(lens))) and a == D.assign_streams(lens, world)
        loads = [sum(lens[i] for i in r) for r in a]
        if world == 2:
            assert max(loads) - min(loads) <= 480000 * 2          # LPT keeps the ranks within one long stream of each other
    assert D.assign_streams([], 2) == [[], []]


def _stream_worker(rank, world, port, lens, q):
    import numpy as np
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    owned = D.assign_streams(lens, world)[rank]
    arrays = [np.full(n, float(i), np.float32) if i in owned else None for i, n in enumerate(lens)]   # a rank only loads what it owns...
    arrays = [a if a is not None or n == 0 else None for a, n in zip(arrays, lens)]
    known = [np.zeros(n, np.float32) if a is None else a for a, n in zip(arrays, lens)]                # ...but every rank knows the lengths

    def fake_streams(local_arrays, ids):
        out = []
        for x, i in zip(local_arrays, ids):
            segs = []
            for k, seek in enumerate(range(0, len(x), 480000)):
                segs.append(SimpleNamespace(id=k, seek=seek + (1 << 33) * (i == 2), start=seek / 16000.0, end=min(len(x), seek + 480000) / 16000.0,
                                            tokens=[i, k, int(x[0]) if len(x) else -1, rank + 100], temperature=0.2 * k))
            out.append(segs)
        return out

    res = D.transcribe_streams_sharded(known, torch.device("cpu"), fake_streams)
    if rank == 0:
        q.put([[(g.id, g.seek, g.start, g.end, g.tokens, round(g.temperature, 4)) for g in segs] for segs in res])
    dist.barrier()
    dist.destroy_process_group()


def test_streams_sharded_gather_world2():
    lens = [480000 * 3 + 5, 480000, 480000 * 2, 0, 1234]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_stream_worker, args=(r, 2, port, lens, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assign = D.assign_streams(lens, 2)
    owner = {i: r for r, ids in enumerate(assign) for i in ids}
    assert len(out) == len(lens) and out[3] == []
    for i, n in enumerate(lens):
        assert len(out[i]) == (n + 479999) // 480000
        for k, (sid, seek, start, end, toks, temp) in enumerate(out[i]):
            assert sid == k and seek == k * 480000 + (1 << 33) * (i == 2)            # 64-bit seeks survive the int32 packing
            assert abs(start - k * 30.0) < 1e-4 and toks == [i, k, i, owner[i] + 100]   # transcribed by its owner, from the audio only the owner loaded
            assert temp == round(0.2 * k, 4)
