"""Edge-only collectives for the data-parallel hot path.

Windows are independent units (the reference fans them out as tasks, WhisperKit.swift:741-809), so the model needs no
collective.  Only the edges move data between ranks: rank 0 scatters PCM shards and gathers token IDs.  Works with
`nccl` (GPU tensors, NVLink) and `gloo` (CPU tensors; used by the CPU tests)."""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist

TOKEN_ROW = 228  # [n_tokens, tokens[226], pad]


def shard_bounds(n_windows: int, world: int, rank: int) -> Tuple[int, int]:
    """Static contiguous split, order preserving; the first (n % world) ranks take one extra window."""
    base, rem = divmod(n_windows, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def scatter_windows(all_pcm: Optional[torch.Tensor], n_windows: int, stride: int, device, src: int = 0) -> torch.Tensor:
    """rank `src` holds all_pcm [n_windows, stride] (pinned host memory or already on `device`); every rank returns its contiguous shard
    on `device`.  Point-to-point sends (NCCL: ncclSend / ncclRecv over NVLink), no padded copies: rank `src` stages and sends the other
    ranks' shards in rank order - the copy of shard r + 1 from the host runs while shard r is on the wire - and takes its own shard last."""
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_bounds(n_windows, world, rank)
    if rank != src:
        out = torch.empty(hi - lo, stride, dtype=torch.float32, device=device)
        if hi > lo:
            dist.recv(out, src=src)
        return out
    keep = []
    for r in range(world):
        if r == src:
            continue
        a, b = shard_bounds(n_windows, world, r)
        if b > a:
            part = all_pcm[a:b].to(device, non_blocking=True)
            keep.append(part)      # stays alive until the send has been enqueued behind the copy
            dist.send(part, dst=r)
    return all_pcm[lo:hi].to(device, non_blocking=True)


def pack_tokens(results, device) -> torch.Tensor:
    """results -> int32 [n, TOKEN_ROW] (row = [count, ids...]).  `results` is a ctypes array of wk_decode_result (its first 227 words are
    exactly that row) or a list of objects with .tokens."""
    import ctypes
    import numpy as np
    if isinstance(results, ctypes.Array):
        n = len(results)
        words = ctypes.sizeof(results) // 4 // max(n, 1)
        flat = np.frombuffer(results, dtype=np.int32).reshape(n, words)
        t = torch.zeros(n, TOKEN_ROW, dtype=torch.int32)
        t[:, :227] = torch.from_numpy(flat[:, :227].copy())
        return t.to(device)
    t = torch.zeros(len(results), TOKEN_ROW, dtype=torch.int32)
    for i, r in enumerate(results):
        toks = list(r.tokens)[: TOKEN_ROW - 2]
        t[i, 0] = len(toks)
        t[i, 1:1 + len(toks)] = torch.tensor(toks, dtype=torch.int32)
    return t.to(device)


def gather_tokens(local: torch.Tensor, n_windows: int, dst: int = 0) -> Optional[List[List[int]]]:
    """Gathers per-rank packed token rows on `dst` (ncclSend / ncclRecv straight into the rows of one buffer), restoring the global
    window order."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if rank != dst:
        if local.shape[0] > 0:
            dist.send(local.contiguous(), dst=dst)
        return None
    buf = torch.empty(n_windows, TOKEN_ROW, dtype=torch.int32, device=local.device)
    for r in range(world):
        a, b = shard_bounds(n_windows, world, r)
        if b <= a:
            continue
        if r == dst:
            buf[a:b] = local
        else:
            dist.recv(buf[a:b], src=r)
    rows = buf.cpu().numpy()
    return [row[1:1 + int(row[0])].tolist() for row in rows]


def transcribe_sharded(all_pcm: Optional[torch.Tensor], n_windows: int, stride: int, device,
                       transcribe_local: Callable[[torch.Tensor], list], stages: Optional[dict] = None) -> Optional[List[List[int]]]:
    """scatter -> per-rank transcribe -> gather.  `transcribe_local(pcm_shard)` returns one result per window.  `stages` (a dict)
    accumulates this rank's wall-clock milliseconds per stage - scatter (host-to-device copies + sends), compute, pack, gather (+ unpack
    on the destination) - so a bench can show where the end-to-end time of rank 0 goes."""
    import time

    def mark(name, t0):
        if stages is not None:
            if device.type == "cuda":
                torch.cuda.synchronize(device)
            stages[name] = stages.get(name, 0.0) + (time.perf_counter() - t0) * 1000.0
        return time.perf_counter()

    t = time.perf_counter()
    shard = scatter_windows(all_pcm, n_windows, stride, device)
    t = mark("scatter", t)
    res = transcribe_local(shard)
    t = mark("compute", t)
    packed = pack_tokens(res, device)
    t = mark("pack", t)
    out = gather_tokens(packed, n_windows)
    mark("gather_unpack", t)
    if stages is not None:
        stages["_calls"] = stages.get("_calls", 0) + 1
    return out


# ---------------------------------------------------------------------------------------------------------------------------------
# Long-form streams: whole audio streams are the independent units (each advances by its own seek
# loop), so ranks take whole streams - no collective inside the path, only the edges move data.
# ---------------------------------------------------------------------------------------------------------------------------------
def assign_streams(n_samples: List[int], world: int) -> List[List[int]]:
    """Longest-processing-time-first: streams sorted by length (ties by index) go to the currently lightest rank (ties by rank).
    Deterministic, so every rank derives the same assignment without communication.  Returns stream indices per rank, ascending."""
    loads = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in sorted(range(len(n_samples)), key=lambda k: (-int(n_samples[k]), k)):
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += max(int(n_samples[i]), 1)
    return [sorted(v) for v in out]


SEG_HEAD = 8  # [stream, id, seek_lo, seek_hi, start(f32 bits), end(f32 bits), n_tokens, temperature(f32 bits)]


def pack_segments(per_stream: dict, device) -> torch.Tensor:
    """per_stream {stream index: [segment objects with .id .seek .start .end .tokens .temperature]} -> one int32 row list:
    every segment is SEG_HEAD header words followed by its tokens; row 0 holds the total length."""
    words: List[int] = [0]
    f2i = lambda v: int(torch.tensor(float(v), dtype=torch.float32).view(torch.int32).item())  # noqa: E731
    for s in sorted(per_stream):
        for g in per_stream[s]:
            seek = int(g.seek)
            words += [int(s), int(g.id), seek & 0x7FFFFFFF, seek >> 31, f2i(g.start), f2i(g.end), len(g.tokens), f2i(getattr(g, "temperature", 0.0))]
            words += [int(t) for t in g.tokens]
    words[0] = len(words)
    return torch.tensor(words, dtype=torch.int32, device=device)


def unpack_segments(buf: torch.Tensor) -> dict:
    from types import SimpleNamespace
    w = buf.cpu()
    n = int(w[0])
    i2f = lambda v: float(torch.tensor(int(v), dtype=torch.int32).view(torch.float32).item())  # noqa: E731
    out: dict = {}
    i = 1
    while i < n:
        s, sid, lo, hi, st, en, nt, tp = (int(v) for v in w[i:i + SEG_HEAD])
        toks = [int(v) for v in w[i + SEG_HEAD:i + SEG_HEAD + nt]]
        out.setdefault(s, []).append(SimpleNamespace(stream=s, id=sid, seek=(hi << 31) | lo, start=i2f(st), end=i2f(en), tokens=toks, temperature=i2f(tp)))
        i += SEG_HEAD + nt
    return out


def transcribe_streams_sharded(audio_arrays: list, device, transcribe_local: Callable[[list, List[int]], list], dst: int = 0) -> Optional[list]:
    """Every rank holds (or can load) the audio of its own streams: `audio_arrays[i]` may be None on ranks that do not own stream i, but
    the LENGTHS must be known everywhere (they decide the assignment).  `transcribe_local(arrays, stream_ids)` -> one segment list per
    local stream (e.g. whisperkit_b200.longform.transcribe_streams).  Rank `dst` returns the segment lists in global stream order."""
    world, rank = dist.get_world_size(), dist.get_rank()
    lengths = torch.tensor([0 if a is None else len(a) for a in audio_arrays], dtype=torch.int64, device=device)
    dist.all_reduce(lengths, op=dist.ReduceOp.MAX)          # the only collective before the work: 8 bytes per stream
    mine = assign_streams([int(v) for v in lengths.cpu()], world)[rank]
    local = transcribe_local([audio_arrays[i] for i in mine], mine) if mine else []
    packed = pack_segments({i: segs for i, segs in zip(mine, local)}, device)
    size = torch.tensor([packed.numel()], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(size) for _ in range(world)]
    dist.all_gather(sizes, size)
    cap = max(int(s.item()) for s in sizes)
    pad = torch.zeros(cap, dtype=torch.int32, device=device)
    pad[: packed.numel()] = packed
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    merged: dict = {}
    for b in bufs:
        merged.update(unpack_segments(b))
    return [merged.get(i, []) for i in range(len(audio_arrays))]


# ---------------------------------------------------------------------------------------------------------------------------------
# The same edges through the library's own C ABI (csrc/comm.cu: ncclSend / ncclRecv issued by libwkb200, no torch tensors in the data
# path) - what a Swift host would call.  torch.distributed is only the rendezvous that carries the 128-byte NCCL id to the other ranks.
# ---------------------------------------------------------------------------------------------------------------------------------
class Comm:
    def __init__(self, lib, rank: int, world: int, device_index: int):
        import ctypes as C
        self.lib, self.rank, self.world = lib, rank, world
        ident = (C.c_uint8 * 128)()
        if rank == 0:
            from ._lib import check
            check(lib.wk_comm_unique_id(ident))
        t = torch.tensor(list(ident), dtype=torch.uint8, device=torch.device("cuda", device_index))
        dist.broadcast(t, src=0)
        ident = (C.c_uint8 * 128)(*[int(v) for v in t.cpu()])
        self.handle = C.c_void_p()
        from ._lib import check
        check(lib.wk_comm_create(ident, rank, world, device_index, C.byref(self.handle)))

    def shard_bounds(self, n_windows: int, rank: Optional[int] = None) -> Tuple[int, int]:
        return shard_bounds(n_windows, self.world, self.rank if rank is None else rank)

    def scatter_windows(self, all_pcm_ptr: Optional[int], n_windows: int, stride: int, shard_ptr: int, root: int = 0) -> int:
        """all_pcm_ptr: address of [n_windows][stride] f32 on the root (host pinned or device), ignored elsewhere; shard_ptr: device buffer."""
        import ctypes as C
        from ._lib import check
        n_local = C.c_int64()
        check(self.lib.wk_comm_scatter_windows(self.handle, C.c_void_p(all_pcm_ptr or 0), n_windows, stride, root, C.c_void_p(shard_ptr), C.byref(n_local)))
        return int(n_local.value)

    def gather_results(self, local_results, n_local: int, n_windows: int, all_results=None, root: int = 0) -> None:
        from ._lib import check
        check(self.lib.wk_comm_gather_results(self.handle, local_results, n_local, n_windows, root, all_results))

    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.wk_comm_free(self.handle)
            self.handle = None
