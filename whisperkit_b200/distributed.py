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
    """rank `src` holds all_pcm [n_windows, stride] on `device`; every rank returns its contiguous shard."""
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_bounds(n_windows, world, rank)
    per = max(shard_bounds(n_windows, world, r)[1] - shard_bounds(n_windows, world, r)[0] for r in range(world))
    out = torch.empty(per, stride, dtype=torch.float32, device=device)
    chunks = None
    if rank == src:
        chunks = []
        for r in range(world):
            a, b = shard_bounds(n_windows, world, r)
            c = torch.zeros(per, stride, dtype=torch.float32, device=device)
            c[: b - a] = all_pcm[a:b]
            chunks.append(c)
    dist.scatter(out, chunks, src=src)
    return out[: hi - lo]


def pack_tokens(results, device) -> torch.Tensor:
    """results: list of objects with .tokens -> int32 [n, TOKEN_ROW] (row = [count, ids...])."""
    t = torch.zeros(len(results), TOKEN_ROW, dtype=torch.int32)
    for i, r in enumerate(results):
        toks = list(r.tokens)[: TOKEN_ROW - 2]
        t[i, 0] = len(toks)
        t[i, 1:1 + len(toks)] = torch.tensor(toks, dtype=torch.int32)
    return t.to(device)


def gather_tokens(local: torch.Tensor, n_windows: int, dst: int = 0) -> Optional[List[List[int]]]:
    """Gathers per-rank packed token rows on `dst`, restoring the global window order."""
    world, rank = dist.get_world_size(), dist.get_rank()
    per = max(shard_bounds(n_windows, world, r)[1] - shard_bounds(n_windows, world, r)[0] for r in range(world))
    pad = torch.zeros(per, TOKEN_ROW, dtype=torch.int32, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    out: List[List[int]] = []
    for r in range(world):
        a, b = shard_bounds(n_windows, world, r)
        rows = bufs[r][: b - a].cpu()
        for row in rows:
            n = int(row[0])
            out.append([int(v) for v in row[1:1 + n]])
    return out


def transcribe_sharded(all_pcm: Optional[torch.Tensor], n_windows: int, stride: int, device,
                       transcribe_local: Callable[[torch.Tensor], list]) -> Optional[List[List[int]]]:
    """scatter -> per-rank transcribe -> gather.  `transcribe_local(pcm_shard)` returns one result per window."""
    shard = scatter_windows(all_pcm, n_windows, stride, device)
    res = transcribe_local(shard)
    return gather_tokens(pack_tokens(res, device), n_windows)
