"""Multi-GPU layout: independent chunks shard across ranks with no collective inside the coding
loop; the only exchange is the gather of per-chunk token / bit counts and, for writing one file,
of the compressed bytes (a few bits per token).  Works on NCCL (GPU tensors) and gloo (CPU tensors).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def chunk_range(n_chunks: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of the chunks rank owns (first n % world ranks get one more)."""
    base, extra = divmod(n_chunks, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_index(ntok: torch.Tensor, nbits: torch.Tensor, n_chunks: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All ranks get the global (ntok, nbits) int64 [n_chunks] from their local shards (rank order = chunk order)."""
    world = dist.get_world_size(group)
    width = -(-n_chunks // world)
    local = torch.zeros((2, width), dtype=torch.int64, device=ntok.device)
    local[0, : ntok.numel()] = ntok.to(torch.int64)
    local[1, : nbits.numel()] = nbits.to(torch.int64)
    everyone = torch.zeros((world, 2, width), dtype=torch.int64, device=ntok.device)
    dist.all_gather_into_tensor(everyone.view(-1), local.view(-1), group=group)
    g_ntok, g_nbits = [], []
    for r in range(world):
        b, e = chunk_range(n_chunks, r, world)
        g_ntok.append(everyone[r, 0, : e - b])
        g_nbits.append(everyone[r, 1, : e - b])
    return torch.cat(g_ntok), torch.cat(g_nbits)


def gather_payload(local_bytes: torch.Tensor, g_nbits: torch.Tensor, n_chunks: int, dst: int = 0, group=None) -> Optional[bytes]:
    """Concatenate every rank's (already concatenated) chunk bytes on rank dst; None elsewhere."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = []
    for r in range(world):
        b, e = chunk_range(n_chunks, r, world)
        sizes.append(int(((g_nbits[b:e] + 7) // 8).sum().item()))
    width = max(max(sizes), 1)
    padded = torch.zeros(width, dtype=torch.uint8, device=local_bytes.device)
    padded[: local_bytes.numel()] = local_bytes
    everyone = torch.zeros((world, width), dtype=torch.uint8, device=local_bytes.device)
    dist.all_gather_into_tensor(everyone.view(-1), padded, group=group)
    if rank != dst:
        return None
    host = everyone.cpu().numpy()
    return b"".join(host[r, : sizes[r]].tobytes() for r in range(world))


def concat_streams(streams: Sequence[bytes], device) -> torch.Tensor:
    buf = np.frombuffer(b"".join(streams), dtype=np.uint8).copy() if streams else np.zeros(0, dtype=np.uint8)
    return torch.from_numpy(buf).to(device)
