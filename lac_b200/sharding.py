"""Multi-GPU layout: independent chunks shard across ranks with no collective inside the coding
loop; the only exchange is the gather of per-chunk token / bit counts and, for writing one file,
of the compressed bytes (a few bits per token).  Works on NCCL (GPU tensors) and gloo (CPU tensors).

Chunks are coded in model batches of exactly `batch_streams` streams (the last batch padded with empty
streams): the predictor then sees the same batch shape whether a file is written by 1 GPU and read by 8 or the
other way round, which is what makes the logits -- and with them the decode -- bit-reproducible.  Batches, not
chunks, are what is dealt to the ranks: rank r owns a contiguous range of batches.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import container


def chunk_range(n_chunks: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of the items rank owns (first n % world ranks get one more)."""
    base, extra = divmod(n_chunks, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def batch_spans(n_chunks: int, batch_streams: int, world: int) -> List[Tuple[int, int]]:
    """Per rank the [begin, end) CHUNK range it owns: whole batches, dealt contiguously."""
    n_batches = -(-n_chunks // batch_streams) if n_chunks else 0
    spans = []
    for r in range(world):
        b, e = chunk_range(n_batches, r, world)
        spans.append((min(n_chunks, b * batch_streams), min(n_chunks, e * batch_streams)))
    return spans


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def gather_index(ntok: torch.Tensor, nbits: torch.Tensor, n_chunks: int, group=None,
                 spans: Optional[Sequence[Tuple[int, int]]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All ranks get the global (ntok, nbits) int64 [n_chunks] from their local shards (rank order = chunk order).
    spans: the chunk range of every rank (default: chunk_range)."""
    rank, world = _world(group)
    if spans is None:
        spans = [chunk_range(n_chunks, r, world) for r in range(world)]
    if world == 1:
        return ntok.to(torch.int64), nbits.to(torch.int64)
    width = max(max(e - b for b, e in spans), 1)
    local = torch.zeros((2, width), dtype=torch.int64, device=ntok.device)
    local[0, : ntok.numel()] = ntok.to(torch.int64)
    local[1, : nbits.numel()] = nbits.to(torch.int64)
    everyone = torch.zeros((world, 2, width), dtype=torch.int64, device=ntok.device)
    dist.all_gather_into_tensor(everyone.view(-1), local.view(-1), group=group)
    g_ntok = [everyone[r, 0, : e - b] for r, (b, e) in enumerate(spans)]
    g_nbits = [everyone[r, 1, : e - b] for r, (b, e) in enumerate(spans)]
    return torch.cat(g_ntok), torch.cat(g_nbits)


def gather_payload(local_bytes: torch.Tensor, g_nbits: torch.Tensor, n_chunks: int, dst: int = 0, group=None,
                   spans: Optional[Sequence[Tuple[int, int]]] = None) -> Optional[bytes]:
    """Concatenate every rank's (already concatenated) chunk bytes on rank dst; None elsewhere."""
    rank, world = _world(group)
    if spans is None:
        spans = [chunk_range(n_chunks, r, world) for r in range(world)]
    if world == 1:
        return local_bytes.cpu().numpy().tobytes()
    sizes = [int(((g_nbits[b:e] + 7) // 8).sum().item()) for b, e in spans]
    width = max(max(sizes), 1)
    padded = torch.zeros(width, dtype=torch.uint8, device=local_bytes.device)
    padded[: local_bytes.numel()] = local_bytes
    everyone = torch.zeros((world, width), dtype=torch.uint8, device=local_bytes.device)
    dist.all_gather_into_tensor(everyone.view(-1), padded, group=group)
    if rank != dst:
        return None
    host = everyone.cpu().numpy()
    return b"".join(host[r, : sizes[r]].tobytes() for r in range(world))


def gather_tokens(local: torch.Tensor, n_chunks: int, chunk_tokens: int, spans: Sequence[Tuple[int, int]],
                  dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Decoded tokens int32 [local chunks, chunk_tokens] of every rank -> [n_chunks, chunk_tokens] on rank dst."""
    rank, world = _world(group)
    if world == 1:
        return local
    width = max(max(e - b for b, e in spans), 1)
    padded = torch.zeros((width, chunk_tokens), dtype=torch.int32, device=local.device)
    padded[: local.shape[0]] = local
    everyone = torch.zeros((world, width, chunk_tokens), dtype=torch.int32, device=local.device)
    dist.all_gather_into_tensor(everyone.view(-1), padded.view(-1), group=group)
    if rank != dst:
        return None
    return torch.cat([everyone[r, : e - b] for r, (b, e) in enumerate(spans)])


def concat_streams(streams: Sequence[bytes], device) -> torch.Tensor:
    buf = np.frombuffer(b"".join(streams), dtype=np.uint8).copy() if streams else np.zeros(0, dtype=np.uint8)
    return torch.from_numpy(buf).to(device)


# ------------------------------------------------------------------ the sharded job (host logic only)
# encode_batch(tokens int32 [B, chunk_tokens], ntok int32 [B]) -> (streams: list of B bytes, nbits: list of B ints)
# decode_batch(streams: list of B bytes, ntok int32 [B])       -> tokens int32 [B, chunk_tokens] (numpy)
EncodeBatch = Callable[[np.ndarray, np.ndarray], Tuple[List[bytes], Sequence[int]]]
DecodeBatch = Callable[[List[bytes], np.ndarray], np.ndarray]


def split_chunks(tokens, chunk_tokens: int) -> Tuple[np.ndarray, np.ndarray]:
    toks = np.ascontiguousarray(tokens, dtype=np.int32)
    n_chunks = (len(toks) + chunk_tokens - 1) // chunk_tokens
    padded = np.zeros(n_chunks * chunk_tokens, dtype=np.int32)
    padded[: len(toks)] = toks
    ntok = np.full(n_chunks, chunk_tokens, dtype=np.int32)
    if n_chunks:
        ntok[-1] = len(toks) - (n_chunks - 1) * chunk_tokens
    return padded.reshape(n_chunks, chunk_tokens), ntok


def compress_sharded(tokens, chunk_tokens: int, batch_streams: int, encode_batch: EncodeBatch, prec: int, vocab: int,
                     device, tag: int = 0, group=None, quantiser: int = container.QUANT_LQ32) -> Optional[bytes]:
    """Every rank passes the SAME token array; rank r codes its batches; rank 0 returns the LACB file (None
    elsewhere).  One gather of the index and one of the payload, nothing inside the coding loop."""
    rank, world = _world(group)
    chunks, ntok = split_chunks(tokens, chunk_tokens)
    n_chunks = len(ntok)
    spans = batch_spans(n_chunks, batch_streams, world)
    b, e = spans[rank]
    streams: List[bytes] = []
    nbits: List[int] = []
    for c0 in range(b, e, batch_streams):
        c1 = min(e, c0 + batch_streams)
        bt = np.zeros((batch_streams, chunk_tokens), dtype=np.int32)   # padded to the full batch shape
        bn = np.zeros(batch_streams, dtype=np.int32)
        bt[: c1 - c0] = chunks[c0:c1]
        bn[: c1 - c0] = ntok[c0:c1]
        s, nb = encode_batch(bt, bn)
        streams += list(s[: c1 - c0])
        nbits += [int(x) for x in nb[: c1 - c0]]
    l_ntok = torch.from_numpy(ntok[b:e].astype(np.int64)).to(device)
    l_nbits = torch.tensor(nbits, dtype=torch.int64, device=device)
    g_ntok, g_nbits = gather_index(l_ntok, l_nbits, n_chunks, group, spans)
    payload = gather_payload(concat_streams(streams, device), g_nbits, n_chunks, 0, group, spans)
    if rank != 0:
        return None
    return container.pack_payload(payload, g_ntok.cpu().numpy(), g_nbits.cpu().numpy(), prec, vocab, chunk_tokens,
                                  quantiser, batch_streams, tag)


def decompress_sharded(blob: bytes, decode_batch: DecodeBatch, device, group=None) -> Optional[np.ndarray]:
    """Every rank passes the SAME file; rank 0 returns the tokens (None elsewhere)."""
    rank, world = _world(group)
    c = container.unpack(blob)
    B = c.batch_streams or max(c.n_chunks, 1)
    spans = batch_spans(c.n_chunks, B, world)
    b, e = spans[rank]
    all_streams = c.streams()
    local = np.zeros((e - b, c.chunk_tokens), dtype=np.int32)
    for c0 in range(b, e, B):
        c1 = min(e, c0 + B)
        bs = all_streams[c0:c1] + [b""] * (B - (c1 - c0))
        bn = np.zeros(B, dtype=np.int32)
        bn[: c1 - c0] = c.ntok[c0:c1]
        local[c0 - b:c1 - b] = decode_batch(bs, bn)[: c1 - c0]
    got = gather_tokens(torch.from_numpy(local).to(device), c.n_chunks, c.chunk_tokens, spans, 0, group)
    if rank != 0:
        return None
    got = got.cpu().numpy()
    flat = [got[i, : int(c.ntok[i])] for i in range(c.n_chunks)]
    return np.concatenate(flat) if flat else np.zeros(0, dtype=np.int32)
