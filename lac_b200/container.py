"""Chunk / stream container ("LACB" v1): the framing the reference lacks.

The reference has no file format: A_from_bin keeps emitting symbols while its bit window allows
(arith_code.py:336-340) and ACSampler's demo loops until the bits run out
(arithmetic_coding.py:266-281), so token counts must travel next to the bitstreams.  Text is cut
into independent chunks (streams); each chunk is one coder stream with its own flush.

Layout (little endian):
    0   4  magic  b"LACB"
    4   2  version (1)
    6   1  precision of the coder (AC(..., prec))
    7   1  quantiser id (1 = LQ32 tables on a fixed total 2^32, 0 = caller-supplied tables)
    8   4  vocabulary size
    12  4  nominal tokens per chunk
    16  8  number of chunks N
    24  8  total tokens
    32  N x (uint32 tokens in chunk, uint32 bits in chunk)      -- the index
    ..  chunk payloads, each byte aligned ((bits + 7) // 8 bytes), in chunk order
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

MAGIC = b"LACB"
VERSION = 1
HEADER = struct.Struct("<4sHBBIIQQ")
QUANT_LQ32 = 1


@dataclass
class Container:
    prec: int
    quantiser: int
    vocab: int
    chunk_tokens: int
    ntok: np.ndarray      # uint32 [N]
    nbits: np.ndarray     # uint32 [N]
    payload: bytes        # concatenated chunk bytes

    @property
    def n_chunks(self) -> int:
        return len(self.ntok)

    def offsets(self) -> np.ndarray:
        """Byte offsets of the chunks inside payload, int64 [N + 1]."""
        off = np.zeros(self.n_chunks + 1, dtype=np.int64)
        np.cumsum((self.nbits.astype(np.int64) + 7) // 8, out=off[1:])
        return off

    def streams(self) -> List[bytes]:
        off = self.offsets()
        return [self.payload[off[i]:off[i + 1]] for i in range(self.n_chunks)]


def pack(streams: Sequence[bytes], ntok: Sequence[int], nbits: Sequence[int], prec: int, vocab: int,
         chunk_tokens: int, quantiser: int = QUANT_LQ32) -> bytes:
    ntok = np.asarray(ntok, dtype=np.uint32)
    nbits = np.asarray(nbits, dtype=np.uint32)
    if not (len(streams) == len(ntok) == len(nbits)):
        raise ValueError("streams, ntok and nbits must have the same length")
    for i, s in enumerate(streams):
        if len(s) != (int(nbits[i]) + 7) // 8:
            raise ValueError(f"chunk {i}: {len(s)} bytes for {int(nbits[i])} bits")
    head = HEADER.pack(MAGIC, VERSION, prec, quantiser, vocab, chunk_tokens, len(streams), int(ntok.sum(dtype=np.uint64)))
    index = np.stack([ntok, nbits], axis=1).astype("<u4").tobytes()
    return head + index + b"".join(streams)


def unpack(blob: bytes) -> Container:
    if len(blob) < HEADER.size:
        raise ValueError("truncated container")
    magic, ver, prec, quant, vocab, chunk_tokens, n, total = HEADER.unpack_from(blob, 0)
    if magic != MAGIC or ver != VERSION:
        raise ValueError("not a LACB v1 container")
    idx_end = HEADER.size + 8 * n
    if len(blob) < idx_end:
        raise ValueError("truncated index")
    index = np.frombuffer(blob, dtype="<u4", count=2 * n, offset=HEADER.size).reshape(n, 2)
    c = Container(prec, quant, vocab, chunk_tokens, index[:, 0].copy(), index[:, 1].copy(), bytes(blob[idx_end:]))
    if int(c.ntok.sum(dtype=np.uint64)) != total:
        raise ValueError("index does not add up to the total token count")
    if len(c.payload) != int(c.offsets()[-1]):
        raise ValueError("payload length does not match the index")
    return c
