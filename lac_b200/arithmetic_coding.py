"""Host-side mirror of the reference's arithmetic_coding.py (ACSampler / packbits / unpackbits), with
the coding done by the CUDA library.

The reference ACSampler is driven one `sample(pdf)` call at a time through callbacks
(arithmetic_coding.py:59-127).  Here the same coder is exposed over whole token sequences:

    s = ACSampler(precision=48)
    bits = s.compress(cdfs, tokens)                 # == everything the reference hands to compress_output,
                                                    #    flush_compress() included (bit-exact)
    toks = s.expand(cdfs, data_bytes, n)            # value-based decoder

`scaled_cdf(pdf)` is the reference's own float64 table construction (arithmetic_coding.py:59-72), kept
on the host because it is the reference's quantisation, not ours; the LLM path uses LQ32 on the GPU.
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Sequence

import numpy as np


class packbits:
    """arithmetic_coding.py:200-214 (MSB first; flush() zero-pads the last byte)."""

    def __init__(self, byte_callback):
        self.state = 1
        self.byte_callback = byte_callback

    def __call__(self, bit):
        self.state = (self.state << 1) | bit
        if self.state >> 8:
            self.byte_callback(self.state & 255)
            self.state >>= 8

    def flush(self):
        while self.state > 1:
            self(0)


def unpackbits(byte_generator: Iterable[int]) -> Iterator[int]:
    for byte in byte_generator:
        for b in range(8):
            yield (byte >> (7 - b)) & 1


class ACSampler:
    def __init__(self, precision: int = 48):
        self.precision = precision

    @property
    def one(self) -> int:
        return 1 << self.precision

    def get_lop_bias(self, pdf):
        return sum(pdf) / (self.one / 2 - len(pdf))          # arithmetic_coding.py:65-72

    def scaled_cdf(self, pdf) -> np.ndarray:
        pdf = np.array(pdf, dtype=np.float64)                 # arithmetic_coding.py:59-63
        pdf += self.get_lop_bias(pdf)
        pdf *= self.one / np.sum(pdf)
        return np.cumsum(pdf).astype(np.uint64)

    def compress(self, cdfs, tokens: Sequence[int], flush=True) -> bytes:
        """cdfs: uint64 inclusive cumulative tables [T, V] (or [V] shared).  flush=True is the reference's
        flush_compress (bit-exact, tail may be undecodable), flush="safe" always round-trips."""
        import torch
        from . import coder
        cdfs = np.ascontiguousarray(cdfs, dtype=np.uint64)
        toks = np.ascontiguousarray(tokens, dtype=np.int32)[None]
        enc = coder.StreamEncoder(1, prec=self.precision, capacity_bytes=toks.shape[1] * 8 + 64)
        enc.acs_encode_tables(torch.from_numpy(cdfs.view(np.int64)).cuda(), torch.from_numpy(toks).cuda(), finish=flush)
        streams, nbits = enc.bitstreams()
        self.last_nbits = int(nbits[0])
        return streams[0]

    def expand(self, cdfs, data: bytes, n: int) -> List[int]:
        import torch
        from . import coder
        cdfs = np.ascontiguousarray(cdfs, dtype=np.uint64)
        dec = coder.StreamDecoder([bytes(data)], prec=self.precision)
        return dec.acs_decode_tables(torch.from_numpy(cdfs.view(np.int64)).cuda(), n).cpu().numpy()[0].astype(int).tolist()
