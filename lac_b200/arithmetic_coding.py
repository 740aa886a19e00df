"""Host-side mirror of the reference's arithmetic_coding.py (ACSampler / Region / CarryBuffer / packbits /
unpackbits), with the coding done by the CUDA library.

The reference ACSampler is driven one `sample(pdf)` / `sample_scaled_cdf(cdf)` call at a time through callbacks
(arithmetic_coding.py:9-124).  The mirror keeps that protocol:

    compress:  sampler.compress_tokens = iter(tokens); sampler.compress_output = bit_callback
               sampler.on_compress_done = ...;  while not sampler.compress_done: sampler.sample(pdf)
               (the reference's own compress_base_ten / to_bin loops, :234-266 / :306-321, run unchanged)
    expand:    sampler.decompress_bits = bits;  sampler.decompress_output = token_callback; ... sample(pdf)

and adds whole-sequence calls (compress(cdfs, tokens) / expand(cdfs, data, n): one GPU call each).

Where the work happens.  Region (low, high) lives in a lac_enc_state / lac_dec_state ON THE DEVICE and every token
is narrowed and renormalised by a kernel (lac_acs_encode_tables / lac_acs_decode_tables with T = 1).  The host
reads the state back and hands the kernel's output to the callbacks:
  * compress: the stream bytes on the device are carry-resolved, which is what CarryBuffer does (:180-208); the
    bits not yet handed to compress_output are handed over whenever Region.definite holds after a token (the
    reference checks after every bit, so the same bits arrive in the same order, at most one token later);
  * expand: (d_bits, d_bits_ulp), the window of code values the bits read so far allow (:99-109), is the bit
    READER's state and is kept on the host; the device decodes the token at both ends of the window (two decoder
    states sharing the region) and another bit is pulled from decompress_bits while they disagree.

`scaled_cdf(pdf)` is the reference's own float64 table construction (:57-62), kept on the host because it is the
reference's quantisation, not ours; the LLM path uses LQ32 on the GPU.

Decoder semantics: the token returned is the one whose encoder interval contains the code
value.  The reference's lookup (bisect_left with key=Region.map, :98) disagrees with its own encoder at interval
boundaries and round-trips only part of its streams; the mirror decodes what the encoder coded.
"""
from __future__ import annotations

import math
from typing import Iterable, Iterator, List, Optional, Sequence

import numpy as np


class packbits:
    """arithmetic_coding.py:212-225 (MSB first; flush() zero-pads the last byte)."""

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
    """arithmetic_coding.py:227-230"""
    for byte in byte_generator:
        for b in range(8):
            yield (byte >> (7 - b)) & 1


class Region:
    """Read-only view of the coder interval (arithmetic_coding.py:128-178); the interval itself is device state,
    refreshed by the sampler after every kernel call."""

    def __init__(self, precision: int = 48):
        self.precision = precision
        self.reset()

    def __repr__(self):
        return f"Region(prec={self.precision},[{self.low / self.one} {(self.high + 1) / self.one}])"

    def reset(self):
        self.low = 0
        self.high = self.one - 1

    @property
    def one(self):
        return 1 << self.precision

    @property
    def span(self):
        return self.high - self.low + 1

    @property
    def entropy(self):
        return self.precision - math.log2(self.span)

    def map(self, v, d=None):
        d = d if d is not None else self.one
        return self.low + (self.span * v) // d

    def unmap(self, v, d=None):
        d = d if d is not None else self.one
        return (v - self.low) * d // self.span

    def entropy_of(self, l, h, d=None):
        return math.log2(self.span) - math.log2(self.map(h, d) - self.map(l, d))

    @property
    def definite(self):
        return self.high < self.one


class CarryBuffer:
    """Bits emitted by the device coder but not yet handed to compress_output (arithmetic_coding.py:180-208): buf is
    their value, bits their count.  The carries themselves are resolved in the device stream."""

    def __init__(self):
        self.reset()

    def __repr__(self):
        return f"CarryBuffer({bin(self.buf | (1 << self.bits))[3:]})"

    def reset(self):
        self.buf = 0
        self.bits = 0


class ACSampler:
    def __init__(self, precision: int = 48):
        self.precision = precision
        self.region = Region(precision)
        self.accumulator = CarryBuffer()
        self.compress_tokens = None
        self.compress_output = None
        self.decompress_bits = iter(())
        self.decompress_output = None
        self.bits_per_token = None
        self.on_decompress_done = None
        self.on_compress_done = None
        self._enc = None
        self._dec = None
        self._cap = 1 << 12
        self.reset()

    def __repr__(self):
        mode = "compressing" if self.compress_tokens else "expanding"
        return f"ACSampler({mode},{self.region},{self.accumulator})"

    # ---- the reference's properties (:31-44)
    @property
    def one(self) -> int:
        return 1 << self.precision

    @property
    def decompress_bits(self):
        return self._decompress_bits

    @decompress_bits.setter
    def decompress_bits(self, bits):
        self._decompress_bits = iter(bits) if bits is not None else iter(())
        self.decompress_done = False

    @property
    def compress_tokens(self):
        return self._compress_tokens

    @compress_tokens.setter
    def compress_tokens(self, toks):
        self._compress_tokens = iter(toks) if toks is not None else None
        self.compress_done = False

    def reset(self):
        """arithmetic_coding.py:45-49"""
        self.region.reset()
        self.accumulator.reset()
        self.d_bits = 0
        self.d_bits_ulp = self.region.one
        self._delivered = 0
        if self._enc is not None:
            self._enc.reset()

    # ---- table construction: the reference's expressions (:57-72), host side
    def get_lop_bias(self, pdf):
        return sum(pdf) / (self.one / 2 - len(pdf))

    def scaled_cdf(self, pdf) -> np.ndarray:
        pdf = np.array(pdf, dtype=np.float64)
        pdf += self.get_lop_bias(pdf)
        pdf *= self.one / np.sum(pdf)
        return np.cumsum(pdf).astype(np.uint64)

    def sample(self, pdf):
        return self.sample_scaled_cdf(self.scaled_cdf(pdf))

    # ---- device plumbing
    def _encoder(self):
        if self._enc is None:
            from . import coder
            self._enc = coder.StreamEncoder(1, prec=self.precision, capacity_bytes=self._cap)
        if (self._enc_bits() + 4 * self.precision + 80) // 8 >= self._cap:
            from . import coder
            old, self._cap = self._enc, self._cap * 2
            self._enc = coder.StreamEncoder(1, prec=self.precision, capacity_bytes=self._cap)
            self._enc.state.copy_(old.state)
            self._enc.out[:, : old.cap].copy_(old.out)
        return self._enc

    def _enc_bits(self) -> int:
        return self.accumulator.bits + self._delivered

    def _read_enc(self):
        st = self._enc.state.cpu().numpy().view(np.int64).reshape(-1)
        status = int(st[3]) & 0xFFFFFFFF
        if status & 2:
            raise IndexError("token outside the cdf")
        if status & 4:
            raise AssertionError("cdf has unencodable token (pdf = 0). Perhaps try using get_lop_bias or adding "
                                 "an arange to the cdf.")
        self.region.low, self.region.high = int(st[0]), int(st[1])
        return int(st[2])

    def _deliver(self, nbits: int, force: bool):
        """Hand the device stream's bits [delivered, nbits) to compress_output once no carry can reach them."""
        pending = nbits - self._delivered
        if pending and (force or self.region.definite):
            first = self._delivered // 8
            data = bytes(self._enc.out[0, first:(nbits + 7) // 8].cpu().numpy())
            bits = np.unpackbits(np.frombuffer(data, dtype=np.uint8))[self._delivered - 8 * first: nbits - 8 * first]
            self._delivered = nbits
            pending = 0
            if self.compress_output:
                for b in bits.tolist():
                    self.compress_output(b)
        self.accumulator.bits = pending
        if pending:
            first = self._delivered // 8
            data = bytes(self._enc.out[0, first:(nbits + 7) // 8].cpu().numpy())
            v = int.from_bytes(data, "big") >> ((8 - nbits % 8) % 8)
            self.accumulator.buf = v & ((1 << pending) - 1)
        else:
            self.accumulator.buf = 0

    def flush_compress(self):
        """arithmetic_coding.py:50-56: the middle-third step, drain the carry buffer, reset the region."""
        enc = self._encoder()
        enc.acs_flush()
        nbits = self._read_enc()
        self._deliver(nbits, force=True)
        self.region.reset()

    def sample_scaled_cdf(self, cdf):
        """One token through the coder (arithmetic_coding.py:73-124): compress mode takes it from compress_tokens,
        expand mode decodes it from decompress_bits."""
        import torch
        cdf = np.ascontiguousarray(cdf, dtype=np.uint64)
        if self.compress_tokens:
            try:
                tok = next(self.compress_tokens)
            except StopIteration:
                self.compress_done = True
                if self.on_compress_done:
                    self.on_compress_done()
                tok = 0
            enc = self._encoder()
            span_before = self.region.span
            before_bits = self._enc_bits()
            d_cdf = torch.from_numpy(cdf.view(np.int64)).to(enc.device)
            enc.acs_encode_tables(d_cdf, torch.tensor([[int(tok)]], dtype=torch.int32, device=enc.device))
            nbits = self._read_enc()
            if self.bits_per_token:
                k = nbits - before_bits
                self.bits_per_token(math.log2(span_before) - math.log2(self.region.span) + k)
            self._deliver(nbits, force=False)
            return tok
        # ---- expand
        from . import coder
        if self._dec is None:
            self._dec = coder.StreamDecoder([b"\x00" * 32, b"\xff" * 32], prec=self.precision)
        dec = self._dec
        d_cdf = torch.from_numpy(cdf.view(np.int64)).to(dec.device)
        span_before = self.region.span
        while True:
            st = np.zeros((2, 5), dtype=np.int64)
            st[:, 0], st[:, 1] = self.region.low, self.region.high
            st[0, 2], st[1, 2] = self.d_bits, self.d_bits + max(self.d_bits_ulp, 1) - 1
            st[:, 3] = self.precision
            dec.state.copy_(torch.from_numpy(st.view(np.uint8).reshape(2, -1)))
            toks = dec.acs_decode_tables(d_cdf, 1, check_status=False).cpu().numpy()[:, 0]
            if int(toks[0]) == int(toks[1]):
                break
            try:
                bit = next(self.decompress_bits)
            except StopIteration:
                self.decompress_done = True
                if self.on_decompress_done:
                    self.on_decompress_done()
                bit = 0
            self.d_bits_ulp >>= 1
            self.d_bits += bit * self.d_bits_ulp
        out = dec.state.cpu().numpy().view(np.int64).reshape(2, 5)
        if (int(out[0, 4]) | int(out[1, 4])) & 0xFFFFFFFF & 4:
            raise IndexError("code value outside the cdf")
        tok = int(toks[0])
        self.region.low, self.region.high = int(out[0, 0]), int(out[0, 1])
        self.d_bits = int(out[0, 2])
        self.d_bits_ulp = int(out[1, 2]) - int(out[0, 2]) + 1
        if self.bits_per_token:
            k = int(out[0, 3]) - self.precision
            self.bits_per_token(math.log2(span_before) - math.log2(self.region.span) + k)
        if self.decompress_output:
            self.decompress_output(tok)
        return tok

    # ---- whole sequences in one GPU call
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
