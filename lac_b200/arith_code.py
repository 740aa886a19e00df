"""Host-side mirror of the reference's arith_code.py interface (same class and method names, same
argument meaning), with the coding done by the CUDA library through the C ABI.

What maps to what (pramasoul/lac arith_code.py):
    Predictor / CDFPredictor / ProbPredictor   :63-135   table providers: .dist, .minp, .accept(), .copy()
    AC(predictor, prec).to_bin / .from_bin     :137-146
    A_to_bin.run / bits / encode / __call__    :147-231  -> lac_ac_encode_tables (one GPU call per run)
    A_from_bin.run / decode                    :233-345  -> lac_ac_decode_tables
    group_bits / ungroup_bits                  :347-362  byte layout of every stream

The uniform base class Predictor(n) (floor-mapped ranges, the default AC() = AC(Predictor(3), 16)) goes through
lac_ac_encode_uniform / lac_ac_decode_uniform and is bit-exact with the reference as well.

Differences that cannot be avoided, both documented in DESIGN.md:
  * the reference decoder has no length framing (it emits symbols while its bit window allows and
    then guesses in flush()); here run()/decode() take the number of symbols to produce;
  * the per-symbol Python methods symbol_to_range / val_to_symbol are not part of this mirror: they
    are what the kernels implement (table_range / table_symbol in csrc/coder_kernels.cu).
There is no CPU coding path in this module: without the CUDA library every coding call raises.
"""
from __future__ import annotations

import itertools
from typing import Iterable, Iterator, List, Optional, Sequence

import numpy as np


# ------------------------------------------------------------------ predictors (table providers)
class Predictor:
    """Uniform predictor over n symbols (arith_code.py:63-74): symbol s maps to
    [floor(s w / n), floor((s + 1) w / n)).  Coded by the dedicated uniform kernels, bit-exact with the reference
    (tests/golden/ac_uniform.npz)."""

    def __init__(self, n: int):
        self.n = n

    def val_to_symbol(self, v, denom):
        return (v * self.n) // denom

    def symbol_to_range(self, s, denom):
        return (s * denom) // self.n, ((s + 1) * denom) // self.n

    def accept(self, symbol):
        pass

    def copy(self):
        return self


class CDFPredictor(Predictor):
    """Fixed inclusive cumulative table (arith_code.py:75-114)."""

    def __init__(self, dist):
        self._dist = dist
        self._minp = min(filter(lambda v: v > 0, self.pdf_iter))

    @property
    def dist(self):
        return self._dist

    @property
    def minp(self):
        return self._minp

    @property
    def pdf_iter(self):
        d = self.dist
        return itertools.chain([d[0]], (d[i + 1] - d[i] for i in range(len(d) - 1)))


class ProbPredictor(CDFPredictor):
    """Adaptive model defined by prob(symbol) (arith_code.py:115-135)."""

    def __init__(self, n: int):
        self.n = n
        self.dcache: Optional[list] = None

    def prob(self, symbol):
        return 1

    def calc_dist(self):
        p = 0
        self.dcache = []
        for s in range(self.n):
            p += self.prob(s)
            self.dcache.append(p)
        return self.dcache

    @property
    def dist(self):
        if self.dcache is None:
            return self.calc_dist()
        return self.dcache

    @property
    def minp(self):
        return min(filter(lambda v: v > 0, self.pdf_iter))

    def accept(self, symbol):
        self.dcache = None

    def copy(self):
        return self


def materialise_tables(predictor, symbols: Sequence[int]):
    """Walk a predictor along `symbols` exactly as A_to_bin.receive_symbol does (table, then accept):
    returns (dist int64 [T, V], minp int64 [T])."""
    tabs, minps = [], []
    for s in symbols:
        tabs.append(np.asarray(predictor.dist, dtype=np.int64))
        minps.append(int(predictor.minp))
        predictor.accept(s)
    if not tabs:
        return np.zeros((0, len(predictor.dist)), dtype=np.int64), np.zeros(0, dtype=np.int64)
    return np.stack(tabs), np.asarray(minps, dtype=np.int64)


# ------------------------------------------------------------------ bit packing (arith_code.py:347-362)
def group_bits(bits: Iterable[int], b: int = 8) -> Iterator[int]:
    r = 1
    for v in bits:
        r = (r << 1) | v
        if r >> b:
            yield r ^ (1 << b)
            r >>= b
    if r > 1:
        while r >> b == 0:
            r <<= 1
        yield r ^ (1 << b)


def ungroup_bits(groups: Iterable[int], b: int = 8) -> Iterator[int]:
    for g in groups:
        for i in range(b):
            yield (g >> (b - i - 1)) & 1


def _bits_of(data: bytes, nbits: int) -> List[int]:
    a = np.unpackbits(np.frombuffer(data, dtype=np.uint8))[:nbits]
    return a.astype(int).tolist()


# ------------------------------------------------------------------ the coder pair
ternary = Predictor(3)  # arith_code.py:143


class AC:
    def __init__(self, predictor=ternary, prec: int = 16, wrap64: bool = False):
        self.predictor = predictor
        self.precision = prec
        self.wrap64 = wrap64  # reproduce Llama_AC's numpy-int64 overflow in fudged_dist (llama_compress.py:29)

    def __repr__(self) -> str:
        return f"AC({self.predictor!r} at {self.precision} bits)"

    @property
    def to_bin(self):
        return A_to_bin(self.predictor.copy(), self.precision, self.wrap64)

    @property
    def from_bin(self):
        return A_from_bin(self.predictor.copy(), self.precision, self.wrap64)


class A_to_bin:
    """Encoder.  run()/bits()/encode() code the whole symbol sequence in one GPU call."""

    def __init__(self, predictor, prec: int = 16, wrap64: bool = False):
        self.predictor = predictor
        self.precision = prec
        self.wrap64 = wrap64
        self.emitted_bits = 0

    def _encode(self, symbols, stop):
        import torch
        from . import coder
        symbols = [int(s) for s in symbols]
        T = len(symbols)
        enc = coder.StreamEncoder(1, prec=self.precision, capacity_bytes=T * 8 + 64)
        uniform = type(self.predictor) is Predictor
        if not uniform:
            dist, minp = materialise_tables(self.predictor, symbols)
        if T == 0:
            if stop:
                enc.finish()
        elif uniform:
            enc.encode_uniform(torch.tensor([symbols], dtype=torch.int32, device="cuda"), self.predictor.n,
                               finish=bool(stop))
        else:
            enc.encode_tables(torch.from_numpy(dist).cuda(), torch.tensor([symbols], dtype=torch.int32, device="cuda"),
                              torch.from_numpy(minp).cuda(), finish=bool(stop), wrap64=self.wrap64)
        streams, nbits = enc.bitstreams()
        self.emitted_bits = int(nbits[0])
        return streams[0], int(nbits[0])

    def run(self, symbols, stop=1):
        data, n = self._encode(symbols, stop)
        yield from _bits_of(data, n)

    bits = run  # the reference's bits() only differs in when carries are resolved; the bit string is the same

    def encode(self, symbols, stop=1):
        data, n = self._encode(symbols, stop)
        return (int.from_bytes(data, "big") >> (8 * len(data) - n)) if n else 0, n

    def compress(self, symbols, stop=1) -> bytes:
        """bytes(group_bits(self.bits(symbols, stop))) without the Python bit loop."""
        return self._encode(symbols, stop)[0]


class A_from_bin:
    """Decoder.  `count` symbols are produced (the reference has no length framing)."""

    def __init__(self, predictor, prec: int = 16, wrap64: bool = False):
        self.predictor = predictor
        self.precision = prec
        self.wrap64 = wrap64

    def decompress(self, data: bytes, count: int) -> List[int]:
        import torch
        from . import coder
        dec = coder.StreamDecoder([bytes(data)], prec=self.precision)
        if type(self.predictor) is Predictor:
            return dec.decode_uniform(self.predictor.n, count).cpu().numpy()[0].astype(int).tolist()
        out: List[int] = []
        static = not hasattr(self.predictor, "dcache") or type(self.predictor).accept is Predictor.accept
        if static:  # fixed table: one GPU call for all symbols
            dist = torch.as_tensor(np.asarray(self.predictor.dist, dtype=np.int64)).cuda()
            minp = torch.tensor([int(self.predictor.minp)], dtype=torch.int64, device="cuda")
            return dec.decode_tables(dist, minp, count, wrap64=self.wrap64).cpu().numpy()[0].astype(int).tolist()
        for _ in range(count):  # adaptive model: the next table depends on the symbol just decoded
            dist = torch.as_tensor(np.asarray(self.predictor.dist, dtype=np.int64)).cuda()
            minp = torch.tensor([int(self.predictor.minp)], dtype=torch.int64, device="cuda")
            s = int(dec.decode_tables(dist, minp, 1, wrap64=self.wrap64).cpu().numpy()[0, 0])
            self.predictor.accept(s)
            out.append(s)
        return out

    def run(self, bits, stop=1, count: Optional[int] = None):
        if count is None:
            raise ValueError("A_from_bin.run needs count=: the reference decoder's open-ended flush() is not reproduced")
        data = bytes(group_bits(bits))
        yield from self.decompress(data, count)

    def decode(self, bits: int, length: int, stop=1, count: Optional[int] = None):
        def biter():
            n = length
            while n:
                n -= 1
                yield (bits >> n) & 1
        yield from self.run(biter(), stop, count)
