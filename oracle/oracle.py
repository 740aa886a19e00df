"""ctypes front-end for oracle/lac_oracle.c.

TEST INFRASTRUCTURE ONLY (see the header of lac_oracle.c): imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by
anything under lac_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblac_oracle.so")

ERRORS = {
    -1: "bad argument",
    -2: "output buffer too small",
    -3: "unknown symbol (arith_code.py:100-101)",
    -4: "predictor range does not correspond to val (arith_code.py:277-278)",
    -5: "carry out of first bit",
    -6: "max() of empty range (arith_code.py:312)",
    -7: "ZeroDivisionError (arith_code.py:305-307)",
    -8: "IndexError in ACSampler lookup (arithmetic_coding.py:114-115)",
}


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__(f"oracle error {code}: {ERRORS.get(code, '?')}")
        self.code = code


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "lac_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        i64, i32, p = C.c_int64, C.c_int32, C.c_void_p
        L.orc_ac_encode.argtypes = [C.c_int, p, i64, i64, p, C.c_int, p, i64, C.c_int, p, i64, p, p, C.c_int]
        L.orc_ac_decode.argtypes = [C.c_int, p, i64, i64, p, C.c_int, p, i64, C.c_int, i64, p, i64, p, C.c_int]
        L.orc_ac_decode_n.argtypes = [C.c_int, p, i64, i64, p, C.c_int, p, i64, i64, p, C.c_int]
        L.orc_acs_encode.argtypes = [C.c_int, p, i64, i64, C.c_int, p, i64, C.c_int, p, i64, p]
        L.orc_acs_decode.argtypes = [C.c_int, p, i64, i64, C.c_int, p, i64, i64, p]
        L.orc_ac_encode_pairs.argtypes = [C.c_int, p, p, i64, C.c_int, p, i64, p]
        L.orc_lq32_cdf.argtypes = [p, i64, C.c_int, i64, p]
        L.orc_lq32_lookup.argtypes = [p, i64, C.c_int, i64, p, p, p]
        L.orc_pack_bits.argtypes = [p, i64, p]
        L.orc_pack_bits.restype = i64
        L.orc_unpack_bits.argtypes = [p, i64, p]
        L.orc_unpack_bits.restype = None
        L.orc_cdf_minp.argtypes = [p, C.c_int]
        L.orc_cdf_minp.restype = i64
        L.orc_llama_minp.argtypes = [p, C.c_int]
        L.orc_llama_minp.restype = i64
        L.orc_ref_calc_dist.argtypes = [p, C.c_int, p]
        L.orc_ref_calc_dist.restype = None
        L.orc_ref_roundtrip_bulk.argtypes = [p, p, i64, i64, C.c_int, C.c_int, C.c_int, p]
        L.orc_ref_roundtrip_bulk.restype = i64
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _ptr(a):
    if hasattr(a, "ctypes_ptr"):
        return a.ctypes_ptr()
    return a.ctypes.data_as(C.c_void_p)


class _Uniform:
    """Stands for the reference's uniform base class Predictor(n) (arith_code.py:64-74)."""

    def __init__(self, n):
        self.n = int(n)

    def ctypes_ptr(self):
        return None


def uniform(n):
    """Pass as `dist` to ac_encode / ac_decode / ac_decode_n for AC(Predictor(n), prec)."""
    return _Uniform(n)


def _tables(dist, minp, kind="cdf"):
    """dist: [V] shared or [T, V] per-position int64 inclusive cumulative tables."""
    if isinstance(dist, _Uniform):
        return dist, 0, 1, dist.n, np.ones(1, dtype=np.int64)
    dist = np.ascontiguousarray(dist, dtype=np.int64)
    if dist.ndim == 1:
        stride, ntab, V = 0, 1, dist.shape[0]
        rows = dist[None]
    else:
        ntab, V = dist.shape
        stride = V
        rows = dist
    if minp is None:
        f = lib().orc_cdf_minp if kind == "cdf" else lib().orc_llama_minp
        minp = np.array([f(_ptr(r), V) for r in rows], dtype=np.int64)
    minp = np.ascontiguousarray(np.atleast_1d(minp), dtype=np.int64)
    return dist, stride, ntab, V, minp


def ac_encode(dist, syms, prec=16, stop=1, minp=None, kind="cdf", return_state=False, wrap64=False):
    """Bits of arith_code.A_to_bin(CDFPredictor(dist), prec).bits(syms, stop)."""
    dist, stride, ntab, V, minp = _tables(dist, minp, kind)
    syms = np.ascontiguousarray(syms, dtype=np.int32)
    cap = int(len(syms)) * (prec + 2) + 4 * prec + 64
    bits = np.zeros(cap, dtype=np.uint8)
    nb = C.c_int64(0)
    st = np.zeros(3, dtype=np.int64)
    rc = lib().orc_ac_encode(prec, _ptr(dist), stride, ntab, _ptr(minp), V, _ptr(syms), len(syms),
                             int(stop), _ptr(bits), cap, C.byref(nb), _ptr(st), int(wrap64))
    if rc:
        raise OracleError(rc)
    out = bits[: nb.value].copy()
    return (out, st) if return_state else out


def ac_decode(dist, bits, prec=16, stop=1, minp=None, kind="cdf", max_syms=0, cap=None, wrap64=False):
    """list(arith_code.A_from_bin(CDFPredictor(dist), prec).run(bits, stop)); returns (symbols, rc)."""
    dist, stride, ntab, V, minp = _tables(dist, minp, kind)
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    cap = cap or (len(bits) * 4 + 4096)
    out = np.zeros(cap, dtype=np.int32)
    n = C.c_int64(0)
    rc = lib().orc_ac_decode(prec, _ptr(dist), stride, ntab, _ptr(minp), V, _ptr(bits), len(bits),
                             int(stop), int(max_syms), _ptr(out), cap, C.byref(n), int(wrap64))
    return out[: n.value].copy(), rc


def ac_decode_n(dist, bits, n, prec=16, minp=None, kind="cdf", wrap64=False):
    dist, stride, ntab, V, minp = _tables(dist, minp, kind)
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    out = np.zeros(n, dtype=np.int32)
    rc = lib().orc_ac_decode_n(prec, _ptr(dist), stride, ntab, _ptr(minp), V, _ptr(bits), len(bits), n, _ptr(out), int(wrap64))
    if rc:
        raise OracleError(rc)
    return out


def _acs_tables(cdf):
    cdf = np.ascontiguousarray(cdf, dtype=np.uint64)
    if cdf.ndim == 1:
        return cdf, 0, 1, cdf.shape[0]
    return cdf, cdf.shape[1], cdf.shape[0], cdf.shape[1]


def acs_encode(cdf, toks, prec=48, flush=1):
    """Bits ACSampler(prec) hands to compress_output for toks, then flush_compress()."""
    cdf, stride, ntab, V = _acs_tables(cdf)
    toks = np.ascontiguousarray(toks, dtype=np.int32)
    cap = len(toks) * (prec + 2) + 4 * prec + 64
    bits = np.zeros(cap, dtype=np.uint8)
    nb = C.c_int64(0)
    rc = lib().orc_acs_encode(prec, _ptr(cdf), stride, ntab, V, _ptr(toks), len(toks), int(flush),
                              _ptr(bits), cap, C.byref(nb))
    if rc:
        raise OracleError(rc)
    return bits[: nb.value].copy()


def acs_decode(cdf, bits, n, prec=48):
    """n calls of ACSampler.sample_scaled_cdf in expand mode, literal (quirks included)."""
    cdf, stride, ntab, V = _acs_tables(cdf)
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    out = np.zeros(n, dtype=np.int32)
    rc = lib().orc_acs_decode(prec, _ptr(cdf), stride, ntab, V, _ptr(bits), len(bits), n, _ptr(out))
    return out, rc


def ac_encode_pairs(lo, hi, prec=48, stop=1):
    lo = np.ascontiguousarray(lo, dtype=np.uint32)
    hi = np.ascontiguousarray(hi, dtype=np.uint64)
    cap = len(lo) * (prec + 2) + 4 * prec + 64
    bits = np.zeros(cap, dtype=np.uint8)
    nb = C.c_int64(0)
    rc = lib().orc_ac_encode_pairs(prec, _ptr(lo), _ptr(hi), len(lo), int(stop), _ptr(bits), cap, C.byref(nb))
    if rc:
        raise OracleError(rc)
    return bits[: nb.value].copy()


def lq32_cdf(logits):
    """Exclusive cumulative LQ32 tables, uint32 [rows, V]; total 2^32 implicit."""
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    rows, V = logits.shape
    cum = np.zeros((rows, V), dtype=np.uint32)
    rc = lib().orc_lq32_cdf(_ptr(logits), rows, V, V, _ptr(cum))
    if rc:
        raise OracleError(rc)
    return cum


def lq32_lookup(logits, syms):
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    syms = np.ascontiguousarray(syms, dtype=np.int32)
    rows, V = logits.shape
    lo = np.zeros(rows, dtype=np.uint32)
    hi = np.zeros(rows, dtype=np.uint64)
    rc = lib().orc_lq32_lookup(_ptr(logits), rows, V, V, _ptr(syms), _ptr(lo), _ptr(hi))
    if rc:
        raise OracleError(rc)
    return lo, hi


def lq32_to_dist(cum):
    """Exclusive uint32 LQ32 table(s) -> int64 inclusive tables as CDFPredictor wants them."""
    cum = np.asarray(cum, dtype=np.uint32).astype(np.int64)
    out = np.empty_like(cum)
    out[..., :-1] = cum[..., 1:]
    out[..., -1] = 1 << 32
    return out


def pack_bits(bits):
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    out = np.zeros((len(bits) + 7) // 8, dtype=np.uint8)
    lib().orc_pack_bits(_ptr(bits), len(bits), _ptr(out))
    return out


def unpack_bits(data):
    data = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
    bits = np.zeros(len(data) * 8, dtype=np.uint8)
    lib().orc_unpack_bits(_ptr(data), len(data), _ptr(bits))
    return bits


def ref_calc_dist(logits_row):
    x = np.ascontiguousarray(logits_row, dtype=np.float32)
    out = np.zeros(x.shape[0], dtype=np.int64)
    lib().orc_ref_calc_dist(_ptr(x), x.shape[0], _ptr(out))
    return out


def ref_roundtrip_bulk(logits, syms, prec=48, threads=0):
    """Reference algorithm (C port) encode+decode of [streams, T, V] logits; returns (mismatches, bits)."""
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    syms = np.ascontiguousarray(syms, dtype=np.int32)
    S, T, V = logits.shape
    tot = C.c_int64(0)
    bad = lib().orc_ref_roundtrip_bulk(_ptr(logits), _ptr(syms), S, T, V, prec, threads, C.byref(tot))
    return int(bad), int(tot.value)


def num_threads():
    return int(lib().orc_num_threads())
