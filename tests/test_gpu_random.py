"""Randomised GPU parity (run with -m gpu on a B200): shapes, strides, alignments and precisions drawn at random,
every case through the raw C ABI and against the oracle, bit for bit.  Covers what the fixed-shape tests do not:
padded rows (row / token / stream strides larger than the data), base pointers off by 1 .. 3 floats (the scalar-load
path with the same segmentation), ragged streams, several calls per stream, workspaces of arbitrary size."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():  # collected on CPU boxes, deselected by -m "not gpu"
    pytest.skip("no CUDA device", allow_module_level=True)

from lac_b200 import _ffi  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def _case(rng):
    V = int(rng.choice([rng.integers(1, 64), rng.integers(64, 5000), rng.integers(5000, 40000),
                        rng.integers(40000, 140000)]))
    big = V > 20000
    S = int(rng.integers(1, 4 if big else 9))
    T = int(rng.integers(1, 6 if big else 20))
    pad_v = int(rng.choice([0, 0, 1, 3, 4, 8]))            # row pitch V + pad_v
    pad_t = int(rng.choice([0, 0, 1, 2]))                   # extra rows between streams
    off = int(rng.choice([0, 0, 0, 1, 2, 3]))               # base pointer offset in floats
    prec = int(rng.choice([34, 40, 48, 56, 60]))
    scale = float(rng.choice([0.3, 2.0, 6.0, 25.0]))
    return V, S, T, pad_v, pad_t, off, prec, scale


@pytest.mark.parametrize("seed", range(12))
def test_random_shapes_strides_and_alignments(seed):
    rng = np.random.default_rng(1000 + seed)
    L = _ffi.lib()
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(5):
        V, S, T, pad_v, pad_t, off, prec, scale = _case(rng)
        tok_stride = V + pad_v
        stream_stride = (T + pad_t) * tok_stride
        # the padded buffer is filled with NaN where no logit lives: nothing may be read into a result from there
        host = np.full(off + S * stream_stride + 8, np.nan, dtype=np.float32)
        logits = (rng.standard_normal((S, T, V)) * scale).astype(np.float32)
        if rng.random() < 0.3:
            logits[rng.integers(0, S), rng.integers(0, T), rng.integers(0, V)] = -np.inf
        for s in range(S):
            for t in range(T):
                b = off + s * stream_stride + t * tok_stride
                host[b:b + V] = logits[s, t]
        dev = torch.from_numpy(host).cuda()
        base = dev.data_ptr() + 4 * off
        syms_h = rng.integers(0, V, (S, T)).astype(np.int32)
        ntok_h = rng.integers(0, T + 1, S).astype(np.int32) if rng.random() < 0.5 else np.full(S, T, dtype=np.int32)
        syms, ntok = torch.from_numpy(syms_h).cuda(), torch.from_numpy(ntok_h).cuda()
        cap = T * 8 + 64
        enc_state = torch.zeros((S, _ffi.ENC_STATE_BYTES), dtype=torch.uint8, device="cuda")
        out = torch.zeros((S, cap), dtype=torch.uint8, device="cuda")
        ws_rows = int(rng.choice([0, 1, S, S * T, 3 * S * T]))
        ws_bytes = int(L.lac_workspace_bytes(max(ws_rows, 1), V)) if ws_rows else 0
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device="cuda")
        wsp = (ws.data_ptr(), ws_bytes) if ws_rows else (None, 0)
        _ffi.check(L.lac_enc_init(enc_state.data_ptr(), S, prec, st))
        # the tokens go in as one call or as two calls split at a random point (state carried, ntok shifted by the caller)
        cut = int(rng.integers(0, T + 1)) if rng.random() < 0.5 else T
        for t0, t1 in ((0, cut), (cut, T)):
            if t1 <= t0:
                continue
            n = torch.clamp(ntok - t0, 0, t1 - t0).to(torch.int32)
            _ffi.check(L.lac_ac_encode_logits_f32(base + 4 * t0 * tok_stride, S, t1 - t0, stream_stride, tok_stride, V,
                                                  syms.data_ptr() + 4 * t0, T, n.data_ptr(), enc_state.data_ptr(),
                                                  out.data_ptr(), cap, 0, prec, *wsp, st))
        _ffi.check(L.lac_ac_encode_logits_f32(None, S, 0, 0, 0, V, None, 0, None, enc_state.data_ptr(), out.data_ptr(),
                                              cap, 1, prec, None, 0, st))                       # flush-only call
        torch.cuda.synchronize()
        state = enc_state.cpu().numpy().view(np.int64).reshape(S, 4)
        host_out = out.cpu().numpy()
        streams = []
        for s in range(S):
            assert int(state[s, 3]) & 0xFFFFFFFF == 0, (seed, V, S, T, "status")
            n = int(ntok_h[s])
            lo, hi = orc.lq32_lookup(logits[s, :n], syms_h[s, :n]) if n else (np.zeros(0, np.uint32), np.zeros(0, np.uint64))
            want = orc.pack_bits(orc.ac_encode_pairs(lo, hi, prec=prec)).tobytes()
            got = host_out[s, : (int(state[s, 2]) + 7) // 8].tobytes()
            assert got == want, (seed, V, S, T, pad_v, pad_t, off, prec, s)
            streams.append(got)
        # decode through the same strides
        offs = np.zeros(S + 1, dtype=np.int64)
        np.cumsum([len(b) for b in streams], out=offs[1:])
        data = torch.from_numpy(np.frombuffer(b"".join(streams) + b"\0" * 16, dtype=np.uint8).copy()).cuda()
        d_offs = torch.from_numpy(offs).cuda()
        dec_state = torch.zeros((S, _ffi.DEC_STATE_BYTES), dtype=torch.uint8, device="cuda")
        back = torch.full((S, T), -1, dtype=torch.int32, device="cuda")
        _ffi.check(L.lac_dec_init(dec_state.data_ptr(), S, prec, data.data_ptr(), d_offs.data_ptr(), st))
        _ffi.check(L.lac_ac_decode_logits_f32(base, S, T, stream_stride, tok_stride, V, ntok.data_ptr(),
                                              dec_state.data_ptr(), data.data_ptr(), d_offs.data_ptr(), back.data_ptr(),
                                              T, prec, *wsp, st))
        got = back.cpu().numpy()
        for s in range(S):
            n = int(ntok_h[s])
            assert np.array_equal(got[s, :n], syms_h[s, :n]), (seed, V, S, T, s)
            assert (got[s, n:] == -1).all()                      # nothing written past a stream's token count
        # the table builder on the same padded rows (row stride = token pitch, first stream only)
        cum = torch.empty((T, V), dtype=torch.int32, device="cuda")
        _ffi.check(L.lac_cdf_build_f32(base, T, V, tok_stride, cum.data_ptr(), *wsp, st))
        assert np.array_equal(cum.cpu().numpy().view(np.uint32), orc.lq32_cdf(logits[0])), (seed, V, "build")


@pytest.mark.parametrize("V,S,T", [(32000, 5, 1), (32000, 3, 3), (1000, 7, 20), (70001, 2, 4), (128256, 3, 2)])
def test_nothing_is_written_outside_the_declared_buffers(V, S, T):
    """Guard bytes around every output of the logits-driven calls -- the workspace beyond its declared size, the
    bitstream buffers (per-token fused path and slice path), the decoded symbols, the pairs, the table -- stay intact.
    (compute-sanitizer is not available on the GPU pool; this is the bounds check of our own.)"""
    rng = np.random.default_rng(V + S + T)
    L = _ffi.lib()
    st = torch.cuda.current_stream().cuda_stream
    G = 0xA5
    logits = torch.from_numpy((rng.standard_normal((S, T, V)) * 4).astype(np.float32)).cuda()
    syms = torch.from_numpy(rng.integers(0, V, (S, T)).astype(np.int32)).cuda()
    ws_bytes = int(L.lac_workspace_bytes(S * T, V))
    ws = torch.full((ws_bytes + 4096,), G, dtype=torch.uint8, device="cuda")
    cap = T * 8 + 64
    out = torch.full((S + 2, cap), G, dtype=torch.uint8, device="cuda")
    enc_state = torch.full((S + 2, _ffi.ENC_STATE_BYTES), G, dtype=torch.uint8, device="cuda")
    _ffi.check(L.lac_enc_init(enc_state[1:].data_ptr(), S, 48, st))
    _ffi.check(L.lac_ac_encode_logits_f32(logits.data_ptr(), S, T, T * V, V, V, syms.data_ptr(), T, None,
                                          enc_state[1:].data_ptr(), out[1:].data_ptr(), cap, 1, 48, ws.data_ptr(),
                                          ws_bytes, st))
    torch.cuda.synchronize()
    assert (ws[ws_bytes:] == G).all(), "workspace overrun (encode)"
    assert (out[0] == G).all() and (out[S + 1] == G).all(), "bitstream buffer overrun"
    assert (enc_state[0] == G).all() and (enc_state[S + 1] == G).all(), "encoder state overrun"
    nbits = enc_state[1:S + 1].cpu().numpy().view(np.int64).reshape(S, 4)[:, 2]
    host_out = out[1:S + 1].cpu().numpy()
    for s in range(S):
        used = (int(nbits[s]) + 7) // 8
        assert (host_out[s, used:] == G).all(), "bytes written past the stream's end"
    streams = [host_out[s, : (int(nbits[s]) + 7) // 8].tobytes() for s in range(S)]
    offs = np.zeros(S + 1, dtype=np.int64)
    np.cumsum([len(b) for b in streams], out=offs[1:])
    data = torch.from_numpy(np.frombuffer(b"".join(streams) + b"\0" * 16, dtype=np.uint8).copy()).cuda()
    d_offs = torch.from_numpy(offs).cuda()
    dec_state = torch.full((S + 2, _ffi.DEC_STATE_BYTES), G, dtype=torch.uint8, device="cuda")
    back = torch.full((S + 2, T), -7, dtype=torch.int32, device="cuda")
    ws.fill_(G)
    _ffi.check(L.lac_dec_init(dec_state[1:].data_ptr(), S, 48, data.data_ptr(), d_offs.data_ptr(), st))
    _ffi.check(L.lac_ac_decode_logits_f32(logits.data_ptr(), S, T, T * V, V, V, None, dec_state[1:].data_ptr(),
                                          data.data_ptr(), d_offs.data_ptr(), back[1:].data_ptr(), T, 48, ws.data_ptr(),
                                          ws_bytes, st))
    torch.cuda.synchronize()
    assert (ws[ws_bytes:] == G).all(), "workspace overrun (decode)"
    assert torch.equal(back[1:S + 1], syms) and (back[0] == -7).all() and (back[S + 1] == -7).all()
    assert (dec_state[0] == G).all() and (dec_state[S + 1] == G).all()
    rows = S * T
    pairs = torch.full((rows + 2, 2), -7, dtype=torch.int32, device="cuda")
    ws.fill_(G)
    _ffi.check(L.lac_cdf_lookup_f32(logits.data_ptr(), rows, V, V, syms.data_ptr(), pairs[1:].data_ptr(), None,
                                    ws.data_ptr(), ws_bytes, st))
    cum = torch.full((rows + 2, V), -7, dtype=torch.int32, device="cuda")
    _ffi.check(L.lac_cdf_build_f32(logits.data_ptr(), rows, V, V, cum[1:].data_ptr(), ws.data_ptr(), ws_bytes, st))
    torch.cuda.synchronize()
    assert (ws[ws_bytes:] == G).all(), "workspace overrun (lookup / build)"
    assert (pairs[0] == -7).all() and (pairs[rows + 1] == -7).all()
    assert (cum[0] == -7).all() and (cum[rows + 1] == -7).all()
    assert np.array_equal(cum[1:rows + 1].cpu().numpy().view(np.uint32), orc.lq32_cdf(logits.view(rows, V).cpu().numpy()))
