"""The reference-facing Python API (same names as arith_code.py / arithmetic_coding.py /
llama_compress.py).  CPU part: table walking and bit packing.  GPU part: the mirrors produce the
reference's golden bits."""
import os

import numpy as np
import pytest

from lac_b200 import arith_code as ac
from lac_b200 import arithmetic_coding as acs


class AdaptiveCounts(ac.ProbPredictor):
    """Same 'simple adaptive frequency model' the goldens were made with (tests/golden/make_golden.py)."""

    def __init__(self, n, counts=None):
        super().__init__(n)
        self.counts = [0] * n if counts is None else counts

    def prob(self, symbol):
        return 1 + self.counts[symbol]

    def accept(self, symbol):
        self.counts[symbol] += 1
        return super().accept(symbol)

    def copy(self):
        return AdaptiveCounts(self.n, list(self.counts))


def test_materialise_tables_walks_like_receive_symbol():
    dist, minp = ac.materialise_tables(AdaptiveCounts(4), [2, 2, 0])
    assert dist.tolist() == [[1, 2, 3, 4], [1, 2, 4, 5], [1, 2, 5, 6]]
    assert minp.tolist() == [1, 1, 1]
    d, m = ac.materialise_tables(ac.CDFPredictor([3, 5, 6, 10]), [0, 3])
    assert d.tolist() == [[3, 5, 6, 10]] * 2 and m.tolist() == [1, 1]


def test_bit_packing_matches_reference_layout():
    from oracle import oracle as orc
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 8, 9, 77):
        bits = rng.integers(0, 2, n).tolist()
        assert bytes(ac.group_bits(bits)) == orc.pack_bits(np.array(bits, dtype=np.uint8)).tobytes()
        assert list(ac.ungroup_bits(bytes(ac.group_bits(bits))))[:n] == bits
        out = bytearray()
        pk = acs.packbits(out.append)
        for b in bits:
            pk(b)
        pk.flush()
        assert bytes(out) == bytes(ac.group_bits(bits))
        assert list(acs.unpackbits(out))[:n] == bits


def test_scaled_cdf_is_the_reference_expression():
    from oracle import ref_quant
    p = np.random.default_rng(1).dirichlet(np.ones(50))
    assert np.array_equal(acs.ACSampler(48).scaled_cdf(p), ref_quant.acs_cdf(p, 48))


# ------------------------------------------------------------------ GPU
torch = pytest.importorskip("torch")
gpu = pytest.mark.gpu
needs_gpu = pytest.mark.skipif(not torch.cuda.is_available(), reason="no CUDA device")


@gpu
@needs_gpu
def test_mirror_ac_on_goldens(golden_dir):
    g = np.load(os.path.join(golden_dir, "ac_small.npz"))
    names = list(g["names"])[::9]
    for nm in names:
        prec, stop = int(g[f"{nm}/prec"]), int(g[f"{nm}/stop"])
        dist, syms = [int(x) for x in g[f"{nm}/dist"]], [int(x) for x in g[f"{nm}/syms"]]
        coder_ = ac.AC(ac.CDFPredictor(dist), prec)
        want = [int(b) for b in g[f"{nm}/bits"]]
        assert list(coder_.to_bin.bits(syms, stop)) == want, nm
        r, n = coder_.to_bin.encode(syms, stop)
        assert n == len(want) and r == int("0" + "".join(map(str, want)), 2)
        if stop and syms:
            assert list(coder_.from_bin.run(want, stop, count=len(syms))) == syms, nm


@gpu
@needs_gpu
def test_mirror_default_coder_is_the_uniform_ternary_predictor(golden_dir):
    """The reference's AC() is AC(Predictor(3), 16): same defaults, same bits, same round trip here."""
    assert ac.AC().predictor.n == 3 and ac.AC().precision == 16
    g = np.load(os.path.join(golden_dir, "ac_uniform.npz"))
    n_default = 0
    for nm in list(g["names"])[::4]:
        prec, stop, n = int(g[f"{nm}/prec"]), int(g[f"{nm}/stop"]), int(g[f"{nm}/n"])
        syms = [int(x) for x in g[f"{nm}/syms"]]
        want = [int(b) for b in g[f"{nm}/bits"]]
        coder_ = ac.AC() if (n, prec) == (3, 16) else ac.AC(ac.Predictor(n), prec)
        n_default += (n, prec) == (3, 16)
        assert list(coder_.to_bin.bits(syms, stop)) == want, nm
        if stop and syms:
            assert list(coder_.from_bin.run(want, stop, count=len(syms))) == syms, nm
    assert n_default >= 2


@gpu
@needs_gpu
def test_mirror_adaptive_model_matches_reference_bytes(golden_dir):
    g = np.load(os.path.join(golden_dir, "ac_adaptive.npz"))
    data = g["data"].tolist()
    coder_ = ac.AC(AdaptiveCounts(256), int(g["prec"]))
    comp = coder_.to_bin.compress(data)
    assert comp == g["comp"].tobytes()
    head = coder_.from_bin.decompress(comp, 300)   # adaptive decode is one GPU call per symbol
    assert head == data[:300]


@gpu
@needs_gpu
def test_mirror_acsampler_on_goldens(golden_dir):
    g = np.load(os.path.join(golden_dir, "acs_small.npz"))
    from oracle import oracle as orc
    for nm in list(g["names"])[::3]:
        prec, cdf, toks = int(g[f"{nm}/prec"]), g[f"{nm}/cdf"], g[f"{nm}/toks"].tolist()
        s = acs.ACSampler(prec)
        assert s.compress(cdf, toks) == orc.pack_bits(g[f"{nm}/bits"]).tobytes()
        assert s.last_nbits == len(g[f"{nm}/bits"])
        assert s.expand(cdf, s.compress(cdf, toks, flush="safe"), len(toks)) == toks


@gpu
@needs_gpu
@pytest.mark.parametrize("use_graphs", [True, False])
def test_llama_compress_roundtrip_and_size(use_graphs):
    """configs[2] in miniature: Llama-style random-init model, vocab 32000, ragged chunks, full-shape batches (the
    last one padded with empty streams), container; per-token step replayed from a CUDA graph."""
    from lac_b200 import container, llama_compress as lc
    chunk, B = 24, 4
    cfg = lc.LlamaConfig("t", 32000, 64, 2, 4, 2, 128, 64)
    model = lc.LlamaModel(cfg, n_streams=B, max_len=chunk, seed=3)
    rng = np.random.default_rng(5)
    toks = rng.integers(0, cfg.vocab, 5 * chunk + 7).astype(np.int32)
    comp = lc.LlamaCompressor(model, chunk_tokens=chunk, use_graphs=use_graphs)
    blob = comp.compress(toks)
    c = container.unpack(blob)
    assert c.n_chunks == 6 and c.ntok.tolist() == [chunk] * 5 + [7] and c.batch_streams == B
    back = comp.decompress(blob)
    assert np.array_equal(back, toks)
    # a second compressor instance (fresh graphs, same seed) reads the file, and writes the same file
    comp2 = lc.LlamaCompressor(lc.LlamaModel(cfg, n_streams=B, max_len=chunk, seed=3), chunk_tokens=chunk,
                               use_graphs=use_graphs)
    assert np.array_equal(comp2.decompress(blob), toks)
    assert comp2.compress(toks) == blob
    # random tokens under a random-init model (logit std ~3) cost about log2(V) + sigma^2 / (2 ln 2) ~ 21 bits each
    bits = float(c.nbits.sum())
    assert 0.9 * np.log2(cfg.vocab) * len(toks) < bits < 2.5 * np.log2(cfg.vocab) * len(toks)
    with pytest.raises(AssertionError):
        comp.compress(np.array([1, 2, cfg.vocab], dtype=np.int32))
    # a file for another predictor is refused, a damaged payload is reported
    other = lc.LlamaCompressor(lc.LlamaModel(lc.LlamaConfig("u", 32000, 64, 1, 4, 2, 128, 64), B, chunk, seed=3),
                               chunk_tokens=chunk, use_graphs=use_graphs)
    with pytest.raises(ValueError):
        other.decompress(blob)


@gpu
@needs_gpu
def test_llama_model_in_distribution_text_compresses():
    """Tokens SAMPLED from the model (in-distribution text) must cost about the model's entropy, far below
    log2(V): the coder is driven by the model, not by a uniform table."""
    import torch
    from lac_b200 import container, llama_compress as lc
    chunk, B = 32, 8
    cfg = lc.LlamaConfig("t", 32000, 64, 2, 4, 2, 128, 64)
    model = lc.LlamaModel(cfg, n_streams=B, max_len=chunk, seed=4)
    g = torch.Generator(device="cuda").manual_seed(0)
    model.reset()
    prev = torch.full((B,), lc.BOS, dtype=torch.int64, device="cuda")
    cols, ent = [], 0.0
    for t in range(chunk):
        logits = model.step(prev, lc._bucket_for(t, chunk))
        p = torch.softmax(logits.double(), -1)
        prev = torch.multinomial(p.float(), 1, generator=g).squeeze(1)
        ent += float(-(p * torch.log2(p.clamp_min(1e-300))).sum(-1).sum())
        cols.append(prev.to(torch.int32))
        model.pos.add_(1)
    toks = torch.stack(cols, 1).cpu().numpy().reshape(-1)
    comp = lc.LlamaCompressor(model, chunk_tokens=chunk, use_graphs=False)
    blob = comp.compress(toks)
    bits = float(container.unpack(blob).nbits.sum())
    assert bits < 1.25 * ent + 64 * B and bits < 0.8 * np.log2(cfg.vocab) * len(toks)
    assert np.array_equal(comp.decompress(blob), toks)
