"""The llama_compress.py path (reference llama_compress.py:1-61) on the GPU: an autoregressive model
predicts every token, the logits go through LQ32 and the batched range coder, independent chunks are
independent streams, batches of chunks shard over the GPUs of a box.

Reference behaviour kept:
  * every chunk starts from a reset model that has seen the BOS token 1 (Llama_AC.reset, :20-23);
  * the table for position t is computed from the logits after tokens < t (calc_dist, :24-30);
  * coder precision 48 (r(..., prec=48), :4);
  * an unknown token raises AssertionError("unknown symbol") (:51-52).
Reference behaviour replaced: tables are LQ32 (total 2^32) instead of cumsum(clip(softmax * 2^60, 2)) --
bound in DESIGN.md section 3 -- and chunks carry their token counts in the LACB container.  The reference's
context roll-over at n_ctx (accept, :31-39) does not arise: a chunk is at most max_len tokens.

Losslessness needs bit-identical logits when compressing and decompressing.  The reference gets that by
evaluating llama.cpp token by token in both directions; here both directions replay the SAME captured CUDA graph
of `model.step` with the same batch shape (chunks are coded in batches of exactly `batch_streams` streams), so
the same kernels run in the same order whether a file is written on one GPU and read on eight or the other way
round.  (A prefill-style encoder would be faster but is not bit-reproducible against a stepwise decoder.)

The per-token step -- model forward, LQ32 row summaries, fused symbol-range + range-coder kernel (or the decoder's
serial pass), position update -- is ONE CUDA graph replay with no host-fed inputs: the tokens of the step are
gathered on the device from the token matrix with the device-side position.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _ffi, coder, container, sharding

BOS = 1  # llama_compress.py:21 self.past = [1]


@dataclass(frozen=True)
class LlamaConfig:
    name: str
    vocab: int
    dim: int
    layers: int
    heads: int
    kv_heads: int
    ffn: int
    max_len: int = 2048

    def describe(self) -> str:
        return (f"{self.name}:v{self.vocab}:d{self.dim}:l{self.layers}:h{self.heads}:kv{self.kv_heads}:f{self.ffn}:"
                f"n{self.max_len}")

    @property
    def params(self) -> int:
        hd = self.dim // self.heads
        per_layer = self.dim * (self.heads + 2 * self.kv_heads) * hd + self.dim * self.dim + 3 * self.dim * self.ffn
        return self.layers * per_layer + 2 * self.vocab * self.dim


CONFIGS: Dict[str, LlamaConfig] = {
    # miniature for tests
    "tiny": LlamaConfig("tiny", 32000, 256, 2, 4, 2, 512, 2048),
    # configs[2]: "Llama-style ~1B random-init model (vocab 32000)" -- TinyLlama-1.1B geometry
    "1b": LlamaConfig("1b", 32000, 2048, 22, 32, 4, 5632, 2048),
    # configs[3]: Llama-3 8B geometry (vocab 128256)
    "8b": LlamaConfig("8b", 128256, 4096, 32, 32, 8, 14336, 2048),
    # configs[4]: a 1B-class predictor with the 128256 vocabulary (Llama-3.2-1B geometry) for 8192 concurrent streams
    "1b-128k": LlamaConfig("1b-128k", 128256, 2048, 16, 32, 8, 8192, 2048),
}

_BUCKETS = (128, 256, 512, 1024, 2048, 4096, 8192)


class LlamaModel(torch.nn.Module):
    """Llama-style decoder (RMSNorm, rotary GQA attention with a KV cache, SwiGLU), random-init, bf16 weights.
    step(tokens int64 [S]) -> fp32 logits [S, vocab] for the next position; the position lives on the device.
    All shapes are static for a given attention bucket, so a step can be captured into a CUDA graph."""

    def __init__(self, cfg: LlamaConfig, n_streams: int, max_len: Optional[int] = None, seed: int = 0,
                 device="cuda", dtype=torch.bfloat16):
        super().__init__()
        self.cfg, self.S = cfg, int(n_streams)
        self.max_len = int(max_len or cfg.max_len)
        self.device, self.dtype = torch.device(device), dtype
        g = torch.Generator(device=self.device).manual_seed(seed)
        hd = cfg.dim // cfg.heads
        self.hd = hd

        def w(*shape, scale):
            return (torch.randn(*shape, generator=g, device=self.device, dtype=torch.float32) * scale).to(dtype)
        d = cfg.dim
        self.emb = w(cfg.vocab, d, scale=1.0)
        self.wqkv = [w(d, (cfg.heads + 2 * cfg.kv_heads) * hd, scale=d ** -0.5) for _ in range(cfg.layers)]
        self.wo = [w(d, d, scale=d ** -0.5) for _ in range(cfg.layers)]
        self.w13 = [w(d, 2 * cfg.ffn, scale=d ** -0.5) for _ in range(cfg.layers)]
        self.w2 = [w(cfg.ffn, d, scale=cfg.ffn ** -0.5) for _ in range(cfg.layers)]
        self.head = w(d, cfg.vocab, scale=d ** -0.5 * 3.0)   # logit std ~3: a peaked, LLM-like next-token law
        inv = 1.0 / (10000 ** (torch.arange(0, hd, 2, device=self.device).float() / hd))
        ang = torch.arange(self.max_len, device=self.device).float()[:, None] * inv[None]
        self.cos = torch.cat([ang.cos(), ang.cos()], -1).to(dtype)    # rotate-half form, [max_len, hd]
        self.sin = torch.cat([ang.sin(), ang.sin()], -1).to(dtype)
        # KV cache [layer][S, kv_heads, max_len, hd]
        self.kc = [torch.zeros((self.S, cfg.kv_heads, self.max_len, hd), device=self.device, dtype=dtype)
                   for _ in range(cfg.layers)]
        self.vc = [torch.zeros_like(k) for k in self.kc]
        self.pos = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.ar = torch.arange(self.max_len, device=self.device)

    def reset(self):
        self.pos.zero_()

    def _norm(self, x):
        return torch.nn.functional.rms_norm(x, (x.shape[-1],), eps=1e-6)

    @staticmethod
    def _rope(x, c, s):  # x [S, heads, hd]; c, s [1, 1, hd]
        half = x.shape[-1] // 2
        rot = torch.cat([-x[..., half:], x[..., :half]], -1)
        return x * c + rot * s

    @torch.no_grad()
    def step(self, tokens: torch.Tensor, bucket: int) -> torch.Tensor:
        """One position for all S streams; attends over the first `bucket` cache slots (bucket > pos).  The
        attention is torch's scaled_dot_product_attention over the KV-cache slice with a position mask (measured
        against a bmm formulation and flash_attn_with_kvcache in tools/attn_bench.py: 0.10 vs 0.29 / 0.19 ms per
        layer at 256 streams x 2048 positions, deterministic)."""
        cfg, S, hd = self.cfg, self.S, self.hd
        H, K = cfg.heads, cfg.kv_heads
        pos = self.pos
        c = self.cos.index_select(0, pos).view(1, 1, hd)
        s = self.sin.index_select(0, pos).view(1, 1, hd)
        mask = (self.ar[:bucket] <= pos).view(1, 1, 1, bucket)
        x = self.emb.index_select(0, tokens)
        for l in range(cfg.layers):
            qkv = (self._norm(x) @ self.wqkv[l]).view(S, H + 2 * K, hd)
            qk = self._rope(qkv[:, :H + K], c, s)
            self.kc[l].index_copy_(2, pos, qk[:, H:].unsqueeze(2))
            self.vc[l].index_copy_(2, pos, qkv[:, H + K:].unsqueeze(2))
            o = torch.nn.functional.scaled_dot_product_attention(
                qk[:, :H].unsqueeze(2), self.kc[l][:, :, :bucket], self.vc[l][:, :, :bucket], attn_mask=mask,
                enable_gqa=True)
            x = x + o.reshape(S, H * hd) @ self.wo[l]
            h13 = self._norm(x) @ self.w13[l]
            x = x + (torch.nn.functional.silu(h13[:, :cfg.ffn]) * h13[:, cfg.ffn:]) @ self.w2[l]
        return (self._norm(x) @ self.head).float()


def _bucket_for(pos: int, max_len: int) -> int:
    for b in _BUCKETS:
        if pos < b:
            return min(b, max_len)
    return max_len


class StepEngine:
    """The per-token loop of one batch: holds the model, the static device buffers and one captured CUDA graph per
    (direction, attention bucket).  A step is `graph.replay()`; nothing is fed from the host."""

    def __init__(self, model: LlamaModel, chunk_tokens: int, prec: int, use_graphs: bool = True):
        self.m, self.T, self.prec, self.use_graphs = model, int(chunk_tokens), int(prec), use_graphs
        S, V, dev = model.S, model.cfg.vocab, model.device
        if self.T > model.max_len:
            raise ValueError("chunk_tokens exceeds the model's max_len")
        self.tok = torch.zeros((S, self.T), dtype=torch.int32, device=dev)      # encode: input; decode: output
        self.ntok = torch.zeros(S, dtype=torch.int32, device=dev)
        self.live = torch.zeros(S, dtype=torch.int32, device=dev)
        self.sym = torch.zeros(S, dtype=torch.int32, device=dev)
        self.ws = coder.Workspace(S, V, dev)
        self.cap = self.T * 8 + 64
        self.enc: Optional[coder.StreamEncoder] = None
        self.dec: Optional[coder.StreamDecoder] = None
        self.graphs: Dict = {}
        self.stream = torch.cuda.Stream(device=dev)

    # ---- one step, expressed on device tensors only
    def _prev_tokens(self):
        pos = self.m.pos
        prev = self.tok.index_select(1, (pos - 1).clamp_(min=0)).squeeze(1).to(torch.int64)
        return torch.where(pos > 0, prev, torch.full_like(prev, BOS))

    def _enc_step(self, bucket):
        logits = self.m.step(self._prev_tokens(), bucket)
        pos = self.m.pos
        self.sym.copy_(self.tok.index_select(1, pos).squeeze(1))
        self.live.copy_((self.ntok > pos).to(torch.int32))
        self.enc.encode_step(logits, self.sym, self.live, self.ws)
        pos.add_(1)

    def _dec_step(self, bucket):
        logits = self.m.step(self._prev_tokens(), bucket)
        pos = self.m.pos
        self.live.copy_((self.ntok > pos).to(torch.int32))
        self.dec.decode_step(logits, self.live, self.ws, out=self.sym)
        self.tok.index_copy_(1, pos, torch.where(self.live > 0, self.sym, torch.zeros_like(self.sym)).unsqueeze(1))
        pos.add_(1)

    def _run(self, kind, n_steps):
        fn = self._enc_step if kind == "enc" else self._dec_step
        cur = torch.cuda.current_stream(self.m.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            for t in range(n_steps):
                bucket = _bucket_for(t, self.m.max_len)
                if not self.use_graphs:
                    fn(bucket)
                    continue
                key = (kind, bucket)
                if key not in self.graphs:
                    # warm-up run outside capture (cuBLAS handles / workspaces), then rewind what it changed
                    snap = self._snapshot(kind)
                    fn(bucket)
                    self._restore(kind, snap)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self.stream):
                        fn(bucket)
                    self._restore(kind, snap)
                    self.graphs[key] = g
                self.graphs[key].replay()
        cur.wait_stream(self.stream)

    def _snapshot(self, kind):
        """Everything a step changes that the next step reads: position, coder state, decoded tokens, and -- encode
        side -- the stream bytes (a carry ripples into bytes already written, so a repeated step would add it twice)."""
        st = self.enc.state if kind == "enc" else self.dec.state
        out = self.enc.out.clone() if kind == "enc" else None
        return (self.m.pos.clone(), st.clone(), self.tok.clone(), out)

    def _restore(self, kind, snap):
        st = self.enc.state if kind == "enc" else self.dec.state
        self.m.pos.copy_(snap[0])
        st.copy_(snap[1])
        self.tok.copy_(snap[2])
        if snap[3] is not None:
            self.enc.out.copy_(snap[3])

    # ---- whole batches
    def encode_batch(self, tokens: np.ndarray, ntok: np.ndarray):
        S = self.m.S
        assert tokens.shape == (S, self.T) and ntok.shape == (S,)
        self.tok.copy_(torch.from_numpy(np.ascontiguousarray(tokens, dtype=np.int32)))
        self.ntok.copy_(torch.from_numpy(np.ascontiguousarray(ntok, dtype=np.int32)))
        if self.enc is None:
            self.enc = coder.StreamEncoder(S, prec=self.prec, capacity_bytes=self.cap, device=self.m.device)
        else:
            self.enc.reset()
        self.m.reset()
        self._run("enc", int(ntok.max()) if len(ntok) else 0)
        self.enc.finish()
        return self.enc.bitstreams()

    def decode_batch(self, streams: List[bytes], ntok: np.ndarray) -> np.ndarray:
        S = self.m.S
        assert len(streams) == S and ntok.shape == (S,)
        self.ntok.copy_(torch.from_numpy(np.ascontiguousarray(ntok, dtype=np.int32)))
        self.tok.zero_()
        if self.dec is None:
            self.dec = coder.StreamDecoder(streams, prec=self.prec, device=self.m.device,
                                           capacity_bytes=S * self.cap)
        else:
            self.dec.reset(streams)
        self.m.reset()
        self._run("dec", int(ntok.max()) if len(ntok) else 0)
        st = self.dec.status()
        if st & _ffi.LAC_ST_TRUNC:
            raise _ffi.LacError(_ffi.LAC_E_STREAM, "truncated or foreign bitstream (wrong model for this file?)")
        return self.tok.cpu().numpy()


class LlamaCompressor:
    """compress(token ids) -> LACB bytes, decompress(bytes) -> token ids.  With torch.distributed initialised every
    rank calls both with the same arguments; rank 0 gets the result, the others None."""

    def __init__(self, model: LlamaModel, chunk_tokens: int = 2048, prec: int = coder.DEFAULT_PREC,
                 use_graphs: bool = True, group=None):
        self.model, self.chunk, self.prec, self.group = model, int(chunk_tokens), int(prec), group
        self.vocab = model.cfg.vocab
        self.engine = StepEngine(model, chunk_tokens, prec, use_graphs)
        self.tag = container.model_tag(model.cfg.describe())

    def compress(self, tokens) -> Optional[bytes]:
        toks = np.ascontiguousarray(tokens, dtype=np.int32)
        if toks.size and (int(toks.min()) < 0 or int(toks.max()) >= self.vocab):
            bad = int(toks.max()) if int(toks.max()) >= self.vocab else int(toks.min())
            raise AssertionError("unknown symbol", bad)  # llama_compress.py:51-52
        return sharding.compress_sharded(toks, self.chunk, self.model.S, self.engine.encode_batch, self.prec,
                                         self.vocab, self.model.device, self.tag, self.group)

    def decompress(self, blob: bytes) -> Optional[np.ndarray]:
        c = container.unpack(blob)
        if c.vocab != self.vocab or c.quantiser != container.QUANT_LQ32:
            raise ValueError("container was written for a different vocabulary / quantiser")
        if c.batch_streams != self.model.S or c.chunk_tokens != self.chunk:
            raise ValueError(f"container was written with batches of {c.batch_streams} streams x {c.chunk_tokens} "
                             f"tokens; this compressor runs {self.model.S} x {self.chunk}")
        if c.tag and c.tag != self.tag:
            raise ValueError("container was written with a different predictor configuration")
        return sharding.decompress_sharded(blob, self.engine.decode_batch, self.model.device, self.group)


# ------------------------------------------------------------------ the reference's own entry points
from .arith_code import AC, ProbPredictor  # noqa: E402


class Llama_AC(ProbPredictor):
    """Drop-in for the reference's Llama_AC (llama_compress.py:14-61): a predictor around a llama_cpp.Llama-like
    object -- anything with reset(), eval(tokens), n_ctx() and _scores[-1] (the fp32 logits row of the current
    position, numpy or torch).  AC(Llama_AC(llm), 48).to_bin / .from_bin then code through the LQ32 kernels: the
    row goes to the device once per token (logits_row) and lac_ac_encode_logits_f32 / lac_ac_decode_logits_f32 do
    calc_dist + symbol_to_range / val_to_symbol + the coder step.  The tables are LQ32's, not the reference's
    cumsum(clip(softmax * 2^60, 2)) (whose re-scaling wraps int64), so the bitstreams are not
    interchangeable with the reference's; calc_dist() returns the LQ32 table in the reference's layout (inclusive
    cumulative, total 2^32) for callers that look at it."""

    def __init__(self, llm, maxtoks=2048):
        super().__init__(0)
        self.llm = llm
        self.overlap = 2
        self.reset()

    def reset(self):                                       # :20-23
        self.past = [BOS]
        self.llm.reset()
        self.llm.eval([BOS])
        self.dcache = None

    def logits_row(self) -> torch.Tensor:
        row = self.llm._scores[-1]
        if not isinstance(row, torch.Tensor):
            row = torch.from_numpy(np.ascontiguousarray(row, dtype=np.float32))
        return row.to(device="cuda", dtype=torch.float32).contiguous().view(-1)

    def calc_dist(self):                                   # :24-30, LQ32 instead of the float cumsum
        cum = coder.cdf_build(self.logits_row().view(1, -1))
        self.dcache = coder.cdf_to_dist(cum)[0]
        return self.dcache

    @property
    def minp(self):                                        # :43-45
        d = self.dist
        return int(min(int(d[0]), int(np.min(np.diff(d)))))

    def accept(self, symbol):                              # :31-39
        self.past.append(symbol)
        if len(self.past) == self.llm.n_ctx():
            self.past = self.past[self.llm.n_ctx() - self.llm.n_ctx() // self.overlap:]
            self.llm.reset()
            self.llm.eval(self.past)
        else:
            self.llm.eval([symbol])
        return super().accept(symbol)

    def copy(self):                                        # :40-41
        return Llama_AC(self.llm)


_llm = None


def r(model_path="../../llama.cpp/models/llama-2-7b.ggmlv3.q5_1.bin", prec=48):
    """llama_compress.py:4-10: AC(Llama_AC(llama_cpp.Llama(model_path, n_ctx=512)), prec)."""
    import llama_cpp
    global _llm
    if _llm is None:
        _llm = llama_cpp.Llama(model_path=model_path, n_ctx=512)
    return AC(Llama_AC(_llm), prec)
