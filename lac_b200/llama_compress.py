"""The llama_compress.py path (reference llama_compress.py:1-61) on the GPU: an autoregressive model
predicts every token, the logits go through LQ32 and the batched range coder, independent chunks are
independent streams.

Reference behaviour kept:
  * every chunk starts from a reset model that has seen the BOS token 1 (Llama_AC.reset, :18-21);
  * the table for position t is computed from the logits after tokens < t (calc_dist, :24-30);
  * coder precision 48 (r(..., prec=48), :4).
Reference behaviour replaced: tables are LQ32 (total 2^32) instead of cumsum(clip(softmax * 2^60, 2)) --
bound in DESIGN.md section 3 -- and chunks carry their token counts in the LACB container.

Losslessness needs bit-identical logits when compressing and decompressing.  The reference gets that by
evaluating llama.cpp token by token in both directions; here both directions call the SAME
`model.step(tokens)` incremental forward with the same batch shape, so the same kernels run in the same
order.  (A prefill-style encoder would be faster but is not bit-reproducible against a stepwise decoder.)
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch

from . import coder, container

BOS = 1  # llama_compress.py:19 self.past = [1]


class TinyLlama(torch.nn.Module):
    """Small Llama-style decoder (RMSNorm, rotary attention with a KV cache, SwiGLU), random-init, used
    as the stand-in predictor.  step(tokens[S]) -> fp32 logits [S, vocab] for the next position."""

    def __init__(self, vocab=32000, dim=256, layers=2, heads=4, max_len=2049, seed=0, dtype=torch.float32):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.vocab, self.dim, self.heads, self.max_len = vocab, dim, heads, max_len
        hd = dim // heads

        def w(*shape, scale):
            return torch.nn.Parameter((torch.randn(*shape, generator=g) * scale).to(dtype), requires_grad=False)
        self.emb = w(vocab, dim, scale=1.0)
        self.blocks = torch.nn.ParameterList()
        for _ in range(layers):
            self.blocks.extend([w(dim, 3 * dim, scale=dim ** -0.5), w(dim, dim, scale=dim ** -0.5),
                                w(dim, 4 * dim, scale=dim ** -0.5), w(dim, 4 * dim, scale=dim ** -0.5),
                                w(4 * dim, dim, scale=(4 * dim) ** -0.5)])
        self.head = w(dim, vocab, scale=dim ** -0.5 * 4.0)
        inv = 1.0 / (10000 ** (torch.arange(0, hd, 2).float() / hd))
        ang = torch.arange(max_len).float()[:, None] * inv[None]
        self.register_buffer("cos", ang.cos())
        self.register_buffer("sin", ang.sin())
        self.layers = layers
        self.cache = None
        self.pos = 0

    def reset(self, n_streams: int):
        dev = self.emb.device
        hd = self.dim // self.heads
        self.cache = [torch.zeros((2, n_streams, self.heads, self.max_len, hd), device=dev, dtype=self.emb.dtype)
                      for _ in range(self.layers)]
        self.pos = 0

    @staticmethod
    def _norm(x):
        return x * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + 1e-6).to(x.dtype)

    def _rope(self, x):  # x [S, H, hd]
        c, s = self.cos[self.pos].to(x.dtype), self.sin[self.pos].to(x.dtype)
        a, b = x[..., 0::2], x[..., 1::2]
        return torch.stack([a * c - b * s, a * s + b * c], dim=-1).flatten(-2)

    @torch.no_grad()
    def step(self, tokens: torch.Tensor) -> torch.Tensor:
        S, H, hd = tokens.shape[0], self.heads, self.dim // self.heads
        x = self.emb[tokens.long()]
        for l in range(self.layers):
            wqkv, wo, w1, w3, w2 = self.blocks[5 * l:5 * l + 5]
            q, k, v = (self._norm(x) @ wqkv).view(S, 3, H, hd).unbind(1)
            q, k = self._rope(q), self._rope(k)
            kc, vc = self.cache[l][0], self.cache[l][1]
            kc[:, :, self.pos], vc[:, :, self.pos] = k, v
            att = torch.einsum("shd,shtd->sht", q, kc[:, :, : self.pos + 1]) / math.sqrt(hd)
            x = x + torch.einsum("sht,shtd->shd", att.softmax(-1), vc[:, :, : self.pos + 1]).reshape(S, self.dim) @ wo
            h = self._norm(x)
            x = x + (torch.nn.functional.silu(h @ w1) * (h @ w3)) @ w2
        self.pos += 1
        return (self._norm(x) @ self.head).float().contiguous()


class LlamaCompressor:
    """compress(token ids) -> LACB bytes, decompress(bytes) -> token ids."""

    def __init__(self, model, vocab: int, chunk_tokens: int = 2048, prec: int = coder.DEFAULT_PREC,
                 max_streams: int = 1024, device="cuda"):
        self.model, self.vocab, self.chunk, self.prec, self.max_streams = model, vocab, chunk_tokens, prec, max_streams
        self.device = torch.device(device)

    def _batches(self, n_chunks):
        for b in range(0, n_chunks, self.max_streams):
            yield b, min(n_chunks, b + self.max_streams)

    def compress(self, tokens) -> bytes:
        toks = np.ascontiguousarray(tokens, dtype=np.int32)
        if toks.size and (int(toks.min()) < 0 or int(toks.max()) >= self.vocab):
            raise AssertionError("unknown symbol", int(toks.max() if toks.max() >= self.vocab else toks.min()))  # arith_code.py:104-105
        n_chunks = (len(toks) + self.chunk - 1) // self.chunk
        padded = np.zeros(n_chunks * self.chunk, dtype=np.int32)
        padded[: len(toks)] = toks
        padded = padded.reshape(n_chunks, self.chunk)
        ntok = np.full(n_chunks, self.chunk, dtype=np.int32)
        if n_chunks:
            ntok[-1] = len(toks) - (n_chunks - 1) * self.chunk
        streams: List[bytes] = []
        nbits: List[int] = []
        for b, e in self._batches(n_chunks):
            S = e - b
            d_tok = torch.from_numpy(padded[b:e]).to(self.device)
            d_ntok = torch.from_numpy(ntok[b:e]).to(self.device)
            enc = coder.StreamEncoder(S, prec=self.prec, capacity_bytes=self.chunk * 8 + 64, device=self.device)
            self.model.reset(S)
            prev = torch.full((S,), BOS, dtype=torch.int32, device=self.device)
            for t in range(int(ntok[b:e].max())):
                logits = self.model.step(prev)                       # what the decoder will also compute
                live = (d_ntok > t).to(torch.int32)                   # ragged tail chunk
                enc.encode_logits(logits.unsqueeze(1), d_tok[:, t:t + 1].contiguous(), ntok=live)
                prev = d_tok[:, t].contiguous()
            enc.finish()
            s, nb = enc.bitstreams()
            streams += s
            nbits += [int(x) for x in nb]
        return container.pack(streams, ntok, nbits, self.prec, self.vocab, self.chunk)

    def decompress(self, blob: bytes) -> np.ndarray:
        c = container.unpack(blob)
        if c.vocab != self.vocab or c.quantiser != container.QUANT_LQ32:
            raise ValueError("container was written for a different vocabulary / quantiser")
        all_streams = c.streams()
        out = np.zeros((c.n_chunks, c.chunk_tokens), dtype=np.int32)
        for b, e in self._batches(c.n_chunks):
            S = e - b
            d_ntok = torch.from_numpy(c.ntok[b:e].astype(np.int32)).to(self.device)
            dec = coder.StreamDecoder(all_streams[b:e], prec=c.prec, device=self.device)
            self.model.reset(S)
            prev = torch.full((S,), BOS, dtype=torch.int32, device=self.device)
            got = torch.zeros((S, c.chunk_tokens), dtype=torch.int32, device=self.device)
            for t in range(int(c.ntok[b:e].max())):
                logits = self.model.step(prev)
                live = (d_ntok > t).to(torch.int32)
                sym = dec.decode_logits(logits.unsqueeze(1), ntok=live).squeeze(1)
                sym = torch.where(live.bool(), sym, torch.zeros_like(sym))
                got[:, t] = sym
                prev = sym
            out[b:e] = got.cpu().numpy()
        flat = [out[i, : int(c.ntok[i])] for i in range(c.n_chunks)]
        return np.concatenate(flat) if flat else np.zeros(0, dtype=np.int32)
