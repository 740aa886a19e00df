"""Micro-benchmark of the decode-step attention variants for the stand-in Llama model (model forward is PyTorch's
job in this repo; this only picks the cheapest deterministic formulation).  python tools/attn_bench.py"""
import math
import time

import torch
import torch.nn.functional as F

dev = "cuda"
S, H, K, hd, Lmax = 256, 32, 4, 64, 2048
G = H // K
torch.manual_seed(0)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    b, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return b.elapsed_time(e) / reps


for bucket in (512, 2048):
    pos = torch.tensor([bucket - 7], device=dev)
    q = torch.randn(S, H, hd, device=dev, dtype=torch.bfloat16)
    # layout A: [S, K, L, hd]
    kc = torch.randn(S, K, Lmax, hd, device=dev, dtype=torch.bfloat16)
    vc = torch.randn(S, K, Lmax, hd, device=dev, dtype=torch.bfloat16)
    ar = torch.arange(Lmax, device=dev)

    def bmm_path():
        kk, vv = kc[:, :, :bucket], vc[:, :, :bucket]
        mask = (ar[:bucket] > pos).view(1, 1, 1, bucket)
        att = torch.matmul(q.view(S, K, G, hd), kk.transpose(-1, -2)).float() / math.sqrt(hd)
        att = att.masked_fill(mask, float("-inf")).softmax(-1).to(torch.bfloat16)
        return torch.matmul(att, vv).reshape(S, H * hd)

    def sdpa_path():
        kk, vv = kc[:, :, :bucket], vc[:, :, :bucket]
        mask = (ar[:bucket] <= pos).view(1, 1, 1, bucket)
        return F.scaled_dot_product_attention(q.view(S, H, 1, hd), kk, vv, attn_mask=mask, enable_gqa=True).reshape(S, H * hd)

    def sdpa_grouped():  # fold the G query heads of a kv head into the query length: plain MHA with q_len = G
        kk, vv = kc[:, :, :bucket], vc[:, :, :bucket]
        mask = (ar[:bucket] <= pos).view(1, 1, 1, bucket)
        return F.scaled_dot_product_attention(q.view(S, K, G, hd), kk, vv, attn_mask=mask).reshape(S, H * hd)

    print(f"bucket {bucket}: KV bytes {2 * S * K * bucket * hd * 2 / 1e6:.0f} MB")
    for name, fn in (("bmm", bmm_path), ("sdpa_gqa", sdpa_path), ("sdpa_grouped", sdpa_grouped)):
        try:
            o = fn()
            t = timeit(fn)
            o2 = fn()
            print(f"  {name:14s} {t:8.3f} ms  deterministic={torch.equal(o, o2)}  maxdiff_vs_bmm={(o.float() - bmm_path().float()).abs().max().item():.3e}")
        except Exception as ex:
            print(f"  {name:14s} failed: {type(ex).__name__}: {str(ex)[:200]}")
    try:
        from flash_attn import flash_attn_with_kvcache
        kc2 = kc.transpose(1, 2).contiguous()  # [S, L, K, hd]
        vc2 = vc.transpose(1, 2).contiguous()
        seqlens = torch.full((S,), bucket - 6, device=dev, dtype=torch.int32)

        def fa():
            return flash_attn_with_kvcache(q.view(S, 1, H, hd), kc2, vc2, cache_seqlens=seqlens, causal=True).reshape(S, H * hd)
        o = fa()
        t = timeit(fa)
        print(f"  {'flash_kvcache':14s} {t:8.3f} ms  deterministic={torch.equal(o, fa())}  maxdiff_vs_bmm={(o.float() - bmm_path().float()).abs().max().item():.3e}")
        g = torch.cuda.CUDAGraph()
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            fa()
            with torch.cuda.graph(g, stream=st):
                og = fa()
        torch.cuda.current_stream().wait_stream(st)
        g.replay()
        torch.cuda.synchronize()
        print(f"  flash_kvcache under a CUDA graph: equal to eager = {torch.equal(og, o)}")
    except Exception as ex:
        print(f"  flash_kvcache failed: {type(ex).__name__}: {str(ex)[:300]}")
    try:
        from vllm.vllm_flash_attn import flash_attn_with_kvcache as vfa  # noqa: F401
        print("  vllm_flash_attn importable")
    except Exception as ex:
        print(f"  vllm_flash_attn: {type(ex).__name__}: {str(ex)[:120]}")

# the GEMMs of one 1b layer at M = 256 for scale
d, ffn = 2048, 5632
x = torch.randn(S, d, device=dev, dtype=torch.bfloat16)
w13 = torch.randn(d, 2 * ffn, device=dev, dtype=torch.bfloat16)
print("w13 gemm ms", timeit(lambda: x @ w13))
