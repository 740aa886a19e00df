"""numpy restatement of the reference's logits -> integer table quantisation.

TEST INFRASTRUCTURE ONLY.  These are the reference's own numpy expressions, so running
them with the same numpy reproduces the reference's tables exactly.

  calc_dist     llama_compress.py:24-30  (Llama_AC.calc_dist)
  llama_minp    llama_compress.py:43-45  (Llama_AC.minp)
  acs_cdf       arithmetic_coding.py:57-72 (ACSampler.sample + get_lop_bias)
"""
import numpy as np


def calc_dist(logits):
    logits = np.asarray(logits, dtype=np.float32)
    pdf = np.exp(logits)
    pdf /= np.sum(pdf)
    return np.cumsum(np.clip((pdf * (1 << 60)).astype(float), 2, None)).astype(int)


def llama_minp(dist):
    return min(dist[0], np.min(np.diff(dist)))


def acs_cdf(pdf, precision=48):
    one = 1 << precision
    pdf = np.array(pdf, dtype=np.float64)
    pdf += sum(pdf) / (one / 2 - len(pdf))
    pdf *= one / np.sum(pdf)
    return np.cumsum(pdf).astype(np.uint64)


def ideal_bits(dist, syms):
    """Code length in bits of syms under inclusive cumulative tables dist [T, V] (float64)."""
    dist = np.asarray(dist).astype(np.float64)
    pdf = np.diff(np.concatenate([np.zeros((dist.shape[0], 1)), dist], axis=1), axis=1)
    p = pdf[np.arange(len(syms)), syms] / dist[:, -1]
    return float(-np.log2(p).sum())
