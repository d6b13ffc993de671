"""Seeded synthetic inputs shaped like the reference's data: ragged video lengths,
row-L2-normalised features with eps 1e-5 (method/data_provider.py:71-73), zero padding."""
import numpy as np
import torch


def encoded_corpus(Nv, L, D, seed, min_len=None, shared=0.6):
    """Random stand-in for ENCODED frame features (Nv, L, D) + mask (Nv, L) + lengths.
    A per-video shared component makes frames of one video correlated, like real encoder output."""
    g = torch.Generator().manual_seed(seed)
    min_len = L // 2 if min_len is None else min_len
    lengths = torch.randint(min_len, L + 1, (Nv,), generator=g)
    lengths[0] = L
    if Nv > 1:
        lengths[1] = max(1, min_len)
    base = torch.randn(Nv, 1, D, generator=g)
    x = shared * base + torch.randn(Nv, L, D, generator=g)
    mask = (torch.arange(L)[None, :] < lengths[:, None]).float()
    x = x * mask[:, :, None]
    return x.contiguous(), mask.contiguous(), lengths.int()


def encoded_queries(M, D, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(M, D, generator=g)


def raw_videos(Nv, L, Dv, seed, min_len=None):
    rng = np.random.default_rng(seed)
    min_len = L // 2 if min_len is None else min_len
    out = []
    for n in range(Nv):
        ln = L if n == 0 else int(rng.integers(min_len, L + 1))
        a = rng.standard_normal((ln, Dv)).astype(np.float32)
        a = a / (np.linalg.norm(a, axis=-1, keepdims=True) + 1e-5)
        out.append(torch.from_numpy(a.astype(np.float32)))
    return out


def raw_queries(Nq, Dq, seed, min_len=5, max_len=30):
    rng = np.random.default_rng(seed)
    out = []
    for q in range(Nq):
        ln = int(rng.integers(min_len, max_len + 1))
        a = rng.standard_normal((ln, Dq)).astype(np.float32)
        a = a / (np.linalg.norm(a, axis=-1, keepdims=True) + 1e-5)
        out.append(torch.from_numpy(a.astype(np.float32)))
    return out
