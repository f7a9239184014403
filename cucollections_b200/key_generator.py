"""Synthetic key streams with the reference benchmark distributions
(include/cuco/utility/key_generator.cuh:91-232, 268-375), generated on the GPU with torch and FIXED
seeds (the reference seeds with time(), key_generator.cuh:249, which is not reproducible).

  unique(n)            shuffled 0..n-1                                     (:272-274)
  uniform(n, m)        uniform_int[1, n/m]  -> ~63 % distinct at m = 1     (:91-116, 275-284)
  gaussian(n, skew)    normal(n/2, n*skew) redrawn until inside [0, n)     (:130-162)
  dropout(keys, keep)  each key replaced w.p. 1-keep by uniform_int[n, max], then shuffled (:355-375)

The streams are statistically, not bitwise, those of thrust's minstd engine; both implementations
under test always consume the same generated tensor, so parity does not depend on it.
"""
from __future__ import annotations

import torch


def _gen(device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def unique(n: int, dtype=torch.int64, device="cuda", seed=42) -> torch.Tensor:
    return torch.randperm(n, device=device, generator=_gen(device, seed)).to(dtype)


def uniform(n: int, multiplicity: int = 1, dtype=torch.int64, device="cuda", seed=42) -> torch.Tensor:
    hi = max(1, n // multiplicity)
    return torch.randint(1, hi + 1, (n,), device=device, generator=_gen(device, seed), dtype=dtype)


def gaussian(n: int, skew: float = 0.5, dtype=torch.int64, device="cuda", seed=42) -> torch.Tensor:
    g = _gen(device, seed)
    out = torch.empty(n, device=device, dtype=torch.float64)
    todo = torch.arange(n, device=device)
    while todo.numel():
        draw = torch.normal(n / 2.0, n * skew, (todo.numel(),), device=device, generator=g,
                            dtype=torch.float64)
        ok = (draw >= 0) & (draw < n)
        out[todo[ok]] = draw[ok]
        todo = todo[~ok]
    return out.to(dtype)


def dropout(keys: torch.Tensor, keep_prob: float, seed=43) -> torch.Tensor:
    """Replaces a (1 - keep_prob) fraction by keys that cannot be in the build set, then shuffles."""
    n = keys.numel()
    if keep_prob >= 1.0:
        return keys.clone()
    g = _gen(keys.device, seed)
    info = torch.iinfo(keys.dtype)
    drop = torch.rand(n, device=keys.device, generator=g) >= keep_prob
    # stay clear of the usual sentinels at the very top of the range
    foreign = torch.randint(n, info.max - 1, (n,), device=keys.device, generator=g, dtype=keys.dtype)
    out = torch.where(drop, foreign, keys)
    return out[torch.randperm(n, device=keys.device, generator=g)]
