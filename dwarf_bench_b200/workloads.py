"""Synthetic join inputs of SURVEY §8(d), generated on the device with torch (input plumbing, not the hot path).

Keys are produced as signed torch integers holding the same bits as the unsigned columns the engine sees.

  fk_pk        unique build keys (a fixed odd-multiplier bijection of a random permutation), payload = row id;
               probe keys drawn uniformly from the build key set, every probe row matches exactly once
               (BASELINE configs 2, 4, 5 and the 256M x 256M target)
  dup_zipf     every distinct build key repeated `dup` times, probe keys Zipf(s) over the distinct keys
               (BASELINE config 3): every probe row yields `dup` matches
  reference    two independent sorted-unique samples of n values from [0, 10 n) -- the reference's own input
               shape (common/common.cpp:7-20), match rate ~0.1 n (BASELINE config 1)
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

ODD32 = 2654435761            # 0x9E3779B1
ODD64 = 0x9E3779B97F4A7C15


@dataclass
class JoinInput:
    build_keys: torch.Tensor
    build_vals: torch.Tensor
    probe_keys: torch.Tensor
    probe_vals: torch.Tensor
    expected_matches: int
    probe_build_row: torch.Tensor | None = None   # fk_pk: build row each probe row hits (for exact checks)
    unique_build: bool = True

    @property
    def n_build(self) -> int:
        return self.build_keys.numel()

    @property
    def n_probe(self) -> int:
        return self.probe_keys.numel()


def _dtype(key_bytes: int):
    return torch.int32 if key_bytes == 4 else torch.int64


def _scatter_keys(idx: torch.Tensor, key_bytes: int) -> torch.Tensor:
    """Bijection index -> sparse key (odd multiplier mod 2^w, plus an offset so 0 and all-ones are not special)."""
    if key_bytes == 4:
        k = (idx.to(torch.int64) * ODD32 + 12345) & 0xFFFFFFFF
        k = torch.where(k >= 2**31, k - 2**32, k)                 # same bits as the uint32 value
        return k.to(torch.int32)
    return idx.to(torch.int64) * (ODD64 - 2**64) + 12345          # wraps mod 2^64 in two's complement


def fk_pk(n_build: int, n_probe: int, key_bytes: int = 4, seed: int = 7, device="cuda", key_base: int = 0,
          key_space: int | None = None, keep_map: bool = True) -> JoinInput:
    """key_base/key_space let several ranks generate disjoint slices of one global build relation:
    this rank's build keys are indices [key_base, key_base + n_build) of a key space of `key_space` indices,
    and probe keys are drawn from the whole key space."""
    g = torch.Generator(device=device).manual_seed(seed)
    dt = _dtype(key_bytes)
    key_space = key_space or n_build
    perm = torch.randperm(n_build, device=device, generator=g, dtype=torch.int64)
    build_keys = _scatter_keys(perm + key_base, key_bytes)
    build_vals = (torch.arange(n_build, device=device, dtype=torch.int64) + key_base).to(dt)
    target = torch.randint(0, key_space, (n_probe,), device=device, generator=g, dtype=torch.int64)
    probe_keys = _scatter_keys(target, key_bytes)
    probe_vals = torch.arange(n_probe, device=device, dtype=torch.int64).to(dt)
    row = None
    if keep_map and key_space == n_build and key_base == 0:
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(n_build, device=device, dtype=torch.int64)
        row = inv[target]
    return JoinInput(build_keys, build_vals, probe_keys, probe_vals, n_probe, row, True)


def dup_zipf(n_build: int, n_probe: int, dup: int = 4, s: float = 1.0, key_bytes: int = 4, seed: int = 11,
             device="cuda") -> JoinInput:
    g = torch.Generator(device=device).manual_seed(seed)
    dt = _dtype(key_bytes)
    distinct = n_build // dup
    keys = _scatter_keys(torch.arange(distinct, device=device, dtype=torch.int64), key_bytes)
    build_keys = keys.repeat_interleave(dup)[torch.randperm(distinct * dup, device=device, generator=g)]
    build_vals = torch.arange(distinct * dup, device=device, dtype=torch.int64).to(dt)
    w = 1.0 / torch.arange(1, distinct + 1, device=device, dtype=torch.float64) ** s
    cdf = torch.cumsum(w, 0)
    cdf /= cdf[-1].clone()
    u = torch.rand(n_probe, device=device, generator=g, dtype=torch.float64)
    ranks = torch.searchsorted(cdf, u).clamp_(max=distinct - 1)
    probe_keys = keys[ranks]
    probe_vals = torch.arange(n_probe, device=device, dtype=torch.int64).to(dt)
    return JoinInput(build_keys, build_vals, probe_keys, probe_vals, n_probe * dup, None, False)


def reference_shape(n: int, seed: int = 1, device="cuda") -> JoinInput:
    g = torch.Generator(device=device).manual_seed(seed)

    def sample():
        return torch.randperm(10 * n, device=device, generator=g)[:n].sort().values.to(torch.int32)

    ak, av, bk, bv = sample(), sample(), sample(), sample()
    matches = int(torch.isin(bk, ak).sum().item())
    return JoinInput(ak, av, bk, bv, matches, None, True)
