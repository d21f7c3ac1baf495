"""Synthetic join inputs of SURVEY §8(d), generated on the device with torch (input plumbing, not the hot path).

Keys are produced as signed torch integers holding the same bits as the unsigned columns the engine sees.

  fk_pk        unique build keys (a fixed odd-multiplier bijection of a pseudo-random permutation of the key indices),
               payload = key index (a unique id); probe keys drawn uniformly from the build key set, every probe row
               matches exactly once (BASELINE configs 2, 4, 5 and the 256M x 256M target)
  dup_zipf     every distinct build key repeated `dup` times, probe keys Zipf(s) over the distinct keys
               (BASELINE config 3): every probe row yields `dup` matches
  reference    two independent sorted-unique samples of n values from [0, 10 n) -- the reference's own input
               shape (common/common.cpp:7-20), match rate ~0.1 n (BASELINE config 1)

Every input carries the order-independent checksum of the result rows the join must produce (`expected_checksum`):
two wrapping 64-bit sums of independent mixes of (build payload, probe payload) over all result rows, computed from the
generator's own knowledge of which build row every probe row hits.  `checksum_rows` computes the same over an output,
so a full-size run -- on one GPU or sharded over eight -- is checked for the exact row multiset without sorting 2^31 rows.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

ODD32 = 2654435761            # 0x9E3779B1
ODD64 = 0x9E3779B97F4A7C15
CHUNK = 1 << 27               # generation / checksum granularity (bounds the temporaries at 2^31 rows)


def _s64(x: int) -> int:
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


_M1, _M2, _M3, _M4 = (_s64(c) for c in (0x9E3779B97F4A7C15, 0xC2B2AE3D27D4EB4F, 0xD6E8FEB86659FD93, 0xFF51AFD7ED558CCD))


def _mix_sums(b: torch.Tensor, p: torch.Tensor):
    """Two wrapping int64 sums over rows of independent mixes of (b, p); b, p int64 tensors."""
    h = b * _M1 + p * _M2
    h ^= h >> 31
    h *= _M3
    g = (b ^ _M4) * _M2 + (p + 0x632BE5AB) * _M1
    g ^= g >> 29
    g *= _M4
    return int(h.sum().item()), int(g.sum().item())


def _add(a, b):
    return (_s64(a[0] + b[0]), _s64(a[1] + b[1]))


def checksum_rows(out_build: torch.Tensor, out_probe: torch.Tensor, n_rows: int):
    """Order-independent checksum of the first n_rows (build payload, probe payload) result rows (any integer dtype)."""
    tot = (0, 0)
    for lo in range(0, n_rows, CHUNK):
        hi = min(lo + CHUNK, n_rows)
        tot = _add(tot, _mix_sums(out_build[lo:hi].to(torch.int64), out_probe[lo:hi].to(torch.int64)))
    return tot


@dataclass
class JoinInput:
    build_keys: torch.Tensor
    build_vals: torch.Tensor
    probe_keys: torch.Tensor
    probe_vals: torch.Tensor
    expected_matches: int
    probe_build_row: torch.Tensor | None = None   # fk_pk (small inputs): build row each probe row hits (exact checks in tests)
    unique_build: bool = True
    expected_checksum: tuple | None = None        # of this rank's PROBE rows' result rows (sum over ranks = the global join)

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


def _permute_pow2(i: torch.Tensor, bits: int, seed: int) -> torch.Tensor:
    """A pseudo-random PERMUTATION of [0, 2^bits) applied to int64 indices i: odd multiplications and xor-shifts, each
    a bijection on `bits`-bit words.  Stands in for torch.randperm, whose sort refuses more than 2^31 - 1 elements."""
    mask = (1 << bits) - 1
    a1 = (0x9E3779B97F4A7C15 * (2 * seed + 1)) & mask | 1
    a2 = (0xC2B2AE3D27D4EB4F * (2 * seed + 3)) & mask | 1
    s = max(bits // 2, 1)
    x = (i * _s64(a1) + seed * 7919) & mask
    x = x ^ (x >> s)
    x = (x * _s64(a2)) & mask
    x = x ^ (x >> s)
    return x


def fk_pk(n_build: int, n_probe: int, key_bytes: int = 4, seed: int = 7, device="cuda", key_base: int = 0,
          key_space: int | None = None, keep_map: bool = True, probe_val_base: int = 0) -> JoinInput:
    """key_base / key_space let several ranks generate disjoint slices of one global build relation: this rank's build
    keys are the key indices [key_base, key_base + n_build) of a key space of `key_space` indices, in pseudo-random
    order; probe keys are drawn from the whole key space; probe payloads are probe_val_base + row.  The build payload
    is the key index, so the build payload every probe row must find is its target index -- known without knowing which
    rank holds that build row."""
    g = torch.Generator(device=device).manual_seed(seed)
    dt = _dtype(key_bytes)
    key_space = key_space or n_build
    pow2 = n_build >= 2 and (n_build & (n_build - 1)) == 0
    bits = n_build.bit_length() - 1
    perm = None
    build_keys = torch.empty(n_build, dtype=dt, device=device)
    build_vals = torch.empty(n_build, dtype=dt, device=device)
    if not pow2:
        perm = torch.randperm(n_build, device=device, generator=g, dtype=torch.int64)
    for lo in range(0, n_build, CHUNK):
        hi = min(lo + CHUNK, n_build)
        idx = perm[lo:hi] if perm is not None else _permute_pow2(torch.arange(lo, hi, device=device, dtype=torch.int64), bits, seed)
        idx = idx + key_base
        build_keys[lo:hi] = _scatter_keys(idx, key_bytes)
        build_vals[lo:hi] = idx.to(dt)
    probe_keys = torch.empty(n_probe, dtype=dt, device=device)
    probe_vals = torch.empty(n_probe, dtype=dt, device=device)
    small = keep_map and key_space == n_build and key_base == 0 and n_probe <= (1 << 28) and perm is not None
    targets = [] if small else None
    want = (0, 0)
    for lo in range(0, n_probe, CHUNK):
        hi = min(lo + CHUNK, n_probe)
        target = torch.randint(0, key_space, (hi - lo,), device=device, generator=g, dtype=torch.int64)
        pv = torch.arange(lo, hi, device=device, dtype=torch.int64) + probe_val_base
        probe_keys[lo:hi] = _scatter_keys(target, key_bytes)
        probe_vals[lo:hi] = pv.to(dt)
        # what comes out of the join are the dt-typed payloads, sign-extended by checksum_rows: mix the same values
        want = _add(want, _mix_sums(target.to(dt).to(torch.int64), pv.to(dt).to(torch.int64)))
        if small:
            targets.append(target)
    row = None
    if small:
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(n_build, device=device, dtype=torch.int64)
        row = inv[torch.cat(targets)]
    return JoinInput(build_keys, build_vals, probe_keys, probe_vals, n_probe, row, True, want)


def dup_zipf(n_build: int, n_probe: int, dup: int = 4, s: float = 1.0, key_bytes: int = 4, seed: int = 11,
             device="cuda") -> JoinInput:
    g = torch.Generator(device=device).manual_seed(seed)
    dt = _dtype(key_bytes)
    distinct = n_build // dup
    keys = _scatter_keys(torch.arange(distinct, device=device, dtype=torch.int64), key_bytes)
    perm = torch.randperm(distinct * dup, device=device, generator=g)
    build_keys = keys.repeat_interleave(dup)[perm]                 # build row i holds key index perm[i] // dup
    build_vals = torch.arange(distinct * dup, device=device, dtype=torch.int64).to(dt)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(distinct * dup, device=device, dtype=torch.int64)     # rows of key r: inv[r*dup .. r*dup+dup)
    w = 1.0 / torch.arange(1, distinct + 1, device=device, dtype=torch.float64) ** s
    cdf = torch.cumsum(w, 0)
    cdf /= cdf[-1].clone()
    probe_keys = torch.empty(n_probe, dtype=dt, device=device)
    probe_vals = torch.arange(n_probe, device=device, dtype=torch.int64).to(dt)
    want = (0, 0)
    for lo in range(0, n_probe, CHUNK):
        hi = min(lo + CHUNK, n_probe)
        u = torch.rand(hi - lo, device=device, generator=g, dtype=torch.float64)
        ranks = torch.searchsorted(cdf, u).clamp_(max=distinct - 1)
        probe_keys[lo:hi] = keys[ranks]
        pv = probe_vals[lo:hi].to(torch.int64)
        for d in range(dup):
            want = _add(want, _mix_sums(inv[ranks * dup + d].to(dt).to(torch.int64), pv))
    return JoinInput(build_keys, build_vals, probe_keys, probe_vals, n_probe * dup, None, False, want)


def reference_shape(n: int, seed: int = 1, device="cuda") -> JoinInput:
    g = torch.Generator(device=device).manual_seed(seed)

    def sample():
        return torch.randperm(10 * n, device=device, generator=g)[:n].sort().values.to(torch.int32)

    ak, av, bk, bv = sample(), sample(), sample(), sample()
    pos = torch.searchsorted(ak, bk).clamp_(max=n - 1)             # both key columns are sorted and unique
    hit = ak[pos] == bk
    want = _mix_sums(av[pos[hit]].to(torch.int64), bv[hit].to(torch.int64))
    return JoinInput(ak, av, bk, bv, int(hit.sum().item()), None, True, want)
