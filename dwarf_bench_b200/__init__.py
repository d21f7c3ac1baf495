"""dwarf_bench_b200 -- a B200-native hash-join engine behind dwarf_bench's Join dwarf.

The product is `lib/libdwj_b200.so` (hand-written sm_100a CUDA behind the C ABI in include/dwj.h)
and the C++ host framework under `host/` (the reference's Dwarf / Meter / Result / Registry /
RunOptions / makeMeasurements surface and the `dwarf_bench` CLI).  This Python package is the thin
ctypes binding used by tests and bench.py, plus the torch.distributed plumbing of the multi-GPU join.
There is no CPU fallback: importing works anywhere, creating an engine needs a B200.
"""
from .capi import (DwjError, Engine, ExchangeJoinRank, JoinTiming, MultiGpuJoin, lib_path, load_library,  # noqa: F401
                   FLAG_L2_PERSIST, FLAG_NO_PARTITION, FLAG_UNIQUE_BUILD_KEYS, FLAG_UNORDERED_OUTPUT, OUT_ALIGNED, OUT_COUNT, OUT_PAIRS)

__all__ = ["DwjError", "Engine", "ExchangeJoinRank", "MultiGpuJoin", "JoinTiming", "lib_path", "load_library", "FLAG_L2_PERSIST",
           "FLAG_NO_PARTITION", "FLAG_UNIQUE_BUILD_KEYS", "FLAG_UNORDERED_OUTPUT", "OUT_ALIGNED", "OUT_COUNT", "OUT_PAIRS"]
