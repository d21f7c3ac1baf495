#!/bin/bash
# Development: direct probing vs region-partitioned probing vs multi-pass sweep at table sizes around and beyond L2.
echo "== direct (DWJ_PARTITION_MIN_MB=1000000)"
DWJ_PARTITION_MIN_MB=1000000 python tools/probe_sweep.py --build-log2 24 25 "$@" | tail -2
echo "== partition, 32 MB regions (DWJ_SWEEP_MAX=1)"
DWJ_SWEEP_MAX=1 DWJ_PARTITION_MIN_MB=100 python tools/probe_sweep.py --build-log2 24 25 "$@" | tail -2
for mb in 32 64 128; do
  echo "== multi-pass sweep, slices of $mb MB (DWJ_SWEEP_MAX=16)"
  DWJ_SWEEP_MAX=16 DWJ_SWEEP_MB=$mb DWJ_PARTITION_MIN_MB=100 python tools/probe_sweep.py --build-log2 24 25 "$@" | tail -2
done
