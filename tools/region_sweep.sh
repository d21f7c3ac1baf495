#!/bin/bash
# Development: direct probing vs region-partitioned probing at table sizes around and beyond L2.
echo "== direct (DWJ_PARTITION_MIN_MB=1000000)"
DWJ_PARTITION_MIN_MB=1000000 python tools/probe_sweep.py --build-log2 24 26 "$@" | tail -2
for mb in 16 32 64; do
  echo "== regions of $mb MB"
  DWJ_PARTITION_MIN_MB=100 DWJ_REGION_MB=$mb python tools/probe_sweep.py --build-log2 24 26 "$@" | tail -2
done
