#!/bin/bash
# Development: staged-PAIRS shape sweep (DWJ_STAGED_SHAPE) at fixed table sizes, ordered and unordered output.
for s in 0 1 2 3 4; do
  echo "== DWJ_STAGED_SHAPE=$s ordered"
  DWJ_STAGED_SHAPE=$s python tools/probe_sweep.py --build-log2 20 24 "$@" | tail -2
done
for s in 0 2 3; do
  echo "== DWJ_STAGED_SHAPE=$s unordered"
  DWJ_STAGED_SHAPE=$s python tools/probe_sweep.py --build-log2 20 24 --unordered "$@" | tail -2
done
