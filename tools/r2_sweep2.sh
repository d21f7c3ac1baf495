#!/bin/bash
O=gpurun_out/r2_sweep2; mkdir -p $O; rm -f $O/*
timeout 100 python -m pytest tests/test_gpu_join.py -m gpu -q -x -k "partition or region or filter or segments" --timeout 100 2>&1 | tail -2
for W in 8 4; do
  for shape in 0 1; do
    echo "## key bytes $W shape $shape (pair staging)" >> $O/sweep.txt
    DWJ_SCATTER_SHAPE=$shape timeout 120 python tools/partition_sweep.py --rows $((1<<28)) --key-bytes $W --parts 8 32 128 256 512 >> $O/sweep.txt 2>&1
  done
done
cat $O/sweep.txt | grep -v "^rows"
