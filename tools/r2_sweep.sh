#!/bin/bash
# scatter shape / MATCH ranking / RED histogram sweep (development; results -> profiles/r2_partition_sweep.txt)
O=gpurun_out/r2_sweep; mkdir -p $O
for W in 8 4; do
  for shape in 0 1 2 3; do
    for match in 0 1; do
      echo "## key bytes $W shape $shape match $match" >> $O/sweep.txt
      DWJ_SCATTER_SHAPE=$shape DWJ_SCATTER_MATCH=$match timeout 120 python tools/partition_sweep.py --rows $((1<<28)) --key-bytes $W --parts 32 128 256 512 >> $O/sweep.txt 2>&1
    done
  done
  echo "## key bytes $W RED histogram from 2^5" >> $O/sweep.txt
  DWJ_HIST_RED_FROM=5 timeout 120 python tools/partition_sweep.py --rows $((1<<28)) --key-bytes $W --parts 32 128 256 512 >> $O/sweep.txt 2>&1
done
cat $O/sweep.txt
