#!/bin/bash
O=gpurun_out/r2_ncu; mkdir -p $O
timeout 300 ncu --set full --import-source on -k regex:partition_scatter_many -s 2 -c 1 -o $O/ncu_scatter_u64_512 python tools/partition_sweep.py --rows $((1<<26)) --key-bytes 8 --parts 512 > $O/ncu1.log 2>&1
ls -la $O
