#!/bin/bash
O=gpurun_out/r2_dbg; mkdir -p $O
export CUDA_DEVICE_MAX_CONNECTIONS=32 DWJ_XJ_TIMEOUT_MS=5000 DWJ_TEST_WATCHDOG=40 DWJ_PARTITION_MIN_MB=0 DWJ_REGION_MB=0.0625
for cfg in "1 4 direct 1" "2 4 direct 1" "2 8 scatter 1" "4 4 scatter 2" "4 8 direct 1"; do
  tag=$(echo $cfg | tr ' ' '_')
  timeout 70 python tests/mg_worker.py virtual $cfg $O/res_$tag.json > $O/out_$tag.log 2>&1; echo "rc=$?" >> $O/out_$tag.log
done
tail -n 15 $O/out_*.log
