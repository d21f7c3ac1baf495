#!/bin/bash
O=gpurun_out/r2_dbg; mkdir -p $O; rm -f $O/*
run() { tag=$(echo "$1 $DWJ_TEST_CHUNK_ROWS p$DWJ_XJ_NO_PRIORITY" | tr ' ' '_'); CUDA_DEVICE_MAX_CONNECTIONS=32 DWJ_XJ_TIMEOUT_MS=2000 DWJ_TEST_WATCHDOG=50 DWJ_PARTITION_MIN_MB=0 DWJ_REGION_MB=0.0625 timeout 70 python tests/mg_worker.py virtual $1 $O/res_$tag.json > $O/out_$tag.log 2>&1; echo "rc=$?" >> $O/out_$tag.log; echo "== $tag"; tail -n 3 $O/out_$tag.log | cut -c1-500; }
DWJ_TEST_CHUNK_ROWS=9000 run "1 4 direct 1"
DWJ_XJ_NO_PRIORITY=1 DWJ_TEST_CHUNK_ROWS=9000 run "1 4 direct 1"
DWJ_TEST_CHUNK_ROWS=9000 run "2 4 direct 1"
DWJ_TEST_CHUNK_ROWS=9000 run "2 8 scatter 1"
DWJ_TEST_CHUNK_ROWS=9000 run "4 4 scatter 2"
DWJ_TEST_CHUNK_ROWS=9000 run "8 8 direct 1"
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_join.py -m gpu -q -x --timeout 100 2>&1 | tail -15 | cut -c1-300
