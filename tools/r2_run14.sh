#!/bin/bash
O=gpurun_out/r2_run14; mkdir -p $O; rm -f $O/*
timeout 300 python -m pytest tests/test_gpu_join.py tests/test_host_framework.py -m gpu -q --timeout 150 --maxfail=5 -k "groupby or host_tests or cli or dropin or bench_usage" 2>&1 | tail -12 | cut -c1-500
DWARF_BENCH_SEED=3 timeout 60 dwarf_bench_b200/lib/dwarf_bench GroupBy --device=gpu --input_size 268435456 --iterations 3 --groups_count 20 2>&1 | tail -5
DWARF_BENCH_SEED=3 timeout 60 dwarf_bench_b200/lib/dwarf_bench GroupBy --device=gpu --input_size 268435456 --iterations 3 --groups_count 1000000 2>&1 | tail -5
