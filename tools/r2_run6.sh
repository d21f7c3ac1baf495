#!/bin/bash
O=gpurun_out/r2_run6; mkdir -p $O; rm -f $O/*
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_host_framework.py -m gpu -q --timeout 150 --maxfail=3 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log; tail -4 $O/pytest.log | cut -c1-300
timeout 200 python bench.py --workload join_512Mx1G_u64_unique --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > $O/cfg4.json 2> $O/cfg4.err; python -c "
import json; d=json.load(open('$O/cfg4.json')); print('cfg4', d['ms_per_step'], d['phases_ms'])"
