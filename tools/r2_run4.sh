#!/bin/bash
O=gpurun_out/r2_run4; mkdir -p $O; rm -f $O/*
timeout 420 python -m pytest tests -m gpu -q --timeout 150 --maxfail=5 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log; tail -8 $O/pytest.log | cut -c1-300
timeout 400 bash tools/r2_sweep.sh > $O/sweep_stdout.log 2>&1; cp gpurun_out/r2_sweep/sweep.txt $O/ 2>/dev/null
timeout 300 python bench.py --steps 3 --warmup 2 --no-sub-configs --no-e2e --no-cpu-baseline > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "bench rc=$?"; tail -5 $O/bench_cfg5.err | cut -c1-400; cut -c1-1500 $O/bench_cfg5.json
nvidia-smi --query-gpu=memory.used --format=csv
