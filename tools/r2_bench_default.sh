#!/bin/bash
O=gpurun_out/r2_bench; mkdir -p $O; rm -f $O/*
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err ) 2> $O/time.txt; echo "rc=$?"; cat $O/time.txt; tail -5 $O/bench.err | cut -c1-600
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err ) 2> $O/time_ref.txt; cat $O/time_ref.txt; cut -c1-700 $O/bench_ref.json
free -g | head -2
