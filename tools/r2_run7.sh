#!/bin/bash
O=gpurun_out/r2_run7; mkdir -p $O; rm -f $O/*
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 150 --maxfail=3 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log; tail -12 $O/pytest.log | cut -c1-500
