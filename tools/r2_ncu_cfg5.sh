#!/bin/bash
# launch list (time + DRAM bytes per kernel) of one step of the default workload on one GPU (-> profiles/r2_ncu_traffic_cfg5.csv.gz)
O=gpurun_out/r2_ncu5; mkdir -p $O; rm -f $O/*
A="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-sub-configs"
timeout 210 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/ncu_cfg5.csv python bench.py $A > $O/ncu_cfg5.log 2>&1; echo "cfg5 rc=$?"
gzip -f $O/ncu_cfg5.csv; ls -la $O
