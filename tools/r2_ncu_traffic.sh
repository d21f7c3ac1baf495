#!/bin/bash
# full GPU test suite, then the per-kernel DRAM traffic of every bench workload (-> profiles/r2_ncu_traffic_*.csv.gz, profiles/traffic.json)
O=gpurun_out/r2_ncu_traffic; mkdir -p $O; rm -f $O/*
timeout 600 python -m pytest tests -m gpu -q --timeout 200 --maxfail=5 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log; tail -6 $O/pytest.log | cut -c1-300
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for w in join_16Mx256M_u32_unique join_16Mx256M_u32_dup4_zipf join_512Mx1G_u64_unique join_256Mx256M_u32_unique; do
  A="--workload $w --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
  timeout 200 python bench.py $A > $O/plain_$w.json 2> $O/plain_$w.err && timeout 400 ncu --metrics $M --clock-control none --csv --log-file $O/ncu_$w.csv python bench.py $A > $O/ncu_$w.log 2>&1
  echo "$w rc=$?"
done
A="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-sub-configs"
timeout 200 python bench.py $A > $O/plain_cfg5.json 2> $O/plain_cfg5.err && timeout 500 ncu --metrics $M --clock-control none --csv --log-file $O/ncu_cfg5.csv python bench.py $A > $O/ncu_cfg5.log 2>&1; echo "cfg5 rc=$?"
ls -la $O | head -30
