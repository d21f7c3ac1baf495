#!/bin/bash
# one-to-many probe kernel: second CTA shape sweep on BASELINE config 3, then ncu of the winner (per-kernel times + DRAM bytes of
# the step; one full capture of the probe kernel exported to CSV on the box) and the OmniSci dwarf at high multiplicity
O=gpurun_out/r2_csr3; mkdir -p $O; rm -f $O/*
A="--workload join_16Mx256M_u32_dup4_zipf --steps 8 --warmup 3 --no-cpu-baseline --no-e2e"
best=2; bestms=1000000
for sh in 2 4 5 6 7; do
  DWJ_MULTI_SHAPE=$sh timeout 150 python bench.py $A > $O/shape$sh.json 2> $O/shape$sh.err || { echo "shape $sh failed"; tail -2 $O/shape$sh.err; continue; }
  ms=$(python -c "import json; d=json.loads(open('$O/shape$sh.json').read()); print(d['ms_per_step'], d['phases_ms']['probe'], d['roofline']['kernel_ms'], round(d['roofline']['frac'],3))")
  echo "shape $sh: $ms"
  k=$(python -c "import json; d=json.loads(open('$O/shape$sh.json').read()); print(int(d['roofline']['kernel_ms']*1000))")
  if [ "$k" -lt "$bestms" ]; then bestms=$k; best=$sh; fi
done
echo "best shape $best ($bestms us)"
export DWJ_MULTI_SHAPE=$best
for n in 1048576 4194304; do DWARF_BENCH_SEED=3 timeout 120 dwarf_bench_b200/lib/dwarf_bench JoinOmnisci --device=gpu --input_size $n --iterations 3 > $O/omnisci_$n.txt 2>&1; echo "omnisci $n rc=$?"; grep -i "kernel\|build\|probe\|incorrect\|pairs" $O/omnisci_$n.txt | tail -6; done
B="--workload join_16Mx256M_u32_dup4_zipf --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/ncu_traffic_cfg3.csv python bench.py $B > $O/ncu1.log 2>&1; echo "ncu1 rc=$?"
R=/tmp/probe_multi_cfg3
timeout 240 ncu --set full --clock-control none -k regex:probe_pairs_multi -s 1 -c 1 -o $R python bench.py $B > $O/ncu2.log 2>&1; echo "ncu2 rc=$?"
ncu -i $R.ncu-rep --page details --csv > $O/ncu_probe_multi_details.csv 2>/dev/null
ls -la $O | head -30
