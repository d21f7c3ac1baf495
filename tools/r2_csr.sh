#!/bin/bash
# one-to-many table (csrc/csr.cuh) on one B200 -> profiles/r2_csr.md
#   tests that touch engines without the unique-keys flag; BASELINE config 3 per CTA shape of the probe kernel
#   (DWJ_MULTI_SHAPE: 0 = 256x4 rows 4 CTAs/SM, 1 = 256x4 3 CTAs/SM, 2 = 256x2 5 CTAs/SM (default), 3 = 256x1 8 CTAs/SM;
#   the first sweeps also held 512x2x2, 256x2x6, 256x1x6 and 128x2x12); the OmniSci dwarf at high multiplicity; ncu: per-kernel
#   time + DRAM bytes of one step and one full capture of the probe kernel, exported to CSV on the box (gpurun returns <= 64 MiB)
O=gpurun_out/r2_csr; mkdir -p $O; rm -f $O/*
timeout 420 python -m pytest tests/test_gpu_join.py tests/test_host_framework.py -m gpu -q -x --timeout 150 > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log | cut -c1-400
A="--workload join_16Mx256M_u32_dup4_zipf --steps 8 --warmup 3 --no-cpu-baseline --no-e2e"
for sh in 0 1 2 3; do
  DWJ_MULTI_SHAPE=$sh timeout 150 python bench.py $A > $O/shape$sh.json 2> $O/shape$sh.err || { echo "shape $sh failed"; tail -2 $O/shape$sh.err; continue; }
  python -c "import json; d=json.loads(open('$O/shape$sh.json').read()); print('shape $sh:', d['ms_per_step'], d['phases_ms'], d['roofline']['kernel_ms'], round(d['roofline']['frac'],3))"
done
DWJ_PARTITION_MIN_MB=100000 timeout 150 python bench.py $A > $O/nopart.json 2> $O/nopart.err
python -c "import json; d=json.loads(open('$O/nopart.json').read()); print('no partition:', d['ms_per_step'], d['phases_ms'], d['roofline']['kernel_ms'])"
for n in 1048576 4194304; do DWARF_BENCH_SEED=3 timeout 120 dwarf_bench_b200/lib/dwarf_bench JoinOmnisci --device=gpu --input_size $n --iterations 3 > $O/omnisci_$n.txt 2>&1; echo "omnisci $n rc=$?"; tail -4 $O/omnisci_$n.txt; done
B="--workload join_16Mx256M_u32_dup4_zipf --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/ncu_traffic_cfg3.csv python bench.py $B > $O/ncu1.log 2>&1; echo "ncu1 rc=$?"
R=/tmp/probe_multi_cfg3
timeout 240 ncu --set full --clock-control none -k regex:probe_pairs_multi -s 1 -c 1 -o $R python bench.py $B > $O/ncu2.log 2>&1; echo "ncu2 rc=$?"
ncu -i $R.ncu-rep --page details --csv > $O/ncu_probe_multi_details.csv 2>/dev/null
ls -la $O
