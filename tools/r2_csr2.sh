#!/bin/bash
# one-to-many probe kernel: CTA shape sweep on BASELINE config 3 (DWJ_MULTI_SHAPE), with and without region partitioning
O=gpurun_out/r2_csr2; mkdir -p $O; rm -f $O/*
timeout 300 python -m pytest tests/test_gpu_join.py -m gpu -q -x --timeout 120 -k "one_to_many or duplicates or heavy or hash_build_dwarf or aligned_probe_under or golden or fixtures or join_host_with_dup or region_partitioned or empty" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log | cut -c1-300
A="--workload join_16Mx256M_u32_dup4_zipf --steps 8 --warmup 3 --no-cpu-baseline --no-e2e"
best=0; bestms=1000000
for sh in 0 1 2 3; do
  DWJ_MULTI_SHAPE=$sh timeout 150 python bench.py $A > $O/shape$sh.json 2> $O/shape$sh.err || { echo "shape $sh failed"; tail -2 $O/shape$sh.err; continue; }
  ms=$(python -c "import json; d=json.loads(open('$O/shape$sh.json').read()); print(d['ms_per_step'], d['phases_ms']['probe'], d['roofline']['kernel_ms'], round(d['roofline']['frac'],3))")
  echo "shape $sh: $ms"
  k=$(python -c "import json; d=json.loads(open('$O/shape$sh.json').read()); print(int(d['roofline']['kernel_ms']*1000))")
  if [ "$k" -lt "$bestms" ]; then bestms=$k; best=$sh; fi
done
echo "best shape $best ($bestms us)"
DWJ_MULTI_SHAPE=$best DWJ_PARTITION_MIN_MB=100000 timeout 150 python bench.py $A > $O/nopart.json 2> $O/nopart.err
python -c "import json; d=json.loads(open('$O/nopart.json').read()); print('no partition:', d['ms_per_step'], d['phases_ms'], d['roofline']['kernel_ms'])"
