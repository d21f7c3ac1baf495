#!/bin/bash
# one-to-many (CSR) table: tests that touch engines without the unique-keys flag, BASELINE config 3, the OmniSci dwarf at
# high multiplicity, then ncu (per-kernel time + DRAM bytes of the step; one full capture of the probe kernel, exported on the box)
O=gpurun_out/r2_csr; mkdir -p $O; rm -f $O/*
timeout 420 python -m pytest tests/test_gpu_join.py tests/test_host_framework.py -m gpu -q -x --timeout 150 > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest.log | cut -c1-400
A="--workload join_16Mx256M_u32_dup4_zipf --steps 5 --warmup 3 --no-cpu-baseline"
timeout 200 python bench.py $A > $O/cfg3.json 2> $O/cfg3.err; echo "bench rc=$?"; tail -3 $O/cfg3.err | cut -c1-400
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_csr/cfg3.json').read())
    print('cfg3', round(d['value']/1e9,2), 'G', round(d['ms_per_step'],3), 'ms', d['phases_ms'], 'kernel', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3), 'e2e', d.get('e2e',{}).get('value'))
except Exception as ex: print('no json', ex)
PY
for n in 1048576 4194304; do DWARF_BENCH_SEED=3 timeout 120 dwarf_bench_b200/lib/dwarf_bench JoinOmnisci --device=gpu --input_size $n --iterations 3 2>&1 | tail -4; done
DWARF_BENCH_SEED=3 timeout 120 dwarf_bench_b200/lib/dwarf_bench HashBuild --device=gpu --input_size 268435456 --iterations 3 2>&1 | tail -3
B="--workload join_16Mx256M_u32_dup4_zipf --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/ncu_traffic_cfg3.csv python bench.py $B > $O/ncu1.log 2>&1; echo "ncu1 rc=$?"
R=/tmp/probe_multi_cfg3
timeout 240 ncu --set full --clock-control none -k regex:probe_pairs_multi -s 1 -c 1 -o $R python bench.py $B > $O/ncu2.log 2>&1; echo "ncu2 rc=$?"
ncu -i $R.ncu-rep --page details --csv > $O/ncu_probe_multi_details.csv 2>/dev/null
ls -la $O
