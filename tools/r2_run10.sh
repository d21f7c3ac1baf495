#!/bin/bash
O=gpurun_out/r2_run10; mkdir -p $O; rm -f $O/*
timeout 500 python -m pytest tests/test_gpu_join.py -m gpu -q --timeout 200 --maxfail=5 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log; tail -5 $O/pytest.log | cut -c1-300
for w in join_16Mx256M_u32_dup4_zipf join_512Mx1G_u64_unique join_256Mx256M_u32_unique join_16Mx256M_u32_unique; do
  timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/$w.json 2> $O/$w.err; python - $O/$w.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read()); print(d['config']['workload'], round(d['ms_per_step'],3), d['phases_ms'], round(d['roofline']['kernel_ms'],3), round(d['roofline']['frac'],3))
except Exception as ex: print('fail', ex)
PY
done
