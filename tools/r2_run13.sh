#!/bin/bash
O=gpurun_out/r2_run13; mkdir -p $O; rm -f $O/*
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_join.py -m gpu -q --timeout 150 --maxfail=3 -k "compact or filter or virtual_ranks" 2>&1 | tail -4 | cut -c1-400
timeout 200 python bench.py --steps 3 --warmup 2 --no-sub-configs --no-e2e --no-cpu-baseline > $O/local.json 2> $O/local.err; tail -2 $O/local.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_run13/local.json').read()); print(round(d['ms_per_step'],2), d['timeline_ms_last_step_max_over_ranks'], {k:(round(v['ms'],2), round(v['gbs'])) for k,v in d['kernels_last_launch'].items()}, d['roofline']['kernel'], round(d['roofline']['frac'],3))
PY
