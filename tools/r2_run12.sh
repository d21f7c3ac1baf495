#!/bin/bash
O=gpurun_out/r2_run12; mkdir -p $O; rm -f $O/*
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 150 --maxfail=3 -k "compact or virtual_ranks" 2>&1 | tail -3 | cut -c1-400
run() { tag=$1; shift; timeout 200 python bench.py --steps 3 --warmup 2 --no-sub-configs --no-e2e --no-cpu-baseline "$@" > $O/$tag.json 2> $O/$tag.err; tail -2 $O/$tag.err | cut -c1-300; python - "$O/$tag.json" $tag <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read()); print(sys.argv[2], round(d['ms_per_step'],2), d['timeline_ms_last_step_max_over_ranks'], {k:round(v['ms'],2) for k,v in d['kernels_last_launch'].items()})
except Exception as ex: print(sys.argv[2], 'fail', ex)
PY
}
run local
