#!/bin/bash
O=gpurun_out/r2_run8; mkdir -p $O; rm -f $O/*
run() { tag=$1; shift; timeout 200 python bench.py --steps 2 --warmup 2 --no-sub-configs --no-e2e --no-cpu-baseline "$@" > $O/$tag.json 2> $O/$tag.err; python - "$O/$tag.json" $tag <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read()); print(sys.argv[2], round(d['ms_per_step'],2), {k:round(v['ms'],2) for k,v in d['kernels_last_launch'].items()})
except Exception as ex: print(sys.argv[2], 'fail', ex)
PY
}
run base
DWJ_SCATTER_SHAPE_FILTERED=0 run filt_shape0
DWJ_SCATTER_SHAPE_FILTERED=1 run filt_shape1
DWJ_SCATTER_SHAPE_FILTERED=2 run filt_shape2
DWJ_BUILD_NO_AHEAD=1 run no_ahead
