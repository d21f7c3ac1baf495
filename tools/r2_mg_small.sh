#!/bin/bash
N=$1
O=gpurun_out/r2_mgf$N; mkdir -p $O; rm -f $O/*
export DWJ_XJ_TIMEOUT_MS=15000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29529"
( time timeout 420 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/driver_default.json 2> $O/driver_default.err ) 2> $O/time.txt; echo "rc=$? $(grep real $O/time.txt)"
python - $O/driver_default.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print('  ', round(d['value']/1e9,2),'G tuples/s', round(d['ms_per_step'],3),'ms', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d.get('timeline_ms_last_step_max_over_ranks',{}).items() if k!='note'}, 'e2e', d.get('e2e',{}).get('value'))
PY
if [ "$N" = "2" ]; then timeout 500 python -m pytest tests -m gpu -q -k "two_ranks or two_gpus" --timeout 200 2>&1 | tail -3; fi
if [ "$N" = "2" ]; then
for c in 1 4 8; do
  DWJ_XJ_COPY_CTAS=$c timeout 200 $TR bench.py --gpus $N --steps 3 --warmup 2 --no-e2e > $O/copy$c.json 2> $O/copy$c.err
  python - $O/copy$c.json $c <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print('copy ctas/SM', sys.argv[2], round(d['ms_per_step'],2),'ms', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d.get('timeline_ms_last_step_max_over_ranks',{}).items() if k!='note'})
PY
done
fi
