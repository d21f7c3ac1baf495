#!/bin/bash
# round-end check on one B200: the whole GPU suite, smoke(), the driver's default bench command
O=gpurun_out/r2_final; mkdir -p $O; rm -f $O/*
timeout 600 python -m pytest tests -m gpu -q --timeout 250 --maxfail=5 > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err ) 2> $O/time.txt; echo "bench rc=$? $(grep real $O/time.txt)"; tail -3 $O/bench.err | cut -c1-400
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final/bench.json').read())
print('headline', round(d['value']/1e9,2), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']/1e9,3), 'cpu', round(d['cpu_baseline']['value']/1e9,3), 'launches', d.get('gpu_launches'), d.get('clocks'))
print('roofline', {k:d['roofline'].get(k) for k in ('kernel','achieved','frac','share_of_step','traffic')})
for k,v in d['configs'].items():
    print(k, round(v['value']/1e9,2),'G', round(v['ms_per_step'],3),'ms', 'roof', round(v['roofline']['frac'],3), 'traffic', v['roofline']['traffic'], 'e2e', round(v['e2e']['value']/1e9,2) if 'e2e' in v else None)
PY
