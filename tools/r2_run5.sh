#!/bin/bash
O=gpurun_out/r2_run5; mkdir -p $O; rm -f $O/*
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_join.py -m gpu -q --timeout 150 --maxfail=5 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log; tail -4 $O/pytest.log | cut -c1-300
for W in 8 4; do echo "## key bytes $W shape 4" >> $O/sweep4.txt; DWJ_SCATTER_SHAPE=4 timeout 120 python tools/partition_sweep.py --rows $((1<<28)) --key-bytes $W --parts 32 128 512 >> $O/sweep4.txt 2>&1; done
timeout 400 python bench.py --steps 3 --warmup 2 --no-sub-configs --no-e2e --no-cpu-baseline > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "bench rc=$?"; tail -3 $O/bench_cfg5.err | cut -c1-400
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_run5/bench_cfg5.json').read())
print(d['value'], d['ms_per_step'], d['config']['probe_chunks'], d['timeline_ms_last_step_max_over_ranks'], {k:(v['ms'],v['gbs']) for k,v in d['kernels_last_launch'].items()})
PY
# ncu: the many-way scatter, 8-byte keys 512-way and 4-byte keys 128-way, one launch each
timeout 300 ncu --set full --import-source on -k regex:partition_scatter_many -s 2 -c 1 -o $O/ncu_scatter_u64_512 python tools/partition_sweep.py --rows $((1<<27)) --key-bytes 8 --parts 512 > $O/ncu1.log 2>&1
timeout 300 ncu --set full --import-source on -k regex:partition_scatter_many -s 2 -c 1 -o $O/ncu_scatter_u32_128 python tools/partition_sweep.py --rows $((1<<27)) --key-bytes 4 --parts 128 > $O/ncu2.log 2>&1
ls -la $O
