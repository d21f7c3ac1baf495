#!/bin/bash
# final multi-GPU session: the driver's command, then the stream-priority A/B on both workloads
N=${1:-8}
O=gpurun_out/r2_mgf$N; mkdir -p $O; rm -f $O/*
export DWJ_XJ_TIMEOUT_MS=15000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527"
run() { tag=$1; shift; ( time timeout 420 $TR bench.py --gpus $N "$@" > $O/$tag.json 2> $O/$tag.err ) 2> $O/$tag.time; echo "== $tag rc=$? $(grep real $O/$tag.time)"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|NCCL version" $O/$tag.err | tail -3 | cut -c1-400; python - "$O/$tag.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print('  ', round(d['value']/1e9,2),'G tuples/s', round(d['ms_per_step'],3),'ms', 'chunks',d['config'].get('probe_chunks'), {k:(round(v,2) if isinstance(v,float) else v) for k,v in d.get('timeline_ms_last_step_max_over_ranks',{}).items() if k!='note'})
    print('  exch', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d.get('exchange',{}).items() if k in ('remote_bytes_pulled_per_gpu_per_step','pull_window_ms','nvlink_gbs_in_window')}, 'e2e', d.get('e2e',{}).get('value'))
    print('  kern', {k:(round(v['ms'],3),round(v['gbs'])) for k,v in d.get('kernels_last_launch',{}).items()})
except Exception as ex: print('  no json', ex)
PY
}
run driver_default --steps 20 --warmup 5
if [ "$N" = "8" ]; then
DWJ_XJ_PRIORITY=1 run cfg5_prio --steps 8 --warmup 3 --no-e2e
run cfg2_weak --steps 15 --warmup 4 --no-e2e --workload join_16Mx256M_u32_unique
DWJ_XJ_PRIORITY=1 run cfg2_weak_prio --steps 15 --warmup 4 --no-e2e --workload join_16Mx256M_u32_unique
DWARF_BENCH_SEED=5 timeout 120 dwarf_bench_b200/lib/dwarf_bench Join --device=gpu --gpus $N --input_size 4194304 --iterations 3 > $O/cli.txt 2>&1; echo "cli rc=$?"; tail -4 $O/cli.txt
fi
