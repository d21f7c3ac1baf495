#!/bin/bash
# round-2 box facts + baseline numbers of the round-1 engine on this round's box
O=gpurun_out/r2_facts; mkdir -p $O
{ free -g; nproc; nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv; nvidia-smi topo -m; ulimit -l; } > $O/box.txt 2>&1
python - > $O/torch_facts.txt 2>&1 <<'PY'
import torch, time
print(torch.cuda.mem_get_info())
t=time.time(); p=torch.randperm(1<<31, device='cuda'); torch.cuda.synchronize(); print('randperm 2^31 ok', time.time()-t, torch.cuda.max_memory_allocated()/1e9)
del p; torch.cuda.empty_cache()
t=time.time(); x=torch.empty(40<<30, dtype=torch.uint8).pin_memory(); print('pinned 40GiB ok', time.time()-t)
PY
for w in join_512Mx1G_u64_unique join_256Mx256M_u32_unique; do
  python bench.py --workload $w --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > $O/base_$w.json 2> $O/base_$w.err
  DWJ_REGION_MB=64 python bench.py --workload $w --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > $O/r64_$w.json 2> $O/r64_$w.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_cfg4.csv python bench.py --workload join_512Mx1G_u64_unique --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_cfg4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_target.csv python bench.py --workload join_256Mx256M_u32_unique --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_target.log 2>&1
echo done
