#!/bin/bash
# 2-GPU validation: oracle parity of the torchrun path, then the default bench (config 5, strong) and config 2 (weak)
O=gpurun_out/r2_mg2; mkdir -p $O; rm -f $O/*
export DWJ_XJ_TIMEOUT_MS=10000
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -k two_ranks --timeout 200 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log; tail -15 $O/pytest.log | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 2 --steps 5 --warmup 2 --no-e2e > $O/bench2_cfg5.json 2> $O/bench2_cfg5.err; echo "cfg5 rc=$?"; tail -4 $O/bench2_cfg5.err | cut -c1-500; cut -c1-300 $O/bench2_cfg5.json
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --workload join_16Mx256M_u32_unique > $O/bench2_cfg2.json 2> $O/bench2_cfg2.err; echo "cfg2 rc=$?"; tail -4 $O/bench2_cfg2.err | cut -c1-500; cut -c1-300 $O/bench2_cfg2.json
timeout 300 $TR bench.py --gpus 2 --steps 5 --warmup 2 --no-e2e --scatter-pull --workload join_16Mx256M_u32_unique > $O/bench2_cfg2_sp.json 2> $O/bench2_cfg2_sp.err; echo "cfg2 scatter-pull rc=$?"; cut -c1-200 $O/bench2_cfg2_sp.json
