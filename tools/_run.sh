N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/final_mg${N}.json 2> gpurun_out/final_mg${N}.err
python - <<PY
import json
d=json.load(open("gpurun_out/final_mg${N}.json")); print(round(d["value"]/1e9,1), "G/s", round(d["ms_per_step"],3), "ms", d.get("e2e",{}).get("ms_per_step"))
PY
