N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/final_mg${N}.json 2> gpurun_out/final_mg${N}.err
$TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --exchange-chunks 8 > gpurun_out/final_mg${N}_c8.json 2> gpurun_out/final_mg${N}_c8.err
python - <<PY
import json
for f in ("gpurun_out/final_mg${N}.json","gpurun_out/final_mg${N}_c8.json"):
    d=json.load(open(f)); print(f, round(d["value"]/1e9,1), "G/s", round(d["ms_per_step"],3), "ms", d.get("e2e",{}).get("ms_per_step"))
PY
