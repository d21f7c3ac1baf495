N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
DWJ_XCHG_TRACE=1 $TR bench.py --gpus $N --steps 3 --warmup 2 --exchange fold --exchange-chunks 2 --no-e2e > gpurun_out/mg${N}_trace.json 2> gpurun_out/mg${N}_trace.err
grep -o "rank 0\] exchange timeline.*" gpurun_out/mg${N}_trace.err | cut -c1-700 || tail -20 gpurun_out/mg${N}_trace.err
for ch in 2 4; do
  $TR bench.py --gpus $N --steps 10 --warmup 3 --exchange fold --exchange-chunks $ch --no-e2e > gpurun_out/mg${N}_blk_$ch.json 2> gpurun_out/mg${N}_blk_$ch.err
done
$TR bench.py --gpus $N --steps 10 --warmup 3 --exchange fold --exchange-chunks 2 --exchange-layout region --no-e2e > gpurun_out/mg${N}_blk_region2.json 2> gpurun_out/mg${N}_blk_region2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/mg*_blk_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["value"]/1e9,1), "G/s", round(d["ms_per_step"],3), "ms")
    except Exception as ex: print(f, "ERR", ex)
PY
