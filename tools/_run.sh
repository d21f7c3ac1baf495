python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for wl in join_16Mx256M_u32_unique join_256Mx256M_u32_unique join_512Mx1G_u64_unique join_16Mx256M_u32_dup4_zipf; do
  python bench.py --workload $wl --steps 4 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ld_$wl.json 2> gpurun_out/ld.err
done
DWJ_KEEP_LOOKBACK=1 python bench.py --workload join_16Mx256M_u32_unique --steps 4 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ld_lookback_join_16Mx256M_u32_unique.json 2> gpurun_out/ld.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ld_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["phases_ms"].items()}, round(d["roofline"]["kernel_ms"],3), d["config"]["table_regions"])
    except Exception as ex: print(f, "ERR", ex)
PY
tail -3 gpurun_out/ld.err
