python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { # name, env...
  name=$1; shift
  for wl in join_16Mx256M_u32_unique join_256Mx256M_u32_unique join_512Mx1G_u64_unique; do
    env "$@" python bench.py --workload $wl --steps 4 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/la_${name}_$wl.json 2> gpurun_out/la.err
  done
}
run both X=1
run nobuild DWJ_BUILD_NO_AHEAD=1
run noprobe DWJ_PROBE_NO_AHEAD=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/la_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["phases_ms"].items()}, round(d["roofline"]["kernel_ms"],3))
    except Exception as ex: print(f, "ERR", ex)
PY
tail -3 gpurun_out/la.err
