# more channels per GPU: the lane kernels are latency-bound, the throughput kernels scale with the work
mkdir -p gpurun_out
for c in 20000 40000; do
  timeout 400 python bench.py --channels $c --no-cpu-baseline --no-e2e --no-extras > gpurun_out/occ_${c}.json 2> gpurun_out/occ_${c}.err
  tail -2 gpurun_out/occ_${c}.err
  python - gpurun_out/occ_${c}.json $c <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("channels", sys.argv[2], "->", round(d["value"]), "MS/s", round(d["ms_per_step"],2), "ms/step", round(d["realtime_channels"]), d["stage_ms"])
except Exception as ex:
    print("ERR", ex)
PY
done
