# usage: bash tools/gpu_stage_probe.sh "<ENV=VAL ...>" ... : the quick bench once per environment setting
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 400 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/probe_$i.json 2> gpurun_out/probe_$i.err
  python - "$envs" gpurun_out/probe_$i.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print(sys.argv[1], "| step", round(d["ms_per_step"],3), "|", {k: round(v,3) for k,v in d["stage_ms"].items()})
PY
done
