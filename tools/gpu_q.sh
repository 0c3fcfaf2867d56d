mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-extras > gpurun_out/q.json 2> gpurun_out/q.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/q.json').read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],2), d["stage_ms"])
PY
