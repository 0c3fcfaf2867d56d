mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_decim_tc.py -x -q 2>&1 | tail -4
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-extras > gpurun_out/q.json 2> gpurun_out/q.err
tail -2 gpurun_out/q.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/q.json').read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],2), d["stage_ms"], d["decoded"])
PY
