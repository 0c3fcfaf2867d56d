# quick validation: full GPU test-suite + default bench without the CPU legs
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log
tail -4 gpurun_out/q_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/q_bench.json').read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],2), d["stage_ms"])
PY
