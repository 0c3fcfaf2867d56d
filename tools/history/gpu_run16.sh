set -x
mkdir -p gpurun_out
./tools/microbench/lat_bench > gpurun_out/r16_lat_bench.log 2>&1; cat gpurun_out/r16_lat_bench.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r16_pytest.log
tail -6 gpurun_out/r16_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r16_bench.json 2> gpurun_out/r16_bench.err
tail -3 gpurun_out/r16_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r16_bench.json').read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],2), d["e2e"] and round(d["e2e"]["value"]), d["stage_ms"])
print(d["roofline"]["streaming_kernels"])
PY
