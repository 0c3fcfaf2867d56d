set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stage_api.py -x -q > gpurun_out/r11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r11_pytest.log
tail -4 gpurun_out/r11_pytest.log
timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r11_r6.json 2> gpurun_out/r11_r6.err
FMGPU_DECIM_R=4 timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r11_r4.json 2> gpurun_out/r11_r4.err
for f in gpurun_out/r11_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), d["stage_ms"], round(sum(d["stage_ms"].values()),2))
except Exception as ex:
    print(sys.argv[1], "ERR", ex)
PY
done
