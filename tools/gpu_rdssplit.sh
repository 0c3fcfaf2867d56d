set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_streaming.py tests/test_gpu_stage_api.py -x -q > gpurun_out/rs_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/rs_pytest.log
tail -5 gpurun_out/rs_pytest.log
timeout 400 python bench.py --no-cpu-baseline --no-e2e --no-extras > gpurun_out/rs_bench.json 2> gpurun_out/rs_bench.err
tail -3 gpurun_out/rs_bench.err
python - gpurun_out/rs_bench.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), d["stage_ms"], d.get("decoded"))
except Exception as ex:
    print("ERR", ex)
PY
FMGPU_DECIM_MODE=1 FMGPU_SCAN_MODE=1 FMGPU_FIR_MODE=1 FMGPU_DEMOD_MODE=1 timeout 300 python tools/timeline.py --steps 3 > gpurun_out/rs_timeline.json 2> gpurun_out/rs_timeline.txt
tail -3 gpurun_out/rs_timeline.txt
