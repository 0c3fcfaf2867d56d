# TC decimator, A operand from TMEM: shared-memory ring depth vs step time (co-residency with other kernels)
set -x
mkdir -p gpurun_out
FMGPU_TC_RING=3 timeout 400 python -m pytest tests/test_gpu_decim_tc.py -x -q > gpurun_out/ring3_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/ring3_pytest.log
tail -3 gpurun_out/ring3_pytest.log
for ring in 3 5 8 10; do
  FMGPU_TC_RING=$ring timeout 400 python bench.py --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ring${ring}_bench.json 2> gpurun_out/ring${ring}_bench.err
  python - gpurun_out/ring${ring}_bench.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), d["stage_ms"])
except Exception as ex:
    print("ERR", ex)
PY
done
