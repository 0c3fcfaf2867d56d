set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chan_demod_tc.py tests/test_gpu_fir_tc.py -x -q > gpurun_out/cd_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/cd_pytest.log
tail -30 gpurun_out/cd_pytest.log
for dm in 1 0; do
  FMGPU_FIR_MODE=1 FMGPU_DEMOD_MODE=$dm timeout 400 python bench.py --no-cpu-baseline --no-e2e --no-extras > gpurun_out/cd_${dm}_bench.json 2> gpurun_out/cd_${dm}_bench.err
  tail -3 gpurun_out/cd_${dm}_bench.err
  python - gpurun_out/cd_${dm}_bench.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), d["stage_ms"], d.get("decoded"))
except Exception as ex:
    print("ERR", ex)
PY
done
