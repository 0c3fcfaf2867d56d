set -x
mkdir -p gpurun_out
R=/root/repo/fmtuner_sdr_b200
for v in default lt16 d64 lt16d64; do
  if [ $v = default ]; then unset FMGPU_LIB; else export FMGPU_LIB=$R/libfmgpu_$v.so; fi
  timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r7_$v.json 2> gpurun_out/r7_$v.err
done
export FMGPU_LIB=$R/libfmgpu_lt16d64.so
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_streaming.py -x -q > gpurun_out/r7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r7_pytest.log
tail -4 gpurun_out/r7_pytest.log
timeout 300 python tools/timeline.py --steps 3 > gpurun_out/r7_timeline.json 2> gpurun_out/r7_timeline.txt
for f in gpurun_out/r7_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), d["stage_ms"], round(sum(d["stage_ms"].values()),2))
except Exception as ex:
    print(sys.argv[1], "ERR", ex)
PY
done
tail -3 gpurun_out/r7_timeline.txt | cut -c1-3000
