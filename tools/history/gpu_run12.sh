set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r12_pytest.log
tail -6 gpurun_out/r12_pytest.log
timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r12_fused.json 2> gpurun_out/r12_fused.err
FMGPU_FUSE_AGC_FD=0 timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r12_unfused.json 2> gpurun_out/r12_unfused.err
for f in gpurun_out/r12_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), d["stage_ms"], round(sum(d["stage_ms"].values()),2))
except Exception as ex:
    print(sys.argv[1], "ERR", ex)
PY
done
