set -x
mkdir -p gpurun_out
./tools/microbench/ffma2_bench > gpurun_out/r2_ffma2_bench.log 2>&1
cat gpurun_out/r2_ffma2_bench.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -5 gpurun_out/r2_pytest.log
FMGPU_FFMA2=0 timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2_a_scalar.json 2> gpurun_out/r2_a.err
timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2_b_ffma2.json 2> gpurun_out/r2_b.err
timeout 300 python tools/timeline.py --steps 3 > gpurun_out/r2_timeline_async8.json 2> gpurun_out/r2_timeline_async8.txt
timeout 300 python tools/timeline.py --steps 3 --sync-steps > gpurun_out/r2_timeline_sync8.json 2> gpurun_out/r2_timeline_sync8.txt
for f in gpurun_out/r2_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), d["e2e"] and round(d["e2e"]["value"]), d["stage_ms"])
except Exception as ex:
    print(sys.argv[1], "ERR", ex)
PY
done
tail -3 gpurun_out/r2_timeline_async8.txt
