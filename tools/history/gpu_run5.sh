set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r5_pytest.log
tail -25 gpurun_out/r5_pytest.log
timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r5_a_ring.json 2> gpurun_out/r5_a.err
FMGPU_DECIM_RING=0 timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r5_b_noring.json 2> gpurun_out/r5_b.err
timeout 300 python tools/timeline.py --steps 3 > gpurun_out/r5_timeline.json 2> gpurun_out/r5_timeline.txt
for f in gpurun_out/r5_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), d["e2e"] and round(d["e2e"]["value"]), d["stage_ms"])
except Exception as ex:
    print(sys.argv[1], "ERR", ex)
PY
done
tail -3 gpurun_out/r5_timeline.txt | cut -c1-3500
tail -3 gpurun_out/r5_a.err
