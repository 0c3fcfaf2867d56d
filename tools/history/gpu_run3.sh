set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3_pytest.log
tail -15 gpurun_out/r3_pytest.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r3_a_g1.json 2> gpurun_out/r3_a.err
timeout 300 python bench.py --no-cpu-baseline --groups 2 > gpurun_out/r3_b_g2.json 2> gpurun_out/r3_b.err
timeout 300 python bench.py --no-cpu-baseline --sync-steps --no-e2e > gpurun_out/r3_c_sync.json 2> gpurun_out/r3_c.err
timeout 300 python tools/timeline.py --steps 3 > gpurun_out/r3_timeline.json 2> gpurun_out/r3_timeline.txt
for f in gpurun_out/r3_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), d["e2e"] and round(d["e2e"]["value"]), d["stage_ms"])
except Exception as ex:
    print(sys.argv[1], "ERR", ex)
PY
done
tail -3 gpurun_out/r3_timeline.txt | cut -c1-3000
tail -5 gpurun_out/r3_a.err
