set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r13_bench_default.json 2> gpurun_out/r13_bench_default.err
tail -2 gpurun_out/r13_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r13_bench_reference.json 2> gpurun_out/r13_bench_reference.err
timeout 300 python bench.py --sync-steps --no-cpu-baseline > gpurun_out/r13_bench_sync.json 2> gpurun_out/r13_bench_sync.err
timeout 300 python tools/timeline.py --steps 3 > gpurun_out/r13_timeline.json 2> gpurun_out/r13_timeline.txt
for f in gpurun_out/r13_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), d.get("e2e") and round(d["e2e"]["value"]), d.get("cpu_baseline") and round(d["cpu_baseline"]["value"],1), d.get("clocks"))
except Exception as ex:
    print(sys.argv[1], "ERR", ex)
PY
done
