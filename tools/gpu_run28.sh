set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r28_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r28_pytest.log
tail -6 gpurun_out/r28_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r28_bench.json 2> gpurun_out/r28_bench.err
timeout 600 python bench.py --steps 4 --warmup 3 --channels 1250 --blocks 4 --no-e2e --no-cpu-baseline > gpurun_out/r28_bench_small.json 2> gpurun_out/r28_bench_small.err
for v in r28_bench r28_bench_small; do python - $v <<'PY'
import json,sys
try:
    d=json.loads(open(f'gpurun_out/{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), d["stage_ms"])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
