set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r23_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r23_pytest.log
tail -6 gpurun_out/r23_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r23_bench.json 2> gpurun_out/r23_bench.err
for v in s6 s8 s8st8 s8lt16; do
FMGPU_LIB=$PWD/build/libfmgpu_$v.so timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r23_bench_$v.json 2> gpurun_out/r23_bench_$v.err
done
for v in r23_bench r23_bench_s6 r23_bench_s8 r23_bench_s8st8 r23_bench_s8lt16; do python - $v <<'PY'
import json,sys
try:
    d=json.loads(open(f'gpurun_out/{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), d["stage_ms"])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
