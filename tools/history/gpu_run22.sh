set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r22_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r22_pytest.log
tail -6 gpurun_out/r22_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r22_bench.json 2> gpurun_out/r22_bench.err
FMGPU_LIB=$PWD/build/libfmgpu_s6.so timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r22_bench_s6.json 2> gpurun_out/r22_bench_s6.err
FMGPU_LIB=$PWD/build/libfmgpu_s6st8.so timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r22_bench_s6st8.json 2> gpurun_out/r22_bench_s6st8.err
FMGPU_DECIM8=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r22_bench_d8.json 2> gpurun_out/r22_bench_d8.err
for v in r22_bench r22_bench_s6 r22_bench_s6st8 r22_bench_d8; do python - $v <<'PY'
import json,sys
try:
    d=json.loads(open(f'gpurun_out/{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), d["stage_ms"])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
