set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r21_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r21_pytest.log
tail -6 gpurun_out/r21_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r21_bench.json 2> gpurun_out/r21_bench.err
FMGPU_DECIM4=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r21_bench_decim4.json 2> gpurun_out/r21_bench_decim4.err
FMGPU_LIB=$PWD/build/libfmgpu_lane8.so timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r21_bench_lane8.json 2> gpurun_out/r21_bench_lane8.err
tail -3 gpurun_out/r21_bench_lane8.err
for v in r21_bench r21_bench_decim4 r21_bench_lane8; do python - $v <<'PY'
import json,sys
try:
    d=json.loads(open(f'gpurun_out/{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), d["stage_ms"])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
