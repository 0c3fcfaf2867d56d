set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r20_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r20_pytest.log
tail -6 gpurun_out/r20_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r20_bench.json 2> gpurun_out/r20_bench.err
tail -3 gpurun_out/r20_bench.err
FMGPU_LIB=$PWD/build/libfmgpu_st8.so timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r20_bench_st8.json 2> gpurun_out/r20_bench_st8.err
for v in r20_bench r20_bench_st8; do python - $v <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), d["stage_ms"])
PY
done
