set -x
mkdir -p gpurun_out
for v in nt96 nt64; do
FMGPU_LIB=$PWD/build/libfmgpu_$v.so timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r33_bench_$v.json 2> gpurun_out/r33_bench_$v.err
done
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r33_bench.json 2> gpurun_out/r33_bench.err
for v in r33_bench r33_bench_nt96 r33_bench_nt64; do python - $v <<'PY'
import json,sys
try:
    d=json.loads(open(f'gpurun_out/{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), d["stage_ms"])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
