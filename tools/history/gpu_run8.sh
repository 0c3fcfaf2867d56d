set -x
mkdir -p gpurun_out
R=/root/repo/fmtuner_sdr_b200
for v in st32 lt64st32; do
  export FMGPU_LIB=$R/libfmgpu_$v.so
  timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r8_$v.json 2> gpurun_out/r8_$v.err
done
unset FMGPU_LIB
FMGPU_RING_K=4 timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r8_k4.json 2> gpurun_out/r8_k4.err
timeout 300 python bench.py --no-cpu-baseline --no-e2e --blocks 4 > gpurun_out/r8_b4.json 2> gpurun_out/r8_b4.err
timeout 300 python bench.py --no-cpu-baseline --no-e2e --blocks 1 > gpurun_out/r8_b1.json 2> gpurun_out/r8_b1.err
timeout 300 python bench.py --no-cpu-baseline --no-e2e --channels 20000 --blocks 2 > gpurun_out/r8_c20k.json 2> gpurun_out/r8_c20k.err
for f in gpurun_out/r8_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), d["stage_ms"], round(sum(d["stage_ms"].values()),2))
except Exception as ex:
    print(sys.argv[1], "ERR", ex)
PY
done
