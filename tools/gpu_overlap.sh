# what limits the overlap: staging depth / ring depth / decimator ring
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fir_tc.py tests/test_gpu_chan_demod_tc.py -x -q 2>&1 | tail -3
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ov_${name}.json 2> gpurun_out/ov_${name}.err
  python - gpurun_out/ov_${name}.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), d["stage_ms"])
except Exception as ex:
    print("ERR", ex)
PY
}
run base A=1
run nstg2 FMGPU_FT_NSTG=2
run nstg4 FMGPU_FT_NSTG=4
run k4 FMGPU_RING_K=4
run k6 FMGPU_RING_K=6
run k6ring3 FMGPU_RING_K=6 FMGPU_TC_RING=3 FMGPU_FT_NSTG=2
