# can the tensor-core kernels sit beside the lane kernels' CTAs? smaller k_stereo tiles (35 KB instead of
# 59 KB per CTA) and smaller tensor-core kernel footprints
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-extras > gpurun_out/cores_$name.json 2> gpurun_out/cores_$name.err
  python - gpurun_out/cores_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "->", round(d["value"]), round(d["ms_per_step"],2), d["stage_ms"])
except Exception as ex:
    print("ERR", ex)
PY
}
run base A=1
run st8 FMGPU_LIB=$PWD/build/libfmgpu_st8.so
run st8_small FMGPU_LIB=$PWD/build/libfmgpu_st8.so FMGPU_FT_NSTG=2 FMGPU_TC_RING=3
FMGPU_LIB=$PWD/build/libfmgpu_st8.so FMGPU_FT_NSTG=2 FMGPU_TC_RING=3 timeout 300 python bench.py --channels 20000 --no-cpu-baseline --no-e2e --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(\"st8_small_20k ->\", round(d[\"value\"]), round(d[\"ms_per_step\"],2), d[\"stage_ms\"])"
