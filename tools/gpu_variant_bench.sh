# usage: bash tools/gpu_variant_bench.sh <variant> ... : the quick default bench with the shipped library and
# with build/libfmgpu_<variant>.so (fmtuner_sdr_b200.build.build_variant), step + stage times alone
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 120 python bench.py --no-cpu-baseline --no-e2e --no-extras > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err
  python - gpurun_out/var_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "->", round(d["value"]), round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms"].items()})
except Exception as ex:
    print("ERR", sys.argv[2], ex)
PY
}
run base A=1
for v in "$@"; do
  run $v FMGPU_LIB=$PWD/build/libfmgpu_$v.so
done
