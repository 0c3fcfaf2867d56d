# Bounds experiment: the default bench with lane kernels (resp. FIR kernels) doing 1/8 of their work.
# Results of these runs are wrong by construction; only ms_per_step is read.
set -x
mkdir -p gpurun_out
for v in lane8 fir8; do
  FMGPU_LIB=$PWD/build/libfmgpu_$v.so timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/exp_$v.json 2> gpurun_out/exp_$v.err
  python - "$v" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/exp_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], round(d["ms_per_step"],2), d["stage_ms"])
PY
done
