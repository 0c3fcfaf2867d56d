# marginal cost of each stage inside the block pipeline (measurement variant of the library)
mkdir -p gpurun_out
export FMGPU_LIB=$PWD/build/libfmgpu_skip.so
for sk in none rds_resample rds_demod stereo_pll decimate dcblock chan_demod pilot_fir audio_lpf afpost "rds_demod,stereo_pll" "rds_resample,afpost,dcblock"; do
  FMGPU_SKIP=$sk timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 6 > gpurun_out/skip.json 2> gpurun_out/skip.err
  python - "$sk" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/skip.json').read().strip().splitlines()[-1])
    print("skip", sys.argv[1], "->", round(d["ms_per_step"],2), "ms/step")
except Exception as ex:
    print("ERR", sys.argv[1], ex)
PY
done
