# second form of the fused level meter (loads batched before the sums): check + timing only
mkdir -p gpurun_out
V=$PWD/build/libfmgpu_fusedlevel.so
FMGPU_LIB=$V timeout 150 python tools/fused_level_check.py > gpurun_out/fl2_fused.json 2> gpurun_out/fl2_fused.err; echo "check fused rc=$?"
cat gpurun_out/fl2_fused.json
B="python bench.py --decim-mode fp32 --no-cpu-baseline --no-e2e --no-extras --steps 6"
FMGPU_LIB=$V timeout 120 $B > gpurun_out/fl2_bench_fused.json 2> gpurun_out/fl2_bench_fused.err; echo "bench fused rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/fl2_bench_fused.json").read().strip().splitlines()[-1])
print("fused step", round(d["ms_per_step"], 3), "decimate", round(d["stage_ms"]["decimate"], 3))
PY
