# SURVEY §8(f) row 2 as worded, measured: RF level sums inside the FP32 decimator's tile fill
# (build/libfmgpu_fusedlevel.so = build_variant('fusedlevel', ['FMGPU_EXP_FUSED_LEVEL'])) against the
# shipped arrangement (k_decim + the separate k_siglevel pass). Then smoke() and the default bench
# of the shipped library.
mkdir -p gpurun_out
V=$PWD/build/libfmgpu_fusedlevel.so
timeout 150 python tools/fused_level_check.py > gpurun_out/fl_shipped.json 2> gpurun_out/fl_shipped.err; echo "check shipped rc=$?"
FMGPU_LIB=$V timeout 150 python tools/fused_level_check.py > gpurun_out/fl_fused.json 2> gpurun_out/fl_fused.err; echo "check fused rc=$?"
cat gpurun_out/fl_shipped.json gpurun_out/fl_fused.json
B="python bench.py --decim-mode fp32 --no-cpu-baseline --no-e2e --no-extras --steps 6"
timeout 120 $B > gpurun_out/fl_bench_shipped.json 2> gpurun_out/fl_bench_shipped.err; echo "bench shipped rc=$?"
FMGPU_LIB=$V timeout 120 $B > gpurun_out/fl_bench_fused.json 2> gpurun_out/fl_bench_fused.err; echo "bench fused rc=$?"
python - <<'PY'
import json
for t in ("shipped", "fused"):
    try:
        d = json.loads(open(f"gpurun_out/fl_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, "step", round(d["ms_per_step"], 3), "decimate", round(d["stage_ms"]["decimate"], 3))
    except Exception as ex:
        print("ERR", t, ex)
PY
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python bench.py > gpurun_out/fl_final_bench.json 2> gpurun_out/fl_final_bench.err; echo "final bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/fl_final_bench.json").read().strip().splitlines()[-1])
print("final", round(d["value"]), "MS/s", round(d["ms_per_step"], 3), "ms/step; e2e", d.get("e2e", {}).get("value"))
PY
