# Round-2 final measurements on one B200: tests, smoke, benches, timeline, ncu launch list + full
# capture of the top kernels. Everything lands in gpurun_out/ (copied to profiles/ afterwards).
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest.log
tail -3 gpurun_out/r02_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
timeout 600 python bench.py --decim-mode fp32 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_reference_order.json 2> gpurun_out/r02_bench_reference_order.err
FAST="FMGPU_DECIM_MODE=1 FMGPU_SCAN_MODE=1 FMGPU_FIR_MODE=1 FMGPU_DEMOD_MODE=1"
env $FAST timeout 300 python tools/timeline.py --steps 3 > gpurun_out/r02_timeline.json 2> gpurun_out/r02_timeline_10000ch.txt
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
$CMD > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file gpurun_out/r02_launches_ncu.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
python tools/ncu_summarize.py launches gpurun_out/r02_launches_ncu.csv > gpurun_out/r02_launches_ncu.md 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_decim_tc|k_fir_tc|k_dcblock_scan|k_rds|k_stereo|k_resample|k_audio_iir_scan|k_blocksync' -s 40 -c 16 -f -o gpurun_out/r02_top $CMD > gpurun_out/r02_ncu_full.log 2>&1
tail -2 gpurun_out/r02_ncu_full.log
python tools/ncu_summarize.py full gpurun_out/r02_top.ncu-rep gpurun_out/r02_top_kernels_ncu.csv > gpurun_out/r02_ncu_summary.log 2>&1
for f in gpurun_out/r02_bench_default.json gpurun_out/r02_bench_reference_arm.json gpurun_out/r02_bench_reference_order.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), d.get("e2e") and round(d["e2e"]["value"]), d.get("cpu_baseline") and round(d["cpu_baseline"]["value"],1), d.get("clocks"), d.get("stage_ms"))
except Exception as ex:
    print(sys.argv[1], "ERR", ex)
PY
done
