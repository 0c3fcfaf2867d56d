# round-1 final measurements: tests, smoke, default bench (+cpu baseline), reference arm, timeline,
# ncu launch list + full capture of the small profiling command
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r43_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r43_pytest.log
tail -4 gpurun_out/r43_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r43_smoke.log 2>&1; tail -2 gpurun_out/r43_smoke.log
timeout 900 python bench.py > gpurun_out/r43_bench_default.json 2> gpurun_out/r43_bench_default.err
tail -2 gpurun_out/r43_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r43_bench_reference.json 2> gpurun_out/r43_bench_reference.err
timeout 300 python bench.py --sync-steps --no-cpu-baseline > gpurun_out/r43_bench_sync.json 2> gpurun_out/r43_bench_sync.err
timeout 300 python tools/timeline.py --steps 3 > gpurun_out/r43_timeline.json 2> gpurun_out/r43_timeline.txt
CMD="python bench.py --steps 2 --warmup 3 --channels 1250 --blocks 4 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/r43_bench_small.json 2> gpurun_out/r43_bench_small.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/r43_launches.csv $CMD > gpurun_out/r43_ncu_list.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_decim|k_fir_pair|k_chanfir|k_rds|k_stereo|k_agc|k_dcblock|k_freqdem|k_resample|k_audio_iir|k_blocksync' -s 100 -c 26 -o gpurun_out/r43_top $CMD > gpurun_out/r43_ncu_full.log 2>&1
tail -2 gpurun_out/r43_ncu_full.log
for f in gpurun_out/r43_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), d.get("e2e") and round(d["e2e"]["value"]), d.get("cpu_baseline") and round(d["cpu_baseline"]["value"],1), d.get("clocks"), d.get("stage_ms"))
except Exception as ex:
    print(sys.argv[1], "ERR", ex)
PY
done
