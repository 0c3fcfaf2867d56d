set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --channels 1250 --blocks 4 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/r24_plain_small.json 2> gpurun_out/r24_plain_small.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/r24_launches.csv $CMD > gpurun_out/r24_ncu_list.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_decim|k_fir_pair|k_chanfir|k_rds|k_stereo|k_agc|k_dcblock|k_freqdem|k_resample|k_audio_iir|k_blocksync' -s 100 -c 26 -o gpurun_out/r24_top $CMD > gpurun_out/r24_ncu_full.log 2>&1
tail -3 gpurun_out/r24_ncu_full.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r24_plain_small.json').read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],2), d["stage_ms"])
PY
