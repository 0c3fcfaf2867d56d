set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r9_pytest.log
tail -5 gpurun_out/r9_pytest.log
python __graft_entry__.py smoke > gpurun_out/r9_smoke.log 2>&1; tail -2 gpurun_out/r9_smoke.log
CMD="python bench.py --steps 2 --warmup 3 --channels 1250 --blocks 4 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/r9_plain_small.json 2> gpurun_out/r9_plain_small.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/r9_launches.csv $CMD > gpurun_out/r9_ncu_list.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_decim|k_fir_pair|k_chanfir|k_rds|k_stereo|k_agc|k_dcblock|k_freqdem|k_resample|k_audio_iir' -s 100 -c 26 -o gpurun_out/r9_top $CMD > gpurun_out/r9_ncu_full.log 2>&1
tail -3 gpurun_out/r9_ncu_full.log
