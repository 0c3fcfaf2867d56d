# ncu passes on the DEFAULT bench command (10,000 channels x 2 blocks): launch list + a full capture of
# the dominant kernel (k_decim) and the top lane kernel (k_stereo)
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/r43_plain_default.json 2> gpurun_out/r43_plain_default.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/r43_launches_default.csv $CMD > gpurun_out/r43_ncu_list.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_decim|k_stereo|k_fir_pair|k_chanfir' -s 20 -c 5 -o gpurun_out/r43_top_default $CMD > gpurun_out/r43_ncu_full.log 2>&1
tail -2 gpurun_out/r43_ncu_full.log
