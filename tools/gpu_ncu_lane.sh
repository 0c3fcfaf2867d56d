# ncu --set full + source counters of the two lane kernels in the default bench configuration
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_stereo|k_rds$' -s 6 -c 2 -f -o gpurun_out/lane $CMD > gpurun_out/lane_ncu.log 2>&1
tail -3 gpurun_out/lane_ncu.log
ncu -i gpurun_out/lane.ncu-rep --page source --csv > gpurun_out/lane_src.csv 2>/dev/null
ncu -i gpurun_out/lane.ncu-rep --page raw --csv > gpurun_out/lane_raw.csv 2>/dev/null
ls -la gpurun_out/lane*
