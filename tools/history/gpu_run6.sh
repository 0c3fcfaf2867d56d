set -x
mkdir -p gpurun_out
export FMGPU_DECIM_RING=1
CMD="python bench.py --steps 2 --warmup 3 --channels 1250 --blocks 4 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/r6_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_decim|k_fir_pair|k_chanfir|k_rds|k_stereo' -s 60 -c 12 -o gpurun_out/r6_top $CMD > gpurun_out/r6_ncu.log 2>&1
tail -5 gpurun_out/r6_ncu.log
cat gpurun_out/r6_plain.log | cut -c1-1500
