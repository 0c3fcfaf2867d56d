# refresh the ncu --set full row of k_rds (the capture in profiles/ predates __maxnreg__(255))
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
timeout 100 $CMD > /dev/null 2>&1 && \
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'^k_rds$' -s 8 -c 2 -f -o gpurun_out/r02_rds $CMD > gpurun_out/r02_rds_ncu.log 2>&1
tail -2 gpurun_out/r02_rds_ncu.log | cut -c1-200
python tools/ncu_summarize.py full gpurun_out/r02_rds.ncu-rep gpurun_out/r02_rds_ncu.csv > gpurun_out/r02_rds_summary.log 2>&1
cat gpurun_out/r02_rds_ncu.csv | cut -c1-400
