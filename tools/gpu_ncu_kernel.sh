# usage: bash tools/gpu_ncu_kernel.sh <kernel-regex> <tag> [extra bench args]
# one ncu --set full capture of the kernels matching the regex inside a short default-size bench run
K=$1; TAG=$2; shift 2
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline $*"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 2 -o gpurun_out/${TAG} -f $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
