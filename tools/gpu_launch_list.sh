# usage: bash tools/gpu_launch_list.sh <tag> [bench args]: ncu launch list (durations) of a short bench run
TAG=$1; shift
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras $*"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu.log 2>&1
python tools/ncu_summarize.py launches gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.md 2>&1
cat gpurun_out/${TAG}_launches.md | head -40
