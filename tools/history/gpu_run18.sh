set -x
mkdir -p gpurun_out
timeout 300 python tools/timeline.py --steps 4 > gpurun_out/r18_timeline.json 2> gpurun_out/r18_timeline.txt
FMGPU_LIB=$PWD/build/libfmgpu_fir8.so timeout 300 python tools/timeline.py --steps 4 > gpurun_out/r18_timeline_fir8.json 2> gpurun_out/r18_timeline_fir8.txt
tail -2 gpurun_out/r18_timeline.txt gpurun_out/r18_timeline_fir8.txt | cut -c1-3000
