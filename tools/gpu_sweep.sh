set -x
mkdir -p gpurun_out
FMGPU_DECIM_MODE=1 FMGPU_SCAN_MODE=1 FMGPU_FIR_MODE=1 FMGPU_DEMOD_MODE=1 timeout 300 python tools/timeline.py --steps 3 > gpurun_out/tl_timeline.json 2> gpurun_out/tl_timeline.txt
tail -3 gpurun_out/tl_timeline.txt
timeout 1500 python -m pytest tests/test_gpu_fulllength.py -q -k "fast or 1" > gpurun_out/sweep_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/sweep_pytest.log
tail -5 gpurun_out/sweep_pytest.log
