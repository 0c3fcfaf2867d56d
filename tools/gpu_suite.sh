# whole GPU test-suite + timeline of the default bench configuration (fast arithmetic)
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/suite_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/suite_pytest.log
tail -5 gpurun_out/suite_pytest.log
FMGPU_DECIM_MODE=1 FMGPU_SCAN_MODE=1 FMGPU_FIR_MODE=1 timeout 300 python tools/timeline.py --steps 3 > gpurun_out/suite_timeline.json 2> gpurun_out/suite_timeline.txt
tail -3 gpurun_out/suite_timeline.txt
