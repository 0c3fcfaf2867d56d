set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_channelizer.py -x -q > gpurun_out/r10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r10_pytest.log
tail -30 gpurun_out/r10_pytest.log
