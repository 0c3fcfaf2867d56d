# the whole GPU suite + smoke, logs into gpurun_out/
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q -s > gpurun_out/full_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/full_pytest.log
grep -E "passed|failed|rc=|channelizer,|reference application" gpurun_out/full_pytest.log | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
