set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1_pytest.log
tail -5 gpurun_out/r1_pytest.log
FMGPU_DECIM_VARIANT=0 FMGPU_FIR_R=8 timeout 300 python bench.py --sync-steps --no-cpu-baseline > gpurun_out/r1_a_old_sync.json 2> gpurun_out/r1_a.err
timeout 300 python bench.py --sync-steps --no-cpu-baseline > gpurun_out/r1_b_new_sync.json 2> gpurun_out/r1_b.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r1_c_new_async8.json 2> gpurun_out/r1_c.err
timeout 300 python bench.py --no-cpu-baseline --groups 16 > gpurun_out/r1_d_new_async16.json 2> gpurun_out/r1_d.err
timeout 300 python bench.py --no-cpu-baseline --groups 4 --no-e2e > gpurun_out/r1_e_new_async4.json 2> gpurun_out/r1_e.err
for f in gpurun_out/r1_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],2), d["e2e"] and round(d["e2e"]["value"]), d["stage_ms"])
except Exception as ex:
    print(sys.argv[1], "ERR", ex)
PY
done
