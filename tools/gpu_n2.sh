set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r47_bench_n2.json 2> gpurun_out/r47_bench_n2.err
tail -3 gpurun_out/r47_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r47_bench_n2.json').read().strip().splitlines()[-1])
print(d["n_gpus"], round(d["value"]), round(d["ms_per_step"],2), d["e2e"] and round(d["e2e"]["value"]), d["clocks"])
PY
