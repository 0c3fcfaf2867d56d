"""bench.py's reference arm runs on host cores only, so its JSON contract can be checked without a
GPU: one line, the driver's keys, `impl: reference`, a cpu_baseline describing the run and an e2e
object with zero PCIe bytes (tier framing (4) of the task statement)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--steps", "1", "--warmup", "0", "--cpu-seconds", "0.2"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "MS/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("aggregate IQ MS/s") and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("BASELINE config 5")
    cb = d["cpu_baseline"]
    # "reference": the reference's own sources (oracle/_ref/libfmref.so); "port" when it is not built
    from oracle import orc
    assert cb["kind"] == ("reference" if orc.OracleLib.have_ref("ref") else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "MS/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
