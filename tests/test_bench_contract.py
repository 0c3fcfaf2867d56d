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


def test_roofline_record_shape_for_both_flavours():
    """The roofline object of the bench line is a pure function of the measured stage times: check its
    contract keys on the stage times of the committed round-2 runs (no GPU needed). The kernel named at
    the top is the dominant throughput-bound one against the roof that binds it; the largest HBM-bound
    kernel is reported beside it in the contract's own terms (bound "hbm", GB/s, MEASURED_PEAKS)."""
    import bench
    fast = {"decimate": 1.126, "dcblock": 0.505, "chan_demod": 0.658, "rds_resample": 1.246, "rds": 1.871,
            "rds_sync": 0.137, "pilot_fir": 0.494, "stereo_pll": 3.517, "audio_lpf": 0.533, "afpost": 0.628,
            "commit": 0.007}
    ref = {"decimate": 4.019, "dcblock": 0.865, "chanfir": 1.169, "agc": 2.60, "freqdem": 0.423,
           "rds_resample": 1.249, "rds": 2.013, "rds_sync": 0.116, "pilot_fir": 1.682, "stereo_pll": 3.455,
           "audio_lpf": 1.557, "afpost": 0.924, "commit": 0.007}
    for acc, mode, step_ms in ((fast, "tc", 7.92), (ref, "fp32", 14.84)):
        r = bench.roofline_record(dict(acc), 10000, 2, 74.2, mode, step_ms)
        for key in ("kernel", "bound", "achieved", "peak", "unit", "frac", "traffic", "kernels",
                    "dominant_lane_kernel", "largest_hbm_bound_kernel"):
            assert key in r, key
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.0 < r["frac"] < 1.0
        h = r["largest_hbm_bound_kernel"]
        assert h["bound"] == "hbm" and h["unit"] == "GB/s" and h["peak"] > 1000.0
        assert abs(h["frac"] - h["achieved"] / h["peak"]) < 1e-9 and 0.0 < h["frac"] < 1.0
        assert r["dominant_lane_kernel"]["kernel"] == "stereo_pll"
    r = bench.roofline_record(dict(fast), 10000, 2, 74.2, "tc", 7.92)
    assert r["kernel"] == "rds_resample" and r["bound"] == "lsu"
    assert r["largest_hbm_bound_kernel"]["kernel"] == "decimate"
    assert 0.55 < r["largest_hbm_bound_kernel"]["frac"] < 0.70
    r = bench.roofline_record(dict(ref), 10000, 2, 74.2, "fp32", 14.84)
    assert r["kernel"] == "decimate" and r["bound"] == "fp32"
