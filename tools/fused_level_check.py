"""Measurement of SURVEY §8(f) row 2 as worded: the RF level meter's sums taken INSIDE the FP32
decimator's tile fill (measurement variant build/libfmgpu_fusedlevel.so, -DFMGPU_EXP_FUSED_LEVEL)
against the shipped arrangement (k_decim + the separate k_siglevel pass).

Run with FMGPU_LIB pointing at the variant: checks that the fused sums equal numpy's (every sample
counted exactly once, clip counters included), then times one logical block of 10,000 channels:
the fused decimator, the plain decimator (timed from the shipped library by the caller:
FMGPU_LIB unset -> `mode: shipped`) and the separate level pass. Prints one JSON line.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import fmtuner_sdr_b200.engine as fm  # noqa: E402


def sums_numpy(iq: np.ndarray):
    i = iq[0::2].astype(np.int64)
    q = iq[1::2].astype(np.int64)
    hard = int(np.count_nonzero((i <= 1) | (i >= 254) | (q <= 1) | (q >= 254)))
    near = int(np.count_nonzero((i <= 8) | (i >= 247) | (q <= 8) | (q >= 247)))
    return dict(sum_i=int(i.sum()), sum_q=int(q.sum()), sum_ii=int((i * i).sum()),
                sum_qq=int((q * q).sum()), hard_clip=hard, near_clip=near, n_samples=len(i))


def main():
    L = fm.load_library(build=False)
    fused = hasattr(L, "fmgpu_exp_fused_level_read") and bool(os.environ.get("FMGPU_LIB"))
    out = {"mode": "fused" if fused else "shipped", "lib": os.environ.get("FMGPU_LIB", "libfmgpu.so")}
    FIELDS = ("sum_i", "sum_q", "sum_ii", "sum_qq", "hard_clip", "near_clip", "n_samples")

    if fused:
        L.fmgpu_exp_fused_level_read.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.fmgpu_exp_fused_level_read.restype = C.c_int
        # ---- exactness: 37 channels x 3 blocks in one call, then 2 more blocks (history carried)
        C_, B = 37, 3
        eng = fm.Engine(fm.make_config(max_blocks=B), C_, 0)
        n_iq = B * 81920
        rng = np.random.default_rng(4)
        host = np.clip(np.rint(rng.normal(127.5, 70.0, (C_, 2 * n_iq))), 0, 255).astype(np.uint8)
        iq = torch.from_numpy(host).cuda()
        buf = (fm.LevelSums * C_)()
        L.fmgpu_exp_fused_level_read(buf, C_, 1)  # clear
        eng.process_batch(iq.data_ptr(), 2 * n_iq, B, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        eng.process_batch(iq.data_ptr(), 2 * n_iq, 2, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()  # (the first two blocks again: the decimator's history is carried)
        L.fmgpu_exp_fused_level_read(buf, C_, 1)
        bad = 0
        for c in range(C_):
            w1 = sums_numpy(host[c])
            w2 = sums_numpy(host[c, :2 * 2 * 81920])
            for f in FIELDS:
                if getattr(buf[c], f) != w1[f] + w2[f]:
                    bad += 1
        out["exact_channels"] = C_
        out["exact_mismatches"] = bad
        eng.close()
        del iq

    # ---- timing: 10,000 channels, one logical block per launch, stages alone
    C_, B = 10000, 2
    eng = fm.Engine(fm.make_config(max_blocks=B), C_, 0)
    eng.set_stage_overlap(False)
    eng.enable_stage_timing(True)
    n_iq = B * 81920
    g = torch.Generator(device="cuda").manual_seed(1)
    iq = torch.randint(40, 216, (C_, 2 * n_iq), dtype=torch.uint8, device="cuda", generator=g)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        eng.process_batch(iq.data_ptr(), 2 * n_iq, B, stream=stream)
        torch.cuda.synchronize()
    reps, acc = 6, {}
    for _ in range(reps):
        eng.process_batch(iq.data_ptr(), 2 * n_iq, B, stream=stream)
        torch.cuda.synchronize()
        for k, v in eng.stage_times().items():   # the last call's stage times, B launches each
            acc[k] = acc.get(k, 0.0) + v / (reps * B)
    out["decimate_ms_per_launch"] = acc.get("decimate")
    out["stage_ms_per_block"] = {k: round(v, 4) for k, v in acc.items()}

    # the separate level pass, same bytes
    sums = torch.zeros((C_, B, 48), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        eng.signal_level_batch(iq.data_ptr(), 2 * n_iq, B, sums.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.signal_level_batch(iq.data_ptr(), 2 * n_iq, B, sums.data_ptr(),
                               stream=torch.cuda.current_stream().cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    out["siglevel_ms_per_block"] = e0.elapsed_time(e1) / (reps * B)
    eng.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
