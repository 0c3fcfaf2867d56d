#!/usr/bin/env python
"""Device timeline of streamed steps: which stages of which pipeline group run when.

  python tools/timeline.py --channels 10000 --blocks 2 --groups 8 --steps 3 > timeline.json

Stage timing adds two event records per stage, so absolute times are a little longer than the
bench's; the picture of what overlaps what is the point."""
from __future__ import annotations

import argparse
import json
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels", type=int, default=10000)
    ap.add_argument("--blocks", type=int, default=2)
    ap.add_argument("--groups", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--sync-steps", action="store_true")
    args = ap.parse_args()

    import numpy as np
    import torch

    import fmtuner_sdr_b200 as fm

    C, B = args.channels, args.blocks
    n_iq = B * 8192 * 10
    stride = (2 * n_iq + 15) // 16 * 16
    dev = torch.device("cuda", 0)
    eng = fm.Engine(fm.make_config(iq_rate=2_400_000, decimation=10, max_blocks=B, dsp_agc=1), C, 0)
    eng.set_pipeline_groups(args.groups)
    iq_dev = torch.empty((C, stride), dtype=torch.uint8, device=dev)
    rng = np.random.default_rng(1)
    params = [fm.SynthParams(75000.0, 400.0 + 37.0 * (c % 200), 0.8, 700.0 + 53.0 * (c % 150), 0.8, 0.10,
                             0.04, 0.5, float(rng.uniform(10.0, 40.0)), c, 0x1000 + (c & 0xFFF), 0)
              for c in range(C)]
    fm.synth_iq(0, params, 2_400_000, n_iq, iq_dev.data_ptr(), stride)
    torch.cuda.synchronize()
    st = torch.cuda.Stream(device=dev)
    f = eng.process_batch if args.sync_steps else eng.process_batch_async

    def run(k):
        for _ in range(k):
            f(iq_dev.data_ptr(), stride, B, None, 0, None, None, 0, None, None, st.cuda_stream)
        eng.join(st.cuda_stream)
        st.synchronize()

    run(3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record()
        run(args.steps)
        e1.record()
    torch.cuda.synchronize()
    plain_ms = e0.elapsed_time(e1) / args.steps
    eng.enable_stage_timing(True)
    with torch.cuda.stream(st):
        e0.record()
        run(args.steps)
        e1.record()
    torch.cuda.synchronize()
    timed_ms = e0.elapsed_time(e1) / args.steps
    spans = eng.debug_timeline(65536)
    eng.enable_stage_timing(False)
    eng.close()
    out = {"channels": C, "blocks": B, "groups": args.groups, "steps": args.steps,
           "ms_per_step_plain": plain_ms, "ms_per_step_with_events": timed_ms,
           "spans": [{"stage": s, "group": g, "t0": round(a, 3), "t1": round(b, 3)} for s, g, a, b in spans]}
    print(json.dumps(out))
    # compact text view on stderr: per group, the stages in time order
    by_group = {}
    for s, g, a, b in spans:
        by_group.setdefault(g, []).append((a, b, s))
    for g in sorted(by_group):
        row = " ".join(f"{s}[{a:.1f}-{b:.1f}]" for a, b, s in sorted(by_group[g]))
        print(f"g{g}: {row}", file=sys.stderr)
    print(f"plain {plain_ms:.2f} ms/step, with events {timed_ms:.2f} ms/step", file=sys.stderr)


if __name__ == "__main__":
    main()
