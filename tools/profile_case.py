"""Small fixed workload for ncu captures: C channels x B blocks, a few steps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fmtuner_sdr_b200 as fm
C = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
agc = int(sys.argv[4]) if len(sys.argv) > 4 else 1
n_iq = B * 81920
stride = 2 * n_iq
dev = torch.device("cuda", 0)
eng = fm.Engine(fm.make_config(max_blocks=B, dsp_agc=agc), C, 0)
iq = torch.empty((C, stride), dtype=torch.uint8, device=dev)
rng = np.random.default_rng(0)
params = [fm.SynthParams(75000.0, 400.0 + 37.0 * (c % 200), 0.8, 700.0 + 53.0 * (c % 150), 0.8, 0.10, 0.04,
                         0.5, float(rng.uniform(10, 40)), c, 0x1000 + c, 0) for c in range(C)]
fm.synth_iq(0, params, 2_400_000, n_iq, iq.data_ptr(), stride)
torch.cuda.synchronize()
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for _ in range(steps):
        eng.process_batch(iq.data_ptr(), stride, B, stream=st.cuda_stream)
torch.cuda.synchronize()
print("ok launches", eng.launch_count())
