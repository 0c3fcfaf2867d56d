#!/usr/bin/env python
"""bench.py — throughput of the IQ -> audio + RDS hot path on N B200s of one node.

Metric (BASELINE.json): aggregate IQ MS/s demodulated (stereo + RDS); real-time channel
count = value / 2.4. One "step" = one pass of the whole pipeline (decimate, discriminate,
pilot PLL + stereo matrix, 15 kHz low-pass, resample to 32 kHz, de-emphasis, RDS down to
groups) over `blocks` logical blocks (8192 samples @ 240 kHz each) of every channel.

Workload: BASELINE config 5 — the 10,000-channel weak-signal sweep — resident on ONE GPU
(it fits: 3.3 GB of IQ per step), weak-scaled to N GPUs (channels-per-GPU fixed, every rank a
differently seeded sweep): 2.4 MS/s uint8 IQ / 10 -> 240 kHz, SNR 10-40 dB, blend mode c%3,
dsp_agc fast, synthetic multiplexes generated on the device from per-channel seeds. Channels
are independent: they are sharded across ranks with no collective on the data path.

  value    IQ samples/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e      same metric through the C-ABI host call (pinned host IQ in, audio/groups/status out)
  roofline dominant kernel: algorithmic bytes / CUDA-event time vs the measured HBM peak, plus
           its FP32 FMA rate (the FIR kernels are FP32-pipe-bound)
  stage_ms every stage's kernels alone on one stream (the RDS branch as rds_resample / rds /
           rds_sync); the dominant one is the kernel `roofline` describes
  cpu_baseline / --impl reference: the CPU oracle (restated reference pipeline, libm flavour)
           on the host cores, one channel per thread; the decode time of the threads is timed.
Multi-GPU runs give every rank NVML's ideal CPU affinity for its GPU before pinned host memory
is allocated; a rank that cannot pin its buffers takes the e2e leg off for all ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# the engine drives up to 3 streams per pipeline group: give them their own hardware queues
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

IQ_RATE = 2_400_000
DECIM = 10
BLOCK = 8192
BYTES_PER_IQ_SAMPLE_ALG = 2.0 + 8.0 * 32000.0 / IQ_RATE   # 2.107 B
FLOP_PER_IQ_SAMPLE_ALG = 269.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--channels", type=int, default=10000, help="channels per GPU")
    ap.add_argument("--blocks", type=int, default=2, help="logical blocks per step")
    ap.add_argument("--groups", type=int, default=1, help="channel ranges with their own set of stage streams (1 or 2)")
    ap.add_argument("--sync-steps", action="store_true",
                    help="join every step on the caller's stream (fmgpu_process_batch) instead of "
                         "streaming the steps (fmgpu_process_batch_async + one fmgpu_join)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            smax = mx
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            sm = [float(l.split(",")[1]) for _, l in self.lines[-3:] if len(l.split(",")) > 2] or [0.0]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores (one channel per thread)
# --------------------------------------------------------------------------------------
def cpu_arm(iq_rows, n_blocks_per_pass: int, passes: int, threads: int):
    """Each thread decodes its own channel: `passes` passes over its n_blocks_per_pass blocks.
    Returns (IQ samples processed in total, wall seconds)."""
    from oracle import orc
    lib = orc.OracleLib("libm")
    chans = [orc.Channel(lib, orc.make_config(iq_rate=IQ_RATE, decimation=DECIM, dsp_agc=1,
                                              stereo_blend=c % 3)) for c in range(threads)]

    def work(i):
        for _ in range(passes):
            chans[i].process(iq_rows[i % len(iq_rows)])

    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return threads * passes * n_blocks_per_pass * BLOCK * DECIM, dt


def host_signals(n_channels: int, n_blocks: int):
    """config-5-style channels from the oracle-side generator (CPU arm input)."""
    import numpy as np
    from oracle import orc
    rows = [None] * n_channels

    def gen(c):
        s = orc.config3_signal(c, fs_iq=IQ_RATE)
        s.snr_db = 10.0 + 30.0 * ((c * 37) % 100) / 100.0
        rows[c] = s.generate(n_blocks * BLOCK * DECIM)

    ths = [threading.Thread(target=gen, args=(c,)) for c in range(n_channels)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return np.stack(rows)


def run_reference(args, rank: int):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = cores
    iq = host_signals(min(threads, 16), args.blocks)
    # one step = every thread decodes PASSES x `blocks` logical blocks of its own channel (a
    # bounded sample of the 10,000-channel workload: one channel per host thread)
    passes = 8
    for _ in range(max(1, args.warmup)):
        cpu_arm(iq, args.blocks, 1, threads)
    samples, dt = 0, 0.0
    for _ in range(args.steps):
        s, d = cpu_arm(iq, args.blocks, passes, threads)   # d: the threads' decode time only (not the
        samples += s                                        # construction of the pipeline objects)
        dt += d
    value = samples / dt / 1e6
    sample_desc = (f"{threads} channels x {passes} passes x {args.blocks} blocks x {BLOCK * DECIM} IQ "
                   f"samples per step, {args.steps} steps, one channel per thread")
    line = {
        "impl": "reference", "metric": "aggregate IQ MS/s demodulated (stereo+RDS)", "value": value,
        "unit": "MS/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "realtime_channels": value * 1e6 / IQ_RATE,
        "config": workload_config(args, threads),
        "cpu_baseline": {"value": value, "unit": "MS/s", "cores": threads, "kind": "port",
                         "sample": sample_desc,
                         "note": "restated reference pipeline (oracle, libm flavour); liquid-dsp "
                                 "itself is not installable here"},
        "e2e": {"value": value, "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, channels_this_arm: int) -> dict:
    return {
        "workload": "BASELINE config 5 (10,000-channel weak-signal sweep) per GPU, sharded by "
                    f"channel: {args.channels} channels/GPU, 2.4 MS/s uint8 IQ /10 -> 240 kHz, "
                    "SNR 10-40 dB, blend soft/normal/aggressive by c%3, dsp_agc fast, stereo + RDS",
        "channels_per_gpu": args.channels, "blocks_per_step": args.blocks,
        "pipeline_groups": args.groups,
        "step_submission": "joined per step" if args.sync_steps else
                           "streamed (async steps, one join before the closing event)",
        "block_samples": BLOCK, "iq_rate": IQ_RATE, "decimation": DECIM,
        "channels_in_this_arm": channels_this_arm,
        "l2_policy": "inputs larger than L2 (no flush): "
                     f"{args.channels * args.blocks * BLOCK * DECIM * 2 / 1e6:.0f} MB of IQ per step per GPU",
        "parallelism": "channels sharded across ranks, no collective on the data path",
    }


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def bind_host_to_gpu(local_rank: int):
    """Multi-GPU runs: give this rank the CPU cores next to ITS GPU (NVML's ideal affinity) before any
    host buffer exists, so that the pinned buffers of the end-to-end leg are first-touched on the
    GPU's own NUMA node and the H2D copies of the ranks do not all cross one socket link. (Not at
    N = 1: there the CPU baseline wants every core.)"""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return {"count": len(cpus), "first": cpus[0], "last": cpus[-1]}
    except Exception as ex:  # affinity is an optimisation, never a requirement
        return {"error": str(ex)[:120]}


def run_b200(args, rank: int, local_rank: int, world: int):
    import numpy as np
    import torch
    import torch.distributed as dist

    import fmtuner_sdr_b200 as fm
    from fmtuner_sdr_b200 import shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_cpus = bind_host_to_gpu(local_rank) if world > 1 else None
    C, B = args.channels, args.blocks
    n_iq = B * BLOCK * DECIM
    stride = (2 * n_iq + 15) // 16 * 16

    eng = fm.Engine(fm.make_config(iq_rate=IQ_RATE, decimation=DECIM, max_blocks=B, dsp_agc=1), C,
                    local_rank)
    eng.set_pipeline_groups(args.groups)
    for c in range(C):
        if (rank * C + c) % 3 != 1:
            eng.set_blend_mode((rank * C + c) % 3, c)

    # synthetic multiplexes, generated on the device from per-channel seeds
    iq_dev = torch.empty((C, stride), dtype=torch.uint8, device=dev)
    rng = np.random.default_rng(1234 + rank)
    params = []
    for c, g in enumerate(shard.channels_of_rank(rank, world, C)):
        params.append(fm.SynthParams(
            float(rng.choice([22_500.0, 37_500.0, 50_000.0, 60_000.0, 75_000.0])),
            400.0 + 37.0 * (g % 200), 0.8, 700.0 + 53.0 * (g % 150), 0.8, 0.10, 0.04, 0.5,
            float(rng.uniform(10.0, 40.0)), g, 0x1000 + (g & 0xFFF), 0))
    fm.synth_iq(local_rank, params, IQ_RATE, n_iq, iq_dev.data_ptr(), stride)
    torch.cuda.synchronize()

    acap = eng.audio_capacity(B)
    gcap = B + 8
    audio = torch.empty((C, 2, acap), dtype=torch.float32, device=dev)
    n_audio = torch.zeros(C, dtype=torch.int32, device=dev)
    groups = torch.zeros((C, gcap, 16), dtype=torch.uint8, device=dev)
    n_groups = torch.zeros(C, dtype=torch.int32, device=dev)
    status = torch.zeros((C, B, 20), dtype=torch.uint8, device=dev)
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    # Steps are streamed the way the reference's main loop runs block after block
    # (main.cpp:992): every pipeline group orders itself after its own previous step, so one
    # step's serial kernels overlap the next step's FIR kernels; fmgpu_join puts all of it back
    # on the caller's stream before the closing event. --sync-steps joins after every step.
    def step():
        if args.sync_steps:
            eng.process_batch(iq_dev.data_ptr(), stride, B, audio.data_ptr(), acap,
                              n_audio.data_ptr(), groups.data_ptr(), gcap, n_groups.data_ptr(),
                              status.data_ptr(), stream)
        else:
            eng.process_batch_async(iq_dev.data_ptr(), stride, B, audio.data_ptr(), acap,
                                    n_audio.data_ptr(), groups.data_ptr(), gcap,
                                    n_groups.data_ptr(), status.data_ptr(), stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    eng.join(stream)
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    l0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step()
    eng.join(stream)
    ev1.record()
    barrier()
    w1 = time.time()
    launches = eng.launch_count() - l0
    clocks = sampler.stop(w0, w1)
    ms = shard.max_over_ranks(ev0.elapsed_time(ev1), dev)
    samples_per_step_rank = C * n_iq
    value = shard.aggregate_throughput(samples_per_step_rank, args.steps, world, ms)  # MS/s, whole job

    # sanity: the run really decoded (stereo flags + RDS groups present)
    st_host = status.cpu().numpy().view(fm.STATUS_DTYPE).reshape(C, B)
    decoded = {"stereo_channels": int(st_host["stereo"][:, -1].sum()),
               "groups_last_step": int(n_groups.sum().item())}

    # ---- per-stage device times: a separate timed pass with the stages of the block pipeline
    # queued on ONE stream (fmgpu_set_stage_overlap(0)), so that the CUDA events around each stage
    # bracket that stage's kernels running alone (in the timed steps above the stages of successive
    # blocks overlap on the device) --------------------------------------------------------------
    eng.set_stage_overlap(False)

    def sstep():
        eng.process_batch(iq_dev.data_ptr(), stride, B, audio.data_ptr(), acap, n_audio.data_ptr(),
                          groups.data_ptr(), gcap, n_groups.data_ptr(), status.data_ptr(), stream)

    for _ in range(2):
        sstep()
    torch.cuda.synchronize()
    eng.enable_stage_timing(True)
    acc = {}
    reps = 3
    for _ in range(reps):
        sstep()
        torch.cuda.synchronize()
        for k, v in eng.stage_times().items():
            acc[k] = acc.get(k, 0.0) + v / reps
    eng.enable_stage_timing(False)
    eng.set_stage_overlap(True)
    stage_ms = {k: round(v, 4) for k, v in acc.items()}
    serial_sum = sum(acc.values())
    dominant = max(acc, key=acc.get)
    roofline = kernel_roofline(dominant, acc[dominant], C, B, clocks)
    roofline["step_share"] = acc[dominant] / serial_sum
    # the dominant roofline-bound (tile) kernel and the dominant latency-bound (lane) kernel,
    # whichever of the two `dominant` is (lane kernels are reported as
    # lane-steps/s, not as a roofline fraction)
    tile = max((k for k in acc if k in TILE_STAGES), key=acc.get)
    lane = max((k for k in acc if k in LANE_STAGES), key=acc.get)
    roofline["dominant_tile_kernel"] = kernel_roofline(tile, acc[tile], C, B, clocks)
    roofline["dominant_lane_kernel"] = {
        "kernel": lane, "launch_ms": acc[lane] / B, "class": "latency-bound serial recursion, one "
        "lane per channel", "lanes": C, "lane_steps_per_s": C * B * BLOCK / (acc[lane] * 1e-3),
        "realtime_factor": (B * BLOCK / 240000.0) / (acc[lane] * 1e-3)}
    roofline["stage_sum_ms"] = serial_sum
    roofline["overlap_factor"] = serial_sum / (ms / args.steps)
    roofline["overlap_note"] = ("stage_ms are the stages run one after the other (fmgpu_set_stage_"
                                "overlap(0)); in the timed steps the block pipeline overlaps them: "
                                "overlap_factor = their sum / the measured step")
    # the HBM-bound (streaming) kernels of the path against the measured copy bandwidth: the
    # discriminator (8 B in + 4 B out per DSP-rate sample) from the stage pass above, and the RF
    # level meter (2 B in per IQ sample), timed here on its own
    streaming = []
    if "freqdem" in acc:
        fd_bytes = 12.0 * C * BLOCK
        fd_gbs = fd_bytes / (acc["freqdem"] / B * 1e-3) / 1e9
        streaming.append({"kernel": "freqdem", "bytes_per_launch": fd_bytes,
                          "launch_ms": acc["freqdem"] / B, "achieved": fd_gbs, "unit": "GB/s",
                          "frac": fd_gbs / roofline["peak"]})
    sums = torch.zeros((C, B, 48), dtype=torch.uint8, device=dev)
    for _ in range(2):
        eng.signal_level_batch(iq_dev.data_ptr(), stride, B, sums.data_ptr(), stream)
    sl0, sl1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    sl0.record()
    for _ in range(5):
        eng.signal_level_batch(iq_dev.data_ptr(), stride, B, sums.data_ptr(), stream)
    sl1.record()
    torch.cuda.synchronize()
    sl_ms = sl0.elapsed_time(sl1) / 5
    sl_gbs = 2.0 * C * n_iq / (sl_ms * 1e-3) / 1e9
    streaming.append({"kernel": "signal_level", "bytes_per_launch": 2.0 * C * n_iq, "launch_ms": sl_ms,
                      "achieved": sl_gbs, "unit": "GB/s", "frac": sl_gbs / roofline["peak"]})
    roofline["streaming_kernels"] = streaming
    step_alg_bytes = samples_per_step_rank * BYTES_PER_IQ_SAMPLE_ALG
    roofline["whole_step"] = {
        "achieved": step_alg_bytes / (ms / args.steps * 1e-3) / 1e9, "unit": "GB/s",
        "frac": step_alg_bytes / (ms / args.steps * 1e-3) / 1e9 / roofline["peak"],
        "fp32_tflops": samples_per_step_rank * FLOP_PER_IQ_SAMPLE_ALG / (ms / args.steps * 1e-3) / 1e12}

    # ---- end to end through the host-buffer C-ABI call -------------------------------------
    e2e = None
    e2e_ready = not args.no_e2e
    if e2e_ready:
        # the pinned host buffers (3.7 GB per rank at the default size); a rank that cannot get them
        # takes the leg off for every rank (the timing reduction below is a collective)
        alloc_failed = 0.0
        try:
            iq_host = torch.empty((C, stride), dtype=torch.uint8).pin_memory()
            iq_host.copy_(iq_dev)
            # two sets of host output buffers: step k+1 is submitted before step k is waited for
            outs = []
            for _ in range(2):
                outs.append((torch.empty((C, 2, acap), dtype=torch.float32).pin_memory(),
                             torch.zeros(C, dtype=torch.int32).pin_memory(),
                             torch.zeros((C, gcap, 16), dtype=torch.uint8).pin_memory(),
                             torch.zeros(C, dtype=torch.int32).pin_memory(),
                             torch.zeros((C, B, 20), dtype=torch.uint8).pin_memory()))
            na_host = outs[0][1]
        except RuntimeError as ex:
            print(f"rank {rank}: no pinned host buffers for the end-to-end leg: {ex}", file=sys.stderr)
            alloc_failed = 1.0
        if shard.max_over_ranks(alloc_failed, dev) > 0:
            e2e_ready = False
    if e2e_ready:

        def esubmit(k):
            a_h, na_h, g_h, ng_h, st_h = outs[k & 1]
            return eng.submit_host_raw(iq_host.data_ptr(), stride, B, a_h.data_ptr(), acap,
                                       na_h.data_ptr(), g_h.data_ptr(), gcap, ng_h.data_ptr(),
                                       st_h.data_ptr())

        def erun(nsteps):
            # every step: H2D of its IQ from pinned memory, the pipeline, D2H of audio / groups /
            # status; at most two steps in flight (fmgpu_submit_host / fmgpu_wait_host)
            if args.sync_steps:
                for k in range(nsteps):
                    eng.wait_host(esubmit(k))
                return
            pending = esubmit(0)
            for k in range(1, nsteps):
                nxt = esubmit(k)
                eng.wait_host(pending)
                pending = nxt
            eng.wait_host(pending)

        erun(max(2, min(args.warmup, 3)))
        barrier()
        t0 = time.perf_counter()
        erun(args.steps)
        torch.cuda.synchronize()
        dt = shard.max_over_ranks(time.perf_counter() - t0, dev)
        frames = int(na_host.max().item())
        e2e = {"value": world * samples_per_step_rank * args.steps / dt / 1e6, "unit": "MS/s",
               "h2d_bytes_per_step": int(C * 2 * n_iq),
               "d2h_bytes_per_step": int(C * 2 * frames * 4 + C * gcap * 16 + C * B * 20 + 8 * C),
               "ms_per_step": dt / args.steps * 1e3,
               "realtime_channels": world * samples_per_step_rank * args.steps / dt / IQ_RATE}

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) ------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        rows = iq_dev[:min(cores, 16), :2 * n_iq].cpu().numpy()
        cpu_arm(rows, B, 1, cores)                       # warm-up + calibration
        s1, d1 = cpu_arm(rows, B, 1, cores)
        passes = max(1, int(args.cpu_seconds / max(d1, 1e-3)))
        s, d = cpu_arm(rows, B, passes, cores)
        cpu_baseline = {"value": s / d / 1e6, "unit": "MS/s", "cores": cores, "kind": "port",
                        "sample": f"{cores} channels (one per thread) x {passes} passes x {B} blocks "
                                  f"x {BLOCK * DECIM} IQ samples of the same synthetic workload",
                        "seconds": d,
                        "note": "CPU oracle = restated reference pipeline, libm flavour, "
                                "-O3 -mavx2 -mfma; liquid-dsp itself is not installable here"}

    if rank == 0:
        line = {
            "metric": "aggregate IQ MS/s demodulated (stereo+RDS)", "value": value, "unit": "MS/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "realtime_channels": value * 1e6 / IQ_RATE,
            "config": workload_config(args, world * C),
            "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "stage_ms": stage_ms, "decoded": decoded,
            "native_library": os.path.basename(fm.lib_path()),
        }
        if host_cpus is not None:
            line["host_affinity_rank0"] = host_cpus
        print(json.dumps(line), flush=True)
    eng.close()


TILE_STAGES = ("decimate", "chanfir", "freqdem", "pilot_fir", "audio_lpf", "afpost", "rds_resample")
LANE_STAGES = ("dcblock", "agc", "stereo_pll", "rds", "rds_sync")


# DRAM bytes per DSP-rate sample (dram__bytes_read.sum + dram__bytes_write.sum) of one launch of
# each stage, from `ncu --set full` captures: decimate, chanfir, pilot_fir, stereo_pll and audio_lpf
# from the capture of the DEFAULT command (10,000 channels x 8192 samples per launch,
# profiles/r01_top_kernels_ncu_default_10000ch.csv), the others from the 1250-channel command
# (profiles/r01_top_kernels_ncu.csv; there, writes still in L2 when the kernel ended are not counted).
NCU_TRAFFIC_BYTES_PER_SAMPLE = {
    "decimate": 27.8, "chanfir": 15.6, "pilot_fir": 7.6, "audio_lpf": 15.6, "stereo_pll": 15.7,
    "dcblock": 12.1, "agc": 10.7, "freqdem": 9.3, "rds": 2.9, "rds_resample": 4.2, "afpost": 9.5,
}


def kernel_roofline(stage: str, stage_ms: float, C: int, B: int, clocks: dict) -> dict:
    """Algorithmic bytes / flops of ONE LAUNCH of the stage (one logical block of all C channels;
    DESIGN.md section 4) over its average launch duration (stage_ms is the sum over the B blocks of a
    step, measured with the stages serialised)."""
    n = C * BLOCK                # DSP-rate samples per launch
    n_iq = n * DECIM
    launch_ms = stage_ms / B
    alg = {   # stage: (bytes, flops) per launch
        "decimate": (2.0 * n_iq + 8.0 * n, 2.0 * 2 * 280 * n),
        "chanfir": (8.0 * n + 8.0 * n, 2.0 * 2 * 81 * n),
        "pilot_fir": (4.0 * n + 4.0 * n, 2.0 * 305 * n),
        "stereo_pll": (8.0 * n + 8.0 * n, 75.0 * n),
        "audio_lpf": (8.0 * n + 8.0 * n, 2.0 * 2 * 121 * n),
        "rds_resample": (4.0 * n + 4.0 * n * 171.0 / 240.0, 2.0 * 26 * n * 171.0 / 240.0),
        "rds": (4.0 * n * 171.0 / 240.0, 2.0 * (22 + 20) * n * 171.0 / 240.0),
        "rds_sync": (0.01 * n, 0.1 * n),
        "dcblock": (8.0 * n + 8.0 * n, 6.0 * n),
        "agc": (8.0 * n + 8.0 * n, 40.0 * n),
        "freqdem": (8.0 * n + 4.0 * n, 30.0 * n),
        "afpost": (8.0 * n + 8.0 * n * 32.0 / 240.0, 2.0 * 2 * 24 * n * 32.0 / 240.0),
    }.get(stage, (BYTES_PER_IQ_SAMPLE_ALG * n_iq, FLOP_PER_IQ_SAMPLE_ALG * n_iq))
    peak_hbm, how = 6650.0, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak_hbm, how = float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        pass
    t = launch_ms * 1e-3
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    traffic = NCU_TRAFFIC_BYTES_PER_SAMPLE.get(stage)
    return {"kernel": stage, "bound": "hbm", "achieved": alg[0] / t / 1e9, "peak": peak_hbm,
            "peak_source": how, "unit": "GB/s", "frac": alg[0] / t / 1e9 / peak_hbm,
            "traffic": (traffic * n) if traffic is not None else None,
            "traffic_source": "ncu dram bytes per sample (profiles/r01_top_kernels_ncu*.csv) x the "
                              "samples of one launch",
            "algorithmic_bytes": alg[0], "launch_ms": launch_ms, "launches_per_step": B,
            "fp32": {"achieved_tflops": alg[1] / t / 1e12, "peak_tflops": fp32_peak,
                     "peak_source": "148 SM x 128 FMA/clk x 2 at the SM clock sampled under load",
                     "frac": alg[1] / t / 1e12 / fp32_peak},
            "note": "FIR kernels are FP32-pipe-bound (~140 flop/B, SURVEY §8(d)): the hbm fraction "
                    "shows how far the kernel sits from the streaming bound, the fp32 fraction how "
                    "well it uses the pipe that actually limits it"}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, local_rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
