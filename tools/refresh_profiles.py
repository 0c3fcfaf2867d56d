#!/usr/bin/env python
"""Copy the artefacts of one tools/gpu_final.sh + tools/gpu_ncu_default.sh call (gpurun_out/<run>_*)
into profiles/r01_* and print the numbers profiles/r01_summary.md quotes.

  python tools/refresh_profiles.py r38
"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
run = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
for src, dst in (("bench_default.json", "r01_bench_default_10000ch.json"),
                 ("bench_reference.json", "r01_bench_reference_arm.json"),
                 ("bench_sync.json", "r01_bench_sync_steps.json"),
                 ("bench_small.json", "r01_bench_small_1250ch.json"),
                 ("timeline.txt", "r01_timeline_10000ch.txt"),
                 ("launches.csv", "r01_launches_ncu.csv"),
                 ("launches_default.csv", "r01_launches_ncu_default_10000ch.csv")):
    shutil.copy(os.path.join(G, f"{run}_{src}"), os.path.join(P, dst))
tool = os.path.join(ROOT, "tools", "ncu_summarize.py")
subprocess.run([sys.executable, tool, "full", os.path.join(G, f"{run}_top.ncu-rep"),
                os.path.join(P, "r01_top_kernels_ncu.csv")], check=True)
subprocess.run([sys.executable, tool, "full", os.path.join(G, f"{run}_top_default.ncu-rep"),
                os.path.join(P, "r01_top_kernels_ncu_default_10000ch.csv")], check=True)
for name in ("r01_launches_ncu.csv", "r01_launches_ncu_default_10000ch.csv"):
    print(f"\n## {name}")
    subprocess.run([sys.executable, tool, "launches", os.path.join(P, name)], check=True)
for name in ("r01_bench_default_10000ch.json", "r01_bench_small_1250ch.json",
             "r01_bench_sync_steps.json", "r01_bench_reference_arm.json"):
    d = json.loads(open(os.path.join(P, name)).read().strip().splitlines()[-1])
    print(f"\n## {name}: value {d['value']:.0f} {d['unit']}, {d['ms_per_step']:.2f} ms/step, e2e "
          f"{(d.get('e2e') or {}).get('value')}, cpu {(d.get('cpu_baseline') or {}).get('value')}")
    if d.get("stage_ms"):
        print("  stage_ms", d["stage_ms"])
    r = d.get("roofline")
    if r and "frac" in r:
        print(f"  roofline {r['kernel']}: hbm frac {r['frac']:.3f}, fp32 frac {r['fp32']['frac']:.3f}, "
              f"share {r['step_share']:.3f}, traffic {r['traffic']:.3e} / alg {r['algorithmic_bytes']:.3e}")
        print("  streaming", [(k['kernel'], round(k['frac'], 3)) for k in r["streaming_kernels"]])
for name in ("r01_top_kernels_ncu.csv", "r01_top_kernels_ncu_default_10000ch.csv"):
    rows = list(csv.reader(open(os.path.join(P, name))))
    hdr = rows[0]
    print(f"\n## {name}")
    print("| kernel | ms | regs | grid | block | smem KB | rd | wr | dram % | FMA pipe % | issue % | LSU % | warps % | M inst |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for r in rows[2:]:
        def c(k):
            return float(r[hdr.index(k)])
        print(f"| {r[0]} | {c('gpu__time_duration.sum'):.3f} | {int(c('launch__registers_per_thread'))} | "
              f"{int(c('launch__grid_size'))} | {int(c('launch__block_size'))} | "
              f"{c('launch__shared_mem_per_block_dynamic') + c('launch__shared_mem_per_block_static'):.1f} | "
              f"{c('dram__bytes_read.sum'):.1f} | {c('dram__bytes_write.sum'):.1f} | "
              f"{c('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{c('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{c('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{c('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{c('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{c('smsp__inst_executed.sum') / 1e6:.1f} |")
    print("units:", dict(zip(hdr[1:], rows[1][1:]))["dram__bytes_read.sum"], "/",
          dict(zip(hdr[1:], rows[1][1:]))["dram__bytes_write.sum"])
