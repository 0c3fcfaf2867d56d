#!/usr/bin/env python
"""Summaries of the two ncu passes of B200_PROFILING.md for profiles/:

  python tools/ncu_summarize.py launches gpurun_out/rNN_launches.csv
      -> markdown table: kernel, launches, total ms, average ms, share of the serialised time
  python tools/ncu_summarize.py full gpurun_out/rNN_top.ncu-rep profiles/rNN_top_kernels_ncu.csv
      -> one CSV row of key metrics per distinct (kernel, grid) of the `ncu --set full` capture
"""
from __future__ import annotations

import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
]


def short(name: str) -> str:
    name = name.replace("fmgpu::", "").replace("void ", "")
    m = re.match(r"([\w:<>, ]+?)\(", name)
    return (m.group(1) if m else name).strip()


def launches(path: str) -> None:
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    acc: dict[str, list[float]] = {}
    for r in rows[1:]:
        acc.setdefault(short(r[ik]), []).append(float(r[iv].replace(",", "")) / 1e6)
    total = sum(sum(v) for v in acc.values())
    print("| kernel | launches | total ms | avg ms | share |")
    print("|---|---|---|---|---|")
    for k, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
        print(f"| {k} | {len(v)} | {sum(v):.3f} | {sum(v) / len(v):.4f} | {100 * sum(v) / total:.1f}% |")
    print(f"\nserialised total {total:.2f} ms over {sum(len(v) for v in acc.values())} launches")


def full(rep: str, out: str) -> None:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keys = [k for k in KEYS if k in hdr]
    seen = set()
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Kernel Name"] + keys)
        w.writerow([""] + [units[hdr.index(k)] for k in keys])
        for r in rows[2:]:
            ident = (r[hdr.index("Kernel Name")], r[hdr.index("Grid Size")])
            if ident in seen:
                continue
            seen.add(ident)
            w.writerow([short(ident[0]) + " grid" + ident[1]] + [r[hdr.index(k)] for k in keys])
    print(f"{len(seen)} kernels -> {out}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3])
