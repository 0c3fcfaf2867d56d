"""Shared helpers for the parity tests."""
from __future__ import annotations

import numpy as np

from oracle import orc

GROUP_KEYS = ("a", "b", "c", "d", "errors", "block_index")


def rates(which: str):
    """'240k' = 2.4 MS/s / 10 (BASELINE configs 3-5); '256k' = 2.048 MS/s / 8 (unmodified main.cpp)."""
    return {"240k": (2_400_000, 10), "256k": (2_048_000, 8), "1024k": (1_024_000, 4),
            "direct256k": (256_000, 1),
            # DSP rate above 2 x 171 kHz: the RDS resampler's window no longer fits its tile
            "480k": (2_400_000, 5),
            # the smallest and the largest instantiated decimation factors
            "m2": (480_000, 2), "m16": (3_840_000, 16)}[which]


def groups_equal(a: np.ndarray, b: np.ndarray, keys=GROUP_KEYS) -> bool:
    return len(a) == len(b) and all((a[k] == b[k]).all() for k in keys)


def run_oracle(lib, cfg_kwargs: dict, iq: np.ndarray, debug=True):
    ch = orc.Channel(lib, orc.make_config(**cfg_kwargs))
    return ch, ch.process(iq, debug=debug)


def run_engine_chunks(eng, iq_rows: np.ndarray, nblk: int, chunk: int, debug_channel=None):
    """Feed [C, bytes] IQ through process_host `chunk` blocks at a time; concatenate outputs."""
    C = eng.n_channels
    per = eng.iq_bytes_per_block
    audio = [[] for _ in range(C)]
    groups = [[] for _ in range(C)]
    status = []
    dbg = {k: [] for k in ("dec", "mpx", "sl", "sr")}
    for b0 in range(0, nblk, chunk):
        nb = min(chunk, nblk - b0)
        a, na, g, ng, st = eng.process_host(iq_rows[:, b0 * per:(b0 + nb) * per], nb)
        status.append(st)
        for c in range(C):
            audio[c].append(a[c, :, :na[c]])
            gg = g[c, :ng[c]].copy()
            gg["block_index"] += b0
            groups[c].append(gg)
        if debug_channel is not None:
            for i, k in enumerate(("dec", "mpx", "sl", "sr")):
                if k == "dec" and eng.decim == 1:
                    continue
                dbg[k].append(eng.debug_read(i, debug_channel))
    out_audio = [np.concatenate(x, axis=1) for x in audio]
    out_groups = [np.concatenate(x) for x in groups]
    out_status = np.concatenate(status, axis=1)
    out_dbg = {k: (np.concatenate(v) if v else None) for k, v in dbg.items()}
    return out_audio, out_groups, out_status, out_dbg


def snr_db(ref: np.ndarray, test: np.ndarray) -> float:
    err = ref.astype(np.float64) - test.astype(np.float64)
    p = float((ref.astype(np.float64) ** 2).mean())
    e = float((err ** 2).mean())
    if e == 0.0:
        return np.inf
    return 10.0 * np.log10(max(p, 1e-300) / e)
