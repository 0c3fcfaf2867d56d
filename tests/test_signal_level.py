"""RF level meter: the engine reduces IQ bytes to exact integer sums on the
device and finishes them on the host; the reference is src/signal_level.cpp itself, compiled
in place into oracle/_ref (a numpy restatement stands in where the reference is absent)."""
import ctypes as C

import numpy as np
import pytest

import fmtuner_sdr_b200 as fm
from oracle import orc


def _cases():
    rng = np.random.default_rng(9)
    yield "noise", rng.integers(96, 160, 2 * 65536, dtype=np.uint8)
    yield "silent", np.full(2 * 4096, 127, np.uint8)
    x = rng.normal(127.5, 60.0, 2 * 81920)
    yield "clipping", np.clip(np.rint(x), 0, 255).astype(np.uint8)
    yield "dc_offset", (rng.integers(0, 40, 2 * 8192) + np.tile([150, 90], 8192)).astype(np.uint8)
    yield "odd_length", rng.integers(0, 256, 2 * 12345, dtype=np.uint8)


def _sums_numpy(iq):
    i = iq[0::2].astype(np.uint64)
    q = iq[1::2].astype(np.uint64)
    ib, qb = iq[0::2], iq[1::2]
    return fm.LevelSums(int(i.sum()), int(q.sum()), int((i * i).sum()), int((q * q).sum()),
                        int(((ib <= 1) | (ib >= 254) | (qb <= 1) | (qb >= 254)).sum()),
                        int(((ib <= 8) | (ib >= 247) | (qb <= 8) | (qb >= 247)).sum()), iq.size // 2, 0)


def _finish(sums, *args):
    L = fm.load_library()
    out = fm.SignalLevel()
    L.fmgpu_signal_level_finish(C.byref(sums), *args, C.byref(out))
    return out


PARAMS = [(0, 0.0, 0.0, -70.0, -5.0), (20, 0.5, -4.0, -62.0, -12.0), (49, 1.0, 3.0, -40.0, -41.5)]


@pytest.mark.parametrize("gain,comp,bias,floor,ceil", PARAMS)
def test_host_finish_matches_reference(gain, comp, bias, floor, ceil):
    for name, iq in _cases():
        ref = orc.signal_level(iq, gain, comp, bias, floor, ceil)
        got = _finish(_sums_numpy(iq), gain, comp, bias, floor, ceil)
        assert abs(got.dbfs - ref[1]) < 1e-9, name
        assert abs(got.compensated_dbfs - ref[2]) < 1e-9, name
        assert abs(got.level120 - ref[0]) < 1e-4, name
        assert got.hard_clip_ratio == ref[3] and got.near_clip_ratio == ref[4], name


def test_empty_block_defaults():
    out = _finish(fm.LevelSums(), 0, 0.0, 0.0, -70.0, -5.0)
    assert (out.level120, out.dbfs, out.hard_clip_ratio) == (0.0, -120.0, 0.0)


@pytest.mark.gpu
def test_device_sums_are_exact():
    import torch
    C_, B = 37, 3
    eng = fm.Engine(fm.make_config(max_blocks=B), C_, 0)
    n_iq = B * 81920
    rng = np.random.default_rng(4)
    host = np.clip(np.rint(rng.normal(127.5, 70.0, (C_, 2 * n_iq))), 0, 255).astype(np.uint8)
    iq = torch.from_numpy(host).cuda()
    sums = torch.zeros((C_, B, 48), dtype=torch.uint8, device="cuda")
    eng.signal_level_batch(iq.data_ptr(), 2 * n_iq, B, sums.data_ptr())
    torch.cuda.synchronize()
    raw = sums.cpu().numpy()
    for c in (0, 5, 36):
        for b in range(B):
            s = fm.LevelSums.from_buffer_copy(raw[c, b].tobytes())
            want = _sums_numpy(host[c, b * 163840:(b + 1) * 163840])
            for f in ("sum_i", "sum_q", "sum_ii", "sum_qq", "hard_clip", "near_clip", "n_samples"):
                assert getattr(s, f) == getattr(want, f), (c, b, f)
            ref = orc.signal_level(host[c, b * 163840:(b + 1) * 163840])
            got = eng.signal_level_finish(s)
            assert abs(got.dbfs - ref[1]) < 1e-9 and got.hard_clip_ratio == ref[3]
    eng.close()
