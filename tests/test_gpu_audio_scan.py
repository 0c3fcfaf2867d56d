"""The linear first-order recursions as warp-shuffle parallel scans (fmgpu_set_scan_mode(1):
k_dcblock_scan for the I/Q DC blockers, fm_demod.cpp:164-165; k_audio_iir_scan for de-emphasis +
DC blocker at 32 kHz, af_post_processor.cpp:66-71; north_star item 5). A scan adds the terms of a
recursion in a different order than the serial loop, so the results agree with the serial form
(mode 0, bit-identical to the oracle) to float rounding, not bit for bit: audio in a tight
tolerance, lock blocks and RDS groups of clean signals identical."""
import numpy as np
import pytest

import fmtuner_sdr_b200 as fm
from oracle import orc
from tests.common import groups_equal, rates, run_engine_chunks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rate,deemph", [("240k", 0), ("240k", 1), ("256k", 2), ("direct256k", 0)])
def test_scans_equal_serial_recursions(orc_fm, rate, deemph):
    iq_rate, decim = rates(rate)
    C, nblk = 33, 15
    sigs = [orc.config1_signal(fs_iq=iq_rate)]
    for c in range(3):
        s = orc.config3_signal(70 + c, fs_iq=iq_rate)
        s.snr_db = 45.0
        s.dc_i, s.dc_q = 0.02 * (c + 1), -0.015 * (c + 1)     # something for the DC blockers to remove
        sigs.append(s)
    iq = np.stack([s.generate(nblk * 8192 * decim) for s in sigs])[np.arange(C) % 4]
    res = {}
    for mode in (0, 1):
        eng = fm.Engine(fm.make_config(iq_rate=iq_rate, decimation=decim, max_blocks=3, deemphasis=deemph), C, 0)
        eng.set_scan_mode(mode)
        assert eng.scan_mode() == mode
        res[mode] = run_engine_chunks(eng, iq, nblk, 3)     # state carried over 5 calls x 3 blocks
        eng.close()
    (a0, g0, s0, _), (a1, g1, s1, _) = res[0], res[1]
    for c in range(C):
        assert a0[c].shape == a1[c].shape
        assert np.array_equal(s0[c]["stereo"], s1[c]["stereo"]) and np.array_equal(s0[c]["n_audio"], s1[c]["n_audio"])
        assert np.abs(s0[c]["pilot_tenths"] - s1[c]["pilot_tenths"]).max() <= 1
        assert np.allclose(s0[c]["clip_ratio"], s1[c]["clip_ratio"], rtol=0, atol=0)
        assert groups_equal(g0[c], g1[c]), c
        assert np.abs(a0[c] - a1[c]).max() <= 2e-5, (c, np.abs(a0[c] - a1[c]).max())
    ref = orc.Channel(orc_fm, orc.make_config(iq_rate=iq_rate, decimation=decim, deemphasis=deemph)).process(iq[0])
    assert np.array_equal(a0[0][0], ref.left) and np.array_equal(a0[0][1], ref.right)   # mode 0: the oracle
    assert np.abs(a1[0][0] - ref.left).max() <= 2e-5 and len(g1[0]) >= 2
