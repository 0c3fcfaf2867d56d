"""De-emphasis + DC blocker as a warp-shuffle parallel scan (fmgpu_set_audio_iir_mode(1),
k_audio_iir_scan; north_star item 5, af_post_processor.cpp:66-71). The scan adds the terms of the
two first-order recursions in a different order than the serial loop, so it is compared with the
serial form (mode 0, bit-identical to the oracle) in float-rounding tolerance; everything that
does not pass through the filters — frame counts, status, RDS groups — must be identical."""
import numpy as np
import pytest

import fmtuner_sdr_b200 as fm
from oracle import orc
from tests.common import groups_equal, rates, run_engine_chunks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("deemph", [0, 1, 2])      # 50 us, 75 us, off
def test_scan_equals_serial_recursion(orc_fm, deemph):
    iq_rate, decim = rates("240k")
    C, nblk = 33, 12
    iq = np.stack([orc.config3_signal(70 + c, fs_iq=iq_rate).generate(nblk * 8192 * decim)
                   for c in range(4)])[np.arange(C) % 4]
    res = {}
    for mode in (0, 1):
        eng = fm.Engine(fm.make_config(iq_rate=iq_rate, decimation=decim, max_blocks=3, deemphasis=deemph), C, 0)
        eng.set_audio_iir_mode(mode)
        assert eng.audio_iir_mode() == mode
        res[mode] = run_engine_chunks(eng, iq, nblk, 3)     # state carried over 4 calls x 3 blocks
        eng.close()
    (a0, g0, s0, _), (a1, g1, s1, _) = res[0], res[1]
    for c in range(C):
        assert a0[c].shape == a1[c].shape and np.array_equal(s0[c], s1[c]) and groups_equal(g0[c], g1[c])
        assert np.abs(a0[c] - a1[c]).max() <= 1e-6, (c, np.abs(a0[c] - a1[c]).max())
    ref = orc.Channel(orc_fm, orc.make_config(iq_rate=iq_rate, decimation=decim, deemphasis=deemph)).process(iq[0])
    assert np.array_equal(a0[0][0], ref.left) and np.array_equal(a0[0][1], ref.right)
    assert np.abs(a1[0][0] - ref.left).max() <= 1e-6
