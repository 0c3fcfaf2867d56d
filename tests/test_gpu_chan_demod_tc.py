"""Channel filter + quadrature discriminator as one tensor-core kernel (fir_tc.cu,
fmgpu_set_demod_mode(1)): FMDemod::demodulateComplex (fm_demod.cpp:194-199) with the channel filter as an
exact integer contraction and the discriminator in the epilogue. The pre-discriminator AGC
(fm_demod.cpp:196-198) scales y[n] by a positive real gain that arg(y[n] conj(y[n-1])) cannot see, so
mode 1 leaves it out: checked here as "MPX with the AGC on equals MPX of mode 0 with the AGC on".
Not bit-identical to mode 0 (quantised samples, another summation order), so: MPX against mode 0 in
a tolerance that scales with 1 / |y|, the carried discriminator sample across calls, the fallback
for calls whose channels use different channel filters (bit-exact again), and the whole pipeline
against the reference-faithful CPU flavour in north_star's tolerance."""
import numpy as np
import pytest

import fmtuner_sdr_b200 as fm
from oracle import orc
from tests.common import rates, run_engine_chunks, snr_db

pytestmark = pytest.mark.gpu


def _mpx(eng, iq, calls, chans):
    per = eng.iq_bytes_per_block
    out = {c: [] for c in chans}
    b = 0
    for nb in calls:
        eng.process_host(iq[:, b * per:(b + nb) * per], nb)
        for c in chans:
            out[c].append(eng.debug_read(1, c))
        b += nb
    return {c: np.concatenate(v) for c, v in out.items()}


@pytest.mark.parametrize("rate,agc", [("240k", 1), ("240k", 0), ("256k", 1), ("1024k", 1)])
def test_mpx_matches_the_three_kernel_form(rate, agc):
    iq_rate, decim = rates(rate)
    C, calls = 131, (2, 1, 2, 2, 1)          # three row tiles of 64 channels, the last almost empty
    nblk = sum(calls)
    rows = [orc.config3_signal(c, fs_iq=iq_rate).generate(nblk * 8192 * decim) for c in range(5)]
    iq = np.stack([rows[c % len(rows)] for c in range(C)])
    chans = (0, 1, 63, 64, 127, 128, 130)
    res = {}
    for mode in (0, 1):
        eng = fm.Engine(fm.make_config(iq_rate=iq_rate, decimation=decim, max_blocks=2, dsp_agc=agc), C, 0)
        eng.set_demod_mode(mode)
        assert eng.demod_mode() == mode
        res[mode] = _mpx(eng, iq, calls, chans)
        eng.close()
    for c in chans:
        a, b = res[1][c], res[0][c]
        assert a.size == b.size == nblk * 8192
        d = np.abs(a - b)[256:]   # (the filter's first outputs are ~0: their phase is noise in both forms)
        # strong carriers (|y| ~ 0.3 .. 1): the discriminator turns 5e-7 of filter-output difference
        # into ~1e-6 of MPX; every sample, block and call boundaries (carried sample) included
        assert d.max() <= 2e-5, (c, d.max(), int(d.argmax()))
        assert np.sqrt((d.astype(np.float64) ** 2).mean()) <= 2e-6, c


def test_mixed_channel_filters_fall_back_to_the_bit_exact_kernels():
    iq_rate, decim = rates("240k")
    nblk = 3
    iq = np.stack([orc.config3_signal(c, fs_iq=iq_rate).generate(nblk * 8192 * decim) for c in range(4)])
    res = {}
    for mode in (0, 1):
        eng = fm.Engine(fm.make_config(iq_rate=iq_rate, decimation=decim, max_blocks=nblk, dsp_agc=1), 4, 0)
        eng.set_demod_mode(mode)
        eng.set_bandwidth_hz(56000, 2)          # channel 2 gets its own channel filter
        res[mode] = _mpx(eng, iq, (nblk,), (0, 2, 3))
        eng.close()
    for c in (0, 2, 3):
        assert np.array_equal(res[0][c], res[1][c]), c


def test_whole_pipeline_in_tolerance_of_the_faithful_reference():
    """config 1 for 48 blocks and eight weak-signal channels with every fast form on (tensor-core
    decimator, fused channel filter + discriminator, tensor-core pilot / low-pass FIRs, scans): same
    lock block, same clean groups, audio inside north_star's tolerance against the faithful flavour."""
    iq_rate, decim = rates("240k")
    nblk = 48
    faith = orc.OracleLib("ref") if orc.OracleLib.have_ref("ref") else orc.OracleLib("libm")
    sigs = [orc.config1_signal(fs_iq=iq_rate)]
    for c in range(8):
        s = orc.config3_signal(60 + c, fs_iq=iq_rate)
        s.snr_db = 22.0 + 2.5 * c
        sigs.append(s)
    iq = np.stack([s.generate(nblk * 8192 * decim) for s in sigs])
    eng = fm.Engine(fm.make_config(iq_rate=iq_rate, decimation=decim, max_blocks=4, dsp_agc=1), len(sigs), 0)
    eng.set_decimator_mode(1)
    eng.set_scan_mode(1)
    eng.set_fir_mode(1)
    eng.set_demod_mode(1)
    audio, groups, status, _ = run_engine_chunks(eng, iq, nblk, 4)
    eng.close()
    for c in range(len(sigs)):
        cfg = orc.make_config(iq_rate=iq_rate, decimation=decim, dsp_agc=1)
        ref = orc.Channel(faith, cfg).process(iq[c])
        assert np.array_equal(status[c]["stereo"], ref.status["stereo"]), c
        assert np.abs(status[c]["pilot_tenths"] - ref.status["pilot_tenths"]).max() <= 1, c
        clean = ref.groups["errors"] == 0
        assert len(groups[c]) == len(ref.groups), c
        for k in np.flatnonzero(clean):
            assert all(groups[c][k][f] == ref.groups[k][f] for f in ("a", "b", "c", "d", "errors")), (c, k)
        lock = int(np.flatnonzero(ref.status["stereo"])[0])
        s0 = int(ref.status["n_audio"][:lock + 2].sum())
        for x, y in ((audio[c][0], ref.left), (audio[c][1], ref.right)):
            assert np.abs(x[s0:] - y[s0:]).max() <= 1e-4 or snr_db(y[s0:], x[s0:]) >= 90.0, c
    assert orc.decode_ps_rt(groups[0]) == (0x1234, "B200TEST", "FM ON B200")
