"""The stereo decoder's real-tap FIRs on the tensor cores (fmtuner_sdr_b200/csrc/fir_tc.cu,
fmgpu_set_fir_mode(1)): the 19 kHz pilot band-pass and the L/R 15 kHz low-pass
(stereo_decoder.cpp:25-63,172-173,233-239) as exact integer contractions — samples as 24-bit fixed
point, taps as 24-bit integers, int32 sums in TMEM. NOT bit-identical to the FP32 chains (mode 0,
the oracle's summation order), so each filter is checked against a float64 evaluation of the same
FIR over the engine's own input rows (tolerance: the sample quantum times the taps' absolute sum,
plus the final float roundings), against mode 0, and through the whole pipeline against the
reference-faithful CPU flavour in the tolerance north_star states."""
import numpy as np
import pytest

import fmtuner_sdr_b200 as fm
from oracle import orc
from tests.common import rates, run_engine_chunks, snr_db

pytestmark = pytest.mark.gpu


def fir64(x, taps, scale):
    """float64 model: y[n] = scale * sum_i h[i] x[n - i], zeros before the stream."""
    return scale * np.convolve(x.astype(np.float64), taps.astype(np.float64))[:x.size]


@pytest.mark.parametrize("rate", ["240k", "256k"])
def test_pilot_and_lowpass_match_float64_model(rate):
    iq_rate, decim = rates(rate)
    C, nblk = 131, 3                       # two row tiles, the second almost empty
    rows = [orc.config3_signal(c, fs_iq=iq_rate).generate(nblk * 8192 * decim) for c in range(4)]
    rng = np.random.default_rng(11)
    rows.append(rng.integers(0, 256, rows[0].size, dtype=np.uint8))   # noise only: discriminator clicks
    iq = np.stack([rows[c % len(rows)] for c in range(C)])
    chans = (0, 1, 4, 127, 128, 130)
    out = {}
    for mode in (0, 1):
        eng = fm.Engine(fm.make_config(iq_rate=iq_rate, decimation=decim, max_blocks=nblk), C, 0)
        eng.set_fir_mode(mode)
        assert eng.fir_mode() == mode
        pil, _ = eng.design(2)
        aud, aud_scale = eng.design(3)
        eng.process_host(iq, nblk)
        out[mode] = {c: [eng.debug_read(w, c) for w in (1, 4, 5, 6, 2, 3)] for c in chans}
        eng.close()
    for c in chans:
        mpx, pilot, lraw, rraw, lf, rf = out[1][c]
        mpx0, pilot0, lraw0, rraw0, lf0, rf0 = out[0][c]
        assert mpx.size == nblk * 8192 and np.array_equal(mpx, mpx0)
        # pilot band-pass: sample quantum 2^-22 (half of it per sample, times sum |h|)
        want = fir64(mpx, pil, 1.0)
        bound = 2.0 ** -23 * np.abs(pil).sum() + 4e-8
        err = np.abs(pilot.astype(np.float64) - want).max()
        err0 = np.abs(pilot0.astype(np.float64) - want).max()
        assert err <= bound, (c, err, bound)
        assert np.abs(pilot - pilot0).max() <= bound + err0
        # L/R low-pass over the engine's own matrix outputs: quantum 2^-20
        for raw, got, got0, raw0 in ((lraw, lf, lf0, lraw0), (rraw, rf, rf0, rraw0)):
            want = fir64(raw, aud, aud_scale)
            bound = 2.0 ** -21 * np.abs(aud * aud_scale).sum() + 3e-7
            err = np.abs(got.astype(np.float64) - want).max()
            assert err <= bound, (c, err, bound)
            # and the FP32 chain over ITS matrix outputs is the same filter
            assert np.abs(got0.astype(np.float64) - fir64(raw0, aud, aud_scale)).max() <= 2e-6


def test_streamed_calls_carry_the_window_and_small_calls_fall_back():
    """history across calls and ring wrap-around (7 calls of 1-2 blocks into a 3-slot ring), 5 channels
    (one partial row tile); a ragged stage-level length takes the FP32 kernel."""
    iq_rate, decim = rates("240k")
    nblk = 10
    iq = np.stack([orc.config3_signal(20 + c, fs_iq=iq_rate).generate(nblk * 8192 * decim) for c in range(5)])
    res = {}
    for mode in (0, 1):
        eng = fm.Engine(fm.make_config(iq_rate=iq_rate, decimation=decim, max_blocks=2), 5, 0)
        eng.set_fir_mode(mode)
        per = eng.iq_bytes_per_block
        pilots, lfs = [], []
        b = 0
        for nb in (1, 2, 1, 2, 2, 1, 1):
            eng.process_host(iq[:, b * per:(b + nb) * per], nb)
            pilots.append(eng.debug_read(4, 3))
            lfs.append(eng.debug_read(2, 3))
            b += nb
        res[mode] = (np.concatenate(pilots), np.concatenate(lfs))
        eng.close()
    assert np.abs(res[1][0] - res[0][0]).max() <= 4e-7
    assert np.abs(res[1][1] - res[0][1]).max() <= 3e-5   # the PLL sees a pilot that differs by 1e-7


def test_whole_pipeline_in_tolerance_of_the_faithful_reference():
    """config 1 for 48 blocks and eight weak-signal channels through the whole pipeline with every fast
    form on (tensor-core decimator and FIRs, scans): same lock block, same clean groups, audio inside
    north_star's tolerance against the reference-faithful flavour."""
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
