"""The tensor-core decimator (fmtuner_sdr_b200/csrc/decim_tc.cu, fmgpu_set_decimator_mode(1)):
ComplexDecimator::executeComplex (liquid_primitives.cpp:461-499) as a tcgen05 int8 contraction.
It is NOT bit-identical to the FP32 chain (mode 0, the oracle's summation order): it rounds the
exact sum once. So it is checked against a float64 evaluation of the same FIR (tolerance: one
float rounding of the result plus the 2^-27 tap quantisation), against mode 0 (the oracle's
arithmetic; tolerance: the float chain's own rounding noise), and through the whole pipeline
against the reference-faithful CPU flavour in the tolerance north_star states."""
import numpy as np
import pytest

import fmtuner_sdr_b200 as fm
from oracle import orc
from tests.common import groups_equal, rates, run_engine_chunks, snr_db

pytestmark = pytest.mark.gpu


def fir64(iq_bytes, taps, scale, M):
    """float64 model: y[n] = scale * sum_i h[L-1-i] x[nM-(L-1)+i], zeros before the stream."""
    x = (iq_bytes.astype(np.float64).reshape(-1, 2) - 127.5) / 127.5
    x = x[:, 0] + 1j * x[:, 1]
    L = taps.size
    xp = np.concatenate([np.zeros(L - 1, np.complex128), x])
    n_out = x.size // M
    idx = (np.arange(n_out) * M)[:, None] + np.arange(L)[None, :]
    return scale * (xp[idx] @ taps[::-1].astype(np.float64))


@pytest.mark.parametrize("rate", ["240k", "256k", "1024k"])
def test_decimated_iq_matches_float64_model(rate):
    iq_rate, decim = rates(rate)
    C, nblk = 131, 3                       # two row tiles, the second almost empty
    rng = np.random.default_rng(5)
    rows = [orc.config3_signal(c, fs_iq=iq_rate).generate(nblk * 8192 * decim) for c in range(3)]
    rows.append(rng.integers(0, 256, rows[0].size, dtype=np.uint8))          # full-scale noise
    rows.append(np.full(rows[0].size, 255, np.uint8))                          # rail
    iq = np.stack([rows[c % len(rows)] for c in range(C)])
    out = {}
    for mode in (0, 1):
        eng = fm.Engine(fm.make_config(iq_rate=iq_rate, decimation=decim, max_blocks=2), C, 0)
        eng.set_decimator_mode(mode)
        assert eng.decimator_mode() == mode
        taps, scale = eng.design(0)
        dec = {c: [] for c in (0, 1, 3, 4, 127, 128, 130)}
        for b0, nb in ((0, 2), (2, 1)):     # a two-block call, then one more block (history carried)
            per = eng.iq_bytes_per_block
            eng.process_host(iq[:, b0 * per:(b0 + nb) * per], nb)
            for c in dec:
                dec[c].append(eng.debug_read(0, c).view(np.complex64).copy())
        eng.close()
        out[mode] = {c: np.concatenate(v) for c, v in dec.items()}
    for c in out[1]:
        want = fir64(iq[c], taps, scale, decim)
        got = out[1][c].astype(np.complex128)
        exact_chain = out[0][c].astype(np.complex128)
        assert got.size == want.size
        # one rounding of the result (|y| <= 1.5: half an ulp is 6e-8) + quantised taps
        assert np.abs(got - want).max() <= 1.5e-7, (c, np.abs(got - want).max())
        # the FP32 chain itself is further from the exact sum than the tensor-core result
        assert np.abs(got - want).max() <= np.abs(exact_chain - want).max() + 1e-9
        assert np.abs(got - exact_chain).max() <= 3e-6


def test_stage_level_call_and_reset():
    """fmgpu_decimate (ComplexDecimator::executeComplex) through the tensor-core kernel, with a
    reset in between (zero window again) and a ragged length that falls back to the FP32 kernel."""
    iq_rate, decim = rates("240k")
    iq = orc.config1_signal(fs_iq=iq_rate).generate(3 * 8192 * decim)
    eng = fm.Engine(fm.make_config(iq_rate=iq_rate, decimation=decim, max_blocks=1), 2, 0)
    eng.set_decimator_mode(1)
    taps, scale = eng.design(0)
    per = 8192 * decim * 2
    a = eng.executeComplex(iq[:per], 8192, channel=1)
    b = eng.executeComplex(iq[per:per + 4000 * decim * 2], 4000, channel=1)
    c = eng.executeComplex(iq[per + 4000 * decim * 2:per + 4003 * decim * 2], 3, channel=1)  # ragged
    want = fir64(iq[:per + 4003 * decim * 2], taps, scale, decim)
    got = np.concatenate([a, b, c]).astype(np.complex128)
    assert np.abs(got - want).max() <= 3e-6 and np.abs(got[:12192] - want[:12192]).max() <= 1.5e-7
    eng.reset(fm.RESET_DECIM, 1)
    d = eng.executeComplex(iq[2 * per:3 * per], 8192, channel=1)
    assert np.abs(d.astype(np.complex128) - fir64(iq[2 * per:3 * per], taps, scale, decim)).max() <= 1.5e-7
    eng.close()


def test_unsupported_factor_is_refused():
    iq_rate, decim = rates("480k")          # factor 5: the tile advance is not a multiple of 32 bytes
    eng = fm.Engine(fm.make_config(iq_rate=iq_rate, decimation=decim), 1, 0)
    with pytest.raises(fm.EngineError):
        eng.set_decimator_mode(1)
    assert eng.decimator_mode() == 0
    eng.close()


def test_whole_pipeline_in_tolerance_of_the_faithful_reference(orc_fm):
    """config 1 for 48 blocks and eight weak-signal channels through the whole pipeline with the
    tensor-core decimator: same lock block, same groups, audio inside north_star's tolerance
    against the reference-faithful flavour; MPX within 2e-5 of the FP32-chain engine."""
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
