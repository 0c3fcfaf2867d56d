"""Physics-level checks of the oracle's restated liquid-dsp design code
and equality with the engine's own, independently written design code (csrc/design.cpp),
reached through the device-free C-ABI entry point fmgpu_design_host."""
import ctypes as C

import numpy as np
import pytest

import fmtuner_sdr_b200 as fm


def _resp(h, f):
    n = np.arange(h.size)
    return np.abs(np.exp(-2j * np.pi * f * n) @ h.astype(np.float64))


def test_kaiser_lowpass_unity_dc_and_stopband(orc_libm):
    h, sc = orc_libm.design(0, 10, 28, 80.0)          # 2.4 MS/s decimator: 280 taps
    assert h.size == 280 and abs(sc - 0.09) < 1e-6
    assert abs(h.sum() * sc - 1.0) < 2e-3                 # scale 2*fc => unity DC gain
    assert np.allclose(h, h[::-1], atol=1e-7)             # linear phase
    # alias bands that fold onto +-100 kHz after /10 are >= 75 dB down
    assert 20 * np.log10(_resp(h * sc, 0.1 - 0.042)) < -75
    h8, sc8 = orc_libm.design(0, 8, 28, 80.0)
    assert h8.size == 224 and abs(sc8 - 2 * 0.45 / 8) < 1e-7


def test_channel_filter_table(orc_libm):
    # W0 = 194 kHz -> table index 7 -> 81 taps, cutoff 97 kHz (Appendix B.5)
    h, sc = orc_libm.design(1, 256000, 0)
    assert h.size == 81 and abs(sc - 2 * 97000 / 256000) < 1e-6
    h2, _ = orc_libm.design(1, 256000, 56000)             # <= 73 kHz -> 121 taps
    assert h2.size == 121
    h3, sc3 = orc_libm.design(1, 256000, 309000)          # index 0 == initial mode: ctor filter kept
    assert h3.size == 81 and abs(sc3 - 2 * 110000 / 256000) < 1e-6


@pytest.mark.parametrize("rate,taps", [(256000, 325), (240000, 305)])
def test_pilot_bandpass(orc_libm, rate, taps):
    h, sc = orc_libm.design(2, rate)
    assert h.size == taps and sc == 1.0
    assert abs(np.abs(h).sum() - 1.0) < 1e-5              # normalised to sum|h| = 1
    g19 = _resp(h, 19000 / rate)
    assert g19 > 0.5
    assert _resp(h, 15000 / rate) < g19 * 10 ** (-50 / 20)
    assert _resp(h, 23000 / rate) < g19 * 10 ** (-50 / 20)


def test_resampler_steps_and_bank(orc_libm):
    bank, step = orc_libm.design(4, 12, 0, np.float32(32000 / 256000))
    assert bank.size == 32 * 24 and int(step) == 8 << 24    # every 8th input, branch 0
    assert abs(bank.reshape(32, 24)[0].sum() - 1.0) < 0.02
    _, step240 = orc_libm.design(4, 12, 0, np.float32(32000 / 240000))
    assert abs(int(step240) / 2 ** 24 - 7.5) < 1e-6
    _, step_rds = orc_libm.design(4, 13, 0, np.float32(171000 / 240000))
    assert abs(int(step_rds) / 2 ** 24 - 240 / 171) < 1e-6


def test_symsync_bank(orc_libm):
    mf, _ = orc_libm.design(6)
    dmf, _ = orc_libm.design(7)
    assert mf.size == 32 * 18 and dmf.size == 32 * 18
    assert abs(np.abs(mf).max() - (1 - 0.8 + 4 * 0.8 / np.pi)) < 1e-5   # RRC peak at t = 0
    assert abs(np.abs(dmf * mf).max() - 0.06) < 2e-3                      # 0.06 / max|h dh|


CASES = [(2_400_000, 10), (2_048_000, 8), (1_024_000, 4), (256_000, 1)]


@pytest.mark.parametrize("iq_rate,decim", CASES)
def test_engine_design_equals_oracle_design(orc_libm, iq_rate, decim):
    """Two independent implementations of the same design formulas give identical floats."""
    L = fm.load_library()
    L.fmgpu_design_host.restype = C.c_size_t
    L.fmgpu_design_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                    C.POINTER(C.c_float)]
    cfg = fm.make_config(iq_rate=iq_rate, decimation=decim)
    fs = iq_rate // decim

    def eng(which, bw=0):
        buf = np.zeros(65536, np.float32)
        sc = C.c_float(0)
        n = L.fmgpu_design_host(C.byref(cfg), which, bw, buf.ctypes.data, buf.size, C.byref(sc))
        return buf[:n].copy(), sc.value

    pairs = [(eng(2), orc_libm.design(2, fs)), (eng(3), orc_libm.design(3, fs)),
             (eng(4), orc_libm.design(4, 12, 0, np.float32(32000) / np.float32(fs)))]
    if decim > 1:
        tpp = 28 if decim >= 8 else 20
        pairs.append((eng(0), orc_libm.design(0, decim, tpp, 80.0)))
    for bw in (0, 56000, 309000, 36000, 133000):
        pairs.append((eng(1, bw), orc_libm.design(1, fs, bw)))
    for (a, sa), (b, sb) in pairs:
        assert a.size == b.size and a.size > 0
        assert np.array_equal(a, b)
        assert sa == sb
    if fs == 240000:
        for w, ow in ((5, 5), (6, 6), (7, 7)):
            (a, sa), (b, sb) = eng(w), orc_libm.design(ow)
            assert np.array_equal(a, b) and sa == sb
        (a, sa), (b, sb) = eng(8), orc_libm.design(4, 13, 0, np.float32(171000.0) / np.float32(fs))
        assert np.array_equal(a, b) and sa == sb


def test_kaiser_design_against_scipy(orc_libm):
    """Independent pin of liquid_firdes_kaiser as restated (Appendix A.1): the same taps from
    scipy's Kaiser window and numpy's normalised sinc, to float rounding. The beta(As) rule is
    Kaiser's, which scipy.signal.kaiser_beta also implements."""
    from scipy.signal import kaiser_beta
    from scipy.signal.windows import kaiser

    for args, n, fc, att in (((0, 10, 28, 80.0), 280, 0.045, 80.0),     # 2.4 MS/s decimator
                             ((0, 8, 28, 80.0), 224, 0.45 / 8, 80.0),    # 2.048 MS/s decimator
                             ((1, 256000, 0), 81, 97000 / 256000, 60.0)):  # channel filter, 194 kHz
        h, _ = orc_libm.design(*args)
        assert h.size == n
        t = np.arange(n) - (n - 1) / 2.0
        ref = np.sinc(2.0 * fc * t) * kaiser(n, kaiser_beta(att), sym=True)
        assert np.abs(h.astype(np.float64) - ref).max() < 2e-7, args
