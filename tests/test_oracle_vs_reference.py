"""The oracle pinned to the REFERENCE'S OWN float DSP code.

oracle/_ref/libfmref.so holds /root/reference's src/fm_demod.cpp, src/stereo_decoder.cpp,
src/af_post_processor.cpp, src/rds_decoder.cpp, src/dsp/liquid_primitives.cpp and
src/redsea_port/** compiled UNMODIFIED, in place (oracle/Makefile), over oracle/liquid_shim — a
liquid/liquid.h whose definitions sit on oracle/liquid_restated.hpp — behind the same per-block
harness as the restated oracle (oracle_capi.cpp, main.cpp:1232-1308). Every float the restated
pipeline (oracle/pipeline.hpp, the `libm` flavour) produces must EQUAL what the reference's own
classes produce: decimated IQ, MPX, DSP-rate L/R, 32 kHz audio, per-block status, RDS bits and
groups. After this only liquid-dsp's internals remain a restatement.

`ref_contract` is the same sources with gcc's default FMA contraction (what a stock -mfma build
of the reference does): compared in the tolerance north_star states.
"""
import numpy as np
import pytest

from oracle import orc
from tests.common import groups_equal, rates, snr_db

pytestmark = pytest.mark.skipif(not orc.OracleLib.have_ref("ref"),
                                reason="oracle/_ref/libfmref.so not built (no /root/reference)")


@pytest.fixture(scope="module")
def ref():
    return orc.OracleLib("ref")


@pytest.fixture(scope="module")
def ref_contract():
    return orc.OracleLib("ref_contract")


def run(lib, cfg, iq, bits=True):
    ch = orc.Channel(lib, orc.make_config(**cfg))
    if bits:
        ch.enable_bits_tap()
    r = ch.process(iq, debug=True)
    return ch, r, (ch.rds_bits() if bits else None)


def assert_identical(a, b, abits=None, bbits=None):
    if a.dec is not None:
        assert np.array_equal(a.dec, b.dec)
    assert np.array_equal(a.mpx, b.mpx)
    assert np.array_equal(a.sl, b.sl) and np.array_equal(a.sr, b.sr)
    assert np.array_equal(a.left, b.left) and np.array_equal(a.right, b.right)
    assert np.array_equal(a.status, b.status)
    assert groups_equal(a.groups, b.groups)
    if abits is not None:
        assert abits.size > 0 and np.array_equal(abits, bbits)


@pytest.mark.parametrize("rate", ["240k", "256k", "1024k", "direct256k"])
def test_config1_identical(orc_libm, ref, rate):
    """BASELINE config 1 signal, every rate the reference (and the bench) runs."""
    iq_rate, decim = rates(rate)
    iq = orc.config1_signal(fs_iq=iq_rate).generate(20 * 8192 * decim)
    cfg = dict(iq_rate=iq_rate, decimation=decim)
    _, a, ab = run(orc_libm, cfg, iq)
    _, b, bb = run(ref, cfg, iq)
    assert_identical(a, b, ab, bb)
    assert a.status["stereo"][-1] == 1 and len(a.groups) >= 3


@pytest.mark.parametrize("rate,nblk", [("240k", 293), ("256k", 313)])
def test_config1_ten_seconds_identical(orc_libm, ref, rate, nblk):
    """The full 10 s of BASELINE configs 1 / 2 (293 blocks at 240 k, 313 at 256 k): resampler phases
    drifting over seconds, > 110 RDS groups, the stereo lock held for the whole run. The restated
    oracle and the reference's own sources must still agree on every float, bit and group."""
    iq_rate, decim = rates(rate)
    iq = orc.config1_signal(fs_iq=iq_rate).generate(nblk * 8192 * decim)
    cfg = dict(iq_rate=iq_rate, decimation=decim)
    _, a, ab = run(orc_libm, cfg, iq)
    _, b, bb = run(ref, cfg, iq)
    assert_identical(a, b, ab, bb)
    assert len(a.groups) > 100 and a.status["stereo"][-1] == 1


@pytest.mark.parametrize("c", [0, 3, 7, 11, 100, 255])
def test_config3_channels_identical(orc_libm, ref, c):
    """BASELINE config 3: varied deviation / SNR / tones / RDS payloads."""
    iq_rate, decim = rates("240k")
    iq = orc.config3_signal(c, fs_iq=iq_rate).generate(14 * 8192 * decim)
    cfg = dict(iq_rate=iq_rate, decimation=decim)
    _, a, ab = run(orc_libm, cfg, iq)
    _, b, bb = run(ref, cfg, iq)
    assert_identical(a, b, ab, bb)


@pytest.mark.parametrize("snr,agc,blend,deemph,bw", [
    (10.0, 1, 0, 0, 0), (15.0, 1, 1, 1, 0), (20.0, 1, 2, 0, 56000), (25.0, 2, 0, 2, 0),
    (30.0, 1, 1, 0, 133000), (40.0, 2, 2, 1, 311000), (12.0, 0, 1, 0, 36000)])
def test_config5_weak_signal_identical(orc_libm, ref, snr, agc, blend, deemph, bw):
    """BASELINE config 5: weak signals, dsp_agc fast/slow, all blend modes, de-emphasis modes and
    channel bandwidths (81- and 121-tap, 60 and 70 dB filters)."""
    iq_rate, decim = rates("240k")
    sig = orc.config3_signal(int(snr), fs_iq=iq_rate)
    sig.snr_db = snr
    iq = sig.generate(16 * 8192 * decim)
    cfg = dict(iq_rate=iq_rate, decimation=decim, dsp_agc=agc, stereo_blend=blend,
               deemphasis=deemph, bandwidth_hz=bw)
    _, a, ab = run(orc_libm, cfg, iq)
    _, b, bb = run(ref, cfg, iq)
    assert_identical(a, b, ab, bb)


def test_mono_mode_and_force_flags_identical(orc_libm, ref):
    iq_rate, decim = rates("256k")
    iq = orc.config1_signal(fs_iq=iq_rate).generate(10 * 8192 * decim)
    for cfg in (dict(stereo=0), dict(force_mono=1), dict(stereo=0, deemphasis=2)):
        cfg = dict(iq_rate=iq_rate, decimation=decim, **cfg)
        _, a, _ = run(orc_libm, cfg, iq, bits=False)
        _, b, _ = run(ref, cfg, iq, bits=False)
        assert_identical(a, b)
    # force stereo set after construction
    out = []
    for lib in (orc_libm, ref):
        ch = orc.Channel(lib, orc.make_config(iq_rate=iq_rate, decimation=decim))
        ch.set_force_stereo(True)
        out.append(ch.process(iq, debug=True))
    assert_identical(*out)


def test_reset_retune_and_bandwidth_change_identical(orc_libm, ref):
    """The dspRuntime reset handler (main.cpp:686-691), an RDS reset, an XDR 'W' bandwidth change
    and a de-emphasis change between blocks."""
    iq_rate, decim = rates("240k")
    per = 8192 * decim * 2
    iq1 = orc.config1_signal(fs_iq=iq_rate).generate(9 * 8192 * decim)
    iq2 = orc.config3_signal(5, fs_iq=iq_rate).generate(9 * 8192 * decim)
    outs = []
    for lib in (orc_libm, ref):
        ch = orc.Channel(lib, orc.make_config(iq_rate=iq_rate, decimation=decim, dsp_agc=1))
        ch.enable_bits_tap()
        res = [ch.process(iq1[:5 * per], debug=True)]
        ch.set_bandwidth_hz(84000)
        res.append(ch.process(iq1[5 * per:], debug=True))
        ch.reset(dsp=True, rds=True)
        res.append(ch.process(iq2[:4 * per], debug=True))
        ch.set_deemphasis(1)
        ch.set_bandwidth_hz(0)
        ch.reset(dsp=False, rds=True)
        res.append(ch.process(iq2[4 * per:], debug=True))
        outs.append((res, ch.rds_bits()))
    for a, b in zip(outs[0][0], outs[1][0]):
        assert_identical(a, b)
    assert np.array_equal(outs[0][1], outs[1][1])


def test_class_level_ragged_calls_identical(orc_libm, ref):
    """The class methods called directly with ragged lengths (the stage-level C ABI mirrors
    these): ComplexDecimator::executeComplex / execute (uint8), FMDemod::processSplit with the
    mono chain, StereoDecoder::processAudio, AFPostProcessor::process with a capacity cut-off,
    RDSDecoder::process."""
    import ctypes as C
    rng = np.random.default_rng(7)
    iq_rate, decim = rates("256k")
    fs = iq_rate // decim
    iq = orc.config1_signal(fs_iq=iq_rate).generate(6 * 8192 * decim)
    res = []
    for lib in (orc_libm, ref):
        L = lib.lib
        p = lambda a, t=C.c_float: a.ctypes.data_as(C.POINTER(t))
        out = {}
        d = L.orc_decim_create(decim, 28, 80.0)
        d8 = L.orc_decim_create(decim, 28, 80.0)
        dm = L.orc_demod_create(fs, 32000)
        L.orc_demod_set_w0(dm, 194000)
        L.orc_demod_set_bandwidth_mode(dm, 9)
        L.orc_demod_set_deemphasis(dm, 50)
        st = L.orc_stereo_create(fs)
        af = L.orc_afpost_create(fs, 32000)
        L.orc_afpost_set_deemphasis(af, 50)
        rd = L.orc_rds_create(fs)
        chunks = [3000 * decim, 8192 * decim, 1 * decim, 5001 * decim, 777 * decim]
        pos = 0
        dec_all, u8_all, mpx_all, mono_all, l_all, r_all, al, ar, groups = [], [], [], [], [], [], [], [], []
        for n_in in chunks:
            seg = np.ascontiguousarray(iq[2 * pos:2 * (pos + n_in)])
            pos += n_in
            n = n_in // decim
            dec = np.zeros(2 * n, np.float32)
            assert L.orc_decim_execute_complex(d, p(seg, C.c_uint8), n_in, p(dec), n) == n
            u8 = np.zeros(2 * n, np.uint8)
            assert L.orc_decim_execute_u8(d8, p(seg, C.c_uint8), n_in, p(u8, C.c_uint8), n) == n
            mpx = np.zeros(n, np.float32)
            mono = np.zeros(n, np.float32)
            nm = L.orc_demod_process_split_complex(dm, p(dec), p(mpx), p(mono), n)
            l = np.zeros(n, np.float32)
            r = np.zeros(n, np.float32)
            assert L.orc_stereo_process(st, p(mpx), p(l), p(r), n) == n
            cap = max(1, n // 9)    # cuts the block short (af_post_processor.cpp:56)
            ol = np.zeros(cap, np.float32)
            orr = np.zeros(cap, np.float32)
            na = L.orc_afpost_process(af, p(l), p(r), n, p(ol), p(orr), cap)
            g = np.zeros(8, orc.GROUP_DTYPE)
            ng = L.orc_rds_process(rd, p(mpx), n, g.ctypes.data, 8)
            dec_all.append(dec); u8_all.append(u8); mpx_all.append(mpx); mono_all.append(mono[:nm])
            l_all.append(l); r_all.append(r); al.append(ol[:na]); ar.append(orr[:na])
            groups.append(g[:ng])
            out.setdefault("stereo", []).append((L.orc_stereo_is_stereo(st), L.orc_stereo_pilot_tenths(st),
                                                 L.orc_demod_clip_ratio(dm)))
        for k, v in (("dec", dec_all), ("u8", u8_all), ("mpx", mpx_all), ("mono", mono_all),
                     ("l", l_all), ("r", r_all), ("al", al), ("ar", ar), ("groups", groups)):
            out[k] = np.concatenate(v)
        res.append(out)
        for h, fn in ((d, L.orc_decim_destroy), (d8, L.orc_decim_destroy), (dm, L.orc_demod_destroy),
                      (st, L.orc_stereo_destroy), (af, L.orc_afpost_destroy), (rd, L.orc_rds_destroy)):
            fn(h)
    a, b = res
    assert a["stereo"] == b["stereo"]
    for k in ("dec", "u8", "mpx", "mono", "l", "r", "al", "ar"):
        assert a[k].size > 0 and np.array_equal(a[k], b[k]), k
    assert groups_equal(a["groups"], b["groups"], keys=("a", "b", "c", "d", "errors"))
    assert rng is not None


def test_stock_fma_contraction_build_within_tolerance(orc_libm, ref_contract):
    """A stock build of the reference (-mavx2 -mfma, gcc contracts a*b+c outside the dot
    products): same lock block, same RDS bits and groups, audio within north_star's tolerance
    (max-abs <= 1e-4 of full scale or SNR >= 90 dB, after lock)."""
    iq_rate, decim = rates("240k")
    for sig in (orc.config1_signal(fs_iq=iq_rate), orc.config3_signal(9, fs_iq=iq_rate)):
        iq = sig.generate(24 * 8192 * decim)
        cfg = dict(iq_rate=iq_rate, decimation=decim, dsp_agc=1)
        _, a, ab = run(orc_libm, cfg, iq)
        _, b, bb = run(ref_contract, cfg, iq)
        assert np.array_equal(ab, bb) and groups_equal(a.groups, b.groups)
        assert np.array_equal(a.status["stereo"], b.status["stereo"])
        assert np.abs(a.status["pilot_tenths"] - b.status["pilot_tenths"]).max() <= 1
        lock = int(np.flatnonzero(a.status["stereo"])[0])
        s0 = int(a.status["n_audio"][:lock + 2].sum())
        for x, y in ((a.left, b.left), (a.right, b.right)):
            # north_star: "<= 1e-4 of full scale max-abs, OR >= 90 dB SNR"
            assert np.abs(x[s0:] - y[s0:]).max() <= 1e-4 or snr_db(x[s0:], y[s0:]) >= 90.0
            assert snr_db(x[s0:], y[s0:]) >= 80.0
