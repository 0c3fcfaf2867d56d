"""The CPU oracle against (1) its committed golden fixtures, (2) physics-level expectations
for the synthetic multiplex and (3) itself across the two
transcendental flavours (libm — faithful to the reference call sites — vs fm_math, the
flavour the engine is compared with bit for bit)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import orc
from tests.common import groups_equal, rates, snr_db

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name,rate", [("config1_240k", "240k"), ("config1_256k", "256k")])
def test_golden_fixture(orc_fm, name, rate):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    iq_rate, decim = rates(rate)
    iq = orc.config1_signal(fs_iq=iq_rate).generate(12 * 8192 * decim)
    assert np.array_equal(iq[:64], g["iq_head"])
    assert np.array_equal(np.frombuffer(hashlib.sha256(iq.tobytes()).digest(), np.uint8),
                          g["iq_sha256"])
    ch = orc.Channel(orc_fm, orc.make_config(iq_rate=iq_rate, decimation=decim))
    r = ch.process(iq, debug=True)
    assert np.array_equal(r.left, g["left"]) and np.array_equal(r.right, g["right"])
    assert np.array_equal(r.mpx[::64], g["mpx_every64"])
    assert np.array_equal(r.status, g["status"])
    assert groups_equal(r.groups, g["groups"])
    assert np.array_equal(ch.rds_bits(), g["rds_bits"])


@pytest.mark.parametrize("name,kind,c", [("reference_config1_240k", "config1", None),
                                         ("reference_config3_ch7", "config3", 7)])
def test_reference_made_fixture(orc_libm, name, kind, c):
    """Fixtures written by the REFERENCE'S OWN SOURCES run in the build container
    (oracle/_ref/libfmref.so; tests/golden/make_golden.py main_reference): the restated oracle in its
    faithful (libm) flavour must reproduce them wherever the tests run, /root/reference present or not.
    Bit for bit on the image they were made on; a host whose libm rounds sinf / cosf / expf / atan2f
    differently may move floats by an ulp, which the fallback bound allows (integers stay exact)."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    iq_rate, decim = rates("240k")
    sig = orc.config1_signal(fs_iq=iq_rate) if kind == "config1" else orc.config3_signal(c, fs_iq=iq_rate)
    iq = sig.generate(14 * 8192 * decim)
    assert np.array_equal(np.frombuffer(hashlib.sha256(iq.tobytes()).digest(), np.uint8),
                          g["iq_sha256"])
    ch = orc.Channel(orc_libm, orc.make_config(iq_rate=iq_rate, decimation=decim))
    ch.enable_bits_tap()
    r = ch.process(iq, debug=True)
    pairs = [(r.dec[::64], g["dec_every64"]), (r.mpx[::16], g["mpx_every16"]), (r.left, g["left"]),
             (r.right, g["right"])]
    assert all(a.shape == b.shape for a, b in pairs)
    if not all(np.array_equal(a, b) for a, b in pairs):
        assert max(float(np.abs(a - b).max()) for a, b in pairs) < 5e-6
    assert np.array_equal(r.status["stereo"], g["status"]["stereo"])
    assert np.array_equal(r.status["n_audio"], g["status"]["n_audio"])
    assert np.abs(r.status["pilot_tenths"].astype(int) - g["status"]["pilot_tenths"].astype(int)).max() <= 1
    assert groups_equal(r.groups, g["groups"]) and len(r.groups) >= 2
    assert np.array_equal(ch.rds_bits(), g["rds_bits"])


@pytest.fixture(scope="module")
def config1(orc_libm):
    iq_rate, decim = rates("240k")
    iq = orc.config1_signal(fs_iq=iq_rate).generate(40 * 8192 * decim)
    ch = orc.Channel(orc_libm, orc.make_config(iq_rate=iq_rate, decimation=decim))
    return iq, ch.process(iq, debug=True)


def test_mpx_scale_and_pilot_lock(config1):
    _, r = config1
    # +-75 kHz deviation <-> MPX +-1: the composite peaks at 0.43+0.43+0.10+0.04 ~ 1
    assert 0.85 < np.abs(r.mpx[20000:]).max() < 1.05
    # stereo flag needs 6 good blocks (stereo_decoder.cpp:6,263-270): lock at block 6..8
    first = int(np.flatnonzero(r.status["stereo"])[0])
    assert 6 <= first <= 8
    assert r.status["stereo"][first:].all()
    assert (r.status["n_audio"].sum() == r.left.size) and abs(r.left.size - 40 * 8192 * 32000 / 240000) < 2
    assert (np.diff(r.status["pilot_tenths"][first:]) >= -1).all()   # settles monotonically


def test_left_only_tone_separation(config1):
    _, r = config1
    l, rr = r.left[-16000:], r.right[-16000:]
    sep_db = 20 * np.log10(np.sqrt((l ** 2).mean()) / np.sqrt((rr ** 2).mean()))
    assert sep_db > 20.0        # Appendix B.2: finite separation from the (N-1)/2 + 1 delay
    spec = np.abs(np.fft.rfft(l * np.hanning(l.size)))
    assert abs(np.argmax(spec) * 32000 / l.size - 1000.0) < 4.0


def test_rds_payload_decodes(config1):
    _, r = config1
    assert len(r.groups) >= 10
    assert orc.decode_ps_rt(r.groups) == (0x1234, "B200TEST", "FM ON B200")
    assert (r.groups["errors"][1:] == 0).all()


def test_deemphasis_corner(orc_libm):
    # 50 us: alpha = dt/(tau+dt) at 32 kHz = 0.384615; 10 kHz tone ~ -9.6 dB vs 400 Hz
    amp = {}
    for f in (400.0, 10000.0):
        sig = orc.Signal(fs_iq=2_400_000, tone_l_hz=f, tone_l_amp=0.5, tone_r_hz=f, tone_r_amp=0.5,
                         rds_amp=0.0)
        iq = sig.generate(14 * 81920)
        r = orc.Channel(orc_libm, orc.make_config()).process(iq)
        amp[f] = np.sqrt((r.left[-8000:] ** 2).mean())
    ratio_db = 20 * np.log10(amp[10000.0] / amp[400.0])
    assert -12.0 < ratio_db < -8.0


@pytest.mark.parametrize("rate", ["240k", "256k"])
def test_libm_and_fm_flavours_agree(orc_libm, orc_fm, rate):
    """The only deviation of the engine-comparable oracle from the libm-faithful one is the
    transcendental kernels; end to end that costs < 1e-4 of full scale and no RDS bit."""
    iq_rate, decim = rates(rate)
    iq = orc.config1_signal(fs_iq=iq_rate).generate(24 * 8192 * decim)
    a = orc.Channel(orc_libm, orc.make_config(iq_rate=iq_rate, decimation=decim))
    b = orc.Channel(orc_fm, orc.make_config(iq_rate=iq_rate, decimation=decim))
    ra, rb = a.process(iq, debug=True), b.process(iq, debug=True)
    assert np.abs(ra.mpx - rb.mpx).max() < 2e-6
    assert np.abs(ra.left - rb.left).max() < 1e-4 and np.abs(ra.right - rb.right).max() < 1e-4
    assert snr_db(ra.left[8000:], rb.left[8000:]) > 90.0
    assert np.array_equal(ra.status["stereo"], rb.status["stereo"])
    assert np.abs(ra.status["pilot_tenths"] - rb.status["pilot_tenths"]).max() <= 1
    assert np.array_equal(a.rds_bits(), b.rds_bits())
    assert groups_equal(ra.groups, rb.groups)


def test_block_boundaries_are_the_callers(orc_fm):
    """Appendix B.12: results depend on the logical block length only through the per-block
    stereo logic; FIR / IIR / RDS outputs do not depend on how the stream is cut."""
    iq_rate, decim = rates("240k")
    iq = orc.config1_signal(fs_iq=iq_rate).generate(8 * 8192 * decim)
    a = orc.Channel(orc_fm, orc.make_config(block_samples=8192)).process(iq, debug=True)
    b = orc.Channel(orc_fm, orc.make_config(block_samples=4096)).process(iq, debug=True)
    assert np.array_equal(a.mpx, b.mpx)
    assert np.array_equal(a.groups[["a", "b", "c", "d", "errors"]], b.groups[["a", "b", "c", "d", "errors"]])


def test_empty_and_null_arguments(orc_libm):
    L = orc_libm.lib
    d = L.orc_decim_create(8, 28, 80.0)
    assert L.orc_decim_execute_complex(d, None, 0, None, 0) == 0
    L.orc_decim_destroy(d)
    s = L.orc_stereo_create(256000)
    assert L.orc_stereo_process(s, None, None, None, 0) == 0
    L.orc_stereo_destroy(s)
    a = L.orc_afpost_create(256000, 32000)
    assert L.orc_afpost_process(a, None, None, 0, None, None, 0) == 0
    L.orc_afpost_destroy(a)


@pytest.mark.parametrize("iq_rate,decim", [(2_400_000, 10), (2_048_000, 8)])
def test_front_end_against_scipy_float64_model(orc_libm, iq_rate, decim):
    """Independent pin of stages a1/a2: the decimating FIR, the I/Q DC blockers, the
    channel filter and the quadrature discriminator rebuilt from scipy.signal.lfilter in float64
    — firdecim_crcf emits output n right after input n*M (Appendix A.7), iirfilt dc_blocker is
    b = [1, -1], a = [1, -(1 - alpha)], freqdem is arg(conj(y[n-1]) y[n]) / (2 pi kf) — must give
    the oracle's decimated IQ and MPX to float32 rounding."""
    from scipy.signal import lfilter

    fs = iq_rate // decim
    iq = orc.config1_signal(fs_iq=iq_rate).generate(3 * 8192 * decim)
    ref = orc.Channel(orc_libm, orc.make_config(iq_rate=iq_rate, decimation=decim)).process(iq, debug=True)
    x = ((iq[0::2].astype(np.float64) - 127.5) + 1j * (iq[1::2].astype(np.float64) - 127.5)) / 127.5
    h, sc = orc_libm.design(0, decim, 28, 80.0)
    dec = sc * lfilter(h.astype(np.float64), [1.0], x)[0::decim]
    assert np.abs(dec - ref.dec.astype(np.complex128)).max() < 5e-6
    w = lfilter([1.0, -1.0], [1.0, -(1.0 - 0.0005)], dec)
    hc, sc1 = orc_libm.design(1, fs, 0)
    z = sc1 * lfilter(hc.astype(np.float64), [1.0], w)
    zp = np.concatenate([[0.0], z[:-1]])
    mpx = np.angle(np.conj(zp) * z) / (2.0 * np.pi * 75000.0 / fs)
    assert np.abs(mpx - ref.mpx).max() < 2e-4   # of a +-1.6 swing


def test_audio_tail_against_scipy_float64_model(orc_libm):
    """Independent pin of stage a9 (AFPostProcessor::process) at 2.048 MS/s / 8, where 256 kHz ->
    32 kHz is exactly every 8th input through branch 0 of the 32-branch resampler: branch 0 as an
    FIR (scipy.signal.lfilter, float64), decimation by 8, de-emphasis iirfilt b = [alpha],
    a = [1, -(1 - alpha)] with alpha = dt / (tau + dt) at 50 us, DC blocker alpha = 0.005 must give
    the oracle's 32 kHz audio from its own DSP-rate L/R."""
    from scipy.signal import lfilter

    iq_rate, decim = 2_048_000, 8
    iq = orc.config1_signal(fs_iq=iq_rate).generate(4 * 8192 * decim)
    ref = orc.Channel(orc_libm, orc.make_config(iq_rate=iq_rate, decimation=decim)).process(iq, debug=True)
    bank, step = orc_libm.design(4, 12, 0, np.float32(32000 / 256000))
    assert int(step) == 8 << 24
    h0 = bank.reshape(32, 24)[0].astype(np.float64)[::-1]   # stored oldest-sample-first
    alpha = (1 / 32000) / (50e-6 + 1 / 32000)
    for dsp, out in ((ref.sl, ref.left), (ref.sr, ref.right)):
        y = lfilter(h0, [1.0], dsp.astype(np.float64))[0::8]
        y = lfilter([alpha], [1.0, -(1.0 - alpha)], y)
        y = lfilter([1.0, -1.0], [1.0, -(1.0 - 0.005)], y)
        assert y.size == out.size
        assert np.abs(y - out).max() < 2e-6


def test_stereo_decoder_against_float64_model(orc_libm):
    """Independent pin of stage a6 (StereoDecoder::processAudio) once locked and fully blended: pilot
    band-pass as an FIR (scipy), an ideal float64 second-order PLL with liquid's loop constants
    (alpha = bw, beta = sqrt(bw), bw = 0.01; error = pilot * sin(theta)), L-R = 2 d cos(2 theta) on the
    MPX delayed by (N - 1) / 2 + 1 samples, the two 15 kHz low-pass filters. The only float32 effect
    modelled is the blend recursion itself: b += (1 - b) * attack stalls at 1 - 9.1e-4 in float32
    (the reference behaves the same), so L - R carries that gain."""
    import math

    from scipy.signal import lfilter

    iq_rate, decim, nblk = 2_048_000, 8, 80
    fs = iq_rate // decim
    iq = orc.config1_signal(fs_iq=iq_rate).generate(nblk * 8192 * decim)
    ref = orc.Channel(orc_libm, orc.make_config(iq_rate=iq_rate, decimation=decim)).process(iq, debug=True)
    assert ref.status["stereo"][-1] == 1
    mpx = ref.mpx.astype(np.float64)
    hp, scp = orc_libm.design(2, fs)
    ha, sca = orc_libm.design(3, fs)
    pilot = scp * lfilter(hp.astype(np.float64), [1.0], mpx)
    theta, dtheta = 0.0, 2.0 * math.pi * 19000.0 / fs
    c2 = np.empty(mpx.size)
    for i in range(mpx.size):
        err = pilot[i] * math.sin(theta)
        dtheta += 0.01 * err
        theta += 0.1 * err + dtheta
        if theta > math.pi:
            theta -= 2.0 * math.pi
        c2[i] = math.cos(2.0 * theta)
    delay = (hp.size - 1) // 2 + 1
    d = np.concatenate([np.zeros(delay), mpx[:-delay]])
    # float32 fixed point of the blend recursion (normal mode: attack tau 0.12 s)
    attack = np.float32(1.0) - np.exp(np.float32(-1.0) / (np.float32(0.120) * np.float32(fs)))
    b = np.float32(0.99)
    while True:
        nb = b + (np.float32(1.0) - b) * attack
        if nb == b:
            break
        b = nb
    assert 0.9985 < float(b) < 0.9995
    lr = 2.0 * d * c2 * float(b)
    sl = sca * lfilter(ha.astype(np.float64), [1.0], 0.5 * (d + lr))
    sr = sca * lfilter(ha.astype(np.float64), [1.0], 0.5 * (d - lr))
    k0 = (nblk - 2) * 8192
    assert np.abs(sl[k0:] - ref.sl[k0:]).max() < 1e-4      # of a 0.42 swing
    assert np.abs(sr[k0:] - ref.sr[k0:]).max() < 1e-4
    assert np.abs((sl + sr)[k0:] - (ref.sl.astype(np.float64) + ref.sr)[k0:]).max() < 2e-6


@pytest.mark.parametrize("mode", [1, 2])
def test_agc_only_scales_the_discriminator_input(orc_libm, mode):
    """Physics pin of the pre-discriminator AGC (fm_demod.cpp:196-199): a positive real gain on y[n]
    cannot move arg(conj(y[n-1]) y[n]), so MPX and audio with dsp_agc fast / slow equal the AGC-off
    result to float32 rounding."""
    iq = orc.config1_signal(fs_iq=2_400_000).generate(4 * 8192 * 10)
    off = orc.Channel(orc_libm, orc.make_config(dsp_agc=0)).process(iq, debug=True)
    on = orc.Channel(orc_libm, orc.make_config(dsp_agc=mode)).process(iq, debug=True)
    assert np.abs(off.mpx - on.mpx).max() < 2e-6
    assert np.abs(off.left - on.left).max() < 2e-6 and np.abs(off.right - on.right).max() < 2e-6


@pytest.mark.parametrize("snr", [12.0, 25.0])
def test_agc_elision_is_invisible_to_the_reference_itself(snr):
    """The fused channel filter + discriminator of the engine's fast flavour (fmgpu_set_demod_mode 1)
    leaves the pre-discriminator AGC out. Justification on the reference's OWN code (libfmref.so when
    /root/reference was present at build time, else the libm restatement that equals it): FMDemod with
    dsp_agc fast against dsp_agc off on weak, noisy channels — the multiplex agrees to one float ulp
    and the RDS groups are the same; the audio moves by ~2e-5, which is what one ulp of MPX does to
    the pilot PLL and blend, i.e. the reference's own sensitivity to rounding."""
    lib = orc.OracleLib("ref") if orc.OracleLib.have_ref("ref") else orc.OracleLib("libm")
    sig = orc.config3_signal(7, fs_iq=2_400_000)
    sig.snr_db = snr
    iq = sig.generate(24 * 8192 * 10)
    off = orc.Channel(lib, orc.make_config(dsp_agc=0)).process(iq, debug=True)
    on = orc.Channel(lib, orc.make_config(dsp_agc=1)).process(iq, debug=True)
    assert np.abs(off.mpx - on.mpx).max() <= 2.4e-7
    assert len(on.groups) >= 4 and len(on.groups) == len(off.groups)
    assert all((on.groups[k] == off.groups[k]).all() for k in ("a", "b", "c", "d", "errors"))
    assert np.abs(off.left - on.left).max() <= 1e-4 and np.abs(off.right - on.right).max() <= 1e-4
