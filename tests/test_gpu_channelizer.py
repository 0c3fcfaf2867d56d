"""BASELINE config 4: the wideband channelizer and the complex-float batch path behind it.

The channelizer has no reference counterpart (the reference tunes one carrier in hardware), so
(1) it is checked against its own published definition evaluated in float64 here,
(2) its pass band / stop band is checked with tones,
(3) a synthetic band with several FM-stereo+RDS carriers goes wideband -> channelizer ->
    fmgpu_process_batch_cf32, and the result must equal, bit for bit, the oracle's reference classes
    (FMDemod::processSplitComplex -> StereoDecoder -> AFPostProcessor, RDSDecoder) fed with the same
    channelizer output, block by block as main.cpp:1285-1293 does."""
import ctypes as C

import numpy as np
import pytest

import fmtuner_sdr_b200 as fm
from oracle import orc
from tests.common import groups_equal

pytestmark = pytest.mark.gpu
F = C.POINTER(C.c_float)


def _f(a):
    return a.ctypes.data_as(F)


def _model(z: "fm.Channelizer", iq_u8: np.ndarray, ks, n_out: int) -> np.ndarray:
    """float64 evaluation of y_k[m] = sum_n h[n] x[mD-n] exp(-j 2 pi f_k (mD-n) / Fs)."""
    h = z.taps().astype(np.float64)
    L, D = h.size, z.decimation
    x = (iq_u8.astype(np.float64) - 127.5) / 127.5
    x = x[0::2] + 1j * x[1::2]
    xp = np.concatenate([np.zeros(L - 1, np.complex128), x])   # input before the first call = 0
    out = np.zeros((len(ks), n_out), np.complex128)
    n = np.arange(L)
    for i, k in enumerate(ks):
        nu = (z.first_center_hz + z.spacing_hz * k) / z.wide_rate
        for m in range(n_out):
            s = m * D - n                                        # absolute sample indices
            seg = xp[(L - 1) + s]
            out[i, m] = np.sum(h * seg * np.exp(-2j * np.pi * nu * s))
    return out


def test_channelizer_matches_float64_definition_across_calls():
    import torch
    dev = torch.device("cuda", 0)
    z = fm.Channelizer()
    assert z.output_rate == 240_000 and z.taps().size == 1600
    rng = np.random.default_rng(4)
    n1, n2 = 3200, 4800
    iq = rng.integers(0, 256, 2 * (n1 + n2), dtype=np.uint8)
    iq_dev = torch.from_numpy(iq).to(dev)
    n_out = (n1 + n2) // 100
    out = torch.zeros((100, n_out, 2), dtype=torch.float32, device=dev)
    z.process(iq_dev.data_ptr(), n1, out.data_ptr(), n_out)
    z.process(iq_dev.data_ptr() + 2 * n1, n2, out.data_ptr() + 8 * (n1 // 100), n_out)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    got = got[..., 0] + 1j * got[..., 1]
    ks = [0, 37, 50, 99]
    ref = _model(z, iq, ks, n_out)
    err = np.abs(got[ks] - ref).max()
    assert err < 2e-5, err            # FP32 accumulation of 1600 terms, outputs of magnitude <= 1
    # a sub-range of the channels (what one rank of a sharded job extracts) gives the same rows
    z2 = fm.Channelizer()
    part = torch.zeros((13, n_out, 2), dtype=torch.float32, device=dev)
    z2.process(iq_dev.data_ptr(), n1 + n2, part.data_ptr(), n_out, ch_first=36, ch_count=13)
    torch.cuda.synchronize()
    assert torch.equal(part, out[36:49])
    z.close()
    z2.close()


def test_channelizer_pass_band_and_stop_band():
    import torch
    dev = torch.device("cuda", 0)
    z = fm.Channelizer()
    n = 100 * 400
    t = np.arange(n)
    f5 = z.first_center_hz + 5 * z.spacing_hz + 30_000.0      # 30 kHz inside channel 5
    x = 0.7 * np.exp(2j * np.pi * f5 * t / z.wide_rate)
    iq = np.empty(2 * n, np.uint8)
    iq[0::2] = np.clip(np.round(127.5 + 127.5 * x.real), 0, 255)
    iq[1::2] = np.clip(np.round(127.5 + 127.5 * x.imag), 0, 255)
    iq_dev = torch.from_numpy(iq).to(dev)
    out = torch.zeros((100, n // 100, 2), dtype=torch.float32, device=dev)
    z.process(iq_dev.data_ptr(), n, out.data_ptr(), n // 100)
    torch.cuda.synchronize()
    y = out.cpu().numpy()
    y = (y[..., 0] + 1j * y[..., 1])[:, 32:]                   # past the filter's start-up
    p = (np.abs(y) ** 2).mean(axis=1)
    assert abs(np.sqrt(p[5]) - 0.7) < 0.01                      # unity pass-band gain
    # the tone comes out of channel 5 at +30 kHz
    ph = np.unwrap(np.angle(y[5]))
    assert abs((ph[-1] - ph[0]) / (2 * np.pi * (len(ph) - 1)) * 240_000 - 30_000.0) < 50.0
    others = np.delete(p, [4, 5, 6])
    assert 10 * np.log10(others.max() / p[5]) < -45.0           # u8 quantisation floor ~ -50 dB
    z.close()


def _oracle_chain(lib, x_cf: np.ndarray, nblk: int, N: int = 8192):
    """main.cpp:1285-1293 on the reference classes of the oracle, block by block."""
    L = lib.lib
    fs = 240_000
    dm = L.orc_demod_create(fs, 32000)
    st = L.orc_stereo_create(fs)
    af = L.orc_afpost_create(fs, 32000)
    rd = L.orc_rds_create(fs)
    # the calls main.cpp makes after construction, with fm.make_config's defaults (main.cpp:641-710)
    L.orc_demod_set_w0(dm, 194000)
    L.orc_demod_set_agc(dm, 0)
    L.orc_stereo_set_blend(st, 1)
    L.orc_afpost_set_deemphasis(af, 50)
    L.orc_demod_set_deemphasis(dm, 50)
    L.orc_stereo_set_force_mono(st, 0)
    L.orc_demod_set_bandwidth_hz(dm, 0)
    left, right, groups, stereo = [], [], [], []
    mpx = np.zeros(N, np.float32)
    sl, sr = np.zeros(N, np.float32), np.zeros(N, np.float32)
    ol, orr = np.zeros(N, np.float32), np.zeros(N, np.float32)
    gbuf = np.zeros(16, fm.GROUP_DTYPE)
    for b in range(nblk):
        blk = np.ascontiguousarray(x_cf[b * N:(b + 1) * N].view(np.float32))
        L.orc_demod_process_split_complex(dm, _f(blk), _f(mpx), None, N)
        ng = L.orc_rds_process(rd, _f(mpx), N, gbuf.ctypes.data, 16)
        g = gbuf[:ng].copy()
        g["block_index"] = b
        groups.append(g)
        L.orc_stereo_process(st, _f(mpx), _f(sl), _f(sr), N)
        k = L.orc_afpost_process(af, _f(sl), _f(sr), N, _f(ol), _f(orr), N)
        left.append(np.clip(ol[:k], -1.0, 1.0).copy())
        right.append(np.clip(orr[:k], -1.0, 1.0).copy())
        stereo.append(int(L.orc_stereo_is_stereo(st)))
    for h, d in ((dm, L.orc_demod_destroy), (st, L.orc_stereo_destroy), (af, L.orc_afpost_destroy),
                 (rd, L.orc_rds_destroy)):
        d(h)
    return np.concatenate(left), np.concatenate(right), np.concatenate(groups), stereo


def test_config4_band_to_audio_and_rds(orc_fm):
    """A 24 MS/s band with three FM-stereo+RDS carriers among 100 channel slots."""
    import torch
    dev = torch.device("cuda", 0)
    wide, D, N, nblk, chunk = 24_000_000, 100, 8192, 12, 4
    n_wide = nblk * N * D
    z = fm.Channelizer()
    stations = {7: (0x2207, "CH07 FM ", 600.0, 1500.0), 50: (0x1234, "B200TEST", 1000.0, 0.0),
                93: (0x4493, "NINETY3 ", 440.0, 880.0)}
    acc_i = np.zeros(n_wide, np.float32)
    acc_q = np.zeros(n_wide, np.float32)
    for k, (pi, ps, fl, fr) in stations.items():
        s = orc.Signal(fs_iq=float(wide), deviation=60_000.0, tone_l_hz=fl, tone_l_amp=0.8,
                       tone_r_hz=fr, tone_r_amp=0.8 if fr else 0.0, iq_amp=0.28,
                       freq_offset_hz=z.first_center_hz + k * z.spacing_hz, seed=k,
                       rds_bits=orc.rds_encode_groups(orc.rds_groups_ps_rt(pi, ps)))
        u = s.generate(n_wide)
        acc_i += (u[0::2].astype(np.float32) - 127.5) / 127.5
        acc_q += (u[1::2].astype(np.float32) - 127.5) / 127.5
    iq = np.empty(2 * n_wide, np.uint8)
    iq[0::2] = np.clip(np.round(127.5 + 127.5 * acc_i), 0, 255)
    iq[1::2] = np.clip(np.round(127.5 + 127.5 * acc_q), 0, 255)
    del acc_i, acc_q

    C_ = 100
    iq_dev = torch.from_numpy(iq).to(dev)
    stride = nblk * N
    x_dev = torch.zeros((C_, stride, 2), dtype=torch.float32, device=dev)
    eng = fm.Engine(fm.make_config(iq_rate=240_000, decimation=1, max_blocks=chunk), C_, 0)
    acap, gcap = eng.audio_capacity(chunk), chunk + 8
    audio = [[] for _ in range(C_)]
    groups = [[] for _ in range(C_)]
    stereo_last = None
    st_ = torch.cuda.current_stream().cuda_stream
    for b0 in range(0, nblk, chunk):
        # the capture arrives chunk by chunk; channelizer and engine both carry their state
        z.process(iq_dev.data_ptr() + 2 * b0 * N * D, chunk * N * D, x_dev.data_ptr() + 8 * b0 * N,
                  stride, stream=st_)
        a = torch.zeros((C_, 2, acap), dtype=torch.float32, device=dev)
        na = torch.zeros(C_, dtype=torch.int32, device=dev)
        g = torch.zeros((C_, gcap, 16), dtype=torch.uint8, device=dev)
        ng = torch.zeros(C_, dtype=torch.int32, device=dev)
        stt = torch.zeros((C_, chunk, 20), dtype=torch.uint8, device=dev)
        eng.process_batch_cf32(x_dev.data_ptr() + 8 * b0 * N, stride, chunk, a.data_ptr(), acap,
                               na.data_ptr(), g.data_ptr(), gcap, ng.data_ptr(), stt.data_ptr(), st_)
        torch.cuda.synchronize()
        a, na, ng = a.cpu().numpy(), na.cpu().numpy(), ng.cpu().numpy()
        gv = g.cpu().numpy().view(fm.GROUP_DTYPE).reshape(C_, gcap)
        stereo_last = stt.cpu().numpy().view(fm.STATUS_DTYPE).reshape(C_, chunk)["stereo"][:, -1]
        for c in range(C_):
            audio[c].append(a[c, :, :na[c]])
            gg = gv[c, :ng[c]].copy()
            gg["block_index"] += b0
            groups[c].append(gg)
    x_host = x_dev.cpu().numpy().view(np.complex64).reshape(C_, stride)
    eng.close()
    z.close()

    for k, (pi, ps, fl, fr) in stations.items():
        a = np.concatenate(audio[k], axis=1)
        g = np.concatenate(groups[k])
        # decoded: stereo lock and the station's PI in every group
        assert stereo_last[k] == 1, k
        pis = [int(x["a"]) for x in g if ((int(x["errors"]) >> 6) & 3) == 0]   # block A received clean
        assert len(pis) >= 2 and all(x == pi for x in pis), (k, g)
        # and exactly what the reference classes give for the same channelizer output
        rl, rr, rg, rst = _oracle_chain(orc_fm, x_host[k], nblk)
        assert np.array_equal(a[0], rl) and np.array_equal(a[1], rr), k
        assert groups_equal(g, rg, keys=("a", "b", "c", "d", "errors", "block_index")), k
        assert rst[-1] == 1
        # the left tone is where it was put
        spec = np.abs(np.fft.rfft(a[0][-8192:] * np.hanning(8192)))
        assert abs(np.argmax(spec[5:]) + 5 - fl * 8192 / 32000.0) <= 1.5, k
    # an empty slot stays mono and silent of RDS
    for k in (20, 60):
        assert stereo_last[k] == 0 and sum(len(x) for x in groups[k]) == 0, k


def test_polyphase_and_direct_forms_agree_and_polyphase_is_faster():
    """The default evaluation is the polyphase bank (partial sums per commutator phase + one DFT row
    per channel); FMGPU_CHANNELIZER_DIRECT=1 selects the direct form (one complex tap table per
    channel). Same definition, so the same numbers to FP32 rounding — and the polyphase form must be
    the faster one on the full band (config 4: 100 channels, two logical blocks per call)."""
    import os

    import torch
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(12)
    n_in = 2 * 8192 * 100
    iq = torch.from_numpy(rng.integers(0, 256, 2 * n_in, dtype=np.uint8)).to(dev)
    n_out = n_in // 100
    res, ms = {}, {}
    for form in ("polyphase", "direct"):
        if form == "direct":
            os.environ["FMGPU_CHANNELIZER_DIRECT"] = "1"
        try:
            z = fm.Channelizer()
        finally:
            os.environ.pop("FMGPU_CHANNELIZER_DIRECT", None)
        out = torch.zeros((100, n_out, 2), dtype=torch.float32, device=dev)
        z.process(iq.data_ptr(), n_in, out.data_ptr(), n_out)      # first call: zero history
        z.process(iq.data_ptr(), n_in, out.data_ptr(), n_out)      # second call: carried history
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            z.process(iq.data_ptr(), n_in, out.data_ptr(), n_out)
        e1.record()
        torch.cuda.synchronize()
        ms[form] = e0.elapsed_time(e1) / 5
        res[form] = out.cpu().numpy()
        z.close()
    err = np.abs(res["polyphase"] - res["direct"]).max()
    assert err < 2e-5, err
    print(f"channelizer, 100 channels x {n_out} outputs: polyphase {ms['polyphase']:.3f} ms, "
          f"direct {ms['direct']:.3f} ms")
    assert ms["polyphase"] < 0.5 * ms["direct"], ms
