"""BASELINE full-size configurations through size-independent properties (the oracle is far too
slow to decode them entirely): replica equality, channel-permutation equivariance, determinism,
plus exact oracle spot checks on a few channels of each configuration."""
import numpy as np
import pytest

import fmtuner_sdr_b200 as fm
from oracle import orc
from tests.common import groups_equal, rates, run_engine_chunks

pytestmark = pytest.mark.gpu


def _distinct(n, nblk, snr_lo=20.0, snr_hi=60.0, seed=0):
    iq_rate, decim = rates("240k")
    rng = np.random.default_rng(seed)
    rows = []
    for c in range(n):
        s = orc.config3_signal(500 + c, fs_iq=iq_rate)
        s.snr_db = float(rng.uniform(snr_lo, snr_hi))
        rows.append(s.generate(nblk * 8192 * decim))
    return np.stack(rows)


def test_config3_256_channels(orc_fm):
    """BASELINE config 3: 256 channels batched on one GPU."""
    C, nblk, nd = 256, 4, 12
    base = _distinct(nd, nblk)
    rng = np.random.default_rng(1)
    which = rng.integers(0, nd, C)
    which[:nd] = np.arange(nd)
    iq = base[which]
    eng = fm.Engine(fm.make_config(max_blocks=2), C, 0)
    eng.set_pipeline_groups(4)
    audio, groups, status, _ = run_engine_chunks(eng, iq, nblk, 2)
    eng.close()
    # replicas of the same multiplex decode identically wherever they sit in the batch
    for c in range(C):
        r = int(which[c])
        assert np.array_equal(audio[c], audio[r]), c
        assert groups_equal(groups[c], groups[r]), c
        assert np.array_equal(status[c], status[r]), c
    # and the distinct ones equal the oracle exactly
    for r in range(nd):
        ref = orc.Channel(orc_fm, orc.make_config()).process(base[r])
        assert np.array_equal(audio[r][0], ref.left) and np.array_equal(audio[r][1], ref.right), r
        assert np.array_equal(status[r], ref.status) and groups_equal(groups[r], ref.groups), r
    # permuting the channels permutes the results (no cross-channel term)
    perm = rng.permutation(C)
    eng2 = fm.Engine(fm.make_config(max_blocks=4), C, 0)
    audio2, groups2, status2, _ = run_engine_chunks(eng2, iq[perm], nblk, 4)
    eng2.close()
    for i in range(0, C, 7):
        assert np.array_equal(audio2[i], audio[perm[i]])
        assert groups_equal(groups2[i], groups[perm[i]])


def test_config5_10000_channels(orc_fm):
    """BASELINE config 5 at full channel count (one logical block per call, two calls):
    SNR 10-40 dB, blend mode c % 3, dsp_agc fast."""
    C, nblk, nd = 10_000, 2, 9
    base = _distinct(nd, nblk, 10.0, 40.0, seed=5)
    # replica r uses signal r % nd and blend mode (r % nd) % 3 so replicas share all settings
    which = np.arange(C) % nd
    iq = base[which]
    eng = fm.Engine(fm.make_config(max_blocks=1, dsp_agc=1), C, 0)
    eng.set_pipeline_groups(8)
    for c in range(C):
        if (which[c] % 3) != 1:
            eng.set_blend_mode(int(which[c] % 3), c)
    audio, groups, status, _ = run_engine_chunks(eng, iq, nblk, 1)
    assert eng.launch_count() > 0
    eng.close()
    for c in range(nd, C):
        r = int(which[c])
        assert np.array_equal(audio[c], audio[r]), c
        assert np.array_equal(status[c], status[r]), c
        assert groups_equal(groups[c], groups[r]), c
    for r in range(nd):
        ref = orc.Channel(orc_fm, orc.make_config(dsp_agc=1, stereo_blend=r % 3)).process(base[r])
        assert np.array_equal(audio[r][0], ref.left) and np.array_equal(audio[r][1], ref.right), r
        assert np.array_equal(status[r], ref.status), r


def test_device_synth_decodes_and_is_deterministic():
    """The on-device generator used by bench.py: deterministic in its seeds, decodes to stereo
    with the PI it was given."""
    import torch
    C, B, calls = 64, 14, 2
    n_iq = calls * B * 81920
    stride = 2 * n_iq
    outs = []
    for _ in range(2):
        iq = torch.empty((C, stride), dtype=torch.uint8, device="cuda:0")
        params = [fm.SynthParams(75000.0, 400.0 + 37.0 * c, 0.8, 700.0 + 53.0 * c, 0.8, 0.10, 0.04,
                                 0.5, 35.0, c, 0x2000 + c, 0) for c in range(C)]
        fm.synth_iq(0, params, 2_400_000, n_iq, iq.data_ptr(), stride)
        torch.cuda.synchronize()
        outs.append(iq.cpu().numpy())
    assert np.array_equal(outs[0], outs[1])
    eng = fm.Engine(fm.make_config(max_blocks=B), C, 0)
    per = B * 81920 * 2
    pis = [[] for _ in range(C)]
    for k in range(calls):
        a, na, g, ng, st = eng.process_host(outs[0][:, k * per:(k + 1) * per], B, group_cap=16)
        for c in range(C):
            gg = g[c, :ng[c]]
            pis[c].extend(int(x) for x in gg["a"][(gg["errors"] >> 6) == 0])
    eng.close()
    assert st["stereo"][:, -1].all()
    for c in range(C):
        assert len(pis[c]) >= 2 and all(p == 0x2000 + c for p in pis[c]), (c, pis[c])


def test_pcm16_packing_matches_reference_formula():
    """audio_output.cpp:1458-1459 (volume scale) + :1386-1391 (clamp, x32767, truncation)."""
    import torch
    C, B = 5, 2
    iq_rate, decim = rates("240k")
    iq = np.stack([orc.config3_signal(40 + c, fs_iq=iq_rate).generate(B * 8192 * decim) for c in range(C)])
    eng = fm.Engine(fm.make_config(max_blocks=B), C, 0)
    d_iq = torch.from_numpy(iq).cuda()
    acap = eng.audio_capacity(B)
    audio = torch.zeros((C, 2, acap), dtype=torch.float32, device="cuda")
    n_audio = torch.zeros(C, dtype=torch.int32, device="cuda")
    pcm = torch.zeros((C, acap, 2), dtype=torch.int16, device="cuda")
    eng.process_batch(d_iq.data_ptr(), iq.shape[1], B, audio.data_ptr(), acap, n_audio.data_ptr())
    # make the packer see values beyond +-1 too
    audio *= 3.0
    for vol in (0.85, 0.85 * 0.37):
        eng.pack_pcm16(audio.data_ptr(), acap, n_audio.data_ptr(), vol, pcm.data_ptr())
        torch.cuda.synchronize()
        a = audio.cpu().numpy()
        n = int(n_audio[0].item())
        want = np.trunc(np.clip(a[:, :, :n] * np.float32(vol), -1.0, 1.0).astype(np.float32)
                        * np.float32(32767.0)).astype(np.int16)
        got = pcm.cpu().numpy()[:, :n, :]
        assert np.array_equal(got[:, :, 0], want[:, 0]) and np.array_equal(got[:, :, 1], want[:, 1])
    eng.close()


def test_config5_10000_channels_streamed_equals_joined():
    """The bench's own schedule — 10,000 channels, ring of three slots, two blocks per call,
    calls STREAMED (fmgpu_process_batch_async, one join at the end; and fmgpu_submit_host with two
    submissions in flight) — must give, call for call, exactly what joined calls give: audio,
    frame counts, groups, GROUP COUNTS and per-block status. Stages of successive blocks overlap
    only in the streamed runs, so a producer overwriting a ring slot whose tail is still some
    reader's halo, or a count written from the wrong stream, shows up here."""
    import torch
    C, calls, B, nd = 10_000, 8, 2, 9      # 16 blocks = 0.55 s: the first RDS groups appear
    nblk = calls * B
    base = _distinct(nd, nblk, 10.0, 40.0, seed=8)
    dev = torch.device("cuda", 0)
    which = torch.arange(C, device=dev) % nd
    iq_dev = torch.from_numpy(base).to(dev)[which].contiguous()   # [C][nblk * per] in HBM
    stride = iq_dev.stride(0)

    def engine():
        return fm.Engine(fm.make_config(max_blocks=B, dsp_agc=1), C, 0)

    def outputs(eng):
        acap, gcap = eng.audio_capacity(B), B + 8
        return [(torch.zeros((C, 2, acap), dtype=torch.float32, device=dev),
                 torch.zeros(C, dtype=torch.int32, device=dev),
                 torch.zeros((C, gcap, 16), dtype=torch.uint8, device=dev),
                 torch.zeros(C, dtype=torch.int32, device=dev),
                 torch.zeros((C, B, 20), dtype=torch.uint8, device=dev)) for _ in range(calls)], acap, gcap

    def run(streamed):
        eng = engine()
        per = eng.iq_bytes_per_block
        outs, acap, gcap = outputs(eng)
        st = torch.cuda.Stream(device=dev)
        torch.cuda.synchronize()
        with torch.cuda.stream(st):
            for k in range(calls):
                a, na, g, ng, stt = outs[k]
                f = eng.process_batch_async if streamed else eng.process_batch
                f(iq_dev.data_ptr() + k * B * per, stride, B, a.data_ptr(), acap, na.data_ptr(),
                  g.data_ptr(), gcap, ng.data_ptr(), stt.data_ptr(), st.cuda_stream)
                if not streamed:
                    st.synchronize()
            if streamed:
                eng.join(st.cuda_stream)
            st.synchronize()
        eng.close()
        return outs

    joined = run(False)
    streamed = run(True)
    total_groups = 0
    for k in range(calls):
        for i, name in enumerate(("audio", "n_audio", "groups", "n_groups", "status")):
            assert torch.equal(joined[k][i], streamed[k][i]), (k, name)
        total_groups += int(joined[k][3].sum().item())
    assert total_groups > C // 2  # the comparison saw real groups
    # replicas agree inside the streamed run too
    last = streamed[-1]
    assert torch.equal(last[0][nd:2 * nd], last[0][:nd]) and torch.equal(last[3][nd:2 * nd], last[3][:nd])

    # host path, two submissions in flight, 2,000 channels of the same batch
    Ch = 2000
    eng = fm.Engine(fm.make_config(max_blocks=B, dsp_agc=1), Ch, 0)
    per = eng.iq_bytes_per_block
    acap, gcap = eng.audio_capacity(B), B + 8
    iq_pin = iq_dev[:Ch].cpu().pin_memory()
    hs = iq_pin.stride(0)
    houts = [(torch.zeros((Ch, 2, acap), dtype=torch.float32).pin_memory(),
              torch.zeros(Ch, dtype=torch.int32).pin_memory(),
              torch.zeros((Ch, gcap, 16), dtype=torch.uint8).pin_memory(),
              torch.zeros(Ch, dtype=torch.int32).pin_memory(),
              torch.zeros((Ch, B, 20), dtype=torch.uint8).pin_memory()) for _ in range(2)]

    def submit(k):
        a, na, g, ng, stt = houts[k & 1]
        return eng.submit_host_raw(iq_pin.data_ptr() + k * B * per, hs, B, a.data_ptr(), acap,
                                   na.data_ptr(), g.data_ptr(), gcap, ng.data_ptr(), stt.data_ptr())

    def check(k):
        for i, name in enumerate(("audio", "n_audio", "groups", "n_groups", "status")):
            want = joined[k][i][:Ch].cpu()
            got = houts[k & 1][i]
            if name == "audio":
                n = int(joined[k][1][0].item())
                assert torch.equal(got[:, :, :n], want[:, :, :n]), (k, name)
            elif name == "groups":
                m = joined[k][3][:Ch].cpu()
                mask = (torch.arange(gcap)[None, :] < m[:, None])
                assert torch.equal(got[mask], want[mask]), (k, name)
            else:
                assert torch.equal(got, want), (k, name)

    pending = submit(0)
    for k in range(1, calls):
        nxt = submit(k)
        eng.wait_host(pending)
        check(k - 1)
        pending = nxt
    eng.wait_host(pending)
    check(calls - 1)
    eng.close()
