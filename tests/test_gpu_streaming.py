"""Streaming forms of the batched path: fmgpu_process_batch_async + fmgpu_join and
fmgpu_submit_host / fmgpu_wait_host (two submissions in flight) must give exactly what the
synchronous calls give — and what the CPU oracle gives — block after block."""
import numpy as np
import pytest

import fmtuner_sdr_b200 as fm
from oracle import orc
from tests.common import groups_equal, rates, run_engine_chunks

pytestmark = pytest.mark.gpu


def _signals(n, nblk, seed=11):
    iq_rate, decim = rates("240k")
    rng = np.random.default_rng(seed)
    rows = []
    for c in range(n):
        s = orc.config3_signal(900 + c, fs_iq=iq_rate)
        s.snr_db = float(rng.uniform(12.0, 45.0))
        rows.append(s.generate(nblk * 8192 * decim))
    return np.stack(rows)


def test_submit_wait_two_in_flight_equals_sync_and_oracle(orc_fm):
    import torch
    C, nblk, chunk = 70, 6, 1
    iq = _signals(C, nblk)
    ref_eng = fm.Engine(fm.make_config(max_blocks=chunk, dsp_agc=1), C, 0)
    ref_eng.set_pipeline_groups(3)
    audio_s, groups_s, status_s, _ = run_engine_chunks(ref_eng, iq, nblk, chunk)
    ref_eng.close()

    eng = fm.Engine(fm.make_config(max_blocks=chunk, dsp_agc=1), C, 0)
    eng.set_pipeline_groups(3)
    per = eng.iq_bytes_per_block
    acap, gcap = eng.audio_capacity(chunk), chunk + 8
    iq_pin = torch.from_numpy(iq).pin_memory()
    stride = iq_pin.stride(0)
    outs = [(torch.zeros((C, 2, acap), dtype=torch.float32).pin_memory(),
             torch.zeros(C, dtype=torch.int32).pin_memory(),
             torch.zeros((C, gcap, 16), dtype=torch.uint8).pin_memory(),
             torch.zeros(C, dtype=torch.int32).pin_memory(),
             torch.zeros((C, chunk, 20), dtype=torch.uint8).pin_memory()) for _ in range(2)]
    audio = [[] for _ in range(C)]
    groups = [[] for _ in range(C)]
    status = []

    def submit(k):
        a, na, g, ng, st = outs[k & 1]
        return eng.submit_host_raw(iq_pin.data_ptr() + k * chunk * per, stride, chunk, a.data_ptr(),
                                   acap, na.data_ptr(), g.data_ptr(), gcap, ng.data_ptr(),
                                   st.data_ptr())

    def collect(k):
        a, na, g, ng, st = (x.numpy() for x in outs[k & 1])
        gv = g.view(fm.GROUP_DTYPE).reshape(C, gcap)
        status.append(st.view(fm.STATUS_DTYPE).reshape(C, chunk).copy())
        for c in range(C):
            audio[c].append(a[c, :, :na[c]].copy())
            gg = gv[c, :ng[c]].copy()
            gg["block_index"] += k * chunk
            groups[c].append(gg)

    steps = nblk // chunk
    pending = submit(0)
    for k in range(1, steps):
        nxt = submit(k)          # k is queued behind k-1 before k-1 is waited for
        eng.wait_host(pending)
        collect(k - 1)
        pending = nxt
    eng.wait_host(pending)
    collect(steps - 1)
    eng.close()

    status = np.concatenate(status, axis=1)
    for c in range(C):
        a = np.concatenate(audio[c], axis=1)
        assert np.array_equal(a, audio_s[c]), c
        assert groups_equal(np.concatenate(groups[c]), groups_s[c]), c
        assert np.array_equal(status[c], status_s[c]), c
    for c in (0, 33, 69):
        ref = orc.Channel(orc_fm, orc.make_config(dsp_agc=1)).process(iq[c])
        a = np.concatenate(audio[c], axis=1)
        assert np.array_equal(a[0], ref.left) and np.array_equal(a[1], ref.right), c
        assert groups_equal(np.concatenate(groups[c]), ref.groups), c


def test_async_batch_join_equals_sync():
    import torch
    C, nblk = 96, 4
    iq = _signals(12, nblk, seed=3)
    iq = iq[np.arange(C) % 12]
    dev = torch.device("cuda", 0)
    iq_dev = torch.from_numpy(iq).to(dev)
    stride = iq_dev.stride(0)

    def run(async_steps, groups_n):
        eng = fm.Engine(fm.make_config(max_blocks=1), C, 0)
        eng.set_pipeline_groups(groups_n)
        per = eng.iq_bytes_per_block
        acap, gcap = eng.audio_capacity(1), 9
        outs = []
        st = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(st):
            for k in range(nblk):
                a = torch.zeros((C, 2, acap), dtype=torch.float32, device=dev)
                na = torch.zeros(C, dtype=torch.int32, device=dev)
                g = torch.zeros((C, gcap, 16), dtype=torch.uint8, device=dev)
                ng = torch.zeros(C, dtype=torch.int32, device=dev)
                stt = torch.zeros((C, 1, 20), dtype=torch.uint8, device=dev)
                outs.append((a, na, g, ng, stt))
            st.synchronize()
            for k in range(nblk):
                a, na, g, ng, stt = outs[k]
                f = eng.process_batch_async if async_steps else eng.process_batch
                f(iq_dev.data_ptr() + k * per, stride, 1, a.data_ptr(), acap, na.data_ptr(),
                  g.data_ptr(), gcap, ng.data_ptr(), stt.data_ptr(), st.cuda_stream)
            if async_steps:
                eng.join(st.cuda_stream)
            st.synchronize()
        res = [tuple(x.cpu().numpy() for x in o) for o in outs]
        eng.close()
        return res

    sync = run(False, 1)
    for groups_n in (1, 3, 16):
        streamed = run(True, groups_n)
        for k in range(nblk):
            na = sync[k][1]
            assert np.array_equal(na, streamed[k][1])
            for c in range(C):
                assert np.array_equal(sync[k][0][c, :, :na[c]], streamed[k][0][c, :, :na[c]]), (k, c)
            assert np.array_equal(sync[k][3], streamed[k][3])
            ng = sync[k][3]
            for c in range(C):
                assert np.array_equal(sync[k][2][c, :ng[c]], streamed[k][2][c, :ng[c]]), (k, c)
            assert np.array_equal(sync[k][4], streamed[k][4])
