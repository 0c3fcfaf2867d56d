"""One process, several devices ("one host thread + >= 2 streams per GPU"): the C ABI
takes a device index; two engines on two devices driven from two host threads of the same process
must both give the oracle's results, and a channelizer on device 1 must work next to them."""
import threading

import numpy as np
import pytest

import fmtuner_sdr_b200 as fm
from oracle import orc
from tests.common import groups_equal, rates, run_engine_chunks

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


@pytest.mark.skipif(_ndev() < 2, reason="needs two CUDA devices in one box")
def test_two_engines_on_two_devices_in_one_process(orc_fm):
    iq_rate, decim = rates("240k")
    C, nblk = 40, 6
    iq = np.stack([orc.config3_signal(300 + c, fs_iq=iq_rate).generate(nblk * 8192 * decim)
                   for c in range(2 * C)])
    out = {}

    def work(dev):
        eng = fm.Engine(fm.make_config(iq_rate=iq_rate, decimation=decim, max_blocks=2), C, dev)
        out[dev] = run_engine_chunks(eng, iq[dev * C:(dev + 1) * C], nblk, 2)
        eng.close()

    ths = [threading.Thread(target=work, args=(d,)) for d in (0, 1)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for dev in (0, 1):
        audio, groups, status, _ = out[dev]
        for c in (0, 17, C - 1):
            ref = orc.Channel(orc_fm, orc.make_config(iq_rate=iq_rate, decimation=decim)).process(iq[dev * C + c])
            assert np.array_equal(audio[c][0], ref.left) and np.array_equal(audio[c][1], ref.right)
            assert np.array_equal(status[c], ref.status) and groups_equal(groups[c], ref.groups)

    import torch
    z = fm.Channelizer(device=1)
    d1 = torch.device("cuda", 1)
    x = torch.randint(0, 256, (2 * 3200,), dtype=torch.uint8, device=d1)
    y = torch.zeros((100, 32, 2), dtype=torch.float32, device=d1)
    z.process(x.data_ptr(), 3200, y.data_ptr(), 32)
    torch.cuda.synchronize(d1)
    assert float(y.abs().max()) > 0.0
    z.close()
