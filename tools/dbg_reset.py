"""Debug: which phase of test_retune_reset_and_bandwidth_change differs, and where."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fmtuner_sdr_b200 as fm
from oracle import orc
from tests.common import groups_equal, rates, run_engine_chunks

iq_rate, decim = rates("240k")
a = orc.config1_signal(fs_iq=iq_rate, seed=1).generate(10 * 8192 * decim)
b = orc.config3_signal(5, fs_iq=iq_rate).generate(10 * 8192 * decim)
lib = orc.OracleLib("fm")
och = orc.Channel(lib, orc.make_config())
eng = fm.Engine(fm.make_config(max_blocks=5), 1, 0)

def both(iq, tag):
    ref = och.process(iq, debug=True)
    audio, groups, status, dbg = run_engine_chunks(eng, iq.reshape(1, -1), 10, 5, debug_channel=0)
    for name, x, y in (("L", audio[0][0], ref.left), ("R", audio[0][1], ref.right),
                       ("mpx", dbg["mpx"], ref.mpx), ("sl", dbg["sl"], ref.sl), ("dec", dbg["dec"], ref.dec.view(np.float32) if ref.dec is not None else None)):
        if y is None or x is None:
            continue
        n = min(len(x), len(y))
        bad = np.nonzero(x[:n] != y[:n])[0]
        print(tag, name, "len", len(x), len(y), "first mismatch", (int(bad[0]) if bad.size else None), "count", bad.size)
    print(tag, "status equal", np.array_equal(status[0], ref.status), "groups", groups_equal(groups[0], ref.groups))

both(a, "1:a")
och.reset(dsp=True, rds=True); eng.reset(fm.engine.RESET_ALL)
both(b, "2:b after RESET_ALL")
och.set_bandwidth_hz(56000); eng.set_bandwidth_hz(56000)
both(a, "3:a after bandwidth")
och.reset(dsp=True, rds=False); eng.reset(fm.engine.RESET_DSP)
both(b, "4:b after RESET_DSP")
