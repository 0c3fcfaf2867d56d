"""Quick GPU-vs-oracle stage-by-stage comparison (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import orc
import fmtuner_sdr_b200 as fm

def main():
    nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    iq_rate, decim = 2_400_000, 10
    sig = orc.config1_signal(fs_iq=iq_rate)
    iq = sig.generate(nblk * 8192 * decim)
    L = orc.OracleLib("fm")
    ch = orc.Channel(L, orc.make_config(iq_rate=iq_rate, decimation=decim))
    t = time.time(); ref = ch.process(iq, debug=True); print("oracle s", time.time() - t)
    eng = fm.Engine(fm.make_config(iq_rate=iq_rate, decimation=decim, max_blocks=4), 1, 0)
    per = 8192 * decim * 2
    outs = []; dec = []; mpx = []; sl = []; sr = []; stat = []; groups = []
    for b0 in range(0, nblk, 4):
        nb = min(4, nblk - b0)
        a, na, g, ng, st = eng.process_host(iq[b0 * per:(b0 + nb) * per].reshape(1, -1), nb)
        outs.append(a[0, :, :na[0]]); stat.append(st[0])
        gg = g[0, :ng[0]].copy(); gg["block_index"] += b0; groups.append(gg)
        dec.append(eng.debug_read(0)); mpx.append(eng.debug_read(1)); sl.append(eng.debug_read(2)); sr.append(eng.debug_read(3))
    audio = np.concatenate(outs, axis=1); dec = np.concatenate(dec); mpx = np.concatenate(mpx)
    sl = np.concatenate(sl); sr = np.concatenate(sr); stat = np.concatenate(stat); groups = np.concatenate(groups)
    def cmp(name, a, b):
        n = min(a.size, b.size)
        d = np.abs(a[:n].astype(np.complex128) - b[:n].astype(np.complex128))
        bad = np.flatnonzero(d > 0)
        print(f"{name:8s} n={a.size}/{b.size} maxabs={d.max() if n else 0:.3e} first_diff={bad[0] if bad.size else -1} ndiff={bad.size}")
    cmp("dec", dec, ref.dec); cmp("mpx", mpx, ref.mpx); cmp("sl", sl, ref.sl); cmp("sr", sr, ref.sr)
    cmp("audioL", audio[0], ref.left); cmp("audioR", audio[1], ref.right)
    print("status eq:", {k: bool((stat[k] == ref.status[k]).all()) for k in stat.dtype.names})
    print("stereo gpu", stat["stereo"], "ref", ref.status["stereo"])
    print("groups gpu", len(groups), "ref", len(ref.groups), "equal", len(groups) == len(ref.groups) and all((groups[k] == ref.groups[k]).all() for k in ("a", "b", "c", "d", "errors", "block_index")))
    print("launches", eng.launch_count())

main()
