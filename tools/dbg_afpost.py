import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fmtuner_sdr_b200 as fm
from oracle import orc
F = C.POINTER(C.c_float)
_f = lambda a: a.ctypes.data_as(F)
L = orc.OracleLib("fm").lib
for fs in (256000, 240000):
    rng = np.random.default_rng(3)
    a = L.orc_afpost_create(fs, 32000)
    eng = fm.Engine(fm.make_config(iq_rate=fs, decimation=1, block_samples=16384, max_blocks=1, deemphasis=1), 1, 0)
    for i, (n, cap, de) in enumerate([(8192, 8192, None), (1000, 8192, None), (8192, 100, None), (5, 8192, 50), (8192, 1024, None),
                                  (3000, 2, None), (3000, 8192, 0), (3000, 8192, None), (100, 8192, 75), (100, 8192, None)]):
        if de is not None:
            L.orc_afpost_set_deemphasis(a, de); eng.set_deemphasis_us(de)
        l = rng.normal(0, 0.2, n).astype(np.float32); r = rng.normal(0, 0.2, n).astype(np.float32)
        ol = np.zeros(cap, np.float32); orr = np.zeros(cap, np.float32)
        k = L.orc_afpost_process(a, _f(l), _f(r), n, _f(ol), _f(orr), cap)
        gl, gr = eng.afpost(l, r, cap)
        okl = gl.size == k and np.array_equal(gl, ol[:k]); okr = gr.size == k and np.array_equal(gr, orr[:k])
        print(fs, i, n, cap, de, "k", k, gl.size, okl, okr, (gl[:3], ol[:3]) if not okl else "")
