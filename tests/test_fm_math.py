"""fm_math.h (the transcendental kernels shared by engine and oracle-fm) against float64 libm.

The bound stated in fmtuner_sdr_b200/csrc/fm_math.h is checked here, on the domains the
engine uses: NCO phases in [-2pi, 2pi], discriminator atan2 over the whole plane, AGC
exp/log arguments.
"""
import ctypes as C

import numpy as np


def _ulp_err(got, exact):
    exact32 = exact.astype(np.float32)
    ulp = np.spacing(np.abs(exact32)).astype(np.float64)
    ulp = np.maximum(ulp, np.finfo(np.float32).tiny)
    return np.abs(got.astype(np.float64) - exact) / ulp


def _call(lib, name, *arrs):
    n = arrs[0].size
    outs = []
    args = []
    for a in arrs:
        args.append(a.ctypes.data_as(C.POINTER(C.c_float)))
    if name == "sincos":
        s = np.zeros(n, np.float32)
        c = np.zeros(n, np.float32)
        lib.lib.orc_math_sincos(args[0], s.ctypes.data_as(C.POINTER(C.c_float)),
                                c.ctypes.data_as(C.POINTER(C.c_float)), n)
        return s, c
    r = np.zeros(n, np.float32)
    getattr(lib.lib, f"orc_math_{name}")(*args, r.ctypes.data_as(C.POINTER(C.c_float)), n)
    return r


def test_math_flavours(orc_fm, orc_libm):
    assert orc_fm.lib.orc_math_name() == b"fm_math"
    assert orc_libm.lib.orc_math_name() == b"libm"


def test_sincos_accuracy(orc_fm):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-2 * np.pi, 2 * np.pi, 200_000),
                        np.linspace(-7, 7, 100_001), [0.0, np.pi / 2, np.pi, -np.pi]]).astype(np.float32)
    s, c = _call(orc_fm, "sincos", x)
    xs = x.astype(np.float64)
    # absolute error bound near zeros of sin/cos, ulp bound elsewhere
    es = np.abs(s - np.sin(xs))
    ec = np.abs(c - np.cos(xs))
    assert es.max() < 1.5e-7 and ec.max() < 1.5e-7
    big = np.abs(np.sin(xs)) > 0.1
    assert _ulp_err(s[big], np.sin(xs[big])).max() <= 2.0
    big = np.abs(np.cos(xs)) > 0.1
    assert _ulp_err(c[big], np.cos(xs[big])).max() <= 2.0


def test_atan2_accuracy_and_signed_zero(orc_fm):
    rng = np.random.default_rng(1)
    y = rng.normal(size=300_000).astype(np.float32)
    x = rng.normal(size=300_000).astype(np.float32)
    r = _call(orc_fm, "atan2", y, x)
    exact = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    assert np.abs(r - exact).max() < 4e-7
    assert _ulp_err(r, exact)[np.abs(exact) > 0.05].max() <= 4.0
    # IEEE special cases the discriminator can hit on the first sample after a reset
    yy = np.array([0.0, -0.0, 0.0, -0.0, 1.0, -1.0, 0.0, -0.0], np.float32)
    xx = np.array([1.0, 1.0, -1.0, -1.0, 0.0, 0.0, 0.0, -0.0], np.float32)
    got = _call(orc_fm, "atan2", yy, xx)
    want = np.arctan2(yy, xx)
    assert np.allclose(got, want, atol=1e-7)
    assert (np.signbit(got) == np.signbit(want)).all()


def test_exp_log_accuracy(orc_fm):
    rng = np.random.default_rng(2)
    x = rng.uniform(-20, 20, 200_000).astype(np.float32)
    assert _ulp_err(_call(orc_fm, "exp", x), np.exp(x.astype(np.float64))).max() <= 2.0
    v = np.exp(rng.uniform(-13, 13, 200_000)).astype(np.float32)
    lg = _call(orc_fm, "log", v)
    exact = np.log(v.astype(np.float64))
    assert np.abs(lg - exact).max() < 2e-6
    assert _ulp_err(lg, exact)[np.abs(exact) > 0.1].max() <= 2.0


def test_nco_constrain_quantisation(orc_fm, orc_libm):
    # Appendix A.8: 2*pi <-> 2^32; small negative steps quantise to 256-count multiples
    f = orc_fm.lib.orc_nco_constrain
    assert f(0.0) == 0
    assert abs(int(f(np.float32(np.pi))) - 2 ** 31) <= 256
    assert f(np.float32(-1e-6)) % 256 == 0
    for x in (0.1, -0.1, 1e-4, 3.0, -3.0, 6.0):
        assert f(np.float32(x)) == orc_libm.lib.orc_nco_constrain(np.float32(x))
        want = (x / (2 * np.pi)) % 1.0 * 2 ** 32
        assert abs(int(f(np.float32(x))) - want) < 1024
