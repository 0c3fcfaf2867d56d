"""Host logic of the tensor-core FIRs (fmtuner_sdr_b200/csrc/fir_tc.cu) without a GPU:
fmgpu_fir_tc_host_model evaluates the kernel's arithmetic from the SAME tables the kernel loads —
integer taps split into three signed base-256 digits and laid out as the MMA's B operand, samples as
24-bit offset-binary fixed point in three byte planes, the 2^23 offset removed limb-wise, the five limb
sets recombined in float — and is checked here against a float64 FIR of the same samples: the table
builder, the operand-image layout (every digit is read back through it), the offset limbs and the
error bound stated in DESIGN.md 4h. (The GPU tests check the kernel against the same float64 FIR.)"""
import ctypes as C

import numpy as np
import pytest

import fmtuner_sdr_b200 as fm


def _model(taps, scale, shift, x_hist, x):
    L = fm.load_library()
    L.fmgpu_fir_tc_host_model.restype = C.c_size_t
    L.fmgpu_fir_tc_host_model.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_size_t,
                                          C.c_size_t, C.c_void_p]
    t = np.ascontiguousarray(taps, np.float32)
    xx = np.ascontiguousarray(np.concatenate([x_hist, x]), np.float32)
    y = np.zeros(x.size, np.float32)
    n = L.fmgpu_fir_tc_host_model(t.ctypes.data, t.size, scale, shift, xx.ctypes.data, x_hist.size, x.size,
                                  y.ctypes.data)
    return n, y


def _design(which, iq_rate=2_400_000, decim=10):
    L = fm.load_library()
    cfg = fm.make_config(iq_rate=iq_rate, decimation=decim)
    buf = np.zeros(4096, np.float32)
    sc = C.c_float(0)
    L.fmgpu_design_host.restype = C.c_size_t
    n = L.fmgpu_design_host(C.byref(cfg), which, 0, buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(sc))
    return buf[:n].copy(), sc.value


@pytest.mark.parametrize("which,shift,amp,rates", [(2, 22, 1.5, (2_400_000, 10)), (2, 22, 1.5, (2_048_000, 8)),
                                                   (3, 20, 6.0, (2_400_000, 10))])
def test_integer_fir_model_against_float64(which, shift, amp, rates):
    taps, scale = _design(which, *rates)
    rng = np.random.default_rng(which * 7 + shift)
    hist = (amp * rng.uniform(-1, 1, 512)).astype(np.float32)
    x = (amp * rng.uniform(-1, 1, 4096)).astype(np.float32)
    x[100:108] = amp                       # a rail
    x[2000:2100] = 0.0
    n, y = _model(taps, scale, shift, hist, x)
    assert n == x.size
    full = np.concatenate([hist, x]).astype(np.float64)
    want = scale * np.convolve(full, taps.astype(np.float64))[hist.size:hist.size + x.size]
    # half a sample quantum times sum |h| (every sample rounded once), the taps' own quantisation
    # (2^-S relative to the largest tap; S >= 23 bits), three float roundings of the result
    bound = scale * (2.0 ** -(shift + 1)) * np.abs(taps).sum() + 4e-7 * max(1.0, np.abs(want).max())
    err = np.abs(y.astype(np.float64) - want).max()
    assert err <= bound, (err, bound)
    # and it is at least as good as a float32 chain of the same length
    chain = np.zeros(x.size, np.float32)
    h32 = taps[::-1].astype(np.float32)
    f32 = np.concatenate([hist, x]).astype(np.float32)
    for i in range(0, x.size, 257):        # a sample of the outputs: the chain is slow in numpy
        acc = np.float32(0.0)
        w = f32[hist.size + i - taps.size + 1:hist.size + i + 1]
        for a, b in zip(h32, w):
            acc = np.float32(acc + np.float32(a * b))
        chain[i] = np.float32(acc * np.float32(scale))
    idx = np.arange(0, x.size, 257)
    assert np.abs(y[idx].astype(np.float64) - want[idx]).max() <= np.abs(chain[idx].astype(np.float64) - want[idx]).max() + bound


def test_saturation_and_refusals():
    taps, scale = _design(2)
    hist = np.zeros(512, np.float32)
    x = np.full(64, 5.0, np.float32)       # beyond the 2^-22 format's range: saturates at 2 - 2^-22
    n, y = _model(taps, scale, 22, hist, x)
    sat = np.full(64, 2.0 - 2.0 ** -22, np.float64)
    want = np.convolve(np.concatenate([hist.astype(np.float64), sat]), taps.astype(np.float64))[512:576]
    assert n == 64 and np.abs(y - want).max() <= 1e-6
    assert _model(taps, scale, 22, hist, np.zeros(33, np.float32))[0] == 0     # not a multiple of 32
    assert _model(taps, scale, 22, np.zeros(64, np.float32), np.zeros(64, np.float32))[0] == 0   # history too short


@pytest.mark.parametrize("rates,valid_full", [((2_400_000, 10), True), ((2_400_000, 10), False),
                                              ((2_048_000, 8), True), ((1_024_000, 4), True)])
def test_integer_decimator_model_against_float64(rates, valid_full):
    """decim_tc.cu's tables: taps quantised to 2^-26 in four base-128 limbs, I/Q de-interleave inside B,
    the 127.5 offset removed in integers (the table of partial tap sums when the window starts inside
    the zeros after a reset), one float rounding: within 1.5e-7 of the float64 FIR."""
    iq_rate, decim = rates
    L = fm.load_library()
    taps, scale = _design(0, iq_rate, decim)
    n_out = 256
    valid = taps.size - 1 if valid_full else 37
    rng = np.random.default_rng(3)
    iq = rng.integers(0, 256, 2 * (valid + n_out * decim), dtype=np.uint8)
    iq[200:260] = 255
    out = np.zeros(2 * n_out, np.float32)
    L.fmgpu_decim_tc_host_model.restype = C.c_size_t
    L.fmgpu_decim_tc_host_model.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_int,
                                            C.c_void_p]
    n = L.fmgpu_decim_tc_host_model(decim, taps.ctypes.data, taps.size, scale, iq.ctypes.data, valid, n_out,
                                    out.ctypes.data)
    assert n == n_out
    x = (iq.astype(np.float64).reshape(-1, 2) - 127.5) / 127.5
    x = x[:, 0] + 1j * x[:, 1]
    # what is missing of the first windows is ZERO SAMPLES (a reset), not byte 0
    xp = np.concatenate([np.zeros(taps.size - 1 - valid, np.complex128), x])
    idx = (np.arange(n_out) * decim)[:, None] + np.arange(taps.size)[None, :]
    want = scale * (xp[idx] @ taps[::-1].astype(np.float64))
    got = out.view(np.complex64).astype(np.complex128)
    assert np.abs(got - want).max() <= 1.5e-7, np.abs(got - want).max()
