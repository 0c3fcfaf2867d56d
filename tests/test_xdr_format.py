"""The XDR / FM-DX text lines of an RDS stream (fmgpu_xdr_rds_lines, host-side) against the
REFERENCE's own XDRServer::updateRDS compiled in place (oracle/_ref/libxdr_ref.so): PI debounce
states, '?' error marks, the block-B gate — over random group streams with a retune in the middle."""
import ctypes as C
import os

import numpy as np
import pytest

import fmtuner_sdr_b200 as fm
from oracle import orc

REF = os.path.join(os.path.dirname(orc.__file__), "_ref", "libxdr_ref.so")


class XdrState(C.Structure):
    _fields_ = [("pi_buffer", C.c_uint16 * 64), ("pi_error", C.c_uint8 * 8), ("pi_fill", C.c_uint8),
                ("pi_pos", C.c_uint8), ("pi_last_state", C.c_uint8), ("pad", C.c_uint8),
                ("pi_last_value", C.c_uint16)]


def _lib():
    L = fm.load_library()
    L.fmgpu_xdr_rds_init.argtypes = [C.POINTER(XdrState)]
    L.fmgpu_xdr_rds_init.restype = None
    L.fmgpu_xdr_rds_lines.argtypes = [C.POINTER(XdrState), C.c_void_p, C.c_char_p]
    return L


def _streams(seed):
    rng = np.random.default_rng(seed)
    pis = [0x1234, 0x1234, 0x1234, 0x4321, 0xD3C2, 0x1235]
    n = 400
    g = np.zeros(n, fm.GROUP_DTYPE)
    g["a"] = rng.choice(pis, n, p=[0.5, 0.1, 0.1, 0.1, 0.1, 0.1])
    g["b"], g["c"], g["d"] = (rng.integers(0, 65536, n) for _ in range(3))
    # error codes per block: 0 clean, 1 corrected, 3 missing (rds_decoder.cpp:29-41)
    codes = rng.choice([0, 1, 3], (n, 4), p=[0.6, 0.25, 0.15])
    g["errors"] = (codes[:, 0] << 6) | (codes[:, 1] << 4) | (codes[:, 2] << 2) | codes[:, 3]
    g["a"][codes[:, 0] == 3] = 0          # a missing block A is reported as 0
    return g


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libxdr_ref.so not built")
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_lines_equal_reference_xdr_server(seed):
    L = _lib()
    R = C.CDLL(REF)
    R.ref_xdr_create.restype = C.c_void_p
    R.ref_xdr_retune.argtypes = [C.c_void_p]
    R.ref_xdr_destroy.argtypes = [C.c_void_p]
    R.ref_xdr_update.argtypes = [C.c_void_p, C.c_uint16, C.c_uint16, C.c_uint16, C.c_uint16,
                                 C.c_uint8, C.c_char_p, C.c_int]
    h = R.ref_xdr_create()
    st = XdrState()
    L.fmgpu_xdr_rds_init(C.byref(st))
    groups = _streams(seed)
    n_p = n_r = 0
    for i, g in enumerate(groups):
        if i == 250:                       # a retune in mid-stream clears the PI history
            R.ref_xdr_retune(h)
            L.fmgpu_xdr_rds_init(C.byref(st))
        ref_buf = C.create_string_buffer(64)
        k = R.ref_xdr_update(h, int(g["a"]), int(g["b"]), int(g["c"]), int(g["d"]), int(g["errors"]),
                             ref_buf, 2)
        ref = [ref_buf.raw[32 * i:32 * i + 32].split(b"\0")[0] for i in range(k)]
        got_buf = C.create_string_buffer(64)
        one = np.array([g])
        m = L.fmgpu_xdr_rds_lines(C.byref(st), one.ctypes.data, got_buf)
        got = [got_buf.raw[32 * i:32 * i + 32].split(b"\0")[0] for i in range(m)]
        assert got == ref, (g, got, ref)
        n_p += sum(x.startswith(b"P") for x in got)
        n_r += sum(x.startswith(b"R") for x in got)
    R.ref_xdr_destroy(h)
    assert n_p > 50 and n_r > 150      # both kinds of line were exercised


def test_known_lines():
    L = _lib()
    st = XdrState()
    L.fmgpu_xdr_rds_init(C.byref(st))
    g = np.zeros(1, fm.GROUP_DTYPE)
    g["a"], g["b"], g["c"], g["d"], g["errors"] = 0x1234, 0x0408, 0xE0CD, 0x4232, 0
    buf = C.create_string_buffer(64)
    # first clean group: PI seen once -> not yet debounced, only the R line
    assert L.fmgpu_xdr_rds_lines(C.byref(st), g.ctypes.data, buf) == 1
    assert buf.raw.split(b"\0")[0] == b"R0408E0CD423200"
    # second clean group with the same PI: debounced
    assert L.fmgpu_xdr_rds_lines(C.byref(st), g.ctypes.data, buf) == 2
    assert buf.raw[:32].split(b"\0")[0] == b"P1234"
    g["errors"] = 1 << 6                   # block A corrected: one question mark
    L.fmgpu_xdr_rds_lines(C.byref(st), g.ctypes.data, buf)
    assert buf.raw[:32].split(b"\0")[0] == b"P1234?"
