"""The drop-in C++ classes (same names / signatures as the reference headers) driven by a C++
harness that mirrors the reference's main loop; results must equal the CPU oracle exactly."""
import os
import subprocess

import numpy as np
import pytest

from fmtuner_sdr_b200 import build as fmbuild
from oracle import orc
from tests.common import rates

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    fmbuild.build_lib()
    fmbuild.build_dropin()
    out = str(tmp_path_factory.mktemp("dropin") / "dropin_harness")
    pkg = os.path.join(ROOT, "fmtuner_sdr_b200")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(pkg, "dropin"), "-I",
                    os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "dropin_harness.cpp"),
                    "-o", out, "-L", pkg, "-lfmgpu_dropin", "-lfmgpu", f"-Wl,-rpath,{pkg}"],
                   check=True)
    return out


def _check_extra(orc_fm, iq, iq_rate, decim, prefix):
    """The rest of the public class surface (see the tail of dropin_harness.cpp)."""
    import ctypes as C
    L = orc_fm.lib
    F = C.POINTER(C.c_float)
    U8 = C.POINTER(C.c_uint8)
    fs = iq_rate // decim
    raw = np.fromfile(prefix + ".extra.bin", np.uint8)
    n = min(4096, iq.size // 2 // max(1, decim))
    pos = 0
    if decim > 1:
        d = L.orc_decim_create(decim, 28 if decim >= 8 else 20, 80.0)
        q = np.zeros(2 * n, np.uint8)
        k = L.orc_decim_execute_u8(d, iq.ctypes.data_as(U8), n * decim, q.ctypes.data_as(U8), n)
        assert k == n and np.array_equal(raw[:2 * n], q)
        pos = 2 * n
        L.orc_decim_destroy(d)
    fl = raw[pos:].view(np.float32)
    h = L.orc_demod_create(fs, 32000)
    L.orc_demod_set_bandwidth_mode(h, 9)
    L.orc_demod_set_deviation(h, 50000.0)
    a = np.zeros(n, np.float32)
    mpx = np.zeros(n, np.float32)
    ka = L.orc_demod_process_split(h, iq.ctypes.data_as(U8), mpx.ctypes.data_as(F), a.ctypes.data_as(F), n)
    na = n * 32000 // fs - 2
    assert ka >= na and np.array_equal(fl[:na], a[:na])
    b = np.zeros(n, np.float32)
    second = np.ascontiguousarray(iq[2 * n:4 * n])
    L.orc_demod_process_split(h, second.ctypes.data_as(U8), b.ctypes.data_as(F), None, n)
    assert np.array_equal(fl[na:na + n], b)
    c = np.zeros(n, np.float32)
    kc = L.orc_demod_downsample(h, b.ctypes.data_as(F), c.ctypes.data_as(F), n)
    assert np.array_equal(fl[na + n:na + n + kc], c[:kc]) and fl.size == na + n + kc + 2 * n
    conv = (iq[:2 * n].astype(np.float32) - np.float32(127.5)) * np.float32(1.0 / 127.5)
    assert np.array_equal(fl[na + n + kc:], conv)       # factor-1 convert, done on the device
    L.orc_demod_destroy(h)


@pytest.mark.parametrize("rate", ["256k", "direct256k", "240k"])
def test_reference_style_main_loop(harness, orc_fm, tmp_path, rate):
    iq_rate, decim = rates(rate)
    nblk = 12
    iq = orc.config1_signal(fs_iq=iq_rate).generate(nblk * 8192 * decim)
    iq_path = tmp_path / "iq.u8"
    iq.tofile(iq_path)
    prefix = str(tmp_path / "out")
    subprocess.run([harness, str(iq_path), str(iq_rate), str(decim), "8192", prefix], check=True)
    ref = orc.Channel(orc_fm, orc.make_config(iq_rate=iq_rate, decimation=decim)).process(iq)
    status = np.loadtxt(prefix + ".status.txt").reshape(nblk, 5)
    audio = np.fromfile(prefix + ".audio.f32", np.float32)
    left, right, k = [], [], 0
    for b in range(nblk):
        n = int(status[b, 0])
        left.append(audio[k:k + n])
        right.append(audio[k + n:k + 2 * n])
        k += 2 * n
    assert np.array_equal(np.concatenate(left), ref.left)
    assert np.array_equal(np.concatenate(right), ref.right)
    assert np.array_equal(status[:, 1].astype(int), ref.status["stereo"])
    assert np.array_equal(status[:, 2].astype(int), ref.status["pilot_tenths"])
    assert np.allclose(status[:, 3], ref.status["clip_ratio"], rtol=0, atol=1e-9)
    assert np.array_equal(status[:, 4].astype(int), ref.status["n_groups"])
    _check_extra(orc_fm, iq, iq_rate, decim, prefix)
    g = np.loadtxt(prefix + ".groups.txt").reshape(-1, 6).astype(np.int64)
    assert len(g) == len(ref.groups)
    for row, r in zip(g, ref.groups):
        assert tuple(row) == (r["block_index"], r["a"], r["b"], r["c"], r["d"], r["errors"])
