"""The C-ABI shared library: it loads, exports every symbol include/fmgpu.h declares, and
refuses to work without a CUDA device (no CPU fallback). No compute calls are made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import fmtuner_sdr_b200 as fm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "fmgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fmgpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = fm.load_library()
    names = _declared_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/fmgpu.h but not exported: {missing}"


def test_struct_layouts_match_header():
    assert C.sizeof(fm.engine.Config) == 14 * 4
    assert fm.GROUP_DTYPE.itemsize == 16 and fm.STATUS_DTYPE.itemsize == 20
    assert C.sizeof(fm.SynthParams) == 44


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(fm.EngineError) as ei:
        fm.Engine(fm.make_config(), 1, 0)
    assert "no CUDA device" in str(ei.value) or "-2" in str(ei.value)


def test_null_arguments_are_rejected():
    L = fm.load_library()
    h = C.c_void_p()
    assert L.fmgpu_engine_create(None, 1, 0, C.byref(h)) == -1
    assert L.fmgpu_set_bandwidth_hz(None, 0, 0) == -1
    assert L.fmgpu_reset(None, 0, 31) == -1
    assert L.fmgpu_stereo(None, 0, None, None, None, 0) == 0
    assert L.fmgpu_afpost(None, 0, None, None, 0, None, None, 0) == 0
    assert L.fmgpu_rds(None, 0, None, 0, None, 0) == 0
    assert L.fmgpu_decimate(None, 0, None, 0, None, 0) == 0
    assert L.fmgpu_launch_count(None) == 0


def test_product_does_not_reference_the_oracle():
    """The shipped path must not import, include, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "fmtuner_sdr_b200")
    bad = re.compile(r"^\s*(from\s+oracle|import\s+oracle|#\s*include\s*[\"<][^\">]*oracle)|liboracle|libsiggen",
                     re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not bad.search(text), f
