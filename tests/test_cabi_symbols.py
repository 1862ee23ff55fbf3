"""CPU: the product library loads, exports every symbol include/b200fft.h declares, and refuses to
plan without a GPU (no CPU fallback).  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

from util import ROOT

import rustfft_b200 as rb


def _declared():
    hdr = open(os.path.join(ROOT, "include", "b200fft.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b200fft_[a-z_0-9]+)\s*\(", hdr)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge

    if not os.path.exists(rb.DEFAULT_LIB_PATH):
        ge.build()
    return rb.Library()


def test_header_symbols_all_exported(lib):
    names = _declared()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib.c, n), f"{n} declared in include/b200fft.h but not exported"
    assert set(names) == set(rb.Library.SYMBOLS)
    assert lib.version() == "b200fft 0.1 sm_100a"


def test_cubin_is_sm_100a_only():
    out = os.popen(f"/usr/local/cuda/bin/cuobjdump -lelf {rb.DEFAULT_LIB_PATH} 2>/dev/null").read()
    if not out:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_no_gpu_means_no_plan(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; this checks the no-device behaviour")
    assert lib.device_count() == 0
    h = ctypes.c_void_p()
    rc = lib.c.b200fft_plan_create(ctypes.byref(h), 1024, 0, 0, 0)
    assert rc == -2 and not h  # B200FFT_ERR_NO_DEVICE, like FftPlannerAvx::new() -> Err(())
    assert b"no CPU fallback" in lib.c.b200fft_last_error()
    with pytest.raises(rb.FftError):
        rb.FftPlanner(np.complex64, lib=lib)


def test_product_never_references_the_oracle_or_emulator():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rustfft_b200")):
        for f in files:
            if f.endswith((".py", ".h", ".cu", ".inl", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "liboracle" not in txt and "libb200fft_emu" not in txt, f
