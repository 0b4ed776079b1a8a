"""CPU: the C-ABI library loads, exports every symbol include/crane_b200.h declares, and refuses to run
without a GPU (no CPU fallback)."""
import ctypes
import json
import os
import re

import pytest

import crane_b200
from crane_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "crane_b200.h")).read()
    return sorted(set(re.findall(r"CRANE_B200_API[^;(]*?\b(crane_b200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    lib = ctypes.CDLL(crane_b200.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/crane_b200.h but not exported"
    assert sorted(crane_b200._SIGNATURES) == names, "python binding and header disagree"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(crane_b200.CraneB200Error) as e:
        crane_b200.Qwen3Model(synth.TINY_QWEN3)
    assert e.value.code == crane_b200.CUDA_ERROR and "no CPU fallback" in str(e.value)


def test_null_and_bad_config_are_errors_not_crashes():
    lib = crane_b200.load_library()
    h = ctypes.c_void_p()
    assert lib.crane_b200_create(None, 0, ctypes.byref(h)) == crane_b200.INVALID_ARG
    assert lib.crane_b200_forward_step(None, None, 0, 0, None) == crane_b200.INVALID_ARG
    assert lib.crane_b200_num_layers(None) == 0
    lib.crane_b200_destroy(None)
