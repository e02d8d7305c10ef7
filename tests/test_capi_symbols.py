"""The C-ABI shared library builds for sm_100a without a GPU, loads, and exports every entry point include/calico_b200.h
declares. No compute is called here; on a machine without a GPU the device entry points must fail loudly (status 13),
never fall back to a CPU path."""
import ctypes as C
import os
import re

import pytest

import calico_b200
from calico_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "calico_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cb2_[a-z0-9_]+)\s*\(", text)))


def test_header_and_symbol_list_agree():
    assert _declared() == sorted(calico_b200.C_ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(product_lib):
    lib = C.CDLL(product_lib)
    for name in _declared():
        assert hasattr(lib, name), name


def test_library_contains_sm100a_code(product_lib):
    out = os.popen(f"cuobjdump -lelf {product_lib} 2>/dev/null").read()
    assert "sm_100a" in out


def test_host_side_assembly_and_loud_failure_without_gpu(product_lib):
    a = _capi.CApi(product_lib)
    assert "sm_100a" in a.version()
    o = _capi.Options()
    lib_defaults = _capi.Options()
    a.lib.cb2_default_options(C.byref(lib_defaults))
    for name, _ in _capi.Options._fields_:
        if name == "minimizer_progress_to_stdout":
            assert getattr(lib_defaults, name) == 1     # batch_optimizer.cpp:13
            continue
        assert getattr(o, name) == getattr(lib_defaults, name), name
    # status codes mirror the reference's absl codes
    with pytest.raises(_capi.CalicoError) as e:
        a.set_trajectory(6, [0.0, 1.0], [[0.0] * 6])
    assert e.value.code == _capi.INVALID_ARGUMENT
    import numpy as np
    a.add_rigid_body(0, [0, 0, 0, 1], [0, 0, 0], [0], [[0, 0, 0]])
    with pytest.raises(_capi.CalicoError) as e:
        a.add_rigid_body(0, [0, 0, 0, 1], [0, 0, 0], [0], [[0, 0, 0]])       # world_model.cpp:32-35
    assert e.value.code == _capi.INVALID_ARGUMENT
    with pytest.raises(_capi.CalicoError) as e:
        a.add_sensor(0, 1, "cam", np.zeros(8), [0, 0, 0, 1], [0, 0, 0], 0.0, -1.0, 0, 1.0, True, True, True)   # camera.cpp:62-65
    assert e.value.code == _capi.INVALID_ARGUMENT
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        knots = np.arange(-5, 12) * 0.1
        a.set_trajectory(6, knots, np.zeros((11, 6)))
        with pytest.raises(_capi.CalicoError) as e:
            a.upload()
        assert e.value.code == _capi.INTERNAL
        assert "no CPU fallback" in str(e.value)


def test_package_never_references_the_oracle_or_the_emulator():
    """The product path must not import, link or execute anything under oracle/ or tests/emul/."""
    pkg = os.path.join(ROOT, "calico_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if not fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                continue
            with open(os.path.join(dirpath, fn)) as f:
                src = f.read()
            assert "import oracle" not in src and "from oracle" not in src, fn
            assert "liboracle" not in src, fn
            if fn.endswith(".py"):
                assert "libcalico_b200_emul" not in src and "tests/emul" not in src and "tests.emul" not in src, fn
