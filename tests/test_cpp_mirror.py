"""The C++ host mirror (include/calico_b200.hpp) exercised by the reference's own integration test transcribed to it
(tests/cpp/mirror_integration_test.cpp <- calico/test/batch_optimizer_test.cpp:32-213): compiled with g++ and linked against the CUDA
library (GPU) or against the SIMT-emulation build of the same kernel sources (CPU, reduced fixture, a few iterations)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "emul"))


def _build(lib_path, exe):
    libdir, libname = os.path.dirname(lib_path), os.path.basename(lib_path)
    assert libname.startswith("lib") and libname.endswith(".so")
    src = os.path.join(HERE, "cpp", "mirror_integration_test.cpp")
    deps = [src, os.path.join(ROOT, "include", "calico_b200.hpp"), os.path.join(ROOT, "include", "calico_b200.h"), lib_path]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-o", exe, src, "-L", libdir, "-l" + libname[3:-3],
                               "-Wl,-rpath," + libdir, "-lpthread"])
    return exe


@pytest.mark.timeout(900)
def test_cpp_mirror_on_emulated_kernels():
    import build as emul_build
    exe = _build(emul_build.build(), os.path.join(HERE, "cpp", "mirror_integration_test_emul"))
    out = subprocess.run([exe, "2", "2"], capture_output=True, text=True, timeout=850)
    assert out.returncode == 0 and "TEST PASSED" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]


@pytest.mark.gpu
def test_cpp_mirror_reference_integration_test(product_lib):
    exe = _build(product_lib, os.path.join(HERE, "cpp", "mirror_integration_test_gpu"))
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "TEST PASSED" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
