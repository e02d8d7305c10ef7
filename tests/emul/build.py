"""TEST HARNESS ONLY: compiles the kernel sources with g++ against tests/emul/cuda_emul.h (SIMT emulation).
The result, tests/emul/libcalico_b200_emul.so, is loaded by tests/test_emulated_kernels.py and by nothing else."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(_HERE))
CSRC = os.path.join(ROOT, "calico_b200", "csrc")
LIB = os.path.join(_HERE, "libcalico_b200_emul.so")


def build(force=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(_HERE, "cuda_emul.h"), os.path.join(ROOT, "include", "calico_b200.h")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++20", "-fPIC", "-shared", "-DCB2_EMUL", "-Wl,-Bsymbolic", "-x", "c++", "-I", _HERE, "-I", CSRC, "-o", LIB,
                               os.path.join(CSRC, "cb2_host.cu"), "-lpthread"])
    return LIB


def build_functor_check():
    exe = os.path.join(_HERE, "check_functors")
    src = os.path.join(_HERE, "check_functors.cpp")
    deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "oracle", f) for f in os.listdir(os.path.join(ROOT, "oracle")) if f.endswith(".hpp")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++20", "-DCB2_EMUL", "-I", _HERE, "-I", CSRC, "-o", exe, src, "-lpthread"])
    return exe
