"""The C++ host-side mirror of the reference API (sview_fmindex_b200/include/sview_fmindex.hpp):
compiles everywhere (g++ only, no CUDA headers needed), runs on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "readme_example.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "readme_example.bin")
PKG = os.path.join(ROOT, "sview_fmindex_b200")


def _build():
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(PKG, "libsvfm.so")):
        ge.build()
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", SRC, "-o", EXE, f"-L{PKG}", "-lsvfm", f"-Wl,-rpath,{PKG}"])


def test_cpp_facade_compiles_and_links():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_readme_example_runs():
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "readme_example: ok" in out.stdout
