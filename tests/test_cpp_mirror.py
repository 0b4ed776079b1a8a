"""The compiled (C++) host mirror of the reference's ModelBackend interface, include/crane_b200.hpp, built with g++ against the
in-tree library: links on CPU (symbols + loud failure without a GPU), runs a tiny model through the ABI on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "crane_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "test_backend")


def _build():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(LIBDIR, "libcrane_b200.so")):
        g.build()
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_backend.cpp"),
                    "-L", LIBDIR, "-lcrane_b200", f"-Wl,-rpath,{LIBDIR}", "-o", EXE], check=True)


def test_cpp_mirror_links_and_fails_loudly_without_gpu():
    _build()
    r = subprocess.run([EXE, "symbols"], capture_output=True, text=True, timeout=120)
    print(r.stdout)
    assert r.returncode == 0 and "symbols ok" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_on_gpu():
    _build()
    r = subprocess.run([EXE, "gpu"], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "C++ mirror: ok" in r.stdout
