"""Builds and runs the C++ transcription of the reference's unit tests (tests/cpp/test_hgi.cpp)
against include/hgi.hpp + libhgi_b200.so.  Compiling is checked on the CPU; running needs a GPU."""
import os
import subprocess

import pytest

from conftest import ROOT

BIN = os.path.join(ROOT, "build", "test_hgi")


def build():
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    libdir = os.path.join(ROOT, "rustyhgi_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-o", BIN,
                           os.path.join(ROOT, "tests", "cpp", "test_hgi.cpp"),
                           "-L" + libdir, "-l:libhgi_b200.so", "-Wl,-rpath," + libdir])


def test_cpp_host_mirror_compiles_and_links():
    build()
    assert os.path.exists(BIN)


def test_criterion_transcription_compiles():
    """tools/bench_criterion.cpp (the reference's benches/bench.rs against include/hgi.hpp) must keep building."""
    out = os.path.join(ROOT, "build", "bench_criterion")
    libdir = os.path.join(ROOT, "rustyhgi_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-o", out, os.path.join(ROOT, "tools", "bench_criterion.cpp"),
                           "-L" + libdir, "-l:libhgi_b200.so", "-Wl,-rpath," + libdir])
    assert os.path.exists(out)


@pytest.mark.gpu
def test_reference_unit_tests_in_cpp_on_gpu():
    build()
    out = subprocess.run([BIN], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all reference unit tests passed" in out.stdout
