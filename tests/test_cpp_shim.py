"""The reference-shaped C++ shim (include/esvo_b200/esvo_core.hpp) compiles against the C ABI with plain g++,
fails loudly without a GPU (CPU box) and runs the BM -> LM -> cull -> fuse chain on the GPU box."""
import os
import subprocess

import pytest

from conftest import has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "esvo_b200", "_build", "example_mapping_frame")


def _build():
    build = os.path.join(ROOT, "esvo_b200", "_build")
    subprocess.check_call(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "mapping_frame.cpp"),
                           "-L" + build, "-lesvo_b200", "-Wl,-rpath," + build, "-o", EXE])


def test_cpp_shim_compiles_and_fails_loudly_without_gpu(product_lib):
    _build()
    if has_gpu():
        return
    p = subprocess.run([EXE], capture_output=True, text=True)
    assert p.returncode == 2 and "no CPU fallback" in p.stdout


@pytest.mark.gpu
def test_cpp_shim_runs_on_gpu(product_lib):
    _build()
    p = subprocess.run([EXE], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "BM 500 seeds" in p.stdout, p.stdout


def test_cpp_shim_event_frontend_helpers(tmp_path):
    """esvo_core::frontend (selectCloseEvents / samplePoseStamps, esvo_Mapping.cpp:536-603) -- host logic, header-only."""
    exe = str(tmp_path / "frontend_check")
    subprocess.check_call(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "frontend_check.cpp"), "-o", exe])
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr


def test_pipelined_c_abi_example_compiles(product_lib, tmp_path):
    """examples/pipelined_stream.cpp (plain C ABI, S frames in flight): builds with g++; without a GPU it stops at esvo_create."""
    build = os.path.join(ROOT, "esvo_b200", "_build")
    exe = str(tmp_path / "pipelined_stream")
    subprocess.check_call(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "pipelined_stream.cpp"),
                           "-L" + build, "-lesvo_b200", "-Wl,-rpath," + build, "-o", exe])
    if has_gpu():
        return
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 2 and "no CPU fallback" in p.stdout
