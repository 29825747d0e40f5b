"""The C++ side of the drop-in boundary: tests/cpp/test_foundationpose_b200.cpp (shaped like the
reference's simple_tests/src/test_foundationpose.cpp) is compiled with g++ against
foundationpose_cpp_b200/cpp/detection_6d_b200.hpp + include/fpb200.h and linked with libfpb200.so.
CPU: it builds, and without a GPU model construction fails loudly (exit code 2, no CPU fallback).
GPU: Register / error paths / Track loop pass."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "foundationpose_cpp_b200")
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _build(tmp_path) -> str:
    from foundationpose_cpp_b200 import build
    build.build()
    exe = str(tmp_path / "test_foundationpose_b200")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "cpp"),
           os.path.join(ROOT, "tests", "cpp", "test_foundationpose_b200.cpp"), "-o", exe,
           "-L", PKG, "-lfpb200", f"-Wl,-rpath,{PKG}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _weights(tmp_path):
    import fp_helpers as helpers
    return helpers.write_seeded_weights(str(tmp_path))


def test_cpp_adaptor_builds_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    rp, sp = _weights(tmp_path)
    r = subprocess.run([exe, rp, sp], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2, (r.returncode, r.stderr)
    assert "model construction failed" in r.stderr


@pytest.mark.gpu
def test_cpp_register_track_on_gpu(tmp_path):
    exe = _build(tmp_path)
    rp, sp = _weights(tmp_path)
    r = subprocess.run([exe, rp, sp, "100"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Register ok" in r.stdout


# ---- the plugin classes against the REFERENCE's own headers (B200InferCore : inference_core::BaseInferCore, the exact
# CreateFoundationPoseModel(refiner_core, scorer_core, loaders, K, h, w) factory).  The binary is built by
# `make -C oracle boundary` where /root/reference exists (this container; __graft_entry__.build() does it) into
# oracle/_ref/, which travels to the GPU box like the other checker binaries.
REF_BOUNDARY = os.path.join(ROOT, "oracle", "_ref", "test_reference_boundary")


def _boundary_exe():
    if os.path.isdir("/root/reference"):
        from foundationpose_cpp_b200 import build
        build.build()
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "boundary"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    if not os.path.exists(REF_BOUNDARY):
        pytest.skip("oracle/_ref/test_reference_boundary not built (needs the reference tree at build time)")
    return REF_BOUNDARY


def test_reference_boundary_compiles_against_reference_headers(tmp_path):
    import torch
    exe = _boundary_exe()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    rp, sp = _weights(tmp_path)
    r = subprocess.run([exe, rp, sp], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2, (r.returncode, r.stderr)
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_reference_boundary_register_track_syncinfer_on_gpu(tmp_path):
    exe = _boundary_exe()
    rp, sp = _weights(tmp_path)
    r = subprocess.run([exe, rp, sp], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "SyncInfer ok" in r.stdout and "Register ok" in r.stdout
