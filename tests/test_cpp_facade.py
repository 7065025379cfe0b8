"""The C++ facade (include/rqcd/facade.hpp): host code stays C++ and calls CUDA through the C ABI.
CPU: it compiles and links against librqcd.so, and without a GPU it refuses to run (exit 3, no fallback).
GPU: the Forest driver's loop written with the reference's class names matches the oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "rapidquadcoptercollisiondetection_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "forest_facade")


def _build():
    src = os.path.join(ROOT, "tests", "cpp", "forest_facade.cpp")
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    deps = [src, os.path.join(ROOT, "include", "rqcd", "facade.hpp"), os.path.join(ROOT, "include", "rqcd.h")]
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.check_call([
        "g++", "-O2", "-std=c++14", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
        "-L", PKG, "-lrqcd", "-L", os.path.join(ROOT, "oracle"), "-loracle",
        f"-Wl,-rpath,{PKG}", f"-Wl,-rpath,{os.path.join(ROOT, 'oracle')}"])


def test_facade_compiles_and_refuses_without_gpu():
    _build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([EXE, "1"], capture_output=True, text=True)
    assert r.returncode == 3, (r.returncode, r.stderr)
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_forest_loop_through_facade_matches_oracle():
    _build()
    r = subprocess.run([EXE, "50"], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "mismatches 0" in r.stdout
