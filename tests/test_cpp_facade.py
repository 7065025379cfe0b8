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


def _fmt(x):
    return "%g" % x  # default ostream formatting of a double: 6 significant digits


@pytest.mark.gpu
def test_forest_loop_through_facade_matches_oracle(tmp_path, forest_golden):
    """... and the files it writes for batch 0 equal the reference's checked-in data/ForestPerfEval.csv (columns 1-21;
    22-24 are wall-clock timings) and data/ForestPerfEvaltreeParamsCSV.csv field for field, as text."""
    _build()
    r = subprocess.run([EXE, "50", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "mismatches 0" in r.stdout
    rows = [ln.rstrip(",\n").split(",") for ln in open(tmp_path / "ForestPerfEval.csv")]
    gold = forest_golden["csv"]
    assert len(rows) == 100 and all(len(x) == 24 for x in rows)
    for i, row in enumerate(rows):
        want = [_fmt(v) for v in gold[i, :19]] + [str(int(gold[i, 19])), str(int(gold[i, 20]))]
        assert row[:21] == want, (i, row[:21], want)
    trees = [ln.rstrip(",\n").split(",") for ln in open(tmp_path / "ForestPerfEvaltreeParamsCSV.csv")]
    assert [[_fmt(v) for v in t] for t in forest_golden["trees"]] == trees
