"""Static checks of the SHIPPED library's device code (cuobjdump, no GPU): every cubin is sm_100a; the obstacle table is
staged with a TMA bulk copy on an mbarrier (UBLKCP / SYNCS in the SASS of the check kernels); nothing on this path is a
dense contraction, so there is no tensor-core opcode; the hot kernels exist in the variants the dispatcher expects and
stay inside their register budgets (512 resident threads per SM for the 256- and 512-thread CTAs of k_pool and k_check)."""
import os
import re
import shutil
import subprocess

import pytest

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rapidquadcoptercollisiondetection_b200", "librqcd.so")
pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="needs cuobjdump and the built library")


@pytest.fixture(scope="module")
def sass():
    return subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout


@pytest.fixture(scope="module")
def usage():
    out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    res, name = {}, None
    for ln in out.splitlines():
        m = re.match(r"\s*Function (\S+):", ln)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+)", ln)
        if m and name:
            res[name] = (int(m.group(1)), int(m.group(2)))
    return res


def test_every_cubin_is_sm_100a(sass):
    archs = set(re.findall(r"arch = (\S+)", sass))
    assert archs == {"sm_100a"}


def test_tma_bulk_copy_and_no_tensor_core_opcodes(sass):
    per_fn = {}
    name = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            name = m.group(1)
            per_fn[name] = [0, 0]
            continue
        if name and "UBLKCP" in ln:
            per_fn[name][0] += 1
        if name and "SYNCS" in ln:
            per_fn[name][1] += 1
    for kernel in ("k_pool", "k_check"):
        fns = [f for f in per_fn if kernel + "I" in f and "rqcd_fma" not in f]
        assert fns, kernel
        staged = [f for f in fns if per_fn[f][0] > 0 and per_fn[f][1] > 0]
        if kernel == "k_pool":
            assert staged == fns  # every variant stages the table
        else:
            assert staged  # the pairwise variants have no table
    assert not re.search(r"\b(HMMA|IMMA|DMMA|QGMMA|UTC\w*MMA)\b", sass)


def test_hot_kernel_variants_and_register_budgets(usage):
    exact = {k: v for k, v in usage.items() if k.startswith("_ZN4rqcd")}
    pool = {k: v for k, v in exact.items() if "k_poolI" in k}
    # MODE (2) x MOVING (2) x CTA size (256 / 384 / 512) x PLAIN (2)
    assert len(pool) == 24
    for k, (reg, _) in pool.items():
        threads = int(re.search(r"k_poolILi\dELb[01]ELi(\d+)E", k).group(1))
        assert reg * threads * (512 // threads) <= 65536, k
    check = {k: v for k, v in exact.items() if "k_checkI" in k}
    assert len(check) >= 6
    for k, (reg, _) in check.items():
        assert reg <= 128, k  # two CTAs of 256 threads per SM
    feas = [v for k, v in exact.items() if "k_feasILi4E" in k]
    assert feas and feas[0][0] <= 128  # four CTAs of 128 threads per SM
    # the FMA-contracted build of the same kernels (RQCD_F_FAST_FMA) is in the library too
    assert sum(1 for k in usage if k.startswith("_ZN8rqcd_fma") and "k_poolI" in k) == 24
