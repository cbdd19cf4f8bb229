"""The C++ host mirror (include/asw/aswMethods_compat.h) compiles against the C ABI and behaves like the
reference's entry points: empty Mat on failure / without a device, same results as the ctypes path on a GPU."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_compat")


@pytest.fixture(scope="module")
def exe(built):
    src = os.path.join(ROOT, "tests", "cpp", "test_compat.cpp")
    pkg = os.path.join(ROOT, "aswstereomatch_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                           "-L", pkg, "-lasw_b200", f"-Wl,-rpath,{pkg}"])
    return EXE


def test_cpp_shim_compiles_and_fails_loudly_without_gpu(exe):
    import aswstereomatch_b200 as asw
    if asw.load_library().asw_device_count() > 0:
        pytest.skip("a GPU is present")
    out = subprocess.run([exe, "--no-gpu"], capture_output=True, text=True)
    assert out.returncode == 0 and "empty=1" in out.stdout


@pytest.mark.gpu
def test_cpp_shim_matches_ctypes_path(exe, ctx, tmp_path):
    from aswstereomatch_b200.synth import make_pair
    import aswstereomatch_b200 as asw
    H, W, D = 48, 64, 8
    L, R, _ = make_pair(H, W, D, 17)
    L.tofile(tmp_path / "L.bin"); R.tofile(tmp_path / "R.bin")
    r = subprocess.run([exe, str(H), str(W), str(D), str(tmp_path / "L.bin"), str(tmp_path / "R.bin"), str(tmp_path / "o.bin")])
    assert r.returncode == 0
    out = np.fromfile(tmp_path / "o.bin", np.float32).reshape(2, H, W)
    assert np.array_equal(out[0], ctx.stereoMatching(L, R, 0, asw.ADAPTIVE_WEIGHT_GUIDED_FILTER_2, 9, 0, D))
    assert np.array_equal(out[1], ctx.computeAdaptiveWeight(L, R, 30, 20, 0, 7, 0, D))
