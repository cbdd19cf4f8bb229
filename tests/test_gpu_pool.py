"""The multi-device C ABI (asw_pool_*): pair-sharded batches and the disparity-range split with the keys MIN-reduced in
device memory.  A pool of ONE device runs everywhere (no collective needed); the NCCL leg needs >= 2 devices and is skipped
on the single-GPU test box (tools/multi_gpu_check.py runs it under `gpurun --gpus 2`)."""
import numpy as np
import pytest

import aswstereomatch_b200 as asw
from aswstereomatch_b200.synth import make_pair

pytestmark = pytest.mark.gpu


def n_devices():
    return int(asw.load_library().asw_device_count())


@pytest.fixture(scope="module")
def pairs():
    return [make_pair(72, 104, 12, 200 + i)[:2] for i in range(5)]


@pytest.mark.parametrize("n", [1, 2])
def test_pool_batches_equal_single_calls(ctx, pairs, n):
    if n > n_devices():
        pytest.skip("needs %d devices" % n)
    pool = asw.Pool(n)
    try:
        assert pool.size == n
        Ls, Rs = [p[0] for p in pairs], [p[1] for p in pairs]
        for alg, win in ((asw.ADAPTIVE_WEIGHT_GUIDED_FILTER_2, 9), (asw.ADAPTIVE_WEIGHT, 7), (asw.ADAPTIVE_WEIGHT_BLO1, 9)):
            got = pool.stereoMatchingBatch(Ls, Rs, 0, alg, win, 0, 12)
            for (L, R), d in zip(pairs, got):
                assert np.array_equal(d, ctx.stereoMatching(L, R, 0, alg, win, 0, 12, strict=True))
        got = pool.guidedf2_lr_refine_batch(Ls, Rs, 1e-4, 9, 0, 12)
        for (L, R), d in zip(pairs, got):
            assert np.array_equal(d, ctx.guidedf2_lr_refine(L, R, 1e-4, 9, 0, 12))
        # errors keep the reference's contract: out-of-scope algorithm -> status, message available
        with pytest.raises(asw.AswError):
            pool.stereoMatchingBatch(Ls, Rs, 0, asw.SGBM, 9, 0, 12)
        with pytest.raises(asw.AswError):
            pool.stereoMatchingBatch(Ls, [Rs[0][:-1]] + Rs[1:], 0, asw.ADAPTIVE_WEIGHT, 7, 0, 12)
    finally:
        pool.close()


@pytest.mark.parametrize("n", [1, 2])
def test_pool_split_equals_unsplit(ctx, n):
    if n > n_devices():
        pytest.skip("needs %d devices" % n)
    pool = asw.Pool(n)
    try:
        L, R, _ = make_pair(80, 120, 13, 77)
        for alg, win, tol in ((asw.ADAPTIVE_WEIGHT_GUIDED_FILTER_2, 9, 1.0), (asw.ADAPTIVE_WEIGHT, 9, 1.0),
                              (asw.ADAPTIVE_WEIGHT_BILATERAL_GRID, 9, 1.0), (asw.ADAPTIVE_WEIGHT_BLO1, 9, 1.0),
                              (asw.ADAPTIVE_WEIGHT_GEODESIC, 9, 0.9999)):
            full = ctx.stereoMatching(L, R, 0, alg, win, 0, 13, strict=True)
            assert (pool.stereoMatchingSplit(L, R, 0, alg, win, 0, 13) == full).mean() >= tol, alg
        with pytest.raises(asw.AswError):       # the weighted-median method has no split
            pool.stereoMatchingSplit(L, R, 0, asw.ADAPTIVE_WEIGHT_MEDIAN, 7, 0, 13)
    finally:
        pool.close()
