"""Pins oracle/sgbm.py (numpy restatement of cv::StereoSGBM MODE_SGBM, the matcher of esvo_Mapping::InitializationAtTime,
esvo_Mapping.cpp:101-108,444) bit for bit against cv2.  The SGM call itself stays with the node (OpenCV); this oracle is the
parity anchor for a device SGBM."""
import importlib.util
import os

import cv2
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("oracle_sgbm", os.path.join(ROOT, "oracle", "sgbm.py"))
sgbm = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(sgbm)


def _ref(L, R, nd=48, bs=11, p1=None, p2=None, d12=-1, uq=11):
    p1 = 8 * bs * bs if p1 is None else p1
    p2 = 32 * bs * bs if p2 is None else p2
    return cv2.StereoSGBM_create(0, nd, bs, p1, p2, d12, 0, uq).compute(L, R), (0, nd, bs, p1, p2, d12, 0, uq)


def test_sgbm_on_time_surface_pair_reference_parameters():
    """The golden time-surface pair (sparse, mostly zero: ties everywhere) with the reference's own parameters."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "small_rig_frame.npz"))
    L, R = z["ts_left"], z["ts_right"]
    ref, args = _ref(L, R)
    got = sgbm.compute(L, R, *args)
    assert got.dtype == np.int16 and (ref >= 0).sum() > 1000
    assert np.array_equal(ref, got)
    x = np.load(os.path.join(ROOT, "tests", "golden", "extras.npz"))
    assert np.array_equal(x["disp16"], got)          # the SGM input of the initialisation fixture is this very map


@pytest.mark.parametrize("bs,nd,uq,d12", [(11, 48, 11, -1), (5, 32, 0, 1000), (3, 16, 15, 2), (1, 16, 0, -1), (7, 64, 5, 1)])
def test_sgbm_textured_images(bs, nd, uq, d12):
    rng = np.random.default_rng(bs * 100 + nd)
    a = cv2.GaussianBlur(rng.integers(0, 256, (36, 160)).astype(np.uint8), (5, 5), 0)
    L, R = a[:, 0:128].copy(), a[:, 7:135].copy()          # true disparity 7
    L[20:, :] = a[20:, 3:131]                                # and 4 in the lower part
    ref, args = _ref(L, R, nd, bs, None, None, d12, uq)
    got = sgbm.compute(L, R, *args)
    assert np.array_equal(ref, got)
    v = ref[ref >= 0]
    assert v.size > 500 and abs(np.median(v[: v.size // 2]) / 16 - 7) < 1.5


def test_sgbm_flat_and_invalid_band():
    Z = np.zeros((20, 100), np.uint8)
    ref, args = _ref(Z, Z)
    got = sgbm.compute(Z, Z, *args)
    assert np.array_equal(ref, got)
    assert (got[:, :47] == -16).all() and (got[:, 49:] == 0).all()   # x < numDisparities is never matched
