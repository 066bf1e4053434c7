"""Pins the oracle's restatement of the OpenCV operations against Python cv2 (4.13 in this image).
These are the third-party semantics the reference depends on."""
import ctypes as C

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from esvo_b200 import capi, configs


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


@pytest.mark.parametrize("shape", [(260, 346), (37, 53), (5, 7)])
def test_median3_matches_cv2(oracle_lib, shape):
    rng = np.random.default_rng(0)
    src = rng.integers(0, 256, shape, dtype=np.uint8)
    src[rng.random(shape) < 0.6] = 0
    dst = np.empty_like(src)
    oracle_lib.lib.esvo_oracle_op_median3(_p(src, C.c_uint8), _p(dst, C.c_uint8), shape[1], shape[0], 3)
    assert np.array_equal(dst, cv2.medianBlur(src, 3))


def test_convert_u8_round_half_even(oracle_lib):
    v = np.array([0.5, 1.5, 2.5, 254.5, 255.4999, 255.5, 300.0, -3.0, 127.49999999, 127.5], np.float64)
    out = np.empty(v.size, np.uint8)
    oracle_lib.lib.esvo_oracle_op_cvt_u8(_p(v, C.c_double), _p(out, C.c_uint8), C.c_size_t(v.size))
    ref = cv2.convertScaleAbs(np.clip(v, 0, None).reshape(1, -1))  # saturating RNE for v >= 0
    assert list(out[:6]) == [0, 2, 2, 254, 255, 255]
    assert np.array_equal(out[[0, 1, 2, 3, 4, 5, 6, 8, 9]], ref.ravel()[[0, 1, 2, 3, 4, 5, 6, 8, 9]])
    assert out[7] == 0


@pytest.mark.parametrize("k", [3, 5])
def test_gaussian_blur_u8_matches_cv2(oracle_lib, k):
    rng = np.random.default_rng(1)
    for shape in [(260, 346), (31, 17)]:
        src = rng.integers(0, 256, shape, dtype=np.uint8)
        dst = np.empty_like(src)
        oracle_lib.lib.esvo_oracle_op_gauss_u8(_p(src, C.c_uint8), _p(dst, C.c_uint8), shape[1], shape[0], k)
        assert np.array_equal(dst, cv2.GaussianBlur(src, (k, k), 0.0))
    imp = np.zeros((9, 9), np.uint8); imp[4, 4] = 255
    dst = np.empty_like(imp)
    oracle_lib.lib.esvo_oracle_op_gauss_u8(_p(imp, C.c_uint8), _p(dst, C.c_uint8), 9, 9, 5)
    assert dst[4, 4] == 36  # SURVEY A.2 known answer


def test_sobel_matches_cv2(oracle_lib):
    rng = np.random.default_rng(2)
    src = rng.integers(0, 256, (48, 64)).astype(np.float64)
    dx = np.empty_like(src); dy = np.empty_like(src)
    oracle_lib.lib.esvo_oracle_op_sobel(_p(src, C.c_double), _p(dx, C.c_double), _p(dy, C.c_double), 64, 48)
    assert np.array_equal(dx, cv2.Sobel(src, cv2.CV_64F, 1, 0))
    assert np.array_equal(dy, cv2.Sobel(src, cv2.CV_64F, 0, 1))


def test_remap_u8_matches_cv2(oracle_lib):
    rng = np.random.default_rng(3)
    H, W = 260, 346
    src = rng.integers(0, 256, (H, W), dtype=np.uint8)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    mx = (xx + rng.normal(0, 3, (H, W))).astype(np.float32)
    my = (yy + rng.normal(0, 3, (H, W))).astype(np.float32)
    mx[:5] -= 20; my[:, :5] -= 20; mx[-5:] += 30  # force out-of-image taps
    dst = np.empty_like(src)
    oracle_lib.lib.esvo_oracle_op_remap_u8(_p(src, C.c_uint8), _p(dst, C.c_uint8), W, H, _p(mx, C.c_float), _p(my, C.c_float))
    ref = cv2.remap(src, mx, my, cv2.INTER_LINEAR)
    assert np.array_equal(dst, ref)


@pytest.mark.parametrize("rig", ["hkust", "dsec"])
def test_rectification_tables_match_cv2(oracle_lib, rig):
    """initUndistortRectifyMap / undistortPoints / validity mask of both cameras."""
    l, r = configs.rig_calibs(rig)
    b = capi.Backend(oracle_lib, l, r, configs.params_for(rig, oracle_lib))
    arr = configs.rig_arrays(rig)
    W, H = arr["width"], arr["height"]
    for cam, side in ((0, "left"), (1, "right")):
        c = arr[side]
        m1, m2, lut, mask = b.get_rectify_tables(cam)
        r1, r2 = cv2.initUndistortRectifyMap(c["K"], c["D"], c["R"], c["P"], (W, H), cv2.CV_32FC1)
        assert np.abs(m1 - r1).max() < 2e-3 and np.abs(m2 - r2).max() < 2e-3
        raw = np.stack(np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32)), -1).reshape(-1, 1, 2)
        und = cv2.undistortPoints(raw, c["K"], c["D"], R=c["R"], P=c["P"]).reshape(H, W, 2)
        assert np.abs(lut - und).max() < 2e-3
        ones = np.ones((H, W), np.float32)
        mk = (cv2.remap(ones, r1, r2, cv2.INTER_LINEAR) > 0.999).astype(np.uint8) * 255
        # the mask is computed from the maps; allow the handful of pixels whose map differs by an ulp
        assert (mask != mk).mean() < 1e-3


def test_equidistant_rectification_tables_match_cv2_fisheye(oracle_lib):
    """calib/upenn is the reference's equidistant rig: cv::fisheye::initUndistortRectifyMap / undistortPoints
    (TimeSurface.cpp:341-346,380-385, CameraSystem.cpp:76-92)."""
    l, r = configs.rig_calibs("upenn")
    b = capi.Backend(oracle_lib, l, r, configs.params_for("upenn", oracle_lib))
    arr = configs.rig_arrays("upenn")
    W, H = arr["width"], arr["height"]
    for cam, side in ((0, "left"), (1, "right")):
        c = arr[side]
        m1, m2, lut, mask = b.get_rectify_tables(cam)
        r1, r2 = cv2.fisheye.initUndistortRectifyMap(c["K"], c["D"].reshape(4, 1), c["R"], c["P"], (W, H), cv2.CV_32FC1)
        assert np.abs(m1 - r1).max() < 2e-3 and np.abs(m2 - r2).max() < 2e-3
        raw = np.stack(np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32)), -1).reshape(-1, 1, 2)
        und = cv2.fisheye.undistortPoints(raw, c["K"], c["D"].reshape(4, 1), R=c["R"], P=c["P"]).reshape(H, W, 2)
        assert np.abs(lut - und).max() < 2e-3
        ones = np.ones((H, W), np.float32)
        mk = (cv2.remap(ones, r1, r2, cv2.INTER_LINEAR) > 0.1).astype(np.uint8) * 255
        assert (mask != mk).mean() < 1e-3
    assert abs(b.get_derived()["baseline"] - 19.941771812941038 / 199.6530123165822) < 1e-12


def test_denoising_mask_rule_matches_cv2_median():
    """esvo_core::frontend::createDenoisingMask (shim) uses 'at least 5 of the 3x3 neighbours, borders replicated';
    that IS cv::medianBlur(eventMap, mask, 3) on a 0/255 image (esvo_Mapping.cpp:1046-1054)."""
    rng = np.random.default_rng(3)
    img = (rng.random((260, 346)) < 0.35).astype(np.uint8) * 255
    ref = cv2.medianBlur(img, 3)
    pad = np.pad(img // 255, 1, mode="edge").astype(np.int32)
    cnt = sum(pad[dy:dy + 260, dx:dx + 346] for dy in range(3) for dx in range(3))
    assert np.array_equal(ref, np.where(cnt >= 5, 255, 0).astype(np.uint8))
