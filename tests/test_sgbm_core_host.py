"""The device SGBM's arithmetic (esvo_b200/csrc/sgbm_core.h: the per-thread functions sgbm.cu launches) executed on the
host in plain loops (tests/sgbm_host_check.cpp) and pinned bit for bit against cv2.StereoSGBM and oracle/sgbm.py."""
import importlib.util
import os
import struct
import subprocess

import cv2
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_check(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("sgbm") / "sgbm_host_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "sgbm_host_check.cpp"), "-o", exe])

    def run(L, R, nd=48, bs=11, p1=None, p2=None, d12=-1, uq=11):
        L = np.ascontiguousarray(L, np.uint8); R = np.ascontiguousarray(R, np.uint8)
        H, W = L.shape
        p1 = 8 * bs * bs if p1 is None else p1
        p2 = 32 * bs * bs if p2 is None else p2
        fin, fout = exe + ".in", exe + ".out"
        with open(fin, "wb") as f:
            f.write(struct.pack("9i", W, H, nd, bs, p1, p2, d12, 0, uq) + L.tobytes() + R.tobytes())
        assert subprocess.call([exe, fin, fout]) == 0
        ref = cv2.StereoSGBM_create(0, nd, bs, p1, p2, d12, 0, uq).compute(L, R)
        return np.fromfile(fout, np.int16).reshape(H, W), ref
    return run


def test_core_on_time_surface_pair(host_check):
    z = np.load(os.path.join(ROOT, "tests", "golden", "small_rig_frame.npz"))
    got, ref = host_check(z["ts_left"], z["ts_right"])
    assert (ref >= 0).sum() > 1000 and np.array_equal(got, ref)
    spec = importlib.util.spec_from_file_location("oracle_sgbm", os.path.join(ROOT, "oracle", "sgbm.py"))
    sgbm = importlib.util.module_from_spec(spec); spec.loader.exec_module(sgbm)
    assert np.array_equal(got, sgbm.compute(z["ts_left"], z["ts_right"]))


@pytest.mark.parametrize("bs,nd,uq,d12", [(11, 48, 11, -1), (5, 32, 0, 1000), (3, 16, 15, 2), (1, 16, 0, -1), (7, 64, 5, 1), (11, 128, 11, -1)])
def test_core_on_textured_images(host_check, bs, nd, uq, d12):
    rng = np.random.default_rng(bs * 100 + nd)
    a = cv2.GaussianBlur(rng.integers(0, 256, (40, 240)).astype(np.uint8), (5, 5), 0)
    L, R = a[:, 0:200].copy(), a[:, 7:207].copy()
    L[22:, :] = a[22:, 3:203]
    got, ref = host_check(L, R, nd, bs, None, None, d12, uq)
    assert np.array_equal(got, ref)


def test_core_full_size_sparse_pair(host_check):
    """346x260, sparse like a time surface (mostly zeros: every tie-break and the uniqueness test on negative sums matter)."""
    rng = np.random.default_rng(5)
    base = np.zeros((260, 400), np.uint8)
    for _ in range(120):                                      # short bright strokes
        y, x, n = rng.integers(5, 255), rng.integers(5, 360), rng.integers(5, 40)
        dy = rng.integers(-1, 2)
        for k in range(n):
            base[np.clip(y + dy * k // 3, 0, 259), x + k] = rng.integers(60, 255)
    base = cv2.GaussianBlur(base, (3, 3), 0)
    L, R = base[:, 20:366].copy(), base[:, 32:378].copy()      # disparity 12
    got, ref = host_check(L, R)
    assert (ref >= 0).sum() > 20000 and np.array_equal(got, ref)
