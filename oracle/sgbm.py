"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, int arithmetic) of cv::StereoSGBM::compute in MODE_SGBM,
the matcher esvo_Mapping::InitializationAtTime runs on the time-surface pair
(esvo_core/src/esvo_Mapping.cpp:101-108: StereoSGBM::create(0, 48, 11, 8*11*11, 32*11*11, -1, 0, 11); :444 compute()).

OpenCV is a third-party dependency of the reference that is not vendored under /root/reference (README: OpenCV 3.2 /
4.x); the algorithm below is restated from the published implementation (modules/calib3d/src/stereosgbm.cpp:
calcPixelCostBT + computeDisparitySGBM, single pass, 5 aggregation directions, Birchfield-Tomasi cost on the
x-derivative and on the intensities, 16-bit costs, uniqueness + left-right check, 3x3 median) and PINNED bit for bit
against cv2.StereoSGBM_create(...).compute() of this image's cv2 4.13 in tests/test_oracle_sgbm.py.
Two details were settled by that pinning rather than by memory of the source: compute() ends with medianBlur(3), and
L_r subtracts min_k L_r(p-r, k) only (not min + P2) -- invisible in the arg-min, decisive for the uniqueness test on
the flat (all-zero) areas a time surface is full of.  Parity anchor for a future device SGBM; the
product does not import this file.
"""
import numpy as np

DISP_SHIFT = 4
DISP_SCALE = 1 << DISP_SHIFT
MAX_COST = 32767
NR2 = 8


def _sat16(a):
    return np.clip(a, -32768, 32767)


def _pixel_cost_bt(img1, img2, y, minD, maxD, tab, tab_ofs):
    """calcPixelCostBT for row y: returns cost[width1, D] (int32 values that fit CostType)."""
    H, W = img1.shape
    minX1, maxX1 = max(maxD, 0), W + min(minD, 0)
    D, width1 = maxD - minD, maxX1 - minX1
    minX2, maxX2 = max(minX1 - maxD, 0), min(maxX1 - minD, W)
    r1, r2 = img1[y].astype(np.int32), img2[y].astype(np.int32)
    n1 = img1[y - 1].astype(np.int32) if y > 0 else r1
    s1 = img1[y + 1].astype(np.int32) if y < H - 1 else r1
    n2 = img2[y - 1].astype(np.int32) if y > 0 else r2
    s2 = img2[y + 1].astype(np.int32) if y < H - 1 else r2
    t0 = int(tab[tab_ofs])
    # channel 0: clipped x-derivative (3-row Sobel-like), channel 1: the intensity itself
    p1 = np.full((2, W), t0, np.int32)
    p2 = np.full((2, W), t0, np.int32)     # NOT reversed here (OpenCV stores row 2 mirrored; the indexing below undoes it)
    lo = max(min(minX1, minX2) - 1, 1)
    hi = min(max(maxX1, maxX2) + 1, W - 1)
    xs = np.arange(lo, hi)
    def grad(r, n, s):
        return (r[xs + 1] - r[xs - 1]) * 2 + n[xs + 1] - n[xs - 1] + s[xs + 1] - s[xs - 1]
    p1[0, xs] = tab[grad(r1, n1, s1) + tab_ofs]
    p2[0, xs] = tab[grad(r2, n2, s2) + tab_ofs]
    p1[1, xs] = r1[xs]
    p2[1, xs] = r2[xs]
    # the intensity channel's border columns stay tab[0] like in OpenCV (prow[width*c] = prow[width*c + width-1] = tab[0])
    cost = np.zeros((width1, D), np.int32)
    xs1 = np.arange(minX1, maxX1)
    for c in range(2):
        diff_scale = 0 if c == 0 else 2
        a, b = p1[c], p2[c]
        # half-pixel interpolated min / max (integer division truncates toward zero; values are >= 0)
        def minmax(v):
            vl = np.empty_like(v); vr = np.empty_like(v)
            vl[1:] = (v[1:] + v[:-1]) // 2; vl[0] = v[0]
            vr[:-1] = (v[:-1] + v[1:]) // 2; vr[-1] = v[-1]
            return np.minimum(np.minimum(vl, vr), v), np.maximum(np.maximum(vl, vr), v)
        u0, u1 = minmax(a)
        # OpenCV interpolates the mirrored row 2; mirroring swaps left/right neighbours, min/max are symmetric -> same values
        v0, v1 = minmax(b)
        for d in range(minD, maxD):
            xr = xs1 - d                                   # matching column in image 2
            u = a[xs1]; v = b[xr]
            c0 = np.maximum(np.maximum(0, u - v1[xr]), v0[xr] - u)
            c1 = np.maximum(np.maximum(0, v - u1[xs1]), u0[xs1] - v)
            cost[:, d - minD] += np.minimum(c0, c1) >> diff_scale
    return cost


def compute(left, right, min_disparity=0, num_disparities=48, block_size=11, P1=968, P2=3872, disp12_max_diff=-1,
            pre_filter_cap=0, uniqueness_ratio=11):
    """Returns the CV_16S disparity map (disparity * 16, invalid = (minDisparity - 1) * 16)."""
    img1 = np.ascontiguousarray(left, np.uint8); img2 = np.ascontiguousarray(right, np.uint8)
    H, W = img1.shape
    minD, maxD = min_disparity, min_disparity + num_disparities
    sad = block_size if block_size > 0 else 5
    ftzero = max(pre_filter_cap, 15) | 1
    uniq = uniqueness_ratio if uniqueness_ratio >= 0 else 10
    d12 = disp12_max_diff if disp12_max_diff > 0 else 1
    P1 = P1 if P1 > 0 else 2
    P2 = max(P2 if P2 > 0 else 5, P1 + 1)
    minX1, maxX1 = max(maxD, 0), W + min(minD, 0)
    D, width1 = maxD - minD, maxX1 - minX1
    INVALID = (minD - 1) * DISP_SCALE
    disp = np.full((H, W), INVALID, np.int16)
    if minX1 >= maxX1:
        return disp
    SW2 = SH2 = sad // 2
    TAB_OFS = 256 * 4
    k = np.arange(256 + TAB_OFS * 2)
    tab = (np.minimum(np.maximum(k - TAB_OFS, -ftzero), ftzero) + ftzero).astype(np.int32)

    def hsum_row(row):
        """horizontal box sum (2*SW2+1 wide, borders replicated) of the pixel costs of image row `row`."""
        pix = _pixel_cost_bt(img1, img2, row, minD, maxD, tab, TAB_OFS)
        idx = np.clip(np.arange(-SW2, width1 + SW2), 0, width1 - 1)
        cs = np.concatenate([np.zeros((1, D), np.int64), np.cumsum(pix[idx].astype(np.int64), axis=0)])
        return _wrap16(cs[2 * SW2 + 1:] - cs[:-(2 * SW2 + 1)])

    hs = {}
    def hsum(row):
        row = min(max(row, 0), H - 1)
        if row not in hs:
            hs[row] = hsum_row(row)
        return hs[row]

    # L_r of the previous row for directions 1..3 and the running one of direction 0 (borders = 0 like the memset buffers)
    Lprev = np.zeros((4, width1 + 2, D + 2), np.int32)       # [dir, x+1, d+1]; d borders are MAX_COST when read
    minLprev = np.zeros((4, width1 + 2), np.int32)
    C = None
    for y in range(H):
        if y == 0:
            C = _wrap16(hsum(0).astype(np.int64) * (SH2 + 1))
            for kk in range(1, SH2 + 1):
                C = _wrap16(C.astype(np.int64) + hsum(kk))
        else:
            C = _wrap16(C.astype(np.int64) + hsum(y + SH2) - hsum(y - SH2 - 1))
        S = np.zeros((width1, D), np.int32)
        Lcur = np.zeros((4, width1 + 2, D + 2), np.int32)
        minLcur = np.zeros((4, width1 + 2), np.int32)
        for x in range(width1):
            Cp = C[x]
            tot = np.zeros(D, np.int32)
            srcs = ((Lcur[0, x], minLcur[0, x]),               # 0: (x-1, y)
                    (Lprev[1, x], minLprev[1, x]),             # 1: (x-1, y-1)
                    (Lprev[2, x + 1], minLprev[2, x + 1]),     # 2: (x,   y-1)
                    (Lprev[3, x + 2], minLprev[3, x + 2]))     # 3: (x+1, y-1)
            for r, (Lp, mn) in enumerate(srcs):
                delta = int(mn) + P2
                Lp = Lp.copy(); Lp[0] = MAX_COST; Lp[D + 1] = MAX_COST
                L = Cp + np.minimum(np.minimum(Lp[1:D + 1], Lp[0:D] + P1), np.minimum(Lp[2:D + 2] + P1, delta)) - int(mn)
                Lcur[r, x + 1, 1:D + 1] = _wrap16(L)
                minLcur[r, x + 1] = _wrap16(L.min())
                tot = tot + L
            S[x] = _sat16(tot)      # saturate_cast<CostType>(Sp[d] + L0 + L1 + L2 + L3) with Sp[d] == 0
        # second sweep of the row: direction (x+1, y), disparity selection
        disp2 = np.full(W, INVALID, np.int32); disp2cost = np.full(W, MAX_COST, np.int32)
        row = np.full(W, INVALID, np.int32)
        Lnext = np.zeros(D + 2, np.int32); minnext = 0
        for x in range(width1 - 1, -1, -1):
            delta0 = int(minnext) + P2
            Lp = Lnext.copy(); Lp[0] = MAX_COST; Lp[D + 1] = MAX_COST
            L0 = C[x] + np.minimum(np.minimum(Lp[1:D + 1], Lp[0:D] + P1), np.minimum(Lp[2:D + 2] + P1, delta0)) - int(minnext)
            Lnext = np.zeros(D + 2, np.int32); Lnext[1:D + 1] = _wrap16(L0); minnext = _wrap16(L0.min())
            Sp = _sat16(S[x] + L0)
            S[x] = Sp
            best = int(np.argmin(Sp)); minS = int(Sp[best])      # first minimum, like the `<` scan
            dd = np.arange(D)
            if np.any((Sp * (100 - uniq) < minS * 100) & (np.abs(best - dd) > 1)):
                continue
            d = best
            x2 = x + minX1 - d - minD
            if disp2cost[x2] > minS:
                disp2cost[x2] = minS; disp2[x2] = d + minD
            if 0 < d < D - 1:
                denom2 = max(int(Sp[d - 1]) + int(Sp[d + 1]) - 2 * int(Sp[d]), 1)
                num = (int(Sp[d - 1]) - int(Sp[d + 1])) * DISP_SCALE + denom2
                q = abs(num) // (denom2 * 2)                             # C++ int division truncates toward zero
                d = d * DISP_SCALE + (q if num >= 0 else -q)
            else:
                d *= DISP_SCALE
            row[x + minX1] = d + minD * DISP_SCALE
        for x in range(minX1, maxX1):
            d1 = int(row[x])
            if d1 == INVALID:
                continue
            _d, d_ = d1 >> DISP_SHIFT, (d1 + DISP_SCALE - 1) >> DISP_SHIFT
            _x, x_ = x - _d, x - d_
            if (0 <= _x < W and disp2[_x] >= minD and abs(int(disp2[_x]) - _d) > d12 and
                    0 <= x_ < W and disp2[x_] >= minD and abs(int(disp2[x_]) - d_) > d12):
                row[x] = INVALID
        disp[y] = row.astype(np.int16)
        Lprev, minLprev = Lcur, minLcur
    return _median3(disp)      # StereoSGBM::compute ends with medianBlur(disp, disp, 3) (speckle filter off: window 0)


def _median3(a):
    """cv::medianBlur 3x3, BORDER_REPLICATE, on int16."""
    p = np.pad(a, 1, mode="edge")
    H, W = a.shape
    st = np.stack([p[dy:dy + H, dx:dx + W] for dy in range(3) for dx in range(3)])
    return np.sort(st, axis=0)[4].astype(a.dtype)


def _wrap16(a):
    """(CostType) conversion of an int: wraps like a C cast (the costs of this workload never overflow, but stay literal)."""
    a = np.asarray(a).astype(np.int64)
    return (((a + 32768) & 0xFFFF) - 32768).astype(np.int32)
