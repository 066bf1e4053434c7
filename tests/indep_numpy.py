"""Independent numpy/scipy re-derivation of the hot-path arithmetic, written from the REFERENCE sources only
(file:line cited per function) -- not from oracle/ and not from the CUDA kernels.  tests/test_indep_pins.py runs the
C++ oracle against these on the golden inputs, so that oracle and kernels are no longer checked only against a
restatement by the same reading ("what's weak" 1).  TEST INFRASTRUCTURE: plain f64 numpy, clarity over speed.
"""
import numpy as np


# ---- a5: tools::meanStdDev / normalizePatch (esvo_core/include/esvo_core/tools/utils.h:74-92),
#          EventBM::zncc_cost (esvo_core/src/core/EventBM.cpp:317-333)
def zncc_cost(pl, pr):
    pl = np.asarray(pl, np.float64); pr = np.asarray(pr, np.float64)
    n = pl.size

    def normalize(p):
        mean = p.sum() / n
        sub = p - mean
        sigma = np.sqrt((sub * sub).sum() / n) + 1e-6
        return (p - mean) / sigma
    return 0.5 * (1.0 - (normalize(pl) * normalize(pr)).sum() / n)


def bm_search(ts_l, ts_r, x1, dmin, dmax, wx, wy):
    """EventBM::epipolarSearching with step 1 (EventBM.cpp:170-226): integer-aligned patches centred on x1 = floor(rectified
    event) in the left and (x1.x - d, x1.y) in the right image; `cost <= min_cost` keeps the LATER disparity on ties (:198)."""
    hx, hy = (wx - 1) // 2, (wy - 1) // 2
    H, W = ts_l.shape
    x, y = int(x1[0]), int(x1[1])
    pl = ts_l[y - hy:y + hy + 1, x - hx:x + hx + 1]
    best, best_d = 1.0, None          # ZNCC_MAX_ = 1 (EventBM.h)
    for d in range(dmin, dmax + 1):
        x2 = x - d
        if x2 - hx < 1 or y - hy < 1 or x2 + hx >= W - 1 or y + hy >= H - 1:   # isValidPatch (:251-267)
            continue
        pr = ts_r[y - hy:y + hy + 1, x2 - hx:x2 + hx + 1]
        c = zncc_cost(pl, pr)
        if c <= best:
            best, best_d = c, d
    return best_d, best


# ---- a7: PerspectiveCamera::cam2World (esvo_core/src/container/CameraSystem.cpp:120-139): literal 4x4 inverse
def cam2world(P, x, inv_depth):
    z = 1.0 / inv_depth
    Pt = np.zeros((4, 4)); Pt[:3, :] = P; Pt[3, :] = [0, 0, 0, z]
    ps = z * np.linalg.inv(Pt) @ np.array([x[0], x[1], 1.0, 1.0])
    return ps[:3] / ps[3]


# ---- a8: DepthProblem::patchInterpolation (esvo_core/src/core/DepthProblem.cpp:193-262)
def patch_interpolation(img, loc, wx, wy):
    H, W = img.shape
    ulx = int(np.floor(loc[0])) - (wx - 1) // 2; uly = int(np.floor(loc[1])) - (wy - 1) // 2
    drx = int(np.floor(loc[0])) + (wx - 1) // 2; dry = int(np.floor(loc[1])) + (wy - 1) // 2
    if ulx < 0 or uly < 0 or drx >= W or dry >= H:
        return None
    ly, lx = int(np.floor(loc[1])), int(np.floor(loc[0]))
    q1 = (lx + 1) - loc[0]; q2 = loc[0] - lx; q3 = (ly + 1) - loc[1]; q4 = loc[1] - ly
    if uly + wy >= H or ulx + wx >= W:
        return None
    src = img[uly:uly + wy + 1, ulx:ulx + wx + 1].astype(np.float64)
    R = q1 * src[:, :wx] + q2 * src[:, 1:wx + 1]
    return q3 * R[:wy, :] + q4 * R[1:wy + 1, :]


# ---- a6/a7: DepthProblem::warping (:162-191) and operator() for LSnorm = Tdist (:34-160)
def depth_residual(rho, coor, T_left_virtual, Pl, Pr, ts_l, ts_r, wx, wy, nu, scale):
    n = wx * wy
    H, W = ts_l.shape

    def fail():
        w = (nu + 1) / (nu + (255.0 / scale) ** 2)
        return np.full(n, np.sqrt(w) * 255.0)
    p_rv = cam2world(Pl, coor, rho)
    p_left = T_left_virtual[:3, :3] @ p_rv + T_left_virtual[:3, 3]
    h1 = Pl[:, :3] @ p_left + Pl[:, 3]; h2 = Pr[:, :3] @ p_left + Pr[:, 3]
    x1 = h1[:2] / h1[2]; x2 = h2[:2] / h2[2]
    hx, hy = (wx - 1) // 2, (wy - 1) // 2
    for xs in (x1, x2):
        if xs[0] < hx or xs[0] > W - hx or xs[1] < hy or xs[1] > H - hy:
            return fail()
    t1 = patch_interpolation(ts_l, x1, wx, wy)
    t2 = patch_interpolation(ts_r, x2, wx, wy) if t1 is not None else None
    if t1 is None or t2 is None:
        return fail()
    r = (t1 - t2).ravel(); r2 = r * r
    s1, s2, first = scale * scale, -1.0, True
    while first or abs(s2 - s1) / s1 > 0.05:                                    # :96
        if not first:
            s1 = s2
        nz = r != 0
        with np.errstate(over="ignore", divide="ignore"):
            tot = float(np.sum(r2[nz] * (nu + 1) / (nu + r2[nz] / s1)))        # :112-114 (order of summation differs)
        if tot == 0:
            s2 = scale * scale
            break
        s2 = tot / n
        first = False
    with np.errstate(over="ignore", divide="ignore"):
        w = (nu + 1) / (nu + r2 / s2)
    return np.sqrt(w) * r


# ---- a11: DepthPoint::update_studentT (esvo_core/src/container/DepthPoint.cpp:166-188)
def update_student_t(state, rho, s2, var, nu):
    """state = dict(rho, s2, nu, var, age); rho <= -1e-6 means 'new point'."""
    st = dict(state)
    if st["rho"] > -1e-6:
        nu_u = min(nu, st["nu"])
        rho_u = (s2 * st["rho"] + st["s2"] * rho) / (st["s2"] + s2)
        s2_u = (nu_u + (st["rho"] - rho) ** 2 / (st["s2"] + s2)) / (nu_u + 1) * (st["s2"] * s2) / (st["s2"] + s2)
        st.update(rho=rho_u, s2=s2_u, nu=nu_u + 1)
        st["var"] = st["nu"] / (st["nu"] - 2) * st["s2"]
        st["age"] += 1
    else:
        st.update(rho=rho, s2=s2, var=var, nu=nu)
    return st


# ---- a12: DepthFusion::propagate_one_point (esvo_core/src/core/DepthFusion.cpp:18-68), Tdist branch
def propagate_point(p_cam, s2, nu, T_prop_prior, Pl, W, H):
    p = T_prop_prior[:3, :3] @ p_cam + T_prop_prior[:3, 3]
    h = Pl[:, :3] @ p + Pl[:, 3]
    x = h[:2] / h[2]
    if x[0] < 0 or x[0] >= W or x[1] < 0 or x[1] >= H:        # boundaryCheck (DepthFusion.cpp:194-199)
        return None
    rho = 1.0 / p[2]
    den = T_prop_prior[2, :2] @ p_cam[:2] + T_prop_prior[2, 3]
    den /= p_cam[2]
    den += T_prop_prior[2, 2]
    J = T_prop_prior[2, 2] / den ** 2
    s2p = J * J * s2
    return dict(row=int(np.floor(x[1])), col=int(np.floor(x[0])), x=x, rho=rho, s2=s2p, nu=nu, var=nu / (nu - 2) * s2p, p_cam=p)


# ---- tools::cayley2rot (esvo_core/src/tools/cayley.cpp:4-21)
def cayley2rot(c):
    c1, c2, c3 = c
    k = 1 + c1 * c1 + c2 * c2 + c3 * c3
    R = np.array([[1 + c1 * c1 - c2 * c2 - c3 * c3, 2 * (c1 * c2 - c3), 2 * (c1 * c3 + c2)],
                  [2 * (c1 * c2 + c3), 1 - c1 * c1 + c2 * c2 - c3 * c3, 2 * (c2 * c3 - c1)],
                  [2 * (c1 * c3 - c2), 2 * (c2 * c3 + c1), 1 - c1 * c1 - c2 * c2 + c3 * c3]])
    return R / k


# ---- RegProblemLM::computeJ_G (esvo_core/src/core/RegProblemLM.cpp:271-320), general x
def compute_J_G(x):
    c1, c2, c3 = x[:3]
    k = 1 + c1 ** 2 + c2 ** 2 + c3 ** 2; k2 = k * k
    A1 = np.array([[2 * c1 / k - 2 * c1 * (1 + c1 ** 2 - c2 ** 2 - c3 ** 2) / k2, -2 * c2 / k - 2 * c2 * (1 + c1 ** 2 - c2 ** 2 - c3 ** 2) / k2,
                    -2 * c3 / k - 2 * c3 * (1 + c1 ** 2 - c2 ** 2 - c3 ** 2) / k2],
                   [2 * c2 / k - 4 * c1 * (c1 * c2 + c3) / k2, 2 * c1 / k - 4 * c2 * (c1 * c2 + c3) / k2, 2 / k - 4 * c3 * (c1 * c2 + c3) / k2],
                   [2 * c3 / k - 4 * c1 * (c1 * c3 - c2) / k2, -2 / k + 4 * c2 * (c1 * c3 - c2) / k2, 2 * c1 / k - 4 * c3 * (c1 * c3 - c2) / k2]])
    A2 = np.array([[2 * c2 / k - 4 * c1 * (c1 * c2 - c3) / k2, 2 * c1 / k - 4 * c2 * (c1 * c2 - c3) / k2, -2 / k - 4 * c3 * (c1 * c2 - c3) / k2],
                   [-2 * c1 / k - 2 * c1 * (1 - c1 ** 2 + c2 ** 2 - c3 ** 2) / k2, 2 * c2 / k - 2 * c2 * (1 - c1 ** 2 + c2 ** 2 - c3 ** 2) / k2,
                    -2 * c3 / k - 2 * c3 * (1 - c1 ** 2 + c2 ** 2 - c3 ** 2) / k2],
                   [2 / k - 4 * c1 * (c1 + c2 * c3) / k2, 2 * c3 / k - 4 * c2 * (c1 + c2 * c3) / k2, 2 * c2 / k - 4 * c3 * (c1 + c2 * c3) / k2]])
    A3 = np.array([[2 * c3 / k - 4 * c1 * (c2 + c1 * c3) / k2, 2 / k - 4 * c2 * (c2 + c1 * c3) / k2, 2 * c1 / k - 4 * c3 * (c2 + c1 * c3) / k2],
                   [-2 / k - 4 * c1 * (c2 * c3 - c1) / k2, 2 * c3 / k - 4 * c2 * (c2 * c3 - c1) / k2, 2 * c2 / k - 4 * c3 * (c2 * c3 - c1) / k2],
                   [-2 * c1 / k - 2 * c1 * (1 - c1 ** 2 - c2 ** 2 + c3 ** 2) / k2, -2 * c2 / k - 2 * c2 * (1 - c1 ** 2 - c2 ** 2 + c3 ** 2) / k2,
                    2 * c3 / k - 2 * c3 * (1 - c1 ** 2 - c2 ** 2 + c3 ** 2) / k2]])
    J = np.zeros((12, 6))
    J[0:3, 0:3] = A1; J[3:6, 0:3] = A2; J[6:9, 0:3] = A3; J[9:12, 3:6] = np.eye(3)
    return J


def bilinear_1x1(img, loc):
    """RegProblemLM::patchInterpolation (RegProblemLM.cpp:418-487) for a 1x1 patch: value at a sub-pixel location."""
    H, W = img.shape
    lx, ly = int(np.floor(loc[0])), int(np.floor(loc[1]))
    if lx < 0 or ly < 0 or lx >= W or ly >= H or ly + 1 >= H or lx + 1 >= W:
        return None
    q1 = (lx + 1) - loc[0]; q2 = loc[0] - lx; q3 = (ly + 1) - loc[1]; q4 = loc[1] - ly
    s = img[ly:ly + 2, lx:lx + 2].astype(np.float64)
    R = q1 * s[:, 0] + q2 * s[:, 1]
    return q3 * R[0] + q4 * R[1]


def track_reproject(p, T, Pl, mask):
    """RegProblemLM::reprojection + isValidPatch for wx = wy = 1 (RegProblemLM.cpp:380-416)."""
    H, W = mask.shape
    pl = T[:3, :3] @ p + T[:3, 3]
    h = Pl[:, :3] @ pl + Pl[:, 3]
    x = h[:2] / h[2]
    if x[0] < 0 or x[0] > W - 1 or x[1] < 0 or x[1] > H - 1:
        return None
    if mask[int(x[1]), int(x[0])] < 125:
        return None
    return x


def warping_transformation(R_, t_, x):
    """RegProblemLM::getWarpingTransformation (RegProblemLM.cpp:322-346)."""
    dR = cayley2rot(x[:3])
    U, _, Vt = np.linalg.svd(R_.T @ dR.T)
    Rcr = U @ Vt
    T = np.eye(4); T[:3, :3] = Rcr; T[:3, 3] = -Rcr @ (x[3:] + dR @ t_)
    return T


def track_residuals(x, pts, R_, t_, Pl, mask, ts_neg, huber):
    """RegProblemLM::operator() + thread (RegProblemLM.cpp:91-176), 1x1 patches, LSnorm Huber (huber=None: l2)."""
    T = warping_transformation(R_, t_, x)
    out = np.empty(len(pts))
    for i, p in enumerate(pts):
        xs = track_reproject(p, T, Pl, mask)
        r = 255.0
        if xs is not None:
            v = bilinear_1x1(ts_neg, xs)
            if v is not None:
                r = v
        if huber is not None:
            w = huber / r if r > huber else 1.0
            out[i] = np.sqrt(w) * r
        else:
            out[i] = r
    return out


def track_jacobian(pts, R_, t_, Pl, mask, d_du, d_dv):
    """RegProblemLM::df at x = 0 (RegProblemLM.cpp:178-269) with J_G_0 = computeJ_G(0)."""
    JG0 = compute_J_G(np.zeros(6))
    Jc = R_.T @ np.array([[1.0 / Pl[0, 0], 0], [0, 1.0 / Pl[1, 1]], [0, 0]])
    T = np.eye(4); T[:3, :3] = R_.T; T[:3, 3] = -R_.T @ t_
    blk = np.zeros((len(pts), 12))
    for i, p in enumerate(pts):
        xs = track_reproject(p, T, Pl, mask)
        if xs is None:
            continue
        gx, gy = bilinear_1x1(d_du, xs), bilinear_1x1(d_dv, xs)
        if gx is None or gy is None:
            continue
        g = np.array([gx / 8, gy / 8])
        dPi = np.zeros((2, 3))
        dPi[:, :2] = Pl[:2, :2] / p[2]
        z2 = p[2] ** 2
        dPi[0, 2] = -(Pl[0, 0] * p[0] + Pl[0, 1] * p[1] + Pl[0, 3]) / z2
        dPi[1, 2] = -(Pl[1, 0] * p[0] + Pl[1, 1] * p[1] + Pl[1, 3]) / z2
        dT = np.zeros((3, 12))
        dT[:, 0:3] = p[0] * np.eye(3); dT[:, 3:6] = p[1] * np.eye(3); dT[:, 6:9] = p[2] * np.eye(3); dT[:, 9:12] = np.eye(3)
        blk[i] = g @ dPi @ Jc @ dPi @ dT * p[2]
    return -blk @ JG0


# ---- event front-end of esvo_Mapping::dataTransferring (esvo_core/src/esvo_Mapping.cpp:536-603), host logic ----
def ros_to_sec(ns):
    """ros::Time::toSec(): sec + 1e-9 * nsec."""
    s, n = divmod(int(ns), 1000000000)
    return float(s) + 1e-9 * float(n)


def ros_from_sec(t):
    """ros::Time(double): sec = floor(t), nsec = round((t - sec) * 1e9), normalised."""
    sec = int(np.floor(t))
    nsec = int(np.round((t - sec) * 1e9))
    sec += nsec // 1000000000
    nsec %= 1000000000
    return sec * 1000000000 + nsec


def select_close_events(t_events, t_end_ns, half_slice, process_event_num):
    """Indices of the events dataTransferring selects (:562-575): walk back from lower_bound(t_end) until lower_bound(t_begin)
    or PROCESS_EVENT_NUM pushes.  When no event is at/after t_end the reference's first push reads the slot one past the newest
    event (deque::end()); that slot is skipped here but still counts against PROCESS_EVENT_NUM (include/esvo_b200/esvo_core.hpp)."""
    t_events = np.asarray(t_events, np.int64)
    secs = np.array([ros_to_sec(t) for t in t_events])                    # EventBuffer_lower_bound compares toSec() doubles (utils.h:50-55)
    t_begin_ns = ros_from_sec(max(0.0, ros_to_sec(t_end_ns) - 10 * half_slice))
    ev_end = int(np.searchsorted(secs, ros_to_sec(t_end_ns), side="left"))
    ev_begin = int(np.searchsorted(secs, ros_to_sec(t_begin_ns), side="left"))
    out, budget = [], process_event_num
    if ev_end == t_events.size and ev_end != ev_begin and budget > 0:
        ev_end -= 1; budget -= 1
    while ev_end != ev_begin and len(out) < budget:
        out.append(ev_end); ev_end -= 1
    return np.array(out, np.int64)


def select_sgm_events(t_events, t_end_ns, half_slice, process_event_num):
    """The INITIALIZATION branch of dataTransferring (:538-552): window 2 * BM_half_slice_thickness, `<=` on the count (one event
    more than PROCESS_EVENT_NUM); the one-past-the-end slot is skipped like in select_close_events."""
    t_events = np.asarray(t_events, np.int64)
    secs = np.array([ros_to_sec(t) for t in t_events])
    t_begin_ns = ros_from_sec(max(0.0, ros_to_sec(t_end_ns) - 2 * half_slice))
    ev_end = int(np.searchsorted(secs, ros_to_sec(t_end_ns), side="left"))
    ev_begin = int(np.searchsorted(secs, ros_to_sec(t_begin_ns), side="left"))
    out, budget = [], process_event_num + 1
    if ev_end == t_events.size and ev_end != ev_begin:
        ev_end -= 1; budget -= 1
    while ev_end != ev_begin and len(out) < budget:
        out.append(ev_end); ev_end -= 1
    return np.array(out, np.int64)


def sample_pose_stamps(t_end_ns, half_slice):
    """Virtual-view stamps of st_map_ (:585-599): t_begin, then t <- Time(t.toSec() + 0.05 * BM_half_slice_thickness) while <= t_end."""
    t_end = ros_to_sec(t_end_ns)
    t = ros_from_sec(max(0.0, t_end - 10 * half_slice))
    out = []
    while ros_to_sec(t) <= t_end:
        out.append(t)
        t = ros_from_sec(ros_to_sec(t) + 0.05 * half_slice)
    return np.array(out, np.int64)


def pack_point_cloud(p_cam, T_world_frame):
    """publishPointCloud (:909-953): p_world = R p_cam + t as pcl::PointXYZ (f32), DepthMap iteration order."""
    T = np.asarray(T_world_frame, float).reshape(4, 4)
    return (np.asarray(p_cam) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)


# ---- comparison modes of esvo_MVStereo: EventMatcher [26] (esvo_core/src/core/EventMatcher.cpp) ----
def em_zncc_cost(pl, pr):
    """EventMatcher::zncc_cost (:253-274): mean-free patches divided by (Frobenius norm + 1e-6)."""
    pl = np.asarray(pl, np.float64); pr = np.asarray(pr, np.float64)
    ls = pl - pl.mean(); rs = pr - pr.mean()
    ln = ls / (np.linalg.norm(ls) + 1e-6); rn = rs / (np.linalg.norm(rs) + 1e-6)
    return 0.5 * (1.0 - float((ln * rn).sum()))


def event_slicing_for_em(t_left, t_low_ns, t_up_ns, slice_thickness):
    """esvo_MVStereo::eventSlicingForEM (esvo_MVStereo.cpp:1008-1040) over the time-ordered left events vEventsPtr_left_:
    numSlice = floor((t_up - t_low) / thickness); a slice runs from its first event to lower_bound(first stamp + thickness)
    INCLUSIVE (the iterator itself belongs to the slice; end() steps back one), the next one starts behind it.
    Returns (counts, median stamps)."""
    t_left = np.asarray(t_left, np.int64)
    secs = np.array([ros_to_sec(t) for t in t_left])
    num = int(np.floor((ros_to_sec(t_up_ns) - ros_to_sec(t_low_ns)) / slice_thickness))
    counts, med = [], []
    it = 0
    for _ in range(num):
        t_end = ros_to_sec(ros_from_sec(secs[it] + slice_thickness))
        it_end = int(np.searchsorted(secs, t_end, side="left"))
        if it_end == t_left.size:
            it_end -= 1
        n = it_end - it + 1
        counts.append(n); med.append(int(t_left[it + n // 2]))
        it = it_end + 1
        if it == t_left.size:
            break
    return np.array(counts, np.int32), np.array(med, np.int64)


def event_match(left, right, slice_counts, slice_poses, lut_l, lut_r, Pl, Pr, baseline, ts_l, ts_r, T_world_left,
                time_thr, epi_thr, ncc_thr, wx, wy, num_thread):
    """EventMatcher::match_all_HyperThread / match / match_an_event (:60-163,185-251).  left / right: dicts of x, y, t, p
    (right time-ordered).  Returns a list of dicts in the reference's thread-major order and the number of zncc evaluations."""
    H, W = ts_l.shape
    tsl = ts_l.astype(np.float64); tsr = ts_r.astype(np.float64)
    slice_of = np.repeat(np.arange(len(slice_counts)), slice_counts)
    total = min(slice_of.size, left["x"].size)
    r_secs = np.array([ros_to_sec(t) for t in right["t"]])
    T_left_world = np.linalg.inv(np.asarray(T_world_left, float).reshape(4, 4))
    hx, hy = (wx - 1) // 2, (wy - 1) // 2
    evals = 0

    def one(i):
        nonlocal evals
        te = ros_to_sec(left["t"][i])
        low = ros_to_sec(ros_from_sec(te - time_thr / 2)); up = ros_to_sec(ros_from_sec(te + time_thr / 2))
        b = int(np.searchsorted(r_secs, low, side="left")); e = int(np.searchsorted(r_secs, up, side="left"))
        cand = [j for j in range(b, e) if low <= r_secs[j] <= up and bool(right["p"][j]) == bool(left["p"][i])]     # temporal + polarity (:66-88)
        if not cand:
            return None
        xl = lut_l[int(left["y"][i]), int(left["x"][i])]
        epi = []
        for j in cand:                                                                                               # epipolar (:93-108)
            xr = lut_r[int(right["y"][j]), int(right["x"][j])]
            if abs(xl[1] - xr[1]) <= epi_thr and xr[0] < xl[0]:
                epi.append((j, xr))
        if not epi:
            return None
        T_world_rv = np.asarray(slice_poses[slice_of[i]], float).reshape(4, 4)
        T_left_rv = T_left_world @ T_world_rv
        f = Pl[0, 0]
        min_cost, best, best_depth = 1.0, 0, 0.0
        for q, (j, xr) in enumerate(epi):                                                                            # motion check (:110-150)
            depth = baseline * f / (xl[0] - xr[0])
            p_rv = cam2world(Pl, xl, 1.0 / depth)
            p_left = T_left_rv[:3, :3] @ p_rv + T_left_rv[:3, 3]
            h1 = Pl[:, :3] @ p_left + Pl[:, 3]; h2 = Pr[:, :3] @ p_left + Pr[:, 3]
            x1 = h1[:2] / h1[2]; x2 = h2[:2] / h2[2]
            if any(xs[0] < hx or xs[0] > W - hx or xs[1] < hy or xs[1] > H - hy for xs in (x1, x2)):               # warping2 (:276-306)
                continue
            pa = patch_interpolation(tsl, x1, wx, wy)
            pb = patch_interpolation(tsr, x2, wx, wy) if pa is not None else None
            if pa is None or pb is None:
                continue
            cost = em_zncc_cost(pa, pb); evals += 1
            if cost < min_cost:
                min_cost, best, best_depth = cost, q, depth
        if min_cost > ncc_thr:
            return None
        with np.errstate(divide="ignore"):
            inv = np.float64(1.0) / np.float64(best_depth)
        return dict(i=i, j=epi[best][0], x_left=xl, x_right=epi[best][1], t_ns=int(left["t"][i]), T=T_world_rv, inv_depth=float(inv), cost=min_cost)

    out = []
    for tid in range(num_thread):
        for i in range(tid, total, num_thread):
            m = one(i)
            if m is not None:
                out.append(m)
    return out, evals


def vemp_to_points(matches, Pl, age_vis_threshold):
    """esvo_MVStereo::vEMP2vDP (esvo_MVStereo.cpp:1072-1097)."""
    out = []
    for m in matches:
        out.append(dict(row=int(np.floor(m["x_left"][1])), col=int(np.floor(m["x_left"][0])), x=m["x_left"], inv_depth=m["inv_depth"],
                        variance=1e-6, residual=m["cost"], age=int(age_vis_threshold), p_cam=cam2world(Pl, m["x_left"], m["inv_depth"]), T=m["T"]))
    return out


# ---- a1/a2: EventQueueMat (esvo_time_surface/include/esvo_time_surface/TimeSurface.h:28-96) and
#      TimeSurface::createTimeSurfaceAtTime, BACKWARD mode (esvo_time_surface/src/TimeSurface.cpp:52-152); image ops by OpenCV itself
def time_surface_backward(ev, T_ns, decay_ms, W, H, ignore_polarity, median_blur_kernel_size, map1, map2, queue_len=20):
    """ev: dict of x, y, t (ns), p in ARRIVAL order.  Returns (raw mono8 before remap, published mono8 after remap)."""
    import collections
    import cv2
    queues = collections.defaultdict(lambda: collections.deque(maxlen=queue_len))       # insertEvent: push_back, pop_front beyond queueLen
    for x, y, t, p in zip(ev["x"].tolist(), ev["y"].tolist(), ev["t"].tolist(), ev["p"].tolist()):
        if 0 <= x < W and 0 <= y < H:
            queues[(x, y)].append((t, p))
    decay_sec = decay_ms / 1000.0
    m = np.zeros((H, W), np.float64)
    for (x, y), q in queues.items():
        for t, p in reversed(q):                                                          # getMostRecentEventBeforeT: newest first, ts < T
            if t < T_ns:
                if ros_to_sec(t) > 0:
                    d = T_ns - t                                                          # ros::Duration (integer ns), then toSec()
                    dt = float(d // 1000000000) + 1e-9 * float(d % 1000000000)
                    v = np.exp(-dt / decay_sec)
                    if not ignore_polarity:
                        v *= 1.0 if p else -1.0
                    m[y, x] = v
                break
    m = 255.0 * m if ignore_polarity else 255.0 * (m + 1.0) / 2.0
    img = np.clip(np.rint(m), 0, 255).astype(np.uint8)                                    # convertTo(CV_8U): saturate_cast(cvRound), half to even
    if median_blur_kernel_size > 0:
        img = cv2.medianBlur(img, 2 * median_blur_kernel_size + 1)
    out = cv2.remap(img, map1, map2, cv2.INTER_LINEAR)
    return img, out


# ---- a4: EventBM::match_an_event / match_all_HyperThread / match (esvo_core/src/core/EventBM.cpp:80-168,269-315) ----
def event_bm_match_all(ev, ts_l, ts_r, lut_l, mask_l, pose_t, poses, wx, wy, dmin, dmax, step, thr, baseline, P00, num_thread):
    """ev: dict x, y, t of the events handed to createMatchProblem (left camera, raw pixels).  Returns the EventMatchPairs (dicts)
    in the reference's thread-major order and the number of zncc evaluations.  Left-right rig only (bUpDownConfiguration false)."""
    H, W = ts_l.shape
    hx, hy = (wx - 1) // 2, (wy - 1) // 2
    tsl = ts_l.astype(np.float64); tsr = ts_r.astype(np.float64)
    pose_sec = np.array([ros_to_sec(t) for t in pose_t])
    evals = 0

    def valid(x, y):                                                   # isValidPatch (:251-267)
        return not (x - hx < 1 or y - hy < 1 or x + hx >= W - 1 or y + hy >= H - 1)

    def search(x1, patch, start, end, sstep, state):                   # epipolarSearching (:170-226)
        nonlocal evals
        costs = {}
        d = start
        while d <= end:
            x2 = x1[0] - d
            if not valid(x2, x1[1]):
                costs[d] = 1.0
            else:
                c = zncc_cost(patch, tsr[x1[1] - hy:x1[1] + hy + 1, x2 - hx:x2 + hx + 1])
                evals += 1
                costs[d] = c
                if c <= state["min"]:
                    state.update(min=c, bx=x2, bd=d)
            d += sstep
        if sstep > 1:
            lo, hi = state["bd"] - sstep, state["bd"] + sstep
            return lo in costs and hi in costs and costs[lo] < 1.0 and costs[hi] < 1.0 and state["min"] < thr
        return state["min"] < thr

    def one(i):
        xr = lut_l[int(ev["y"][i]), int(ev["x"][i])]
        if xr[0] < 0 or xr[0] > W - 1 or xr[1] < 0 or xr[1] > H - 1:
            return None
        if mask_l[int(xr[1]), int(xr[0])] <= 125:                      # Eigen's (double) index: truncation
            return None
        x1 = (int(np.floor(xr[0])), int(np.floor(xr[1])))
        if not valid(*x1):
            return None
        patch = tsl[x1[1] - hy:x1[1] + hy + 1, x1[0] - hx:x1[0] + hx + 1]
        if (patch < 1).sum() > 0.95 * patch.size:                      # low information-to-noise ratio (:104-109)
            return None
        st = dict(min=1.0, bx=0, bd=0)
        if not search(x1, patch, dmin, dmax, step, st):                # coarse
            return None
        if not search(x1, patch, st["bd"] - (step - 1), st["bd"] + (step - 1), 1, st):     # fine
            return None
        if not st["min"] <= thr:
            return None
        disparity = float(x1[0] - st["bx"])
        k = int(np.searchsorted(pose_sec, ros_to_sec(ev["t"][i]), side="left"))             # StampTransformationMap_lower_bound on toSec()
        if k == len(pose_t):
            return None
        depth = baseline * P00 / disparity
        return dict(i=i, x_left_raw=(float(ev["x"][i]), float(ev["y"][i])), x_left=xr, x_right=(float(st["bx"]), float(x1[1])), t_ns=int(ev["t"][i]),
                    pose=k, inv_depth=1.0 / depth, cost=st["min"], disp=disparity)

    out = []
    n = len(ev["x"])
    for tid in range(num_thread):                                      # match (:300-315): thread tid takes events tid, tid + NT, ...
        for i in range(tid, n, num_thread):
            m = one(i)
            if m is not None:
                out.append(m)
    return out, evals
