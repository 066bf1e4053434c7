"""Independent Python re-derivation of the map-side CONTROL FLOW of the hot path, written from the reference sources only:
DepthFusion::update / fusion (esvo_core/src/core/DepthFusion.cpp:71-192) with its case analysis (create / compatible fuse /
occlusion / replace), SmartGrid::set / clean / getNeighbourhood (esvo_core/include/esvo_core/container/SmartGrid.h:222-243,
308-316,367-386) and DepthRegularization::apply (esvo_core/src/core/DepthRegularization.cpp:19-110).
tests/test_indep_pins.py runs the C++ oracle against it, so that the fusion case analysis, the element order and the
regularisation are no longer checked only against a restatement by the same reading.  TEST INFRASTRUCTURE: plain Python
objects, clarity over speed (a few thousand points).  The arithmetic pieces come from tests/indep_numpy.py.
"""
import copy
import math

import numpy as np

import indep_numpy as ind


class Pt:
    """container::DepthPoint (DepthPoint.cpp:7-35): Student-t and Gaussian fields side by side."""

    def __init__(self, row=0, col=0):
        self.row, self.col = int(row), int(col)
        self.x = np.array([col + 0.5, row + 0.5])
        self.rho, self.var, self.res, self.age = -1.0, 0.0, 0.0, 0
        self.s2, self.nu = 0.0, 0.0                 # never initialised by the reference's constructors; the oracle's choice
        self.p_cam = np.zeros(3)

    def valid(self, *thr):
        if not thr:
            return self.rho > -1e-6                                               # DepthPoint.cpp:215-218
        var_thr, age_thr, rmax, rmin = thr                                        # :220-231
        return self.rho > -1e-6 and self.age >= age_thr and self.var <= var_thr and self.rho <= rmax and self.rho >= rmin

    def update_student_t(self, rho, s2, var, nu):                                 # DepthPoint.cpp:166-188
        st = ind.update_student_t(dict(rho=self.rho, s2=self.s2, nu=self.nu, var=self.var, age=self.age), rho, s2, var, nu)
        self.rho, self.s2, self.nu, self.var, self.age = st["rho"], st["s2"], st["nu"], st["var"], st["age"]

    def copy_from(self, o):                                                       # DepthPoint::copy (:233-245): all but row / col
        self.rho, self.var, self.s2, self.nu = o.rho, o.var, o.s2, o.nu
        self.x, self.p_cam, self.res, self.age = o.x.copy(), o.p_cam.copy(), o.res, o.age


class Grid:
    """container::SmartGrid<DepthPoint>: a pointer grid over an insertion-ordered element list."""

    def __init__(self, rows, cols):
        self.rows, self.cols = rows, cols
        self.cell = {}           # (row, col) -> element
        self.elements = []       # insertion order (_elements)

    def exists(self, r, c):
        return (r, c) in self.cell

    def set(self, r, c, value):                                                   # SmartGrid.h:308-316
        if (r, c) not in self.cell:
            e = Pt(r, c)
            self.elements.append(e)
            self.cell[(r, c)] = e
        self.cell[(r, c)].copy_from(value)

    def assign(self, r, c, value):
        """`dm->get(row, col) = dp_prop`: the implicit assignment copies row_ / col_ too (DepthFusion.cpp:186)."""
        e = self.cell[(r, c)]
        e.copy_from(value)
        e.row, e.col = value.row, value.col

    def clean(self, var_thr, age_thr, rmax, rmin):                                # SmartGrid.h:222-243
        keep = []
        for e in self.elements:
            if e.valid(var_thr, age_thr, rmax, rmin):
                keep.append(e)
            else:
                # the reference clears _grid[temp->row()][temp->col()], i.e. the cell the ELEMENT names, which a replacement
                # (assign) may have pointed elsewhere (a dangling pointer there); like oracle and kernels this restatement
                # erases the cell that really holds the element
                for k, v in list(self.cell.items()):
                    if v is e:
                        del self.cell[k]
        self.elements = keep

    def neighbourhood(self, row, col, radius):                                    # SmartGrid.h:367-386
        out = []
        # `for (int r = row - radius; r <= row + radius; r++)` with size_t row / radius: for row < radius the start wraps to a
        # negative int that the size_t comparison then sees as huge -- the loop body never runs; the same for the columns
        if row < radius or col < radius:
            return out
        for r in range(row - radius, row + radius + 1):
            for c in range(col - radius, col + radius + 1):
                if 0 <= r < self.rows and 0 <= c < self.cols and (r, c) in self.cell and self.cell[(r, c)].valid():
                    out.append(self.cell[(r, c)])
        return out


def fuse_vector(grid, pts, T_world_frame, Pl, W, H, radius):
    """DepthFusion::update (:71-87) of one vector of points (dicts: p_cam, s2, nu, res, age, T_world_cam) into `grid`, Tdist norm.
    Returns the number of fusions."""
    T_frame_world = np.linalg.inv(np.asarray(T_world_frame, float).reshape(4, 4))
    n_fusion = 0
    for p in pts:
        T = T_frame_world @ np.asarray(p["T_world_cam"], float).reshape(4, 4)
        pr = ind.propagate_point(np.asarray(p["p_cam"], float), p["s2"], p["nu"], T, Pl, W, H)             # :18-68
        if pr is None:
            continue
        prop = Pt(pr["row"], pr["col"])
        prop.x = np.asarray(pr["x"], float)
        prop.update_student_t(pr["rho"], pr["s2"], pr["var"], pr["nu"])
        prop.p_cam = np.asarray(pr["p_cam"], float)
        prop.res, prop.age = p["res"], p["age"]
        if radius == 0:                                                                                      # :97-121
            cells = [(prop.row + dy, prop.col + dx) for dy in (0, 1) for dx in (0, 1)]
        else:
            cells = [(prop.row + dy, prop.col + dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
        for (r, c) in cells:
            if c < 0 or c >= W or r < 0 or r >= H:                                                           # boundaryCheck (size_t wrap = out of range)
                continue
            if not grid.exists(r, c):                                                                        # case 1 (:126-145)
                nw = Pt(r, c)
                nw.update_student_t(prop.rho, prop.s2, prop.var, prop.nu)
                nw.res, nw.age = prop.res, prop.age
                nw.p_cam = ind.cam2world(Pl, nw.x, prop.rho)
                grid.set(r, c, nw)
                continue
            cur = grid.cell[(r, c)]
            diff = abs(prop.rho - cur.rho)                                                                   # studentTCompatibleTest (:218-231)
            if diff < 2 * math.sqrt(prop.var) or diff < 2 * math.sqrt(cur.var):                              # case 2.1 (:162-177)
                cur.update_student_t(prop.rho, prop.s2, prop.var, prop.nu)
                cur.age += 1
                cur.res = min(cur.res, prop.res)
                cur.p_cam = ind.cam2world(Pl, cur.x, prop.rho)
                n_fusion += 1
            else:                                                                                            # case 2.2 (:178-188)
                if cur.rho - 2 * math.sqrt(cur.var) > prop.rho:
                    continue
                if prop.var < cur.var and prop.res < cur.res:
                    grid.assign(r, c, prop)
    return n_fusion


def regularize(grid, radius, min_nb, min_close, literal=True):
    """DepthRegularization::apply (:19-110), Tdist branch.  Returns the new grid (dm = dmTmp).
    literal=True: exactly as written -- `dmTmp.set(it->row(), it->col(), *it)` and the neighbourhood are keyed by the coordinates
    the element NAMES, which a replacement in the fusion (case 2.2) may have copied from another pixel: such an element is merged
    into the cell it names and its own cell vanishes.  literal=False: keyed by the cell that holds the element (the oracle's and
    the kernels' documented choice, DESIGN.md deviation 7); the named coordinates stay data."""
    true_cell = {id(e): k for k, e in grid.cell.items()}
    tmp = Grid(grid.rows, grid.cols)
    for it in grid.elements:
        kr, kc = (it.row, it.col) if literal else true_cell[id(it)]
        tmp.set(kr, kc, it)
        new = tmp.cell[(kr, kc)]
        if not literal:
            new.row, new.col = it.row, it.col
        if not it.valid():
            continue
        nbs = grid.neighbourhood(kr, kc, radius)
        is_set = False
        if len(nbs) > min_nb:
            close = [q for q in nbs if q.valid() and (abs(it.rho - q.rho) < 2.0 * math.sqrt(it.var) or abs(it.rho - q.rho) < 2.0 * math.sqrt(q.var))]
            if len(close) > min_close:
                nu_post, rho_post, s2_post = close[0].nu, close[0].rho, close[0].s2
                for q in close[1:]:
                    nu_p, rho_p, s2_p = nu_post, rho_post, s2_post
                    nu_post = min(nu_p, q.nu)
                    rho_post = (q.s2 * rho_p + s2_p * q.rho) / (q.s2 + s2_p)
                    s2_post = (nu_post + (rho_p - q.rho) ** 2 / (s2_p + q.s2)) / (nu_post + 1) * (s2_p * q.s2) / (s2_p + q.s2)
                new.rho = rho_post
                is_set = True
        if not is_set:
            new.rho = -1.0
    return tmp


def naive_propagate_vector(grid, pts, T_world_frame, Pl, W, H):
    """DepthFusion::naive_propagation (DepthFusion.cpp:232-288) with naive_propagate_one_point (:290-327): Gaussian propagation,
    the four pixels around the re-projection, an empty pixel takes the point, an occupied one keeps the closer estimate and is
    replaced only by a propagated point with a smaller residual.  pts: dicts p_cam, var, res, age, T_world_cam."""
    T_frame_world = np.linalg.inv(np.asarray(T_world_frame, float).reshape(4, 4))
    for p in pts:
        T = T_frame_world @ np.asarray(p["T_world_cam"], float).reshape(4, 4)
        pc = np.asarray(p["p_cam"], float)
        pp = T[:3, :3] @ pc + T[:3, 3]
        h = Pl[:, :3] @ pp + Pl[:, 3]
        x = h[:2] / h[2]
        if x[0] < 0 or x[0] >= W or x[1] < 0 or x[1] >= H:
            continue
        prop = Pt(int(math.floor(x[1])), int(math.floor(x[0])))
        prop.x = x
        den = (T[2, :2] @ pc[:2] + T[2, 3]) / pc[2] + T[2, 2]
        J = T[2, 2] / den ** 2
        prop.rho, prop.var = 1.0 / pp[2], max(J * J * p["var"], 1e-6)      # DepthPoint::update on a fresh point + boundVariance (:140-164)
        prop.p_cam, prop.res, prop.age = pp, p["res"], p["age"]
        for dy in (0, 1):
            for dx in (0, 1):
                r, c = prop.row + dy, prop.col + dx
                if c >= W or r >= H:
                    continue
                if not grid.exists(r, c):                                   # case 1
                    nw = Pt(r, c)
                    nw.rho, nw.var = prop.rho, max(prop.var, 1e-6)
                    nw.res, nw.age = prop.res, prop.age
                    nw.p_cam = ind.cam2world(Pl, nw.x, prop.rho)
                    grid.set(r, c, nw)
                else:                                                       # case 2
                    cur = grid.cell[(r, c)]
                    if cur.rho > prop.rho:
                        continue
                    if prop.res < cur.res:
                        grid.assign(r, c, prop)
