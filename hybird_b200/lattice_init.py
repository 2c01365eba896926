"""Host-side lattice initialisation: the state LB::latticeBolzmannInit hands to the first step.

The drop-in shim (hybird_b200/shim/LB_gpu.cpp) lets the reference's own host code build this
state and uploads it; this module builds the same state without the reference for the benchmark,
the smoke test and the GPU tests: types, particle flags, wall nodes, hydrostatic density, initial
velocity, masses.  It restates LB.cpp:324-996 (initializeTypes, initializeLists,
initializeVariables, initializeWalls) and DEM::initializeWalls (DEM.cpp:435-640) for box domains
whose boundaries are the six lattice planes (problemName NONE / demChute); cylinders, objects and
the other hard-coded problem geometries are out of scope.  numpy only - no device code here;
populations are left to lbGpuInit (f=NULL => equilibrium of (n,u), as node::initialize does).

All arrays are in the reference's cell order i = x + X*(y + Y*z).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# D3Q19 velocity set, lattice.h:36-60
CX = np.array([0, 1, -1, 0, 0, 0, 0, 1, -1, -1, 1, 0, 0, 0, 0, 1, -1, 1, -1])
CY = np.array([0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 1, -1, -1, 1, 0, 0, 0, 0])
CZ = np.array([0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, -1, 1])

FLUID, GAS, INTERFACE, PERIODIC, SLIP_STAT, SLIP_DYN, STAT_WALL, DYN_WALL, CURVED = 0, 2, 3, 4, 5, 6, 7, 8, 9
P_BIT, NODE_BIT = 0x10, 0x20

PARTICLE_DTYPE = np.dtype([("x0", "<f8", 3), ("r", "<f8"), ("radiusVec", "<f8", 3), ("clusterIndex", "<u4"),
                           ("particleIndex", "<u4")], align=True)
ELEMENT_DTYPE = np.dtype([("x1", "<f8", 3), ("wGlobal", "<f8", 3), ("compBegin", "<u4"), ("compEnd", "<u4")],
                         align=True)


@dataclass
class Wall:
    """DEM wall created from a lattice boundary (DEM.cpp:435-640)."""
    index: int
    axis: int
    side: int  # 0 = low plane, 1 = high plane
    moving: bool
    slip: bool
    vel: tuple = (0.0, 0.0, 0.0)  # physical units


@dataclass
class LatticeState:
    params: dict
    type_flags: np.ndarray
    solidIndex: np.ndarray
    n: np.ndarray
    u: np.ndarray
    mass: np.ndarray
    visc: np.ndarray
    walls: list = field(default_factory=list)
    parts: np.ndarray = None
    elmts: np.ndarray = None
    comps: np.ndarray = None


def params_from_case(case: dict) -> dict:
    """LB::latticeBoltzmannGet (LB.cpp:85-188): units, sizes, scaled viscosities and force."""
    L = float(case.get("unitLength", 1.0)); T = float(case.get("unitTime", 1.0)); D = float(case.get("unitDensity", 1.0))
    size = [int(math.floor(float(case["lbSize" + a]) / L + 0.5)) for a in "XYZ"]
    # measureUnits::setComposite (node.cpp:476-488)
    accel = L / T / T
    dynVisc = D * L * L / T
    stress = D * L * L / T / T
    speed = L / T
    lbF = [float(case.get("lbF" + a, 0.0)) for a in "XYZ"]
    if case.get("problemName") == "demChute":  # LB.cpp:163-171
        inc = float(case.get("chuteInclination", 0.0))
        g = 9.086
        lbF = [-1.0 * g * math.sin(inc * (math.pi / 180)), -1.0 * 0.0, -1.0 * g * math.cos(inc * (math.pi / 180))]
    return dict(
        size=size, boundary=[int(case["boundary%d" % k]) for k in range(6)],
        lbF=[v / accel for v in lbF],
        initVelocity=[float(case.get("initVelocity" + a, 0.0)) / speed for a in "XYZ"],
        initDynVisc=float(case["initVisc"]) / dynVisc, plasticVisc=float(case["plasticVisc"]) / dynVisc,
        yieldStress=float(case["yieldStress"]) / stress, turbConst=float(case.get("turbConst", 0.0)),
        slipCoefficient=float(case.get("slipCoefficient", 0.0)),
        freeSurface=int(case.get("freeSurfaceSolve", 0)), forceField=int(case.get("forceFieldSolve", 0)),
        nonNewtonian=int(case.get("nonNewtonianSolve", 0)), turbulence=int(case.get("turbulenceSolve", 0)),
        unitLength=L, unitTime=T, unitDensity=D)


def make_walls(params: dict, wall_vel=()) -> list:
    """DEM::initializeWalls (DEM.cpp:435-640): one wall per boundary of type 5..8, indexed in order."""
    walls = []
    for k in range(6):
        b = params["boundary"][k]
        if b in (5, 6, 7, 8):
            walls.append(Wall(index=len(walls), axis=k // 2, side=k % 2, moving=b in (6, 8), slip=b in (5, 6)))
    for w in wall_vel:
        walls[int(w[0])].vel = (float(w[1]), float(w[2]), float(w[3]))
    return walls


def coords(size, planes=None):
    """x, y, z of every cell in the reference's order; with `planes` (global z of each stored plane) z is global."""
    X, Y, Z = size
    nz = Z if planes is None else len(planes)
    i = np.arange(X * Y * nz, dtype=np.int64)
    zl = i // (X * Y)
    return i % X, (i // X) % Y, (zl if planes is None else np.asarray(planes, dtype=np.int64)[zl])


class Neighbors:
    """neighbors[i].d[j] as LB::initializeLatticeBoundaries builds it (LB.cpp:377-472): shell cells point to
    themselves, interior cells get i+ne[j] with a per-axis wrap when that side is periodic; d[0] of interior
    cells is never assigned and stays 0.  One direction is materialised at a time (large domains).

    `planes` restricts the table to a set of global z planes (a slab window plus its stencil margin): indices are
    then local to that set, and a link that leaves the set points back to the cell itself (only margin cells,
    whose results are discarded, have such links)."""

    def __init__(self, size, boundary, planes=None):
        self.size = size
        self.boundary = boundary
        X, Y, Z = size
        self.planes = np.arange(Z, dtype=np.int64) if planes is None else np.asarray(planes, dtype=np.int64)
        self.x, self.y, self.z = coords(size, None if planes is None else self.planes)
        self.zl = np.arange(X * Y * len(self.planes), dtype=np.int64) // (X * Y)
        self.shell = (self.x == 0) | (self.x == X - 1) | (self.y == 0) | (self.y == Y - 1) | (self.z == 0) | (self.z == Z - 1)
        self.idx = np.arange(X * Y * len(self.planes), dtype=np.int64)
        # local plane holding global plane g (-1: not stored)
        self.local_of = np.full(Z, -1, dtype=np.int64)
        self.local_of[self.planes] = np.arange(len(self.planes))

    def __getitem__(self, j):
        X, Y, Z = self.size
        b = self.boundary
        if j == 0:
            return np.where(self.shell, self.idx, 0)
        xs, ys, zs = self.x + CX[j], self.y + CY[j], self.z + CZ[j]
        if CX[j]:
            if b[1] == PERIODIC: xs = np.where(xs == X - 1, 1, xs)
            if b[0] == PERIODIC: xs = np.where(xs == 0, X - 2, xs)
        if CY[j]:
            if b[3] == PERIODIC: ys = np.where(ys == Y - 1, 1, ys)
            if b[2] == PERIODIC: ys = np.where(ys == 0, Y - 2, ys)
        if CZ[j]:
            if b[5] == PERIODIC: zs = np.where(zs == Z - 1, 1, zs)
            if b[4] == PERIODIC: zs = np.where(zs == 0, Z - 2, zs)
        zt = self.local_of[np.clip(zs, 0, Z - 1)]
        link = np.where(zt >= 0, xs + X * (ys + Y * zt), self.idx)
        return np.where(self.shell, self.idx, link)


class Stencil:
    """What `a[nb[j]]` and `out[nb[j][mask]] = True` compute with the table of `Neighbors`, evaluated on shifted 3-D
    views instead of index gathers (a 1024x256x256 lattice is built in seconds instead of minutes).  The periodic wrap is
    served the way the engine does it: the shell planes of a periodic axis are filled with the opposite interior plane
    (x, then y, then z, whole planes, so edges and corners wrap twice / three times as LB.cpp:438-472 does)."""

    def __init__(self, nb: Neighbors):
        self.nb = nb
        X, Y, Z = nb.size
        self.full = len(nb.planes) == Z
        b = nb.boundary
        self.per = (b[0] == PERIODIC, b[2] == PERIODIC, b[4] == PERIODIC)
        # per stored plane l: the stored plane holding the +z / -z neighbour plane (-1: not stored or l is a shell plane)
        g = nb.planes
        def target(cz):
            gz = g + cz
            if self.per[2]:
                gz = np.where(gz == Z - 1, 1, gz)
                gz = np.where(gz == 0, Z - 2, gz)
            t = nb.local_of[np.clip(gz, 0, Z - 1)]
            return np.where((g >= 1) & (g <= Z - 2), t, -1)
        self.zt = {1: target(1), -1: target(-1), 0: np.where((g >= 1) & (g <= Z - 2), np.arange(len(g)), -1)}

    def _filled(self, A):
        X, Y, _ = self.nb.size
        if not (self.per[0] or self.per[1]):
            return A
        E = A.copy()
        if self.per[0]:
            E[:, :, 0] = E[:, :, X - 2]; E[:, :, X - 1] = E[:, :, 1]
        if self.per[1]:
            E[:, 0, :] = E[:, Y - 2, :]; E[:, Y - 1, :] = E[:, 1, :]
        return E

    def gather(self, a, j):
        """a[nb[j]] for j >= 1 (shell cells link to themselves)."""
        X, Y, _ = self.nb.size
        nz = len(self.nb.planes)
        A = a.reshape(nz, Y, X)
        E = self._filled(A)
        out = A.copy()
        cx, cy, cz = int(CX[j]), int(CY[j]), int(CZ[j])
        zt = self.zt[cz]
        ys, xs = slice(1 + cy, Y - 1 + cy), slice(1 + cx, X - 1 + cx)
        if self.full and not (cz and self.per[2]):
            out[1:nz - 1, 1:Y - 1, 1:X - 1] = E[1 + cz:nz - 1 + cz, ys, xs]
        else:
            for l in np.nonzero(zt >= 0)[0]:
                out[l, 1:Y - 1, 1:X - 1] = E[zt[l], ys, xs]
        return out.reshape(-1)

    def scatter_or(self, out, mask, j):
        """out[nb[j][mask]] = True."""
        X, Y, _ = self.nb.size
        nz = len(self.nb.planes)
        O = out.reshape(nz, Y, X)
        M = mask.reshape(nz, Y, X)
        shell = self.nb.shell.reshape(nz, Y, X)
        O |= M & shell  # shell cells link to themselves
        if j == 0:  # d[0] of interior cells is never assigned: it stays 0 (LB.cpp:377-387)
            if (M & ~shell).any():
                out[0] = True
            return
        cx, cy, cz = int(CX[j]), int(CY[j]), int(CZ[j])
        zt = self.zt[cz]
        T = np.zeros((nz, Y, X), dtype=bool)
        ys, xs = slice(1 + cy, Y - 1 + cy), slice(1 + cx, X - 1 + cx)
        for l in np.nonzero(zt >= 0)[0]:
            T[zt[l], ys, xs] |= M[l, 1:Y - 1, 1:X - 1]
        # targets on the shell plane of a periodic axis wrap to the opposite interior plane
        if cy and self.per[1]:
            T[:, Y - 2, :] |= T[:, 0, :]; T[:, 1, :] |= T[:, Y - 1, :]; T[:, 0, :] = False; T[:, Y - 1, :] = False
        if cx and self.per[0]:
            T[:, :, X - 2] |= T[:, :, 0]; T[:, :, 1] |= T[:, :, X - 1]; T[:, :, 0] = False; T[:, :, X - 1] = False
        O |= T


def window_planes(Z, periodic_z, zlo, zhi, margin=2):
    """Global planes a slab window [zlo, zhi) needs for the init stencils: the window itself plus `margin` planes
    on either side, wrapped through the periodic boundary (LB.cpp:438-472) or clipped at a wall."""
    want = set(range(zlo, zhi))
    if periodic_z:
        # the interior planes 1..Z-2 form a ring (plane 0 stands for Z-2, plane Z-1 for 1): margins are counted from
        # the window's interior planes
        inner = [g for g in range(zlo, zhi) if 1 <= g <= Z - 2]
        def wrap(g):
            while g < 1: g += Z - 2
            while g > Z - 2: g -= Z - 2
            return g
        for k in range(1, margin + 1):
            want.add(wrap(min(inner) - k)); want.add(wrap(max(inner) + k))
    else:
        for g in list(range(zlo - margin, zlo)) + list(range(zhi, zhi + margin)):
            if 0 <= g < Z:
                want.add(g)
    return np.array(sorted(want), dtype=np.int64)


def neighbor_table(size, boundary):
    nb = Neighbors(size, boundary)
    return np.stack([nb[j] for j in range(19)])


def expand_elements(elements, unit_length=1.0):
    """Particles / elements as elmt::generateParticles leaves them (elmt.cpp:122-137): a sphere per element, or the 2-4
    spheres of a cluster at x0 + r * prototype (DEM::compositeProperties, DEM.cpp:404-433; the orientation of a fresh run is the
    identity, so project(prototype, q0) is the prototype itself)."""
    from .dem_init import prototypes
    protos = prototypes()
    nE = len(elements)
    nP = sum(int(el.get("size", 1)) for el in elements)
    parts = np.zeros(nP, PARTICLE_DTYPE)
    elmts = np.zeros(nE, ELEMENT_DTYPE)
    a = 0
    for e, el in enumerate(elements):
        size = int(el.get("size", 1))
        elmts[e]["x1"] = el["x1"]; elmts[e]["wGlobal"] = el["w"]; elmts[e]["compBegin"] = a
        for i in range(size):
            x0 = np.array(el["x0"], dtype=np.float64)
            if size > 1:
                xa = x0 + float(el["radius"]) * np.array(protos[size][i], dtype=np.float64)
                parts[a]["radiusVec"] = xa - x0
                x0 = xa
            parts[a]["x0"] = x0; parts[a]["r"] = el["radius"]; parts[a]["clusterIndex"] = e; parts[a]["particleIndex"] = a
            a += 1
        elmts[e]["compEnd"] = a
    return parts, elmts, np.arange(nP, dtype=np.uint32)


def advance_kinematic(parts, elmts, x0_elmt, dt):
    """Prescribed rigid motion used by tests/bench: x0 += x1*dt; single-sphere elements."""
    x0_elmt += elmts["x1"] * dt
    parts["x0"] = x0_elmt[parts["clusterIndex"]] + parts["radiusVec"]
    return parts


def regions_from_case(case: dict, prm: dict):
    """The gas regions of a case as (kind, gasInside, values) in the order build_state applies them (lbGpuInitBox)."""
    out = []
    if case.get("problemName") == "demChute":  # LB.cpp:612-627
        out.append((2, 1, [0.025 / prm["unitLength"] + 0.5]))
    if "fluid_box" in case:
        out.append((0, 0, [float(v) for v in case["fluid_box"]]))
    if "gas_box" in case:
        out.append((0, 1, [float(v) for v in case["gas_box"]]))
    if "gas_sphere" in case:
        out.append((1, 1, [float(v) for v in case["gas_sphere"]]))
    if "fluid_sphere" in case:
        out.append((1, 0, [float(v) for v in case["fluid_sphere"]]))
    return out


def build_state(case: dict, parts=None, window=None, reduce_max=None) -> LatticeState:
    """LB::latticeBolzmannInit (LB.cpp:190-219) for a box case (see oracle/cases.py for the keys).

    window=(zlo, zhi): build only the global planes [zlo, zhi) (a slab plus its two halo planes) -- the arrays
    returned cover exactly those planes, in the reference's cell order within the window; the rest of the lattice
    is never materialised.  reduce_max(v) must then return the element-wise maximum of the 3-vector v over all
    ranks (the hydrostatic reference height is a global quantity, LB.cpp:909-944)."""
    prm = params_from_case(case)
    X, Y, Z = prm["size"]
    bnd = prm["boundary"]
    for b in bnd:
        if b not in (4, 5, 6, 7, 8):
            raise ValueError("boundary type %d not supported" % b)
    L = prm["unitLength"]
    speed = L / prm["unitTime"]
    planes = None
    if window is not None:
        zlo, zhi = int(window[0]), int(window[1])
        if not (0 <= zlo < zhi <= Z):
            raise ValueError("window [%d, %d) outside the lattice" % (zlo, zhi))
        planes = window_planes(Z, bnd[4] == PERIODIC, zlo, zhi)
    nb = Neighbors(prm["size"], bnd, planes)
    x, y, z = nb.x, nb.y, nb.z
    N = x.size
    t = np.zeros(N, dtype=np.uint8)  # initializeNodes: all fluid

    def is_wall(a):
        return (a >= 5) & (a <= 9)
    # initializeLatticeBoundaries (LB.cpp:394-432): per axis `if lo ... else if hi`, solid wins in corners
    for axis, c, n_ in ((0, x, X), (1, y, Y), (2, z, Z)):
        lo = (c == 0) & ~is_wall(t)
        t[lo] = bnd[2 * axis]
        hi = (c == n_ - 1) & ~is_wall(t)  # `else if`: a cell is never on both planes of one axis
        t[hi] = bnd[2 * axis + 1]

    # initializeParticleBoundaries (LB.cpp:475-495): active cells, highest particle index wins
    pflag = np.zeros(N, dtype=bool)
    solid = np.zeros(N, dtype=np.uint32)
    if parts is not None and len(parts):
        # cells of each sphere's bounding box only (the reference tests every cell against every particle); the
        # exact test is tVect::insideSphere (vector.cpp:153-158); ascending particle order keeps "last wins"
        for k in range(len(parts)):
            c0 = parts[k]["x0"] / L
            r = parts[k]["r"] / L
            bx = np.arange(max(0, math.floor(c0[0] - r)), min(X - 1, math.ceil(c0[0] + r)) + 1, dtype=np.int64)
            by = np.arange(max(0, math.floor(c0[1] - r)), min(Y - 1, math.ceil(c0[1] + r)) + 1, dtype=np.int64)
            bz = np.arange(max(0, math.floor(c0[2] - r)), min(Z - 1, math.ceil(c0[2] + r)) + 1, dtype=np.int64)
            bz = bz[nb.local_of[bz] >= 0]
            if not (bx.size and by.size and bz.size):
                continue
            gz, gy, gx = np.meshgrid(bz, by, bx, indexing="ij")
            dx, dy, dz = gx - c0[0], gy - c0[1], gz - c0[2]
            ids = (gx + X * (gy + Y * nb.local_of[gz]))[(dx * dx + dy * dy + dz * dz) < r * r]
            ids = ids[(t[ids] == FLUID) | (t[ids] == INTERFACE)]
            pflag[ids] = True
            solid[ids] = parts[k]["particleIndex"]

    # initializeWallBoundaries (LB.cpp:497-533): DEM walls in index order; later walls override
    walls = make_walls(prm, case.get("wall_vel", ()))
    for w in walls:
        c, n_ = ((x, X), (y, Y), (z, Z))[w.axis]
        if w.side == 0:
            p = 0.5 * L / L
            sel = (1.0 * (c - p)) < 0.0
        else:
            p = (float(n_) - 1.5) * L / L
            sel = (-1.0 * (c - p)) < 0.0
        solid[sel] = w.index
        t[sel] = (SLIP_DYN if w.moving else SLIP_STAT) if w.slip else (DYN_WALL if w.moving else STAT_WALL)

    # initializeInterface (LB.cpp:605-815): gas region, then the two closure loops
    gas = np.zeros(N, dtype=bool)
    if case.get("problemName") == "demChute":  # LB.cpp:612-627
        gas |= ((z - (0.025 / L + 0.5)) * 1.0) > 0.0
    if "fluid_box" in case:
        b = case["fluid_box"]
        gas |= ~((x >= b[0]) & (x <= b[1]) & (y >= b[2]) & (y <= b[3]) & (z >= b[4]) & (z <= b[5]))
    if "gas_box" in case:
        b = case["gas_box"]
        gas |= (x >= b[0]) & (x <= b[1]) & (y >= b[2]) & (y <= b[3]) & (z >= b[4]) & (z <= b[5])
    for key, inside_is_gas in (("gas_sphere", True), ("fluid_sphere", False)):
        if key in case:
            s = case[key]
            d2 = (x - s[0]) * (x - s[0]) + (y - s[1]) * (y - s[1]) + (z - s[2]) * (z - s[2])
            ins = d2 < s[3] * s[3]
            gas |= ins if inside_is_gas else ~ins
    t[(t == FLUID) & gas] = GAS
    st3 = Stencil(nb)
    if (t == GAS).any():
        fluid = t == FLUID
        is_gas = t == GAS
        near_gas = np.zeros(N, dtype=bool)
        for j in range(1, 19):
            near_gas |= st3.gather(is_gas, j)
        t[fluid & near_gas] = INTERFACE
        is_fluid = t == FLUID
        near_fluid = np.zeros(N, dtype=bool)
        for j in range(1, 19):
            near_fluid |= st3.gather(is_fluid, j)
        t[(t == INTERFACE) & ~near_fluid] = GAS

    # initializeVariables (LB.cpp:909-944): hydrostatic density below the highest active cell
    active = (t == FLUID) | (t == INTERFACE)
    lbF = prm["lbF"]
    n = np.zeros(N); u = np.zeros((N, 3)); mass = np.zeros(N); visc = np.zeros(N)
    if window is not None:
        own = active & (z >= zlo) & (z < zhi)
        loc = np.array([x[own].max(), y[own].max(), z[own].max()] if own.any() else [-1.0, -1.0, -1.0], dtype=np.float64)
        glob = np.asarray(reduce_max(loc) if reduce_max is not None else loc, dtype=np.float64)
        maxP = tuple(float(v) for v in glob) if glob.max() >= 0 else (0.0, 0.0, 0.0)
    elif active.any():
        maxP = (float(x[active].max()), float(y[active].max()), float(z[active].max()))
    else:
        maxP = (0.0, 0.0, 0.0)
    dot = ((x - maxP[0]) * lbF[0] + (y - maxP[1]) * lbF[1]) + (z - maxP[2]) * lbF[2]
    dens = 1.0 + 3.0 * 1.0 * 1.0 * 1.0 * dot
    fl = t == FLUID
    itf = t == INTERFACE
    n[fl] = dens[fl]; mass[fl] = 1.0
    n[itf] = 1.0; mass[itf] = 0.5 * 1.0
    visc[active] = prm["initDynVisc"]
    for k in range(3):  # node::initialize: u = lbmDt*F/2/n + velocity
        u[active, k] = lbF[k] * 1.0 / 2.0 / n[active] + prm["initVelocity"][k]
    node = active.copy()

    # initializeWalls (LB.cpp:946-996): a node for every wall cell linked (j=0..18) from a non-wall cell
    nonwall = ~is_wall(t)
    linked = np.zeros(N, dtype=bool)
    for j in range(19):
        st3.scatter_or(linked, nonwall, j)
    wall_node = linked & is_wall(t)
    n[wall_node] = 1.0
    wvel = np.zeros((len(walls), 3))
    for w in walls:
        if w.moving:  # wall::getSpeed with omega = 0 (utils.cpp:38-50)
            wvel[w.index] = [w.vel[0] / speed, w.vel[1] / speed, w.vel[2] / speed]
    dyn = wall_node & ((t == DYN_WALL) | (t == SLIP_DYN))
    if dyn.any():
        u[dyn] = wvel[solid[dyn]]
    node |= wall_node

    tf = (t | (pflag.astype(np.uint8) * P_BIT) | (node.astype(np.uint8) * NODE_BIT)).astype(np.uint8)
    prm["nWalls"] = len(walls)
    if window is not None:
        # keep the window's planes, in global order
        keep = np.concatenate([np.arange(X * Y, dtype=np.int64) + X * Y * int(nb.local_of[g]) for g in range(zlo, zhi)])
        tf, solid, n, u, mass, visc = tf[keep], solid[keep], n[keep], u[keep], mass[keep], visc[keep]
    return LatticeState(params=prm, type_flags=tf, solidIndex=solid, n=n, u=u, mass=mass, visc=visc, walls=walls)
