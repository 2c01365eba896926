"""Host-side mirror of what DEM::discreteElementGet + DEM::discreteElementInit leave behind for a box problem with
single-sphere elements (DEM.cpp:13-83, 186-296, 435-640, 1270-1312; elmt.cpp:13-87): the dict lbGpuDemInit takes through
`LB.demInit` -- material constants, sub-step length, neighbour-table range, per-element mass and inertia, the plane walls of
the lattice boundaries.  Same expressions in the same order as the reference (checked number by number against the values
the unmodified reference holds after its own initialisation, tests/test_dem_port.py).  Periodic pairs of lattice boundaries make
periodic DEM boundaries (DEM::initializePbcs, DEM.cpp:937-988; listed under `pbcs` for completeness), whose ghost particles the
device-side DEM builds for single spheres."""
from __future__ import annotations

import math

from . import lattice_init as li


def covered(case: dict, params: dict) -> bool:
    """True when the device-side DEM (lbGpuDem*) covers this case: spheres and clusters of 2-4 spheres between plane walls,
    periodic pairs of boundaries for single spheres (ghost particles, DEM.cpp:1586-1660), box geometry, an imposed number of
    sub-steps."""
    b = params["boundary"]
    els = case.get("elements", [])
    periodic = any(v == 4 for v in b)
    return (len(els) > 0 and all(1 <= int(e["size"]) <= 4 for e in els) and (not periodic or all(int(e["size"]) == 1 for e in els)) and
            all((b[2 * a] == 4) == (b[2 * a + 1] == 4) for a in range(3)) and
            case.get("problemName", "NONE") == "NONE" and int(case.get("multiStep", 1)) > 0 and
            float(case.get("demInitialRepeat", 0.0)) == 0.0)


def prototypes():
    """DEM::compositeProperties (DEM.cpp:404-433): sphere i of an element of `size`, unit = radius.  The reference writes the
    triangle's y = -1/2 with integers, i.e. 0: kept, the reference's inertia and contacts are those of that shape."""
    s2, s3, s6 = math.sqrt(2), math.sqrt(3), math.sqrt(6)
    return {1: [(0.0, 0.0, 0.0)],
            2: [(0.5, 0.0, 0.0), (-0.5, 0.0, 0.0)],
            3: [(0.0, 1.0, 0.0), (-s3 / 2, 0.0, 0.0), (s3 / 2, 0.0, 0.0)],
            4: [(0.0, 0.0, 1.0), (0.0, 2.0 * s2 / 3.0, -1.0 / 3.0), (2.0 * s6 / 6.0, -2.0 * s2 / 6.0, -1.0 / 3.0),
                (-2.0 * s6 / 6.0, -2.0 * s2 / 6.0, -1.0 / 3.0)]}


def dem_from_case(case: dict, params: dict | None = None) -> dict:
    prm = params or li.params_from_case(case)
    if not covered(case, prm):
        raise ValueError("dem_from_case: the device-side DEM covers spheres / clusters of 2-4 spheres in a box (periodic pairs of "
                         "boundaries: spheres only) with an imposed multiStep")
    L, T, D = prm["unitLength"], prm["unitTime"], prm["unitDensity"]
    accel = L / T / T
    # material (DEM.cpp:13-62); the HERTZIAN case of the switch falls through into LINEAR's damping coefficient
    young, poisson = float(case["youngMod"]), float(case["poisson"])
    rest = float(case["restitution"])
    p = dict(
        contactModel=1 if case.get("contactModel", "none") == "HERTZIAN" else 0,
        multiStep=int(case["multiStep"]),
        knConst=2.0 / 3.0 * young / (1 - poisson * poisson),
        ksConst=2.0 * young / (2.0 - poisson) / (1.0 + poisson),
        dampCoeff=-1.0 * math.sqrt(2.0) * math.log(rest) / math.sqrt((math.log(rest) * math.log(rest) + math.pi)),
        viscTang=float(case["viscTang"]), linearStiff=float(case["linearStiff"]),
        frictionCoefPart=float(case["frictionCoefPart"]), frictionCoefWall=float(case["frictionCoefWall"]),
        numVisc=float(case.get("numVisc", 0.0)),
        demF=[v * accel for v in prm["lbF"]],                      # DEM.cpp:193: demF = lbF * unit.Accel
        deltat=T / float(int(case["multiStep"])),                  # DEM.cpp:213
    )
    density = float(case["density"])
    elmts = []
    protos = prototypes()
    for e in case["elements"]:
        r = float(e["radius"]); size = int(e["size"])
        single = 4.0 / 3.0 * density * math.pi * r * r * r           # elmt::initialize (elmt.cpp:63-75)
        m = size * single
        inertia = [size * 2.0 / 5.0 * single * r * r * 1.0] * 3
        for i in range(size):  # Huygens-Steiner: + singleMass r^2 prototype.transport()
            px, py, pz = protos[size][i]
            tr = (py * py + pz * pz, pz * pz + px * px, px * px + py * py)
            inertia = [inertia[k] + single * r * r * tr[k] for k in range(3)]
        elmts.append(dict(size=size, radius=r, m=m, I=inertia, x0=[float(v) for v in e["x0"]],
                          x1=[float(v) for v in e["x1"]], w0=[float(v) for v in e["w"]]))
    # neighbour-table range (DEM::initNeighborParameters, DEM.cpp:1270-1312)
    max_rad = max(e["radius"] for e in elmts)
    widths = []
    for k in range(3):
        dem_size = float(prm["size"][k]) * L
        w = min(max_rad * 5.0, dem_size)
        n_cells = int(math.ceil(dem_size / w))
        widths.append(dem_size / float(n_cells))
    p["nebrRange"] = max(max_rad * 3.0, 0.5 * min(widths[0], min(widths[1], widths[2])))
    p["maxDisp"] = 0.5 * p["nebrRange"]
    # walls (DEM::initializeWalls, DEM.cpp:435-640)
    walls = []
    for w in li.make_walls(prm, case.get("wall_vel", ())):
        n = [0.0, 0.0, 0.0]; pnt = [0.0, 0.0, 0.0]
        n[w.axis] = 1.0 if w.side == 0 else -1.0
        pnt[w.axis] = 0.5 * L if w.side == 0 else (float(prm["size"][w.axis]) - 1.5) * L
        walls.append(dict(n=n, p=pnt, vel=list(w.vel), omega=[0.0, 0.0, 0.0], rotCenter=[0.0, 0.0, 0.0], moving=int(w.moving)))
    # periodic DEM boundaries (DEM::initializePbcs, DEM.cpp:937-988): one per periodic pair of lattice boundaries
    pbcs = []
    for a in range(3):
        if prm["boundary"][2 * a] == 4:
            pp = [0.0, 0.0, 0.0]; vv = [0.0, 0.0, 0.0]
            pp[a] = 0.5 * L
            vv[a] = (float(prm["size"][a]) - 2.0) * L
            pbcs.append(dict(p=pp, v=vv))
    return dict(params=p, elmts=elmts, walls=walls, pbcs=pbcs, counts=dict(pbcs=len(pbcs), cylinders=0, objects=0))
