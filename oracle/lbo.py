"""ctypes binding of the C restatement oracle + reader for ref_harness dumps.

TEST INFRASTRUCTURE: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg only.  The product package (hybird_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liblboracle.so")
Q = 19


class LboParams(C.Structure):
    _fields_ = [("size", C.c_int32 * 3), ("boundary", C.c_int32 * 6), ("lbF", C.c_double * 3),
                ("initDynVisc", C.c_double), ("plasticVisc", C.c_double), ("yieldStress", C.c_double),
                ("turbConst", C.c_double), ("slipCoefficient", C.c_double),
                ("freeSurface", C.c_int32), ("forceField", C.c_int32), ("nonNewtonian", C.c_int32),
                ("turbulence", C.c_int32),
                ("unitLength", C.c_double), ("unitTime", C.c_double), ("unitDensity", C.c_double)]


PARTICLE_DTYPE = np.dtype([("x0", "<f8", 3), ("r", "<f8"), ("radiusVec", "<f8", 3), ("clusterIndex", "<u4"),
                           ("particleIndex", "<u4")], align=True)
ELEMENT_DTYPE = np.dtype([("x1", "<f8", 3), ("wGlobal", "<f8", 3), ("compBegin", "<u4"), ("compEnd", "<u4")],
                         align=True)
assert PARTICLE_DTYPE.itemsize == 64 and ELEMENT_DTYPE.itemsize == 56


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < max(os.path.getmtime(os.path.join(HERE, f)) for f in ("lb_oracle.c", "lb_oracle.h")):
        subprocess.run(["make", "-C", HERE, "port"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp = C.c_void_p
        L.lbo_create.restype = vp
        L.lbo_create.argtypes = [C.POINTER(LboParams)] + [vp] * 7
        L.lbo_destroy.argtypes = [vp]
        L.lbo_free_surface_step.argtypes = [vp]
        L.lbo_coupling_step.argtypes = [vp, C.c_int, vp, C.c_uint32, vp, C.c_uint32, vp]
        L.lbo_step.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint32, vp, vp, vp, vp, C.c_uint32]
        for nm in ("type_flags", "solid_index", "f", "fs", "n", "u", "hydro_force", "mass", "visc", "shear_rate"):
            getattr(L, "lbo_" + nm).restype = vp
            getattr(L, "lbo_" + nm).argtypes = [vp]
        L.lbo_nodes.restype = C.c_uint32
        L.lbo_nodes.argtypes = [vp]
        L.lbo_neighbor.restype = C.c_uint32
        L.lbo_neighbor.argtypes = [vp, C.c_uint32, C.c_int]
        L.lbo_count_type.restype = C.c_uint32
        L.lbo_count_type.argtypes = [vp, C.c_int]
        L.lbo_set_threads.argtypes = [C.c_int]
        L.lbo_set_curves.argtypes = [vp, C.c_uint32, vp, vp]
        L.lbo_set_mass_target.argtypes = [vp, C.c_double]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_params(p: dict) -> LboParams:
    P = LboParams()
    P.size[:] = [int(v) for v in p["size"]]
    P.boundary[:] = [int(v) for v in p["boundary"]]
    P.lbF[:] = [float(v) for v in p["lbF"]]
    for k in ("initDynVisc", "plasticVisc", "yieldStress", "turbConst", "slipCoefficient", "unitLength", "unitTime",
              "unitDensity"):
        setattr(P, k, float(p[k]))
    for k in ("freeSurface", "forceField", "nonNewtonian", "turbulence"):
        setattr(P, k, int(p[k]))
    return P


class Oracle:
    """One LB state advanced by the C restatement, same call surface as the reference's LB class."""

    def __init__(self, params: dict, type_flags, solid_index, n, u, mass, visc, f=None, threads=1):
        self.L = lib()
        self.L.lbo_set_threads(int(threads))
        self.params = dict(params)
        self.P = make_params(params)
        self.N = int(np.prod(params["size"]))
        a = lambda x, dt: np.ascontiguousarray(x, dtype=dt)
        self._keep = [a(type_flags, np.uint8), a(solid_index, np.uint32), None if f is None else a(f, np.float64),
                      a(n, np.float64), a(u, np.float64), a(mass, np.float64), a(visc, np.float64)]
        self.h = self.L.lbo_create(C.byref(self.P), *[_ptr(x) for x in self._keep])
        if not self.h:
            raise RuntimeError("lbo_create failed")
        self._keep = None
        self.nWalls = int(params.get("nWalls", 0))

    def set_curves(self, cells, delta):
        cells = np.ascontiguousarray(cells, dtype=np.uint32)
        delta = np.ascontiguousarray(delta, dtype=np.float64)
        assert delta.size == Q * cells.size
        self.L.lbo_set_curves(self.h, cells.size, _ptr(cells), _ptr(delta))

    def set_mass_target(self, total_mass):
        self.L.lbo_set_mass_target(self.h, float(total_mass))

    def close(self):
        if self.h:
            self.L.lbo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _view(self, name, dtype, width):
        p = getattr(self.L, "lbo_" + name)(self.h)
        n = self.N * width
        buf = (C.c_uint8 * (n * np.dtype(dtype).itemsize)).from_address(p)
        arr = np.frombuffer(buf, dtype=dtype, count=n)
        return arr.reshape(self.N, width) if width > 1 else arr

    # state views (zero-copy; valid until close)
    type_flags = property(lambda s: s._view("type_flags", np.uint8, 1))
    solid_index = property(lambda s: s._view("solid_index", np.uint32, 1))
    f = property(lambda s: s._view("f", np.float64, Q))
    fs = property(lambda s: s._view("fs", np.float64, Q))
    n = property(lambda s: s._view("n", np.float64, 1))
    u = property(lambda s: s._view("u", np.float64, 3))
    hydro_force = property(lambda s: s._view("hydro_force", np.float64, 3))
    mass = property(lambda s: s._view("mass", np.float64, 1))
    visc = property(lambda s: s._view("visc", np.float64, 1))
    shear_rate = property(lambda s: s._view("shear_rate", np.float64, 1))

    def latticeBoltzmannFreeSurfaceStep(self):
        self.L.lbo_free_surface_step(self.h)

    def latticeBoltzmannCouplingStep(self, newNeighborList, elmts, parts, components):
        self.L.lbo_coupling_step(self.h, int(bool(newNeighborList)), _ptr(parts), len(parts), _ptr(elmts), len(elmts),
                                 _ptr(components))

    def latticeBolzmannStep(self, elmts=None, parts=None, components=None):
        nE = 0 if elmts is None else len(elmts)
        nP = 0 if parts is None else len(parts)
        F = np.zeros((nE, 3)); M = np.zeros((nE, 3)); V = np.zeros(nE); Wf = np.zeros((self.nWalls, 3))
        self.L.lbo_step(self.h, _ptr(parts), nP, _ptr(elmts), nE, _ptr(F), _ptr(M), _ptr(V), _ptr(Wf), self.nWalls)
        return F, M, V, Wf

    def neighbor(self, i, j):
        return self.L.lbo_neighbor(self.h, int(i), int(j))


# ---------------------------------------------------------------------------------------------
# ref_harness output readers
# ---------------------------------------------------------------------------------------------
def read_state(path):
    """Parse one PREFIX_stateNNNNNN.bin written by oracle/ref_harness.cpp::dumpState."""
    raw = np.memmap(path, dtype=np.uint8, mode="r")  # full-size dumps are several GB: mapped, not read
    assert bytes(raw[:7]) == b"HBDUMP2", bytes(raw[:8])
    hdr = np.frombuffer(raw, dtype="<u4", count=8, offset=8)
    X, Y, Z, step, nE, nW, withNb, nP = [int(v) for v in hdr]
    N = X * Y * Z
    off = 8 + 32
    out = dict(size=(X, Y, Z), step=step, nElmts=nE, nWalls=nW, nParts=nP)

    def take(dtype, count, shape=None):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=off)
        off += a.nbytes
        return a.reshape(shape) if shape else a
    out["type"] = take("u1", N)
    out["flags"] = take("u1", N)
    out["solidIndex"] = take("<u4", N)
    lite = bool(withNb & 2)
    withNb &= 1
    if not lite:
        out["f"] = take("<f8", 19 * N, (N, 19))
    out["fs"] = take("<f8", 19 * N, (N, 19))
    out["n"] = take("<f8", N)
    out["u"] = take("<f8", 3 * N, (N, 3))
    out["hydroForce"] = take("<f8", 3 * N, (N, 3))
    out["mass"] = take("<f8", N)
    out["visc"] = take("<f8", N)
    out["shearRate"] = take("<f8", N)
    if withNb:
        out["neighbors"] = take("<u4", 19 * N, (N, 19))
    if bytes(raw[off:off + 7]) == b"CURVES1":  # curved-wall cells and the mass target (trailer)
        off += 8
        nc = int(take("<u4", 1)[0])
        out["curve_cells"] = take("<u4", nc)
        out["curve_delta"] = take("<f8", 19 * nc, (nc, 19))
        out["totalMass"] = float(take("<f8", 1)[0])
    # combined byte as the oracle / CUDA engine use it: t | p<<4 | node<<5
    out["type_flags"] = (out["type"] | ((out["flags"] & 1) << 4) | ((out["flags"] & 2) << 4)).astype(np.uint8)
    return out


def read_log(path):
    """PREFIX_log.txt header -> params dict usable by Oracle / the CUDA engine."""
    hdr = {}
    with open(path) as fh:
        lines = [ln for ln in fh if ln.startswith("#")]
    toks = " ".join(ln[1:] for ln in lines).split()
    i = 0
    def nums(k, cnt, conv=float):
        j = toks.index(k)
        return [conv(t) for t in toks[j + 1:j + 1 + cnt]]
    hdr["size"] = nums("size", 3, int)
    hdr["nElmts"] = nums("elmts", 1, int)[0]
    hdr["nParts"] = nums("parts", 1, int)[0]
    hdr["nWalls"] = nums("walls", 1, int)[0]
    hdr["unitLength"], hdr["unitTime"], hdr["unitDensity"] = nums("unitLength", 1)[0], nums("unitTime", 1)[0], nums("unitDensity", 1)[0]
    hdr["lbF"] = nums("lbF", 3)
    hdr["initDynVisc"] = nums("initVisc", 1)[0]
    hdr["plasticVisc"] = nums("plasticVisc", 1)[0]
    hdr["yieldStress"] = nums("yieldStress", 1)[0]
    hdr["turbConst"] = nums("turbConst", 1)[0]
    hdr["slipCoefficient"] = nums("slip", 1)[0]
    hdr["initVelocity"] = nums("initVelocity", 3)
    hdr["boundary"] = nums("boundary", 6, int)
    j = toks.index("flags")
    hdr["freeSurface"], hdr["forceField"], hdr["nonNewtonian"], hdr["turbulence"] = \
        int(toks[j + 2]), int(toks[j + 4]), int(toks[j + 6]), int(toks[j + 8])
    if "totalMass" in toks:
        hdr["totalMass"] = nums("totalMass", 1)[0]
        hdr["enforceMass"] = nums("enforceMass", 1, int)[0]
    return hdr


def read_forces(path, nElmts, nWalls):
    """PREFIX_forces.bin -> per step (F[nE,3], M[nE,3], V[nE], wallF[nW,3])."""
    raw = np.fromfile(path, dtype="<f8")
    per = 7 * nElmts + 3 * nWalls
    if per == 0:
        return []
    steps = raw.size // per
    out = []
    for s in range(steps):
        blk = raw[s * per:(s + 1) * per]
        e = blk[:7 * nElmts].reshape(nElmts, 7)
        out.append((e[:, 0:3].copy(), e[:, 3:6].copy(), e[:, 6].copy(), blk[7 * nElmts:].reshape(nWalls, 3).copy()))
    return out


def read_particle_trace(path):
    """PREFIX_parts.bin -> per step (parts[PARTICLE_DTYPE], elmts[ELEMENT_DTYPE], components[u4], newNeighborList)."""
    with open(path, "rb") as fh:
        raw = fh.read()
    off = 0
    out = []
    while off < len(raw):
        nP, nE, flag = [int(v) for v in np.frombuffer(raw, "<u4", 3, off)]
        off += 12
        parts = np.zeros(nP, PARTICLE_DTYPE)
        for i in range(nP):
            d = np.frombuffer(raw, "<f8", 7, off); off += 56
            uu = np.frombuffer(raw, "<u4", 2, off); off += 8
            parts[i]["x0"] = d[0:3]; parts[i]["r"] = d[3]; parts[i]["radiusVec"] = d[4:7]
            parts[i]["clusterIndex"] = uu[0]; parts[i]["particleIndex"] = uu[1]
        elmts = np.zeros(nE, ELEMENT_DTYPE)
        comps = []
        for e in range(nE):
            d = np.frombuffer(raw, "<f8", 6, off); off += 48
            nc = int(np.frombuffer(raw, "<u4", 1, off)[0]); off += 4
            cc = np.frombuffer(raw, "<u4", nc, off); off += 4 * nc
            elmts[e]["x1"] = d[0:3]; elmts[e]["wGlobal"] = d[3:6]
            elmts[e]["compBegin"] = len(comps); comps.extend(int(c) for c in cc); elmts[e]["compEnd"] = len(comps)
        out.append((parts, elmts, np.asarray(comps, dtype=np.uint32), bool(flag)))
    return out
