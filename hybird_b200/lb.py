"""Host mirror of the reference's `LB` class (LB.h:32-215) on top of the C ABI.

Same method names, argument meaning and call order as the reference so that a driver written
against hybird's goCycle (hybird.cpp:35-66) reads the same:

    lb = LB(params); lb.latticeBolzmannInit(state)
    each cycle:  lb.latticeBoltzmannFreeSurfaceStep()            # if freeSurface
                 lb.latticeBoltzmannCouplingStep(newNeighborList, elmts, particles, components)
                 F, M, V, wallF = lb.latticeBolzmannStep(elmts, particles)

The free-surface and coupling calls only record the request; the device work of one cycle is
issued by latticeBolzmannStep through a single lbGpuStep call, in the reference's order.
Errors follow the reference's convention for this path (ASSERT -> message + exit(1), macros.h:8)
translated to Python: LbGpuError carrying the library's message.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi


class LB:
    def __init__(self, params: dict, device: int = -1):
        self.params = dict(params)
        self.lib = abi.load_library()
        self.P = abi.make_params(params, device)
        self.h = C.c_void_p()
        # cells in the host arrays: the whole lattice, or the planes of this handle's slabs plus one plane below and
        # above when the other slabs live in other processes (include/lbgpu.h, LbGpuParams)
        X, Y, Z = (int(v) for v in params["size"])
        G = int(params.get("nSlabs", 1))
        nLocal = int(params.get("nLocalSlabs", G))
        first = int(params.get("slabIndex", 0))
        if G > 1 and nLocal < G:
            inner = Z - 2
            zlo = 1 + (first * inner) // G
            zhi = 1 + ((first + nLocal) * inner) // G
            self.planes = (zlo - 1, zhi + 1)
        else:
            self.planes = (0, Z)
        self.N = X * Y * (self.planes[1] - self.planes[0])
        self.nWalls = int(params.get("nWalls", 0))
        self.freeSurface = bool(params["freeSurface"])
        self._fs_requested = False
        self._couple = None
        self._last = (np.zeros(0, abi_particle_dtype()), np.zeros(0, abi_element_dtype()), np.zeros(0, np.uint32))
        self.time = 0

    # -- LB::latticeBolzmannInit (LB.h:158) -----------------------------------------------------
    def latticeBolzmannInit(self, type_flags, solidIndex, n, u, mass, visc, f=None):
        a = lambda x, dt: np.ascontiguousarray(x, dtype=dt)
        tf, si = a(type_flags, np.uint8), a(solidIndex, np.uint32)
        n_, u_, m_, v_ = a(n, np.float64), a(u, np.float64), a(mass, np.float64), a(visc, np.float64)
        f_ = None if f is None else a(f, np.float64)
        for arr, cnt in ((tf, self.N), (si, self.N), (n_, self.N), (u_, 3 * self.N), (m_, self.N), (v_, self.N)):
            if arr.size != cnt:
                raise ValueError("latticeBolzmannInit: array of %d elements, expected %d" % (arr.size, cnt))
        if f_ is not None and f_.size != abi.Q * self.N:
            raise ValueError("latticeBolzmannInit: f must hold 19*N values")
        abi.check(self.lib.lbGpuInit(C.byref(self.P), abi.ptr(tf), abi.ptr(si), abi.ptr(f_), abi.ptr(n_), abi.ptr(u_),
                                     abi.ptr(m_), abi.ptr(v_), C.byref(self.h)))
        return self

    # -- the same, built on the device for a box problem (lbGpuInitBox) ------------------------------
    def latticeBolzmannInitBox(self, case: dict, particles=None):
        """`case`: hybird configuration keys + set-up options as in hybird_b200.workloads (what
        lattice_init.build_state restates on the host)."""
        from . import lattice_init as li
        regions = li.regions_from_case(case, self.params)
        reg = np.zeros(len(regions), dtype=np.dtype([("kind", "<i4"), ("gasInside", "<i4"), ("a", "<f8", 6)], align=True))
        for k, (kind, inside, a) in enumerate(regions):
            reg[k]["kind"] = kind; reg[k]["gasInside"] = inside; reg[k]["a"][:len(a)] = a
        iv = np.ascontiguousarray(self.params.get("initVelocity", [0.0, 0.0, 0.0]), dtype=np.float64)
        wv = np.zeros((6, 3))
        walls = li.make_walls(self.params, case.get("wall_vel", ()))
        for w in walls:
            wv[2 * w.axis + w.side] = w.vel
        parts = None if particles is None or len(particles) == 0 else np.ascontiguousarray(particles)
        abi.check(self.lib.lbGpuInitBox(C.byref(self.P), abi.ptr(iv), abi.ptr(wv), abi.ptr(reg) if len(reg) else None, len(reg),
                                        abi.ptr(parts), 0 if parts is None else len(parts), C.byref(self.h)))
        return self

    # -- LB::curves (LB.h:63-64) as LB::initializeCurved left them; LB::totalMass for problemName DRUM -------
    def setCurves(self, cells, delta):
        cells = np.ascontiguousarray(cells, dtype=np.uint32)
        delta = np.ascontiguousarray(delta, dtype=np.float64)
        if delta.size != abi.Q * cells.size:
            raise ValueError("setCurves: delta must hold 19 values per curved cell")
        abi.check(self.lib.lbGpuSetCurves(self.h, cells.size, abi.ptr(cells), abi.ptr(delta)))
        return self

    def setMassTarget(self, totalMass):
        abi.check(self.lib.lbGpuSetMassTarget(self.h, float(totalMass)))
        return self

    # -- LB::latticeBoltzmannFreeSurfaceStep (LB.h:161) ------------------------------------------
    def latticeBoltzmannFreeSurfaceStep(self):
        self._fs_requested = True

    # -- LB::latticeBoltzmannCouplingStep (LB.h:160) ---------------------------------------------
    def latticeBoltzmannCouplingStep(self, newNeighborList, elmts, particles, components):
        """Returns False: the reference resets dem.newNeighborList through its bool& (LB.cpp:258)."""
        self._couple = (bool(newNeighborList), np.ascontiguousarray(particles), np.ascontiguousarray(elmts),
                        np.ascontiguousarray(components, dtype=np.uint32))
        return False

    # -- LB::latticeBolzmannStep (LB.h:159) ------------------------------------------------------
    def latticeBolzmannStep(self, elmts=None, particles=None, fetch_forces=True, components=None):
        """`elmts`, `particles` as the reference passes them to LB::latticeBolzmannStep; they only matter when no
        coupling step was requested this cycle (demSolve = 0: the flags of the initialisation stay, the direct forcing
        of LB::computeHydroForces still acts on them) -- `components` = the flattened elmts[].components then (they
        include the periodic ghost particles, DEM.cpp:1658, so there is no default; the ones of the last coupling
        step are reused when there was one)."""
        fs = int(self._fs_requested and self.freeSurface)
        if self._couple is not None:
            rescan, parts, els, comps = self._couple
            self._last = (parts, els, comps)
            abi.check(self.lib.lbGpuStep(self.h, fs, 1, int(rescan), abi.ptr(parts), len(parts), abi.ptr(els), len(els),
                                         abi.ptr(comps), len(comps)))
        elif particles is not None and len(particles) and elmts is not None and len(elmts):
            parts, els = np.ascontiguousarray(particles), np.ascontiguousarray(elmts)
            if components is None:
                if not len(self._last[2]):
                    raise ValueError("latticeBolzmannStep: particles without a coupling step need `components`")
                components = self._last[2]
            comps = np.ascontiguousarray(components, dtype=np.uint32)
            self._last = (parts, els, comps)
            abi.check(self.lib.lbGpuStep(self.h, fs, 0, 0, abi.ptr(parts), len(parts), abi.ptr(els), len(els),
                                         abi.ptr(comps), len(comps)))
        else:
            abi.check(self.lib.lbGpuStep(self.h, fs, 0, 0, None, 0, None, 0, None, 0))
        self._fs_requested = False
        self._couple = None
        self.time += 1
        if not fetch_forces:
            return None
        return self.forces()

    def forces(self):
        nE = len(self._last[1])
        F = np.zeros((nE, 3)); M = np.zeros((nE, 3)); V = np.zeros(nE); W = np.zeros((self.nWalls, 3))
        abi.check(self.lib.lbGpuParticleForces(self.h, abi.ptr(F) if nE else None, abi.ptr(M) if nE else None,
                                               abi.ptr(V) if nE else None, abi.ptr(W) if self.nWalls else None))
        return F, M, V, W

    def run(self, steps, free_surface=None):
        """`steps` cycles issued back to back (no host round trip); the particles of the last coupling step, if any,
        stay resident and fixed (lbGpuRun)."""
        fs = self.freeSurface if free_surface is None else free_surface
        abi.check(self.lib.lbGpuRun(self.h, int(fs), int(steps)))
        self.time += int(steps)

    # -- DEM::discreteElementStep on the device (lbGpuDem*): single-sphere elements, plane walls ---------------
    def demInit(self, dem: dict):
        """`dem`: dict(params=..., elmts=[...], walls=[...]) with the reference's DEM member names (what
        DEM::discreteElementInit left: sphereMat constants, deltat, multiStep, nebrRange, maxDisp; per element x0, x1, w0,
        radius, m, I; per wall n, p, vel, omega, rotCenter, moving) -- physical units."""
        p = dem["params"]
        pbcs = dem.get("pbcs") or []
        if pbcs and any(int(e.get("size", 1)) != 1 for e in dem["elmts"]):
            raise ValueError("demInit: periodic DEM boundaries are covered for single spheres only; keep the host DEM and "
                             "latticeBoltzmannCouplingStep / latticeBolzmannStep")
        P = abi.LbGpuDemParams()
        P.nPbc = len(pbcs)
        for b, pb in enumerate(pbcs):
            for q in range(3):
                P.pbcP[b][q] = float(pb["p"][q]); P.pbcV[b][q] = float(pb["v"][q])
        P.contactModel = int(p["contactModel"]); P.multiStep = int(p["multiStep"])
        for k in ("knConst", "ksConst", "dampCoeff", "viscTang", "linearStiff", "frictionCoefPart", "frictionCoefWall", "numVisc",
                  "deltat", "nebrRange", "maxDisp"):
            setattr(P, k, float(p[k]))
        P.demF[:] = [float(v) for v in p["demF"]]
        E = np.zeros(len(dem["elmts"]), dtype=abi.DEM_ELEMENT_DTYPE)
        for k, e in enumerate(dem["elmts"]):
            if not 1 <= int(e.get("size", 1)) <= 4:
                raise ValueError("demInit: elements are spheres or clusters of 2-4 spheres (element %d has size %d)" % (k, e["size"]))
            for f in ("x0", "x1", "w0", "I"):
                E[k][f] = e[f]
            E[k]["radius"] = e["radius"]; E[k]["m"] = e["m"]; E[k]["size"] = int(e.get("size", 1))
        W = np.zeros(len(dem["walls"]), dtype=abi.DEM_WALL_DTYPE)
        for k, w in enumerate(dem["walls"]):
            for f in ("n", "p", "vel", "omega", "rotCenter"):
                W[k][f] = w[f]
            W[k]["moving"] = int(w["moving"])
        abi.check(self.lib.lbGpuDemInit(self.h, C.byref(P), abi.ptr(E), len(E), abi.ptr(W) if len(W) else None, len(W)))
        self._dem_n = len(E)
        nP = int(E["size"].sum()) * (7 if pbcs else 1)
        self._last = (np.zeros(nP, abi_particle_dtype()), np.zeros(len(E), abi_element_dtype()), np.arange(nP, dtype=np.uint32))
        return self

    def demStep(self, hydro=None):
        """dem.discreteElementStep(); `hydro` (nElmts x 7: FHydro, MHydro, -) replaces the device's own forces (tests)."""
        hy = None if hydro is None else np.ascontiguousarray(hydro, dtype=np.float64)
        if hy is not None and hy.size != 7 * self._dem_n:
            raise ValueError("demStep: hydro must hold 7 values per element")
        abi.check(self.lib.lbGpuDemStep(self.h, abi.ptr(hy)))

    def runDem(self, steps, free_surface=None):
        """`steps` goCycles on the device: DEM step, free-surface step, coupling step, LB step (lbGpuRunDem)."""
        fs = self.freeSurface if free_surface is None else free_surface
        abi.check(self.lib.lbGpuRunDem(self.h, int(fs), int(steps)))
        self.time += int(steps)

    def demState(self):
        n = self._dem_n
        x0 = np.zeros((n, 3)); x1 = np.zeros((n, 3)); w0 = np.zeros((n, 3)); info = (C.c_double * 3)()
        abi.check(self.lib.lbGpuDemState(self.h, abi.ptr(x0), abi.ptr(x1), abi.ptr(w0), C.byref(info)))
        return dict(x0=x0, x1=x1, w0=w0, maxDisp=float(info[0]), rebuilds=int(info[1]), longest_list=int(info[2]))

    def demParticles(self):
        """The spheres of the elements (particle::updateCorrected): x0, radiusVec, clusterIndex."""
        n = C.c_uint32()
        abi.check(self.lib.lbGpuDemParticles(self.h, C.byref(n), None, None, None))
        x0 = np.zeros((n.value, 3)); rv = np.zeros((n.value, 3)); ci = np.zeros(n.value, dtype=np.uint32)
        abi.check(self.lib.lbGpuDemParticles(self.h, C.byref(n), abi.ptr(x0), abi.ptr(rv), abi.ptr(ci)))
        return dict(x0=x0, radiusVec=rv, clusterIndex=ci)

    def demContacts(self):
        """elmt::FParticle, FWall, MParticle, MWall of the last DEM sub-step."""
        n = self._dem_n
        out = [np.zeros((n, 3)) for _ in range(4)]
        abi.check(self.lib.lbGpuDemContacts(self.h, *[abi.ptr(a) for a in out]))
        return dict(FParticle=out[0], FWall=out[1], MParticle=out[2], MWall=out[3])

    def synchronize(self):
        abi.check(self.lib.lbGpuSynchronize(self.h))

    def last_step_ms(self):
        ms = C.c_float()
        abi.check(self.lib.lbGpuLastStepMs(self.h, C.byref(ms)))
        return float(ms.value)

    def last_kernel_ms(self):
        """(sum of fused-kernel device times, number of launches) of the last run()/step call."""
        ms, n = C.c_float(), C.c_uint32()
        abi.check(self.lib.lbGpuLastKernelMs(self.h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def launch_count(self):
        v = C.c_uint64()
        abi.check(self.lib.lbGpuLaunchCount(self.h, C.byref(v)))
        return int(v.value)

    PHASES = ("dem", "list_build", "free_surface_update", "coupling", "step_kernels_and_halo", "wall_slots_and_sums", "element_forces_and_type_sync")

    def phase_trace(self, on=True):
        abi.check(self.lib.lbGpuPhaseTrace(self.h, int(bool(on))))

    def phase_ms(self):
        """Mean device time per cycle and phase over the last (<= 64) cycles since phase_trace(True)."""
        ms, n = (C.c_float * 7)(), C.c_uint32()
        abi.check(self.lib.lbGpuPhaseMs(self.h, C.byref(ms), C.byref(n)))
        return {k: float(ms[i]) for i, k in enumerate(self.PHASES)}, int(n.value)

    def graph_info(self):
        """(captures, replays) of the CUDA graph lbGpuRun uses for free-surface cycles without particles."""
        v = (C.c_uint64 * 2)()
        abi.check(self.lib.lbGpuGraphInfo(self.h, C.byref(v)))
        return int(v[0]), int(v[1])

    def counts(self, local=False):
        c = (C.c_uint64 * 4)()
        abi.check((self.lib.lbGpuCountsLocal if local else self.lib.lbGpuCounts)(self.h, C.byref(c)))
        return dict(fluid=int(c[0]), interface=int(c[1]), particle=int(c[2]), steps=int(c[3]))

    def peer_halo(self):
        """True when the per-step halo of this handle goes through peer memory (lbGpuPeerHalo)."""
        on = C.c_int32()
        abi.check(self.lib.lbGpuPeerHalo(self.h, C.byref(on)))
        return bool(on.value)

    # -- what IO reads from lb.types / lb.nodes (IO.cpp:698-831) ---------------------------------
    @staticmethod
    def host_mirrors(N, fields):
        """Host arrays lbGpuFetchFields fills, allocated and touched once (the reference keeps lb.types / lb.nodes for
        the whole run; a caller that exports repeatedly reuses its mirrors)."""
        shapes = dict(type_flags=((N,), np.uint8), solidIndex=((N,), np.uint32), n=((N,), np.float64),
                      u=((N, 3), np.float64), mass=((N,), np.float64), visc=((N,), np.float64),
                      shearRate=((N,), np.float64), hydroForce=((N, 3), np.float64), f=((N, abi.Q), np.float64))
        out = {}
        for k in fields:
            out[k] = np.empty(shapes[k][0], shapes[k][1])
            out[k].fill(0)
        return out

    def fetch(self, fields=("type_flags", "solidIndex", "n", "u", "mass", "visc", "shearRate", "hydroForce", "f"), out=None):
        if out is None:
            out = self.host_mirrors(self.N, fields)
        else:
            out = {k: out[k] for k in fields}
        order = ("type_flags", "solidIndex", "n", "u", "mass", "visc", "shearRate", "hydroForce", "f")
        abi.check(self.lib.lbGpuFetchFields(self.h, *[abi.ptr(out.get(k)) for k in order]))
        return out

    # -- IO's screen export and fluid file without a full-field fetch (IO.cpp:698-895, 969-999) ------------
    def summary(self):
        """max |u| (lattice units), total fluid mass outside particles, active cells, per cent plastic."""
        o = (C.c_double * 4)()
        abi.check(self.lib.lbGpuFluidSummary(self.h, C.byref(o)))
        act = float(o[2])
        return dict(max_speed=float(o[0]) ** 0.5, mass=float(o[1]), active=int(act),
                    plastic_pct=100.0 * float(o[3]) / act if act else float("nan"))

    def write_vti(self, path, dem_solve=True):
        abi.check(self.lib.lbGpuWriteVti(self.h, str(path).encode(), int(bool(dem_solve))))

    # -- checkpoint / restart (no counterpart in the reference) --------------------------------------
    def save_state(self) -> np.ndarray:
        n = C.c_uint64()
        abi.check(self.lib.lbGpuStateBytes(self.h, C.byref(n)))
        blob = np.empty(int(n.value), dtype=np.uint8)
        abi.check(self.lib.lbGpuSaveState(self.h, abi.ptr(blob), n))
        return blob

    def load_state(self, blob, last_particles=None):
        """`last_particles` = (parts, elmts, comps) the saved run coupled last, if forces() is to be called before the
        next coupling step."""
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        abi.check(self.lib.lbGpuLoadState(self.h, abi.ptr(blob), C.c_uint64(blob.size)))
        if last_particles is not None:
            self._last = last_particles
        return self

    def close(self):
        if self.h:
            self.lib.lbGpuFinalize(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def abi_particle_dtype():
    from .lattice_init import PARTICLE_DTYPE
    return PARTICLE_DTYPE


def abi_element_dtype():
    from .lattice_init import ELEMENT_DTYPE
    return ELEMENT_DTYPE
