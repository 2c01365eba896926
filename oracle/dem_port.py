"""Plain-Python restatement of the reference's DEM sub-step for single-sphere elements -- TEST INFRASTRUCTURE (the checker
of the device-side DEM, pinned against the particle traces the unmodified reference records; never imported by the
product).  Follows DEM::discreteElementStep (DEM.cpp:331-376): evalMaxDisp + the neighbour-table trigger (DEM.cpp:1314-1324,
340-346), elmt::predict (elmt.cpp:139-177), particle-particle and wall-particle contacts (DEM.cpp:1668-1717, 1801-1982) with
the LINEAR / HERTZIAN laws (DEM.cpp:2138-2224), Newton's equations (DEM.cpp:1150-1181), elmt::correct (elmt.cpp:179-254).
Quirks kept: contacts are only looked for among the pairs DEM::evalNeighborTable listed at the last rebuild (centres
closer than nebrRange THEN, DEM.cpp:1326-1494; the linked cells only bound the search and are not restated: every cell is
wider than nebrRange, so the table is exactly the set of pairs within range); DEM::evalNearWallTable lists only the FIRST
wall within nebrRange of a particle, at table-rebuild time (DEM.cpp:1496-1513); spheres never rotate their frame (q2 is only integrated for size > 1), so the body frame is the
global one.  No periodic boundaries, cylinders, objects or clusters."""
from __future__ import annotations

import math

import numpy as np


def read_dem_file(path):
    """PREFIX_dem.txt written by oracle/ref_harness.cpp -> dict(params, elmts, walls)."""
    return parse_dem_text(open(path).read())


def parse_dem_text(text):
    prm, elmts, walls, counts, pbcs = {}, [], [], {}, []
    for ln in str(text).splitlines():
        t = ln.split()
        if not t:
            continue
        if t[0] == "contactModel":
            i = 0
            while i < len(t):
                k = t[i]
                if k == "demF":
                    prm[k] = [float(v) for v in t[i + 1:i + 4]]; i += 4
                else:
                    prm[k] = float(t[i + 1]); i += 2
        elif t[0] == "elmt":
            elmts.append(dict(size=int(t[3]), radius=float(t[5]), m=float(t[7]), I=[float(v) for v in t[9:12]],
                              x0=[float(v) for v in t[13:16]], x1=[float(v) for v in t[17:20]], w0=[float(v) for v in t[21:24]]))
        elif t[0] == "wall":
            walls.append(dict(n=[float(v) for v in t[3:6]], p=[float(v) for v in t[7:10]], vel=[float(v) for v in t[11:14]],
                              omega=[float(v) for v in t[15:18]], rotCenter=[float(v) for v in t[19:22]], moving=int(t[23])))
        elif t[0] == "pbc":
            pbcs.append(dict(p=[float(v) for v in t[3:6]], v=[float(v) for v in t[7:10]]))
        elif t[0] == "pbcs":
            counts = {t[i]: int(t[i + 1]) for i in range(0, len(t), 2)}
    prm["contactModel"] = int(prm["contactModel"]); prm["multiStep"] = int(prm["multiStep"])
    return dict(params=prm, elmts=elmts, walls=walls, counts=counts, pbcs=pbcs)


def _cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


def _norm2(a):
    return a[0] * a[0] + a[1] * a[1] + a[2] * a[2]


class DemPort:
    def __init__(self, dem):
        p = dem["params"]
        self.p = p
        self.n = len(dem["elmts"])
        E = dem["elmts"]
        assert all(e["size"] == 1 for e in E), "single-sphere elements only"
        z = lambda: np.zeros((self.n, 3))
        self.x = [np.array([e["x0"] for e in E], dtype=float).reshape(-1, 3), np.array([e["x1"] for e in E], dtype=float).reshape(-1, 3), z(), z(), z(), z()]
        self.xp = [self.x[0].copy(), self.x[1].copy(), z(), z(), z(), z()]
        self.w = [np.array([e["w0"] for e in E], dtype=float).reshape(-1, 3), z(), z(), z(), z(), z()]
        self.wp = [z(), z(), z(), z(), z(), z()]
        self.radius = np.array([e["radius"] for e in E]); self.m = np.array([e["m"] for e in E])
        self.I = np.array([e["I"] for e in E]).reshape(-1, 3)
        self.walls = dem["walls"]
        self.maxDisp = p["maxDisp"]; self.nebrRange = p["nebrRange"]
        self.nearWall = [-1] * self.n
        self.pairs = []  # DEM::neighborTable as (i, j), i < j
        dt = p["deltat"]
        self.c = [dt, dt * dt / 2.0, dt * dt * dt / 6.0, dt * dt * dt * dt / 24.0, dt * dt * dt * dt * dt / 120.0]
        g1 = [95.0 / 288.0, 1.0, 25.0 / 24.0, 35.0 / 72.0, 5.0 / 48.0, 1.0 / 120.0]
        g2 = [3.0 / 16.0, 251.0 / 360.0, 1.0, 11.0 / 18.0, 1.0 / 6.0, 1.0 / 60.0]
        c = self.c
        self.coeff1 = [g1[0] * c[0], g1[1] * c[0] / c[0], g1[2] * c[0] / c[1], g1[3] * c[0] / c[2], g1[4] * c[0] / c[3], g1[5] * c[0] / c[4]]
        self.coeff2 = [g2[0] * c[1], g2[1] * c[1] / c[0], g2[2] * c[1] / c[1], g2[3] * c[1] / c[2], g2[4] * c[1] / c[3], g2[5] * c[1] / c[4]]
        self.FHydro = z(); self.MHydro = z()
        self.rebuilds = 0

    # -- contact laws (DEM.cpp:2138-2224) ---------------------------------------------------------------
    def normal(self, overlap, vreln, effRad, effMass):
        p = self.p
        if p["contactModel"] == 1:
            kn = p["knConst"] * math.sqrt(effRad) * math.sqrt(overlap)
            gamman = 2.0 * p["dampCoeff"] * math.sqrt(kn * effMass)
            return max(kn * overlap + (-gamman * vreln), 0.0)
        gamman = 2.0 * p["dampCoeff"] * math.sqrt(p["linearStiff"] * effMass)
        return max(p["linearStiff"] * overlap + (-gamman * vreln), 0.0)

    def tangential(self, vrelt, fn, effRad, effMass, friction):
        p = self.p
        if p["contactModel"] == 1:
            ks = p["ksConst"] * math.sqrt(effRad) * math.pow(abs(fn), 1.0 / 3.0)
        else:
            ks = p["linearStiff"]
        fsMax = friction * fn
        gammas = 2.0 * p["viscTang"] * math.sqrt(effMass * ks)
        return min(gammas * vrelt, fsMax)

    def substep(self):
        p, n, c = self.p, self.n, self.c
        x, xp, w, wp = self.x, self.xp, self.w, self.wp
        # evalMaxDisp + trigger
        maxVel = 0.0
        for k in range(n):
            v2 = _norm2(x[1][k])
            if v2 > maxVel:
                maxVel = v2
        self.maxDisp += math.sqrt(maxVel) * p["deltat"]
        if self.maxDisp > 0.25 * self.nebrRange:
            self.maxDisp = 0.0
            self.rebuilds += 1
            r2 = self.nebrRange * self.nebrRange
            self.pairs = [(i, j) for i in range(n) for j in range(i + 1, n) if _norm2(x[0][j] - x[0][i]) < r2]
            for k in range(n):  # evalNearWallTable: the first wall within range, at the corrected position
                self.nearWall[k] = -1
                for wi, wl in enumerate(self.walls):
                    if float(np.dot(np.array(wl["n"]), x[0][k] - np.array(wl["p"]))) < self.nebrRange:
                        self.nearWall[k] = wi
                        break
        # predictor
        xp[0] = x[0] + x[1] * c[0] + x[2] * c[1] + x[3] * c[2] + x[4] * c[3] + x[5] * c[4]
        xp[1] = x[1] + x[2] * c[0] + x[3] * c[1] + x[4] * c[2] + x[5] * c[3]
        xp[2] = x[2] + x[3] * c[0] + x[4] * c[1] + x[5] * c[2]
        xp[3] = x[3] + x[4] * c[0] + x[5] * c[1]
        xp[4] = x[4] + x[5] * c[0]
        xp[5] = x[5].copy()
        wp[0] = w[0] + w[1] * c[0] + w[2] * c[1] + w[3] * c[2] + w[4] * c[3] + w[5] * c[4]
        wp[1] = w[1] + w[2] * c[0] + w[3] * c[1] + w[4] * c[2] + w[5] * c[3]
        wp[2] = w[2] + w[3] * c[0] + w[4] * c[1] + w[5] * c[2]
        wp[3] = w[3] + w[4] * c[0] + w[5] * c[1]
        wp[4] = w[4] + w[5] * c[0]
        wp[5] = w[5].copy()
        FP, FW, MP, MW = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
        # particle-particle contacts at the predicted positions
        for i, j in self.pairs:
            if True:
                d = xp[0][j] - xp[0][i]
                sig = self.radius[i] + self.radius[j]
                if _norm2(d) < sig * sig:
                    dist = math.sqrt(_norm2(d))
                    overlap = self.radius[i] + self.radius[j] - dist
                    relVel = xp[1][j] - xp[1][i]
                    en = d / dist
                    vn = float(np.dot(relVel, en))
                    normalRelVel = en * vn
                    effMass = self.m[i] * self.m[j] / (self.m[i] + self.m[j])
                    effRad = self.radius[i] * self.radius[j] / (self.radius[i] + self.radius[j])
                    fn = self.normal(overlap, vn, effRad, effMass)
                    nf = en * fn
                    vecRadI, vecRadJ = self.radius[i] * en, -self.radius[j] * en
                    FP[i] = FP[i] - nf; FP[j] = FP[j] + nf
                    relC = relVel - _cross(wp[0][i], vecRadI) + _cross(wp[0][j], vecRadJ)
                    tang = relC - normalRelVel
                    nt = math.sqrt(_norm2(tang))
                    if nt != 0.0:
                        ft = self.tangential(nt, fn, effRad, effMass, p["frictionCoefPart"])
                        et = tang / nt
                        tf = ft * et
                        MP[i] = MP[i] + _cross(vecRadI, tf); FP[i] = FP[i] + tf
                        MP[j] = MP[j] - _cross(vecRadJ, tf); FP[j] = FP[j] - tf
        # wall contacts (one listed wall per particle)
        for k in range(n):
            wi = self.nearWall[k]
            if wi < 0:
                continue
            wl = self.walls[wi]
            en = np.array(wl["n"])
            dist = float(np.dot(en, xp[0][k] - np.array(wl["p"])))
            overlap = self.radius[k] - dist
            if overlap > 0.0:
                cpv = np.zeros(3)
                if wl["moving"]:
                    dc = xp[0][k] - np.array(wl["rotCenter"])
                    cpv = np.array(wl["vel"]) + _cross(np.array(wl["omega"]), dc - float(np.dot(dc, en)) * en)
                relVel = xp[1][k] - cpv
                vn = float(np.dot(relVel, en))
                normalRelVel = en * vn
                fn = self.normal(2.0 * overlap, vn, self.radius[k], self.m[k])
                nf = en * fn
                vecRadJ = -self.radius[k] * en
                FW[k] = FW[k] + nf
                relC = relVel + _cross(wp[0][k], vecRadJ)
                tang = relC - normalRelVel
                nt = math.sqrt(_norm2(tang))
                if nt != 0.0:
                    ft = self.tangential(nt, fn, self.radius[k], self.m[k], p["frictionCoefWall"])
                    et = tang / math.sqrt(_norm2(tang))
                    tf = ft * et
                    MW[k] = MW[k] - _cross(vecRadJ, tf)
                    FW[k] = FW[k] - tf
        # Newton (DEM.cpp:1150-1181) + corrector
        demF = np.array(p["demF"])
        for k in range(n):
            FVisc = -6.0 * math.pi * p["numVisc"] * self.radius[k] * xp[1][k]
            MVisc = -8.0 * math.pi * p["numVisc"] * self.radius[k] * self.radius[k] * self.radius[k] * wp[0][k]
            x[2][k] = (FVisc + self.FHydro[k] + FP[k] + FW[k]) / self.m[k] + demF
            mom = MVisc + self.MHydro[k] + MP[k] + MW[k]
            I, wl_ = self.I[k], wp[0][k]
            w[1][k] = np.array([(mom[0] + (I[1] - I[2]) * wl_[1] * wl_[2]) / I[0], (mom[1] + (I[2] - I[0]) * wl_[2] * wl_[0]) / I[1],
                                (mom[2] + (I[0] - I[1]) * wl_[0] * wl_[1]) / I[2]])
        c2, c1 = self.coeff2, self.coeff1
        x2c = x[2] - xp[2]
        x[0] = xp[0] + x2c * c2[0]; x[1] = xp[1] + x2c * c2[1]
        x[3] = xp[3] + x2c * c2[3]; x[4] = xp[4] + x2c * c2[4]; x[5] = xp[5] + x2c * c2[5]
        for k in range(6):
            xp[k] = x[k].copy()
        w1c = w[1] - wp[1]
        w[0] = wp[0] + w1c * c1[0]
        w[2] = wp[2] + w1c * c1[2]; w[3] = wp[3] + w1c * c1[3]; w[4] = wp[4] + w1c * c1[4]; w[5] = wp[5] + w1c * c1[5]
        for k in range(6):
            wp[k] = w[k].copy()

    def step(self, FHydro, MHydro):
        """One DEM::discreteElementStep with the hydrodynamic forces of the last LB step (physical units)."""
        self.FHydro = np.asarray(FHydro, dtype=float).reshape(-1, 3); self.MHydro = np.asarray(MHydro, dtype=float).reshape(-1, 3)
        for _ in range(self.p["multiStep"]):
            self.substep()
        return self.x[0].copy(), self.x[1].copy(), self.w[0].copy()


# ---------------------------------------------------------------------------------------------------------------------
# Elements of several spheres (clusters, size 2-4): the general restatement.  Adds to the above the orientation
# quaternions q0..q5 of elmt::predict / elmt::correct (elmt.cpp:139-254), the particles of an element at
# e.x0 + r * project(prototype, q0) (particle::updatePredicted / updateCorrected, elmt.cpp:261-290; prototypes of
# DEM::compositeProperties, DEM.cpp:404-433), per-particle neighbour and wall tables, the lever arms of the contact forces
# (centerDist, DEM.cpp:1841-1890, 1936-1976) and the rotation in the body frame (DEM.cpp:1164-1180; project / newtonAcc /
# quatAcc of vector.cpp:481-504).  TEST INFRASTRUCTURE like everything in this file.
# ---------------------------------------------------------------------------------------------------------------------
def _qmul(q, r):
    """quaternion::multiply (vector.cpp:301-310): q.multiply(r)."""
    return np.array([r[0] * q[0] - r[1] * q[1] - r[2] * q[2] - r[3] * q[3],
                     r[0] * q[1] + r[1] * q[0] - r[2] * q[3] + r[3] * q[2],
                     r[0] * q[2] + r[1] * q[3] + r[2] * q[0] - r[3] * q[1],
                     r[0] * q[3] - r[1] * q[2] + r[2] * q[1] + r[3] * q[0]])


def _qadj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def _qnormalize(q):
    n = math.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])
    return np.array([q[0] / n, q[1] / n, q[2] / n, q[3] / n])


def _project(v, q):
    """project(vec, quat) = quat2vec(q (0, v) q*) (vector.cpp:481-491)."""
    rot = _qmul(_qmul(q, np.array([0.0, v[0], v[1], v[2]])), _qadj(q))
    return np.array([rot[1], rot[2], rot[3]])


def prototypes():
    s3, s2, s6 = math.sqrt(3), math.sqrt(2), math.sqrt(6)
    return {1: [np.zeros(3)],
            2: [np.array([0.5, 0.0, 0.0]), np.array([-0.5, 0.0, 0.0])],
            3: [np.array([0.0, 1.0, 0.0]), np.array([-s3 / 2, -0.5, 0.0]), np.array([s3 / 2, -0.5, 0.0])],  # C++: -1/2 is integer 0!
            4: [np.array([0.0, 0.0, 1.0]), np.array([0.0, 2.0 * s2 / 3.0, -1.0 / 3.0]),
                np.array([2.0 * s6 / 6.0, -2.0 * s2 / 6.0, -1.0 / 3.0]), np.array([-2.0 * s6 / 6.0, -2.0 * s2 / 6.0, -1.0 / 3.0])]}


class DemPortClusters(DemPort):
    def __init__(self, dem):
        E = dem["elmts"]
        sizes = [int(e["size"]) for e in E]
        saved = [e["size"] for e in E]
        for e in E:
            e["size"] = 1
        try:
            super().__init__(dem)
        finally:
            for e, s in zip(E, saved):
                e["size"] = s
        n = self.n
        self.size = sizes
        self.proto = prototypes()
        # DEM.cpp:423-424 writes -1/2 with integers: the y component of the second and third sphere of a triangle is 0
        self.proto[3] = [np.array([0.0, 1.0, 0.0]), np.array([-math.sqrt(3) / 2, 0.0, 0.0]), np.array([math.sqrt(3) / 2, 0.0, 0.0])]
        ident = np.array([1.0, 0.0, 0.0, 0.0])
        self.q = [np.tile(ident, (n, 1))] + [np.zeros((n, 4)) for _ in range(5)]
        self.qp = [self.q[0].copy()] + [np.zeros((n, 4)) for _ in range(5)]
        self.wGlobal = self.w[0].copy(); self.wLocal = self.w[0].copy(); self.wpGlobal = self.w[0].copy(); self.wpLocal = self.w[0].copy()
        # particles (DEM.cpp:236-245, elmt::generateParticles)
        self.cluster, self.pproto = [], []
        for k in range(n):
            for i in range(sizes[k]):
                self.cluster.append(k); self.pproto.append(i)
        self.P = len(self.cluster)
        self.comp = [[p for p in range(self.P) if self.cluster[p] == k] for k in range(n)]
        self.px0 = np.zeros((self.P, 3)); self.px1 = np.zeros((self.P, 3)); self.prv = np.zeros((self.P, 3))
        self.pr = np.array([self.radius[k] for k in self.cluster])
        self._update_particles(corrected=True)
        self.nearWallP = [-1] * self.P
        self.ppairs = []

    def _update_particles(self, corrected):
        x0 = self.x[0] if corrected else self.xp[0]
        x1 = self.x[1] if corrected else self.xp[1]
        q0 = self.q[0] if corrected else self.qp[0]
        wg = self.wGlobal if corrected else self.wpGlobal
        for p in range(self.P):
            k = self.cluster[p]
            self.px0[p] = x0[k]; self.prv[p] = 0.0; self.px1[p] = x1[k]
            if self.size[k] > 1:
                self.px0[p] = x0[k] + self.pr[p] * _project(self.proto[self.size[k]][self.pproto[p]], q0[k])
                self.prv[p] = self.px0[p] - x0[k]
                self.px1[p] = x1[k] + _cross(wg[k], self.prv[p])

    def substep(self):
        p, n, c = self.p, self.n, self.c
        x, xp, w, wp, q, qp = self.x, self.xp, self.w, self.wp, self.q, self.qp
        maxVel = 0.0
        for k in range(n):
            maxVel = max(maxVel, _norm2(x[1][k]))
        self.maxDisp += math.sqrt(maxVel) * p["deltat"]
        if self.maxDisp > 0.25 * self.nebrRange:
            self.maxDisp = 0.0
            self.rebuilds += 1
            r2 = self.nebrRange * self.nebrRange
            self.ppairs = [(a, b) for a in range(self.P) for b in range(a + 1, self.P)
                           if self.cluster[a] != self.cluster[b] and _norm2(self.px0[b] - self.px0[a]) < r2]
            for a in range(self.P):
                self.nearWallP[a] = -1
                for wi, wl in enumerate(self.walls):
                    if float(np.dot(np.array(wl["n"]), self.px0[a] - np.array(wl["p"]))) < self.nebrRange:
                        self.nearWallP[a] = wi
                        break
        # predictor: positions and spins as for spheres, plus the quaternions (c2 coefficients) and the frames
        xp[0] = x[0] + x[1] * c[0] + x[2] * c[1] + x[3] * c[2] + x[4] * c[3] + x[5] * c[4]
        xp[1] = x[1] + x[2] * c[0] + x[3] * c[1] + x[4] * c[2] + x[5] * c[3]
        xp[2] = x[2] + x[3] * c[0] + x[4] * c[1] + x[5] * c[2]
        xp[3] = x[3] + x[4] * c[0] + x[5] * c[1]
        xp[4] = x[4] + x[5] * c[0]
        xp[5] = x[5].copy()
        qp[0] = q[0] + q[1] * c[0] + q[2] * c[1] + q[3] * c[2] + q[4] * c[3] + q[5] * c[4]
        qp[1] = q[1] + q[2] * c[0] + q[3] * c[1] + q[4] * c[2] + q[5] * c[3]
        qp[2] = q[2] + q[3] * c[0] + q[4] * c[1] + q[5] * c[2]
        qp[3] = q[3] + q[4] * c[0] + q[5] * c[1]
        qp[4] = q[4] + q[5] * c[0]
        qp[5] = q[5].copy()
        for k in range(n):
            qp[0][k] = _qnormalize(qp[0][k])
        wp[0] = w[0] + w[1] * c[0] + w[2] * c[1] + w[3] * c[2] + w[4] * c[3] + w[5] * c[4]
        wp[1] = w[1] + w[2] * c[0] + w[3] * c[1] + w[4] * c[2] + w[5] * c[3]
        wp[2] = w[2] + w[3] * c[0] + w[4] * c[1] + w[5] * c[2]
        wp[3] = w[3] + w[4] * c[0] + w[5] * c[1]
        wp[4] = w[4] + w[5] * c[0]
        wp[5] = w[5].copy()
        self.wpGlobal = wp[0].copy()
        self.wpLocal = np.array([_project(self.wpGlobal[k], _qadj(qp[0][k])) for k in range(n)])
        self._update_particles(corrected=False)
        FP, FW, MP, MW = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
        for a, b in self.ppairs:
            d = self.px0[b] - self.px0[a]
            sig = self.pr[a] + self.pr[b]
            if not (_norm2(d) < sig * sig):
                continue
            i, j = self.cluster[a], self.cluster[b]
            dist = math.sqrt(_norm2(d))
            overlap = self.pr[a] + self.pr[b] - dist
            relVel = self.px1[b] - self.px1[a]
            en = d / dist
            vn = float(np.dot(relVel, en))
            normalRelVel = en * vn
            effMass = self.m[i] * self.m[j] / (self.m[i] + self.m[j])
            effRad = self.pr[a] * self.pr[b] / (self.pr[a] + self.pr[b])
            fn = self.normal(overlap, vn, effRad, effMass)
            nf = en * fn
            vecRadI, vecRadJ = self.pr[a] * en, -self.pr[b] * en
            cdI, cdJ = vecRadI + self.prv[a], vecRadJ + self.prv[b]
            FP[i] = FP[i] - nf
            if self.size[i] > 1:
                MP[i] = MP[i] - _cross(cdI, nf)
            FP[j] = FP[j] + nf
            if self.size[j] > 1:
                MP[j] = MP[j] + _cross(cdJ, nf)
            relC = relVel - _cross(self.wpGlobal[i], vecRadI) + _cross(self.wpGlobal[j], vecRadJ)
            tang = relC - normalRelVel
            nt = math.sqrt(_norm2(tang))
            if nt != 0.0:
                ft = self.tangential(nt, fn, effRad, effMass, p["frictionCoefPart"])
                et = tang / nt
                tf = ft * et
                MP[i] = MP[i] + _cross(cdI, tf); FP[i] = FP[i] + tf
                MP[j] = MP[j] - _cross(cdJ, tf); FP[j] = FP[j] - tf
        for a in range(self.P):
            wi = self.nearWallP[a]
            if wi < 0:
                continue
            k = self.cluster[a]
            wl = self.walls[wi]
            en = np.array(wl["n"])
            dist = float(np.dot(en, self.px0[a] - np.array(wl["p"])))
            overlap = self.pr[a] - dist
            if overlap > 0.0:
                cpv = np.zeros(3)
                if wl["moving"]:
                    dc = self.px0[a] - np.array(wl["rotCenter"])
                    cpv = np.array(wl["vel"]) + _cross(np.array(wl["omega"]), dc - float(np.dot(dc, en)) * en)
                relVel = self.px1[a] - cpv
                vn = float(np.dot(relVel, en))
                normalRelVel = en * vn
                fn = self.normal(2.0 * overlap, vn, self.pr[a], self.m[k])
                nf = en * fn
                vecRadJ = -self.pr[a] * en
                cdJ = vecRadJ
                if self.size[k] > 1:
                    cdJ = cdJ + (self.px0[a] - xp[0][k])
                FW[k] = FW[k] + nf
                if self.size[k] > 1:
                    MW[k] = MW[k] + _cross(cdJ, nf)
                relC = relVel + _cross(self.wpGlobal[k], vecRadJ)
                tang = relC - normalRelVel
                nt = math.sqrt(_norm2(tang))
                if nt != 0.0:
                    ft = self.tangential(nt, fn, self.pr[a], self.m[k], p["frictionCoefWall"])
                    et = tang / math.sqrt(_norm2(tang))
                    tf = ft * et
                    MW[k] = MW[k] - _cross(cdJ, tf)
                    FW[k] = FW[k] - tf
        demF = np.array(p["demF"])
        for k in range(n):
            FVisc = -6.0 * math.pi * p["numVisc"] * self.radius[k] * xp[1][k]
            MVisc = -8.0 * math.pi * p["numVisc"] * self.radius[k] * self.radius[k] * self.radius[k] * self.wpGlobal[k]
            x[2][k] = (FVisc + self.FHydro[k] + FP[k] + FW[k]) / self.m[k] + demF
            mom = MVisc + self.MHydro[k] + MP[k] + MW[k]
            momBf = _project(mom, _qadj(qp[0][k]))
            I, wl_ = self.I[k], self.wpLocal[k]
            waBf = np.array([(momBf[0] + (I[1] - I[2]) * wl_[1] * wl_[2]) / I[0], (momBf[1] + (I[2] - I[0]) * wl_[2] * wl_[0]) / I[1],
                             (momBf[2] + (I[0] - I[1]) * wl_[0] * wl_[1]) / I[2]])
            w[1][k] = _project(waBf, qp[0][k])
            if self.size[k] > 1:
                n2 = qp[1][k][0] ** 2 + qp[1][k][1] ** 2 + qp[1][k][2] ** 2 + qp[1][k][3] ** 2
                waQuat = np.array([-2.0 * n2, waBf[0], waBf[1], waBf[2]])
                q[2][k] = 0.5 * _qmul(qp[0][k], waQuat)
        c2, c1 = self.coeff2, self.coeff1
        x2c = x[2] - xp[2]
        x[0] = xp[0] + x2c * c2[0]; x[1] = xp[1] + x2c * c2[1]
        x[3] = xp[3] + x2c * c2[3]; x[4] = xp[4] + x2c * c2[4]; x[5] = xp[5] + x2c * c2[5]
        for k in range(6):
            xp[k] = x[k].copy()
        w1c = w[1] - wp[1]
        w[0] = wp[0] + w1c * c1[0]
        w[2] = wp[2] + w1c * c1[2]; w[3] = wp[3] + w1c * c1[3]; w[4] = wp[4] + w1c * c1[4]; w[5] = wp[5] + w1c * c1[5]
        for k in range(6):
            wp[k] = w[k].copy()
        q2c = q[2] - qp[2]
        q[0] = qp[0] + q2c * c2[0]; q[1] = qp[1] + q2c * c2[1]
        q[3] = qp[3] + q2c * c2[3]; q[4] = qp[4] + q2c * c2[4]; q[5] = qp[5] + q2c * c2[5]
        for k in range(n):
            q[0][k] = _qnormalize(q[0][k])
        for k in range(6):
            qp[k] = q[k].copy()
        self.wGlobal = w[0].copy()
        self.wLocal = np.array([_project(self.wGlobal[k], _qadj(q[0][k])) for k in range(n)])
        self._update_particles(corrected=True)

    def step(self, FHydro, MHydro):
        self.FHydro = np.asarray(FHydro, dtype=float).reshape(-1, 3); self.MHydro = np.asarray(MHydro, dtype=float).reshape(-1, 3)
        for _ in range(self.p["multiStep"]):
            self.substep()
        return self.px0.copy(), self.prv.copy(), self.x[1].copy(), self.wGlobal.copy()


# ---------------------------------------------------------------------------------------------------------------------
# Periodic DEM boundaries (single spheres): at every rebuild of the neighbour table the elements that left the domain are
# shifted back (DEM::pbcShift, DEM.cpp:1554-1584), particles within nebrRange of a periodic plane get a ghost on the other
# side, particles with two ghosts one more in the corner (DEM::createGhosts, DEM.cpp:1586-1660), and the table is built over
# particles and ghosts; between rebuilds a ghost follows its origin (particle::ghostUpdate, elmt.cpp:292-299).  A contact
# acts on the standard particles of the pair only (DEM.cpp:1849, 1856, 1885, 1889).  dem.newNeighborList tells the LB side
# to rescan after every rebuild (DEM.cpp:1414).
# ---------------------------------------------------------------------------------------------------------------------
class DemPortPbc(DemPort):
    def __init__(self, dem):
        super().__init__(dem)
        self.pbcs = []
        for b in dem["pbcs"]:
            p, v = np.array(b["p"], dtype=float), np.array(b["v"], dtype=float)
            n1 = v / math.sqrt(_norm2(v))
            self.pbcs.append(dict(v=v, p1=p, n1=n1, p2=p + v, n2=-1.0 * n1))
        self.ghosts = []       # (origin element, shift vector) in creation order
        self.allpairs = []     # (a, b) particle indices, a < b, over particles + ghosts
        self.new_list = False

    def particle_x(self, a, pred):
        src = self.xp[0] if pred else self.x[0]
        if a < self.n:
            return src[a]
        o, sh = self.ghosts[a - self.n]
        return src[o] + sh

    def origin(self, a):
        return a if a < self.n else self.ghosts[a - self.n][0]

    def lists(self):
        """(x0 of every particle, clusterIndex, components per element) as the LB side receives them."""
        P = self.n + len(self.ghosts)
        x0 = np.array([self.particle_x(a, False) for a in range(P)])
        comps = [[k] + [self.n + g for g, (o, _) in enumerate(self.ghosts) if o == k] for k in range(self.n)]
        return x0, np.array([self.origin(a) for a in range(P)]), comps

    def substep(self):
        p, n, c = self.p, self.n, self.c
        x, xp, w, wp = self.x, self.xp, self.w, self.wp
        maxVel = 0.0
        for k in range(n):
            maxVel = max(maxVel, _norm2(x[1][k]))
        self.maxDisp += math.sqrt(maxVel) * p["deltat"]
        if self.maxDisp > 0.25 * self.nebrRange:
            self.maxDisp = 0.0
            self.rebuilds += 1
            self.new_list = True
            for b in self.pbcs:  # pbcShift
                for k in range(n):
                    left = float(np.dot(b["n1"], x[0][k] - b["p1"]))
                    right = float(np.dot(b["n2"], x[0][k] - b["p2"]))
                    if left < 0.0:
                        x[0][k] = x[0][k] + b["v"]; xp[0][k] = xp[0][k] + b["v"]
                    if right < 0.0:
                        x[0][k] = x[0][k] + (-1.0 * b["v"]); xp[0][k] = xp[0][k] + (-1.0 * b["v"])
            self.ghosts = []
            for b in self.pbcs:  # createGhosts
                for k in range(n):
                    left = float(np.dot(b["n1"], x[0][k] - b["p1"]))
                    right = float(np.dot(b["n2"], x[0][k] - b["p2"]))
                    if left < self.nebrRange:
                        self.ghosts.append((k, b["v"]))
                    elif right < self.nebrRange:
                        self.ghosts.append((k, -1.0 * b["v"]))
            corner = []
            for g1 in range(len(self.ghosts)):
                for g2 in range(g1 + 1, len(self.ghosts)):
                    if self.ghosts[g1][0] == self.ghosts[g2][0]:
                        corner.append((self.ghosts[g1][0], self.ghosts[g1][1] + self.ghosts[g2][1]))
            self.ghosts += corner
            P = n + len(self.ghosts)
            r2 = self.nebrRange * self.nebrRange
            self.allpairs = [(a, b) for a in range(P) for b in range(a + 1, P) if self.origin(a) != self.origin(b) and
                             _norm2(self.particle_x(b, False) - self.particle_x(a, False)) < r2]
            for k in range(n):
                self.nearWall[k] = -1
                for wi, wl in enumerate(self.walls):
                    if float(np.dot(np.array(wl["n"]), x[0][k] - np.array(wl["p"]))) < self.nebrRange:
                        self.nearWall[k] = wi
                        break
        xp[0] = x[0] + x[1] * c[0] + x[2] * c[1] + x[3] * c[2] + x[4] * c[3] + x[5] * c[4]
        xp[1] = x[1] + x[2] * c[0] + x[3] * c[1] + x[4] * c[2] + x[5] * c[3]
        xp[2] = x[2] + x[3] * c[0] + x[4] * c[1] + x[5] * c[2]
        xp[3] = x[3] + x[4] * c[0] + x[5] * c[1]
        xp[4] = x[4] + x[5] * c[0]
        xp[5] = x[5].copy()
        wp[0] = w[0] + w[1] * c[0] + w[2] * c[1] + w[3] * c[2] + w[4] * c[3] + w[5] * c[4]
        wp[1] = w[1] + w[2] * c[0] + w[3] * c[1] + w[4] * c[2] + w[5] * c[3]
        wp[2] = w[2] + w[3] * c[0] + w[4] * c[1] + w[5] * c[2]
        wp[3] = w[3] + w[4] * c[0] + w[5] * c[1]
        wp[4] = w[4] + w[5] * c[0]
        wp[5] = w[5].copy()
        FP, FW, MP, MW = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
        for a, b in self.allpairs:
            if a >= n and b >= n:
                continue  # two ghosts: nobody to push
            i, j = self.origin(a), self.origin(b)
            d = self.particle_x(b, True) - self.particle_x(a, True)
            sig = self.radius[i] + self.radius[j]
            if not (_norm2(d) < sig * sig):
                continue
            dist = math.sqrt(_norm2(d))
            overlap = self.radius[i] + self.radius[j] - dist
            relVel = xp[1][j] - xp[1][i]
            en = d / dist
            vn = float(np.dot(relVel, en))
            normalRelVel = en * vn
            effMass = self.m[i] * self.m[j] / (self.m[i] + self.m[j])
            effRad = self.radius[i] * self.radius[j] / (self.radius[i] + self.radius[j])
            fn = self.normal(overlap, vn, effRad, effMass)
            nf = en * fn
            vecRadI, vecRadJ = self.radius[i] * en, -self.radius[j] * en
            if a < n:
                FP[i] = FP[i] - nf
            if b < n:
                FP[j] = FP[j] + nf
            relC = relVel - _cross(wp[0][i], vecRadI) + _cross(wp[0][j], vecRadJ)
            tang = relC - normalRelVel
            nt = math.sqrt(_norm2(tang))
            if nt != 0.0:
                ft = self.tangential(nt, fn, effRad, effMass, p["frictionCoefPart"])
                et = tang / nt
                tf = ft * et
                if a < n:
                    MP[i] = MP[i] + _cross(vecRadI, tf); FP[i] = FP[i] + tf
                if b < n:
                    MP[j] = MP[j] - _cross(vecRadJ, tf); FP[j] = FP[j] - tf
        for k in range(n):
            wi = self.nearWall[k]
            if wi < 0:
                continue
            wl = self.walls[wi]
            en = np.array(wl["n"])
            dist = float(np.dot(en, xp[0][k] - np.array(wl["p"])))
            overlap = self.radius[k] - dist
            if overlap > 0.0:
                cpv = np.zeros(3)
                if wl["moving"]:
                    dc = xp[0][k] - np.array(wl["rotCenter"])
                    cpv = np.array(wl["vel"]) + _cross(np.array(wl["omega"]), dc - float(np.dot(dc, en)) * en)
                relVel = xp[1][k] - cpv
                vn = float(np.dot(relVel, en))
                normalRelVel = en * vn
                fn = self.normal(2.0 * overlap, vn, self.radius[k], self.m[k])
                nf = en * fn
                vecRadJ = -self.radius[k] * en
                FW[k] = FW[k] + nf
                relC = relVel + _cross(wp[0][k], vecRadJ)
                tang = relC - normalRelVel
                nt = math.sqrt(_norm2(tang))
                if nt != 0.0:
                    ft = self.tangential(nt, fn, self.radius[k], self.m[k], p["frictionCoefWall"])
                    et = tang / math.sqrt(_norm2(tang))
                    tf = ft * et
                    MW[k] = MW[k] - _cross(vecRadJ, tf)
                    FW[k] = FW[k] - tf
        demF = np.array(p["demF"])
        for k in range(n):
            FVisc = -6.0 * math.pi * p["numVisc"] * self.radius[k] * xp[1][k]
            MVisc = -8.0 * math.pi * p["numVisc"] * self.radius[k] * self.radius[k] * self.radius[k] * wp[0][k]
            x[2][k] = (FVisc + self.FHydro[k] + FP[k] + FW[k]) / self.m[k] + demF
            mom = MVisc + self.MHydro[k] + MP[k] + MW[k]
            I, wl_ = self.I[k], wp[0][k]
            w[1][k] = np.array([(mom[0] + (I[1] - I[2]) * wl_[1] * wl_[2]) / I[0], (mom[1] + (I[2] - I[0]) * wl_[2] * wl_[0]) / I[1],
                                (mom[2] + (I[0] - I[1]) * wl_[0] * wl_[1]) / I[2]])
        c2, c1 = self.coeff2, self.coeff1
        x2c = x[2] - xp[2]
        x[0] = xp[0] + x2c * c2[0]; x[1] = xp[1] + x2c * c2[1]
        x[3] = xp[3] + x2c * c2[3]; x[4] = xp[4] + x2c * c2[4]; x[5] = xp[5] + x2c * c2[5]
        for k in range(6):
            xp[k] = x[k].copy()
        w1c = w[1] - wp[1]
        w[0] = wp[0] + w1c * c1[0]
        w[2] = wp[2] + w1c * c1[2]; w[3] = wp[3] + w1c * c1[3]; w[4] = wp[4] + w1c * c1[4]; w[5] = wp[5] + w1c * c1[5]
        for k in range(6):
            wp[k] = w[k].copy()

    def step(self, FHydro, MHydro):
        self.new_list = False
        self.FHydro = np.asarray(FHydro, dtype=float).reshape(-1, 3); self.MHydro = np.asarray(MHydro, dtype=float).reshape(-1, 3)
        for _ in range(self.p["multiStep"]):
            self.substep()
        return self.lists() + (self.x[1].copy(), self.w[0].copy(), self.new_list)
