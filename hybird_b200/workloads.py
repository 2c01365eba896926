"""Workload catalogue: the five BASELINE.json configurations (`cfg1`..`cfg5`), their scaled-down `*_mini` versions
and the branch-coverage cases of the parity tests.

Every case is a plain dict of hybird configuration keys (the `key = value` pairs GetPot reads in the reference:
hybird.cpp:132-223, LB.cpp:85-188, DEM.cpp:13-183) plus set-up options the reference hard-codes per problemName
(initial gas regions, wall velocities, particle list, prescribed particle motion).  `lattice_init.build_state` turns
a case into the arrays lbGpuInit takes; oracle/cases.py writes the same case as a .cfg + particle file for the
unmodified reference.
"""
from __future__ import annotations

import copy

_BASE = dict(
    demSolve=1, lbSolve=1, freeSurfaceSolve=0, forceFieldSolve=1, nonNewtonianSolve=0, turbulenceSolve=0,
    problemName="NONE", demInitialRepeat=0, lbmInitialRepeat=0, maximumTimeSteps=0, maxTime=1e9,
    screenExpTime=0, fluidExpTime=0, partExpTime=0, recycleExpTime=0, objectExpTime=0, saveCount=0,
    unitLength=1.0, unitTime=1.0, unitDensity=1.0,
    lbSizeX=16, lbSizeY=16, lbSizeZ=16, initVisc=1.0 / 6.0,
    initVelocityX=0.0, initVelocityY=0.0, initVelocityZ=0.0, plasticVisc=1.0 / 6.0, yieldStress=0.0,
    lbFX=0.0, lbFY=0.0, lbFZ=0.0,
    boundary0=7, boundary1=7, boundary2=7, boundary3=7, boundary4=7, boundary5=7,
    slipCoefficient=0.0, turbConst=0.0,
    density=2.5, contactModel="LINEAR", youngMod=1.0, poisson=0.3, linearStiff=1.0, restitution=0.9,
    viscTang=0.5, frictionCoefPart=0.3, frictionCoefWall=0.3,
    translateX=0.0, translateY=0.0, translateZ=0.0, scale=1.0, numVisc=0.0, multiStep=1, criticalRatio=0.1,
)

# harness-only keys (not written to the .cfg)
_HARNESS_KEYS = ("name", "elements", "elements_gen", "fluid_box", "gas_box", "fluid_sphere", "gas_sphere", "wall_vel", "motion",
                 "rescan_every", "note")


def make_case(name, **kw):
    c = copy.deepcopy(_BASE)
    c.update(name=name, elements=[], motion="none", rescan_every=0)
    c.update(kw)
    return c


def _sphere_bed(n, lo, hi, rmin, rmax, seed):
    """Random sequential addition of non-overlapping spheres (own 64-bit LCG: no numpy dependency,
    identical on every box)."""
    state = seed & 0xFFFFFFFFFFFFFFFF
    def rnd():
        nonlocal state
        state = (state * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        return (state >> 11) / float(1 << 53)
    out = []
    cell = 2.0 * rmax
    grid = {}
    tries = 0
    while len(out) < n and tries < 200 * n:
        tries += 1
        r = rmin + (rmax - rmin) * rnd()
        p = [lo[k] + r + (hi[k] - lo[k] - 2 * r) * rnd() for k in range(3)]
        key = tuple(int(p[k] // cell) for k in range(3))
        ok = True
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    for (q, rq) in grid.get((key[0] + dx, key[1] + dy, key[2] + dz), ()):
                        if (p[0] - q[0]) ** 2 + (p[1] - q[1]) ** 2 + (p[2] - q[2]) ** 2 < (r + rq) ** 2:
                            ok = False
        if ok:
            grid.setdefault(key, []).append((p, r))
            out.append(dict(size=1, radius=r, x0=p, x1=[0.0, 0.0, 0.0], w=[0.0, 0.0, 0.0]))
    return out


def catalogue():
    C = {}
    # --- cfg 2: pure-fluid channel, periodic x/y, no-slip z, body force along x -----
    C["cfg2"] = make_case("cfg2", lbSizeX=256, lbSizeY=256, lbSizeZ=256, boundary0=4, boundary1=4, boundary2=4,
                          boundary3=4, lbFX=1e-6)
    C["cfg2_mini"] = make_case("cfg2_mini", lbSizeX=24, lbSizeY=20, lbSizeZ=16, boundary0=4, boundary1=4,
                               boundary2=4, boundary3=4, lbFX=1e-6)
    # a harder pure-fluid case: oblique force, initial velocity, all three periodic pairs off/on mixes
    C["channel_oblique"] = make_case("channel_oblique", lbSizeX=18, lbSizeY=14, lbSizeZ=12, boundary0=4, boundary1=4,
                                     lbFX=3e-5, lbFY=-2e-5, lbFZ=-1e-5, initVelocityX=0.01, initVelocityY=-0.02,
                                     initVelocityZ=0.005, initVisc=0.05)
    C["box_noforce"] = make_case("box_noforce", lbSizeX=12, lbSizeY=12, lbSizeZ=12, forceFieldSolve=0, lbFX=1e-3,
                                 initVelocityX=0.02)
    C["periodic_all"] = make_case("periodic_all", lbSizeX=12, lbSizeY=10, lbSizeZ=14, boundary0=4, boundary1=4,
                                  boundary2=4, boundary3=4, boundary4=4, boundary5=4, lbFZ=-2e-5, initVelocityY=0.03)
    # --- viscosity models ------------------------------------------------------------------------
    C["smago_channel"] = make_case("smago_channel", lbSizeX=16, lbSizeY=12, lbSizeZ=14, boundary0=4, boundary1=4,
                                   turbulenceSolve=1, turbConst=0.01, lbFX=5e-5, initVisc=0.01)
    C["bingham_channel"] = make_case("bingham_channel", lbSizeX=16, lbSizeY=12, lbSizeZ=14, boundary0=4, boundary1=4,
                                     nonNewtonianSolve=1, plasticVisc=1.0 / 30.0, yieldStress=1e-5, initVisc=1.0 / 30.0,
                                     lbFX=5e-5)
    C["bingham_smago"] = make_case("bingham_smago", lbSizeX=14, lbSizeY=12, lbSizeZ=12, boundary2=4, boundary3=4,
                                   nonNewtonianSolve=1, turbulenceSolve=1, turbConst=0.02, plasticVisc=0.02,
                                   yieldStress=2e-5, initVisc=0.02, lbFY=4e-5, lbFZ=-1e-5)
    # --- walls -----------------------------------------------------------------------------------
    C["couette_dyn"] = make_case("couette_dyn", lbSizeX=14, lbSizeY=12, lbSizeZ=12, boundary0=4, boundary1=4,
                                 boundary2=4, boundary3=4, boundary4=7, boundary5=8, wall_vel=[(1, 0.03, 0.01, 0.0)],
                                 lbFX=1e-6)
    C["slip_box"] = make_case("slip_box", lbSizeX=14, lbSizeY=12, lbSizeZ=12, boundary0=5, boundary1=5, boundary2=7,
                              boundary3=5, boundary4=5, boundary5=7, slipCoefficient=0.3, lbFX=2e-5, lbFY=1e-5,
                              lbFZ=-3e-5, initVelocityX=0.01)
    C["slip_dyn"] = make_case("slip_dyn", lbSizeX=14, lbSizeY=12, lbSizeZ=12, boundary0=4, boundary1=4, boundary2=5,
                              boundary3=6, boundary4=6, boundary5=8, slipCoefficient=0.4,
                              wall_vel=[(1, 0.02, 0.0, 0.01), (2, 0.01, 0.02, 0.0), (3, -0.02, 0.01, 0.0)],
                              lbFX=2e-5, lbFZ=-1e-5)
    # --- cfg 3: single sphere, DEM-coupled -------------------------------------------
    C["cfg3"] = make_case("cfg3", lbSizeX=128, lbSizeY=128, lbSizeZ=256, lbFZ=-1e-5, initVisc=0.1,
                          elements=[dict(size=1, radius=8.0, x0=[64.0, 64.0, 192.0], x1=[0, 0, 0], w=[0, 0, 0])],
                          motion="dem")
    C["cfg3_mini"] = make_case("cfg3_mini", lbSizeX=24, lbSizeY=24, lbSizeZ=40, lbFZ=-1e-5, initVisc=0.1,
                               elements=[dict(size=1, radius=4.0, x0=[12.0, 12.0, 28.0], x1=[0, 0, 0], w=[0, 0, 0])],
                               motion="dem")
    C["sphere_kin"] = make_case("sphere_kin", lbSizeX=20, lbSizeY=18, lbSizeZ=24, lbFZ=-1e-5, initVisc=0.1,
                                elements=[dict(size=1, radius=3.6, x0=[9.3, 8.7, 15.2], x1=[0.021, -0.013, -0.034],
                                               w=[0.01, 0.02, -0.015])],
                                motion="kin", rescan_every=7)
    C["two_spheres_kin"] = make_case(
        "two_spheres_kin", lbSizeX=26, lbSizeY=18, lbSizeZ=20, boundary0=4, boundary1=4, lbFZ=-2e-5, initVisc=0.08,
        elements=[dict(size=1, radius=3.2, x0=[8.4, 8.9, 10.2], x1=[0.05, 0.0, 0.01], w=[0.0, 0.03, 0.0]),
                  dict(size=1, radius=2.7, x0=[14.6, 9.4, 10.9], x1=[-0.04, 0.01, -0.01], w=[0.02, 0.0, 0.01])],
        motion="kin", rescan_every=5)
    # demSolve = 0 (the reference's default, hybird.cpp:180): goCycle never runs the coupling step, the flags of
    # LB::initializeParticleBoundaries stay and LB::computeHydroForces keeps forcing the flagged cells
    C["sphere_fixed"] = make_case("sphere_fixed", demSolve=0, lbSizeX=20, lbSizeY=18, lbSizeZ=22, boundary0=4, boundary1=4,
                                  lbFX=4e-5, lbFZ=-1e-5, initVisc=0.08,
                                  elements=[dict(size=1, radius=3.4, x0=[9.6, 8.8, 11.3], x1=[0.0, 0.0, 0.0], w=[0.0, 0.0, 0.0]),
                                            dict(size=1, radius=2.1, x0=[14.2, 9.1, 6.4], x1=[0.0, 0.0, 0.0], w=[0.0, 0.0, 0.0])])
    C["cluster_dem"] = make_case(
        "cluster_dem", lbSizeX=24, lbSizeY=22, lbSizeZ=26, lbFZ=-3e-5, initVisc=0.1,
        elements=[dict(size=2, radius=2.6, x0=[11.0, 10.5, 17.0], x1=[0.0, 0.0, 0.0], w=[0.01, 0.02, 0.0]),
                  dict(size=3, radius=2.2, x0=[12.5, 11.0, 8.5], x1=[0.0, 0.0, 0.0], w=[0.0, 0.0, 0.02])],
        motion="dem")
    # DEM in the loop with contacts: sphere-sphere and sphere-wall collisions, three DEM sub-steps per LB step; sphere 3
    # drifts into a corner (DEM::evalNearWallTable lists ONE wall per particle).  LINEAR and HERTZIAN contact models.
    dem_spheres = [dict(size=1, radius=3.2, x0=[10.6, 12.0, 15.0], x1=[0.09, 0.0, 0.0], w=[0.0, 0.0, 0.01]),
                   dict(size=1, radius=3.0, x0=[17.6, 12.6, 15.4], x1=[-0.08, 0.0, 0.0], w=[0.0, 0.02, 0.0]),
                   dict(size=1, radius=3.4, x0=[14.0, 13.0, 4.5], x1=[0.0, 0.01, -0.07], w=[0.0, 0.0, 0.0]),
                   dict(size=1, radius=3.0, x0=[3.9, 3.8, 23.0], x1=[-0.06, -0.05, 0.01], w=[0.01, 0.0, 0.0])]
    C["spheres_dem"] = make_case("spheres_dem", lbSizeX=30, lbSizeY=26, lbSizeZ=30, lbFZ=-2e-5, initVisc=0.1, multiStep=3,
                                 elements=copy.deepcopy(dem_spheres), motion="dem")
    C["spheres_hertz"] = make_case("spheres_hertz", lbSizeX=30, lbSizeY=26, lbSizeZ=30, lbFZ=-2e-5, initVisc=0.1, multiStep=2,
                                   contactModel="HERTZIAN", youngMod=4.0, poisson=0.3, restitution=0.8, viscTang=0.3,
                                   elements=copy.deepcopy(dem_spheres), motion="dem")
    # clusters in contact: a pair of spheres and a triangle collide in mid-fluid, a tetrahedron lands on the floor (lever arms of
    # the contact forces, rotation in the body frame, per-particle neighbour and wall tables)
    C["clusters_hit"] = make_case(
        "clusters_hit", lbSizeX=34, lbSizeY=28, lbSizeZ=32, lbFZ=-4e-5, initVisc=0.05, multiStep=2, density=6.0,
        elements=[dict(size=2, radius=2.4, x0=[11.0, 13.0, 18.0], x1=[0.10, 0.0, 0.0], w=[0.0, 0.0, 0.02]),
                  dict(size=3, radius=2.2, x0=[22.5, 14.0, 18.6], x1=[-0.09, 0.0, 0.0], w=[0.0, 0.01, 0.0]),
                  dict(size=4, radius=2.0, x0=[16.0, 13.0, 6.3], x1=[0.0, 0.01, -0.08], w=[0.01, 0.0, 0.0])],
        motion="dem")
    # periodic DEM boundaries: x and y periodic (two pbcs, so ghost particles in the corners too), walls in z.  Spheres 0 and 1
    # collide ACROSS the x face (a contact between a particle and a ghost), sphere 2 leaves through the x and the y face and
    # re-enters on the other side (DEM::pbcShift), the ghost count changes from rebuild to rebuild, the LB side rescans
    C["spheres_pbc_dem"] = make_case(
        "spheres_pbc_dem", lbSizeX=30, lbSizeY=26, lbSizeZ=30, boundary0=4, boundary1=4, boundary2=4, boundary3=4, lbFZ=-3e-5,
        initVisc=0.04, multiStep=2, density=8.0,
        elements=[dict(size=1, radius=3.0, x0=[4.6, 12.0, 15.0], x1=[-0.12, 0.0, 0.0], w=[0.0, 0.0, 0.01]),
                  dict(size=1, radius=3.2, x0=[24.0, 12.8, 15.6], x1=[0.096, 0.0, 0.0], w=[0.0, 0.02, 0.0]),
                  dict(size=1, radius=2.6, x0=[3.4, 3.2, 21.0], x1=[-0.06, -0.072, 0.0], w=[0.01, 0.0, 0.0]),
                  dict(size=1, radius=3.0, x0=[15.0, 13.0, 5.0], x1=[0.02, 0.05, -0.06], w=[0.0, 0.0, 0.0])],
        motion="dem")
    # a loose bed: twelve heavy spheres with random velocities in a box wider than nebrRange -- several rebuilds of the
    # neighbour table, pairs that enter and leave it, wall contacts on every side
    bed = _sphere_bed(12, (1.0, 1.0, 1.0), (39.0, 35.0, 39.0), 2.2, 3.0, 777)
    for k, e in enumerate(bed):
        e["x1"] = [0.12 * (((k * 7) % 5) - 2) / 2.0, 0.1 * (((k * 3) % 7) - 3) / 3.0, 0.11 * (((k * 5) % 3) - 1)]
        e["w"] = [0.01 * ((k % 3) - 1), 0.0, 0.02 * ((k % 2) - 0.5)]
    C["bed_dem"] = make_case("bed_dem", lbSizeX=40, lbSizeY=36, lbSizeZ=40, lbFZ=-4e-5, initVisc=0.04, multiStep=2, density=8.0,
                             elements=bed, motion="dem")
    # --- cfg 4: free-surface dam break, Bingham --------------------------------------
    C["cfg4"] = make_case("cfg4", lbSizeX=512, lbSizeY=128, lbSizeZ=256, freeSurfaceSolve=1, nonNewtonianSolve=1,
                          lbFZ=-1e-4, plasticVisc=1.0 / 30.0, yieldStress=1e-5, initVisc=1.0 / 30.0,
                          fluid_box=(1, 128, 1, 126, 1, 192))
    # (A/B of the step kernel's variants on cfg4's geometry: the same dam break with a Newtonian fluid, and the full column)
    C["cfg4_newtonian"] = make_case("cfg4_newtonian", lbSizeX=512, lbSizeY=128, lbSizeZ=256, freeSurfaceSolve=1, lbFZ=-1e-4,
                                    initVisc=1.0 / 30.0, fluid_box=(1, 128, 1, 126, 1, 192))
    C["cfg4_mini"] = make_case("cfg4_mini", lbSizeX=40, lbSizeY=12, lbSizeZ=24, freeSurfaceSolve=1, nonNewtonianSolve=1,
                               lbFZ=-1e-4, plasticVisc=1.0 / 30.0, yieldStress=1e-5, initVisc=1.0 / 30.0,
                               fluid_box=(1, 12, 1, 10, 1, 18))
    C["dam_newtonian"] = make_case("dam_newtonian", lbSizeX=30, lbSizeY=10, lbSizeZ=20, freeSurfaceSolve=1,
                                   lbFZ=-2e-4, initVisc=0.02, fluid_box=(1, 10, 1, 8, 1, 14))
    C["droplet"] = make_case("droplet", lbSizeX=20, lbSizeY=20, lbSizeZ=24, freeSurfaceSolve=1, lbFZ=-3e-4,
                             initVisc=0.03, fluid_sphere=(9.5, 9.5, 14.0, 5.2))
    C["bubble_periodic"] = make_case("bubble_periodic", lbSizeX=18, lbSizeY=16, lbSizeZ=20, boundary0=4, boundary1=4,
                                     boundary2=4, boundary3=4, freeSurfaceSolve=1, lbFZ=-2e-4, initVisc=0.03,
                                     gas_sphere=(8.0, 8.0, 8.0, 4.3), gas_box=(0, 17, 0, 15, 15, 19))
    # --- cfg 1: the shipped lbmConfigDevisFluid.cfg values (stand-in empty particle/object files) -
    devis = dict(problemName="demChute", freeSurfaceSolve=1, turbulenceSolve=1, unitLength=1.0e-3, unitTime=2.0e-4,
                 unitDensity=0.846e3, initVisc=0.017258, plasticVisc=50.0, yieldStress=500.0, boundary0=4, boundary1=4,
                 turbConst=0.01, chuteInclination=15.0, density=2230.0, contactModel="HERTZIAN", youngMod=2.0e7,
                 linearStiff=0.0, restitution=0.99, viscTang=0.2, frictionCoefPart=0.5, frictionCoefWall=0.2,
                 multiStep=0, criticalRatio=0.005)
    C["cfg1"] = make_case("cfg1", lbSizeX=0.1, lbSizeY=0.15, lbSizeZ=0.03, **devis)
    C["cfg1_mini"] = make_case("cfg1_mini", lbSizeX=0.02, lbSizeY=0.024, lbSizeZ=0.03, **devis)
    # --- cfg 5: debris flow, free surface + many spheres -----------------------------------------
    C["cfg5_mini"] = make_case(
        "cfg5_mini", lbSizeX=48, lbSizeY=20, lbSizeZ=24, boundary2=4, boundary3=4, freeSurfaceSolve=1, lbFZ=-1e-4,
        lbFX=3e-5, initVisc=0.05, fluid_box=(0, 47, 0, 19, 0, 15),
        elements=_sphere_bed(14, (1.5, 1.5, 1.5), (46.5, 18.5, 13.5), 2.2, 3.0, 12345), motion="kin", rescan_every=6)
    for e in C["cfg5_mini"]["elements"]:
        e["x1"] = [0.02, 0.0, -0.01]
    # --- curved walls (Mei-Luo-Shyy) and LB::enforceMassConservation: the reference's rotating DRUM geometry
    # (DEM.cpp:876-890, LB.cpp:734-748) scaled to a 38x10x50 lattice; unitDensity is chosen so that the mass target
    # fluidMass/unit.Mass (LB.cpp:205-208) is close to the mass the fluid region actually holds
    C["drum_mini"] = make_case(
        "drum_mini", problemName="DRUM", unitLength=0.05, unitTime=8e-4, unitDensity=2049.0, lbSizeX=1.9, lbSizeY=0.5,
        lbSizeZ=2.5, boundary2=4, boundary3=4, freeSurfaceSolve=1, lbFZ=-9.81, initVisc=320.0, plasticVisc=320.0,
        drumSpeed=1.5, fluidMass=400.0)
    C["drum_bingham"] = make_case(
        "drum_bingham", problemName="DRUM", unitLength=0.0625, unitTime=1e-3, unitDensity=1308.0, lbSizeX=1.9,
        lbSizeY=0.5, lbSizeZ=2.5, boundary2=7, boundary3=8, freeSurfaceSolve=1, nonNewtonianSolve=1, lbFZ=-9.81,
        initVisc=200.0, plasticVisc=200.0, yieldStress=20.0, drumSpeed=-1.2, fluidMass=500.0)
    # cfg 5 at full size: 1024x256x256 with the long (flow) axis stored slowest, i.e. as the lattice's z (the engine
    # cuts slabs along z); depth is x (fluid for x <= 160, gravity along -x), y periodic, 20 000 spheres r in [3,4]
    # by random sequential addition (seed 12345) inside the fluid.  The bed is generated on demand (materialise).
    C["cfg5"] = make_case(
        "cfg5", lbSizeX=256, lbSizeY=256, lbSizeZ=1024, boundary2=4, boundary3=4, freeSurfaceSolve=1, lbFX=-1e-4,
        lbFZ=3e-5, initVisc=0.05, fluid_box=(0, 160, 0, 255, 0, 1023), motion="kin", rescan_every=0,
        elements_gen=(20000, (1.0, 1.0, 1.0), (160.0, 255.0, 1023.0), 3.0, 4.0, 12345, [-0.01, 0.0, 0.02]))
    # cfg 5 as the reference would run it: the DEM advances the 20 000 spheres every cycle (contacts, periodic y with ghost
    # particles, walls in x and z) -- on the device through lbGpuRunDem
    C["cfg5_dem"] = dict(C["cfg5"], name="cfg5_dem", motion="dem")
    C["cfg5_mini_dem"] = dict(copy.deepcopy(C["cfg5_mini"]), name="cfg5_mini_dem", motion="dem", rescan_every=0)
    return C


def materialise(case):
    """Generate the sphere bed of a case that only carries its recipe (`elements_gen`)."""
    if "elements_gen" in case:
        case = dict(case)
        n, lo, hi, rmin, rmax, seed, vel = case.pop("elements_gen")
        case["elements"] = _sphere_bed(n, lo, hi, rmin, rmax, seed)
        for e in case["elements"]:
            e["x1"] = list(vel)
    return case
