"""cfg5 with the DEM in the loop on the device: per-cycle diagnostics (python tools/cfg5_dem_probe.py [cycles])"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
from hybird_b200 import slabs, dem_init
cycles = int(sys.argv[1]) if len(sys.argv) > 1 else 12
case = bench.workload_case("cfg5_dem", 1)
lb, info = slabs.build_engine(case, 0, 1, device=0, dist=None, device_init=True)
dem = dem_init.dem_from_case(case, info["params"])
print("elements", len(dem["elmts"]), "pbcs", dem["pbcs"], "nebrRange", dem["params"]["nebrRange"], "walls", len(dem["walls"]), flush=True)
lb.demInit(dem)
print("flagged cells after init", lb.counts()["particle"], flush=True)
for c in range(cycles):
    try:
        lb.runDem(1)
        st = lb.demState(); pt = lb.demParticles()
        cn = lb.counts()
        F = lb.forces()[0]
        if c < 3 or (c + 1) % 10 == 0 or st["rebuilds"] != getattr(sys.modules[__name__], "_rb", 0):
          sys.modules[__name__]._rb = st["rebuilds"]
          print("cycle", c + 1, "particles+ghosts", len(pt["x0"]), "rebuilds", st["rebuilds"], "max|x1| %.4f" % np.abs(st["x1"]).max(), "max|F| %.3g" % np.abs(F).max(),
              "flagged", cn["particle"], "fluid", cn["fluid"], "interface", cn["interface"], flush=True)
    except Exception as e:
        print("cycle", c + 1, "FAILED:", str(e)[:300], flush=True)
        break
