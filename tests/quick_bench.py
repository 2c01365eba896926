"""Quick device timing of one catalogue case:  python tests/quick_bench.py CASE STEPS"""
import sys
import time

import numpy as np

import common
from common import li

sys.path.insert(0, common.ROOT + "/oracle")
import cases  # noqa: E402

name = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
case = cases.catalogue()[name]
t0 = time.time()
prm = li.params_from_case(case)
tr = common.KinematicTrace(case, prm)
state = li.build_state(case, tr.initial_particles() if len(tr.parts) else None)
print("init %.1fs" % (time.time() - t0), prm["size"])
g = common.make_gpu(state)
print("upload %.1fs" % (time.time() - t0))
c0 = g.counts()
g.run(5)
g.synchronize()
for rep in range(3):
    g.run(steps)
    ms = g.last_step_ms() / steps
    c = g.counts()
    act = c["fluid"] + c["interface"]
    print("%s: %.4f ms/step  %.1f MLUPS(active %d)  %.1f GB/s algorithmic" % (name, ms, act / ms / 1e3, act, act * 304 / ms / 1e6))
d = g.fetch(("n", "u", "type_flags"))
a = np.isin(d["type_flags"] & 15, (0, 3))
print("n range", d["n"][a].min(), d["n"][a].max(), "umax", np.abs(d["u"][a]).max())
