"""Quick device timing of one catalogue case:  python tests/quick_bench.py CASE STEPS
Cases with particles are stepped through lbGpuStep with the (kinematic) particle state every step, like goCycle."""
import sys
import time

import numpy as np

import common
from common import li

sys.path.insert(0, common.ROOT + "/oracle")
import cases  # noqa: E402

name = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
case = cases.materialise(dict(cases.catalogue()[name]))
t0 = time.time()
prm = li.params_from_case(case)
tr = common.KinematicTrace(case, prm)
state = li.build_state(case, tr.initial_particles() if len(tr.parts) else None)
print("init %.1fs" % (time.time() - t0), prm["size"], "particles", len(tr.parts))
g = common.make_gpu(state)
print("upload %.1fs" % (time.time() - t0))
fs = bool(state.params["freeSurface"])


def advance(n):
    if len(tr.parts) == 0:
        g.run(n)
        g.synchronize()
        return g.last_step_ms() / n
    t = time.perf_counter()
    for _ in range(n):
        common.cycle(g, state.params, tr)
    g.synchronize()
    return 1e3 * (time.perf_counter() - t) / n


advance(5)
for rep in range(3):
    l0 = g.launch_count()
    ms = advance(steps)
    c = g.counts()
    act = c["fluid"] + c["interface"]
    print("%s: %.4f ms/step  %.1f MLUPS(active %d, interface %d, particle cells %d)  %.1f GB/s algorithmic  %.1f launches/step"
          % (name, ms, act / ms / 1e3, act, c["interface"], c["particle"], act * 304 / ms / 1e6, (g.launch_count() - l0) / steps))
d = g.fetch(("n", "u", "type_flags"))
a = np.isin(d["type_flags"] & 15, (0, 3))
print("n range", d["n"][a].min(), d["n"][a].max(), "umax", np.abs(d["u"][a]).max())
