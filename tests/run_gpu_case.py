"""Ad-hoc GPU-vs-oracle comparison:  python tests/run_gpu_case.py CASE[,CASE...] STEPS"""
import sys

import numpy as np

import common
from common import li

sys.path.insert(0, common.ROOT + "/oracle")
import cases  # noqa: E402


def run(name, steps, every=0):
    case = cases.catalogue()[name]
    prm = li.params_from_case(case)
    tr_o = common.KinematicTrace(case, prm)
    tr_g = common.KinematicTrace(case, prm)
    state = li.build_state(case, tr_o.initial_particles() if len(tr_o.parts) else None)
    o = common.make_oracle(state)
    g = common.make_gpu(state)
    worst = {}
    fmax = 0.0
    for s in range(1, steps + 1):
        Fo = common.cycle(o, state.params, tr_o)
        Fg = common.cycle(g, state.params, tr_g)
        for a, b in zip(Fo, Fg):
            fmax = max(fmax, common.max_rel(a, b))
        if (every and s % every == 0) or s == steps or s == 1:
            rep = common.compare(common.gpu_state(g), common.oracle_state(o))
            for k, v in rep.items():
                worst[k] = max(worst.get(k, 0), v)
            if s == 1:
                print("  step 1:", {k: v for k, v in rep.items() if v})
    print(name, "steps", steps, "force_rel %.3g" % fmax, {k: v for k, v in worst.items() if v} or "IDENTICAL",
          "launches", g.launch_count(), g.counts())
    g.close(); o.close()
    return worst


if __name__ == "__main__":
    names = sys.argv[1].split(",")
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    every = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    for nm in names:
        run(nm, steps, every)
