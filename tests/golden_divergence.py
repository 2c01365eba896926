"""Ad-hoc: first step at which the GPU engine and the C restatement differ on a golden case (inputs replayed from
the fixture).  python tests/golden_divergence.py CASE [STEPS]"""
import sys
import numpy as np
import conftest  # noqa: F401
import common
import golden_util as gu
import lbo

name = sys.argv[1]
g = gu.Golden(name)
steps = int(sys.argv[2]) if len(sys.argv) > 2 else g.steps
from hybird_b200 import LB
lb = LB(dict(g.params)); lb.latticeBolzmannInit(*g.init_arrays()); g.configure(lb)
o = g.configure(lbo.Oracle(dict(g.params), *g.init_arrays()))
X, Y, Z = g.params["size"]
it_o = gu.replay(g, o, None)
for s, F, M, V, W in gu.replay(g, lb, None):
    _, Fo, Mo, Vo, Wo = next(it_o)
    a, b = common.gpu_state(lb), common.oracle_state(o)
    rep = common.compare(a, b)
    werr = float(np.abs(W - Wo).max() / max(np.abs(Wo).max(), 1e-300)) if W.size else 0.0
    ferr = float(np.abs(F - Fo).max() / max(np.abs(Fo).max(), 1e-300)) if F.size else 0.0
    bad = {k: v for k, v in rep.items() if v and (k.endswith("_rel") and v > 1e-12 or k.endswith("mismatch"))}
    if bad or werr > 1e-12 or ferr > 1e-9:
        print("step", s, bad, "wall force rel", werr, "element force rel", ferr)
        if W.size:
            print("W gpu", W.tolist()); print("W ref", Wo.tolist())
        if F.size:
            print("F gpu", F.tolist()); print("F ref", Fo.tolist()); print("V gpu", V.tolist(), "ref", Vo.tolist())
            hf = np.abs(a["hydroForce"] - b["hydroForce"]).max(axis=1)
            print("hydroForce cells differing:", int(np.count_nonzero(hf)), "of", int(np.count_nonzero(b["type_flags"] & 16)), "flagged")
        act = np.isin(b["type_flags"] & 15, (0, 3))
        for k in ("mass", "n", "visc", "f"):
            d = np.abs(a[k] - b[k]); d = d.reshape(len(act), -1).max(axis=1) * act
            idx = np.argsort(-d)[:6]
            for i in idx:
                if d[i] > 0:
                    nb = [int(b["type_flags"][i + dx + X * (dy + Y * dz)] & 15) for dx, dy, dz in zip(common.li.CX, common.li.CY, common.li.CZ)]
                    print("  ", k, "cell", int(i), "xyz", int(i % X), int(i // X % Y), int(i // (X * Y)), "type gpu/ref",
                          int(a["type_flags"][i]), int(b["type_flags"][i]), "gpu", a[k][i] if a[k].ndim == 1 else "-", "ref",
                          b[k][i] if b[k].ndim == 1 else "-", "diff", float(d[i]), "link types", nb)
        break
    if s >= steps:
        print("no divergence in", s, "steps")
        break
