"""CPU: the C restatement (oracle/lb_oracle.c) against the golden vectors generated from the
unmodified reference (tests/golden/make_golden.py).  Bit-exact: type maps after every step,
element / wall forces every step, sha256 of every field over the active cells at the check steps."""
import numpy as np
import pytest

import common
import golden_util as gu

NAMES = gu.names()


def test_fixtures_present():
    assert len(NAMES) >= 22, "golden fixtures missing: run tests/golden/make_golden.py where /root/reference exists"


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_golden(name, oracle_lib):
    import lbo
    g = gu.Golden(name)
    tf, si, n, u, mass, visc = g.init_arrays()
    prm = dict(g.params)
    o = g.configure(lbo.Oracle(prm, tf, si, n, u, mass, visc))  # f=None: equilibrium of (n,u), node::initialize
    active0 = np.isin(tf & 0x0F, (0, 3))
    assert gu.sha(np.array(o.f)[active0] + 0.0) == str(g.z["sha_init_f"]), "initial populations differ"
    assert np.array_equal(o.type_flags & 0x1F, g.types[0])
    for s, F, M, V, W in gu.replay(g, o, None):
        assert np.array_equal(o.type_flags & 0x1F, g.types[s]), "type map differs after step %d" % s
        if g.forces:
            rF, rM, rV, rW = g.forces[s - 1]
            assert np.array_equal(F, rF) and np.array_equal(M, rM) and np.array_equal(V, rV), "element forces, step %d" % s
            assert np.array_equal(W, rW), "wall forces, step %d" % s
        if s in g.check_steps:
            mine = gu.state_hashes(common.oracle_state(o))
            ref = g.hashes(s)
            bad = [k for k in gu.FIELDS if mine[k] != ref[k]]
            assert not bad, "step %d: fields differ from the reference: %s" % (s, bad)
        if s == g.steps:
            assert np.array_equal(np.array(o.n)[np.isin(o.type_flags & 15, (0, 3))],
                                  g.z["final_n"][np.isin(g.types[s] & 15, (0, 3))])
    o.close()


def test_oracle_threads_agree(oracle_lib):
    """The restatement's OpenMP loops are order-free: 1 thread == 4 threads, bit for bit."""
    import lbo
    g = gu.Golden("cfg4_mini")
    res = []
    for th in (1, 4):
        o = lbo.Oracle(dict(g.params), *g.init_arrays(), threads=th)
        for _ in gu.replay(g, o, None):
            pass
        res.append(gu.state_hashes(common.oracle_state(o)))
        o.close()
    lbo.lib().lbo_set_threads(1)
    assert res[0] == res[1]
