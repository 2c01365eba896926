"""CPU: the C restatement against the 1000-step hash fixtures of the free-surface minis (bit-exact: it keeps the
reference's serial summation order) -- north_star's "<= 1e-9 after 1000 steps" pinned on the checker itself."""
import numpy as np
import pytest

import common
import golden_full_util as gfu

NAMES = gfu.names("long")


def test_long_fixtures_present():
    assert len(NAMES) >= 8, "run tests/golden/make_golden_full.py where /root/reference exists"


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_for_1000_steps(name, oracle_lib):
    import lbo
    g = gfu.GoldenFull(name)
    o = g.configure(lbo.Oracle(dict(g.params), *g.init_arrays()))
    rep = gfu.check_run(g, o, lambda e: np.array(e.type_flags), common.oracle_state, exact=True, log=lambda *_: None)
    o.close()
    assert all(not rep[s]["not_bit_identical"] for s in g.dumps)


def test_oracle_matches_reference_at_configuration_size(oracle_lib):
    """cfg1 (the shipped lbmConfigDevisFluid.cfg case, 100x150x30) at FULL size: the host restatement of the
    reference's initialisation reproduces its initial state (hashes), the C restatement its first 100 steps."""
    import lbo
    if "cfg1_full" not in gfu.names("full"):
        pytest.skip("cfg1_full.npz not generated")
    g = gfu.GoldenFull("cfg1_full")
    o = g.configure(lbo.Oracle(dict(g.params), *g.init_arrays(), threads=4))
    rep = gfu.check_run(g, o, lambda e: np.array(e.type_flags), common.oracle_state, exact=True, steps=100, log=lambda *_: None)
    o.close()
    lbo.lib().lbo_set_threads(1)
    assert not rep[100]["not_bit_identical"]
