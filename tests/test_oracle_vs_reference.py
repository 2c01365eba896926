"""CPU, only where oracle/_ref/ref_harness exists (built from /root/reference by oracle/Makefile):
run the unmodified reference live next to the C restatement on identical inputs and demand
bit-identical states.  This is what pins the oracle; tests/golden/ holds the travelling copy."""
import os
import sys

import pytest

import common

sys.path.insert(0, os.path.join(common.ROOT, "oracle"))
import cases  # noqa: E402

pytestmark = pytest.mark.reference

LIVE = ["cfg2_mini", "channel_oblique", "bingham_smago", "slip_dyn", "couette_dyn", "cfg3_mini", "two_spheres_kin",
        "cluster_dem", "cfg4_mini", "bubble_periodic", "cfg1_mini", "cfg5_mini", "drum_mini", "drum_bingham"]


@pytest.mark.parametrize("name", LIVE)
def test_restatement_bit_identical_to_reference(name, oracle_lib, tmp_path):
    if not os.path.exists(cases.REF_HARNESS):
        pytest.skip("oracle/_ref/ref_harness not built (no /root/reference here)")
    import check_vs_ref
    rep = check_vs_ref.run_case(cases.catalogue()[name], 25, workdir=str(tmp_path), dumps=(0, 1, 13, 25), verbose=False)
    assert rep["type_map_mismatch_steps"] == 0
    assert rep["force_rel"] == 0.0 and rep["wall_rel"] == 0.0
    for s, r in rep["states"].items():
        bad = {k: v for k, v in r.items() if v}
        assert not bad, "step %d: %s" % (s, bad)
