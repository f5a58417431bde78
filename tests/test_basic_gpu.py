"""GPU parity (-m gpu): processors::Basic on the device (artp_process_basic) == the CPU restatement, bit for bit, and == the
golden layers produced through OpenCV."""
import os

import numpy as np
import pytest

import cases
from art_planner_b200 import synth
from oracle import basic_oracle as bo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("case", cases.BASIC_CASES, ids=[c[0] for c in cases.BASIC_CASES])
def test_device_masked_elevation_bit_exact(case, maps):
    import art_planner_b200 as ap
    name, mk, scale, p = case
    gold = np.load(os.path.join(ROOT, "tests", "golden", "basic_masks.npz"))
    m = maps(mk)
    trav, obs = synth.make_traversability(m, seed=13)
    chk = ap.StateValidityChecker(cases.PARAMS["yaml"], device=0)
    masked, thr = chk.processBasic(m.elevation, trav, obs, m.res * scale, p)
    ref_m, ref_t = bo.masked_elevation(m.elevation, trav, obs, m.res * scale, p)
    assert np.array_equal(masked.view(np.uint32), ref_m.view(np.uint32))       # bit-exact incl. the -inf pattern
    assert np.array_equal(thr, ref_t)
    n = masked.size
    assert np.array_equal(np.isfinite(masked).ravel(order="F"), np.unpackbits(gold[name + "/finite"])[:n].astype(bool))
    # the produced layer feeds the checker like any other elevation_masked
    import dataclasses
    m2 = dataclasses.replace(m, elevation_masked=masked)
    chk.setMap(m2); chk.updateHeightField()
    poses = synth.make_terrain_poses(m2, 4000, seed=5)
    from oracle import orc
    o = orc.Oracle(cases.PARAMS["yaml"], "port"); o.set_map(m2)
    assert np.array_equal(chk.isValidBatch(poses), o.check_poses(poses))
