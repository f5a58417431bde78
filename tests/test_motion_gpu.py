"""GPU parity (-m gpu) for edge validity (MotionValidator batch) and the closed-form edge cost."""
import numpy as np
import pytest

import cases
from art_planner_b200 import synth

pytestmark = pytest.mark.gpu


def unpack(golden, key, n):
    return np.unpackbits(golden[key])[:n]


@pytest.fixture(scope="module")
def make_checker():
    import art_planner_b200 as ap
    from art_planner_b200 import build
    build.build()

    def mk(pk, m):
        c = ap.StateValidityChecker(cases.PARAMS[pk], device=0)
        c.setMap(m)
        c.updateHeightField()
        return c
    return mk


@pytest.mark.parametrize("mode", [0, 1], ids=["warp+group", "group-only"])
@pytest.mark.parametrize("case", cases.EDGE_CASES, ids=[c[0] for c in cases.EDGE_CASES])
def test_edge_masks_bit_exact(case, mode, golden, maps, make_checker):
    import art_planner_b200 as ap
    name, mk, pk, n, steps, seed = case
    m = maps(mk)
    chk = make_checker(pk, m)
    chk.setMode(mode)
    s1, s2 = synth.make_edges(m, n, seed)
    got = ap.MotionValidator(chk, steps).checkMotionBatch(s1, s2)
    ref = unpack(golden, name + "/mask", n)
    bad = np.nonzero(got != ref)[0]
    assert bad.size == 0, f"{bad.size} mismatches, first {bad[:8]}"


def test_zero_steps_equals_endpoint_check(maps, make_checker):
    import art_planner_b200 as ap
    m = maps("fixture")
    chk = make_checker("yaml", m)
    s1, s2 = synth.make_edges(m, 2000, 77)
    assert np.array_equal(ap.MotionValidator(chk, 0).checkMotionBatch(s1, s2), chk.isValidBatch(s2))
    assert bool(ap.MotionValidator(chk, 3).checkMotion(s1[0], s2[0])) == bool(
        ap.MotionValidator(chk, 3).checkMotionBatch(s1[:1], s2[:1])[0])


@pytest.mark.parametrize("case", cases.EDGE_CASES, ids=[c[0] for c in cases.EDGE_CASES])
def test_path_length_cost(case, golden, maps, make_checker):
    import art_planner_b200 as ap
    name, mk, pk, n, steps, seed = case
    m = maps(mk)
    chk = make_checker(pk, m)
    s1, s2 = synth.make_edges(m, n, seed)
    got = ap.PathLengthObjective(chk).motionCostBatch(s1, s2)
    ref = golden[name + "/cost"]
    if cases.PARAMS[pk].use_directional_cost:
        # directional cost goes through atan2/sin/cos (libm vs CUDA math differ in the last ulp): 1e-12 relative
        assert np.allclose(got, ref, rtol=1e-12, atol=0)
    else:
        assert np.array_equal(got, ref)   # sqrt/div only: bit-exact


def test_edges_device_buffers(maps, make_checker):
    import torch
    import art_planner_b200 as ap
    m = maps("fbm_rough")
    chk = make_checker("yaml", m)
    s1, s2 = synth.make_edges(m, 5000, 41)
    mv = ap.MotionValidator(chk, 20)
    host = mv.checkMotionBatch(s1, s2)
    dev = mv.checkMotionBatch(torch.from_numpy(s1).cuda(), torch.from_numpy(s2).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(host, dev.cpu().numpy())
