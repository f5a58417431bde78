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


@pytest.mark.parametrize("mode", [0, 1], ids=["warp+group", "group-only"])
@pytest.mark.parametrize("case", cases.INTERIOR_CASES, ids=[c[0] for c in cases.INTERIOR_CASES])
def test_edge_interiors_bit_exact(case, mode, golden, maps, make_checker):
    """addValidMilestone's connection loop (prm_motion_cost.cpp:341-372): leading valid interior states per edge."""
    import torch
    import art_planner_b200 as ap
    name, mk, pk, n, seed, dmin, dmax = case
    m = maps(mk)
    chk = make_checker(pk, m)
    chk.setMode(mode)
    s1, s2 = synth.make_edges(m, n, seed, dmin=dmin, dmax=dmax)
    ref = golden[name + "/prefix"].astype(np.int32)
    mv = ap.MotionValidator(chk)
    got, ni = mv.checkEdgeInteriors(s1, s2)                 # counts derived inside the library (n_interp == NULL)
    assert np.array_equal(got, ref)
    got2, _ = mv.checkEdgeInteriors(s1, s2, n_interp=ni)    # caller-provided counts
    assert np.array_equal(got2, ref)
    dev, ni_d = mv.checkEdgeInteriors(torch.from_numpy(s1).cuda(), torch.from_numpy(s2).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(ni_d.cpu().numpy(), ni)
    assert np.array_equal(dev.cpu().numpy(), ref)
    chk.setMode(0)


def test_edge_interiors_against_per_state_checks(maps, make_checker, port_lib):
    """Size-independent property: the prefix equals what isValidBatch says about the interpolated states (oracle
    interpolation), on a batch larger than one 2^20-item round."""
    import art_planner_b200 as ap
    m = maps("fbm_rough")
    chk = make_checker("yaml", m)
    n = 400_000
    s1, s2 = synth.make_edges(m, n, 5, dmin=0.05, dmax=3.4)
    got, ni = ap.MotionValidator(chk).checkEdgeInteriors(s1, s2)
    assert int(ni.sum()) > (1 << 20)
    assert (got <= ni).all()
    # edges with zero interior states are trivially valid connections
    assert (got[ni == 0] == 0).all()
    # cross-check a slice against the CPU oracle
    o = port_lib.Oracle(cases.PARAMS["yaml"], "port")
    o.set_map(m)
    sl = slice(n - 3000, n)
    assert np.array_equal(got[sl], o.check_edge_interiors(s1[sl], s2[sl], None, 0.5))


def test_edge_interiors_empty_and_errors(maps, make_checker):
    import art_planner_b200 as ap
    m = maps("flat")
    chk = make_checker("yaml", m)
    mv = ap.MotionValidator(chk)
    got, ni = mv.checkEdgeInteriors(np.zeros((0, 7)), np.zeros((0, 7)))
    assert got.shape == (0,)
    s1, s2 = synth.make_edges(m, 10, 3)
    got, _ = mv.checkEdgeInteriors(s1, s2, n_interp=np.zeros(10, np.int32))
    assert (got == 0).all()
    with pytest.raises(RuntimeError):
        mv.checkEdgeInteriors(s1, s2, n_interp=np.full(10, -1, np.int32))


def test_single_edge_latency_path(maps, make_checker, port_lib):
    """One checkMotion call at a time (<= 64 states) takes the fused one-launch path: same flags as the oracle."""
    import art_planner_b200 as ap
    m = maps("fbm_rough")
    chk = make_checker("yaml", m)
    o = port_lib.Oracle(cases.PARAMS["yaml"], "port")
    o.set_map(m)
    s1, s2 = synth.make_edges(m, 300, 77)
    for steps in (0, 7, 20, 63):
        ref = o.check_motions(s1, s2, steps)
        mv = ap.MotionValidator(chk, steps)
        got = np.array([mv.checkMotion(s1[i], s2[i]) for i in range(len(s1))], dtype=np.uint8)
        assert np.array_equal(got, ref), steps
    ref = o.check_motions(s1, s2, 20)
    mv = ap.MotionValidator(chk, 20)
    assert np.array_equal(np.concatenate([mv.checkMotionBatch(s1[i:i + 3], s2[i:i + 3]) for i in range(0, 300, 3)]), ref)


@pytest.mark.parametrize("mode", [0, 1], ids=["warp+group", "group-only"])
@pytest.mark.parametrize("case", cases.SEGMENT_CASES, ids=[c[0] for c in cases.SEGMENT_CASES])
def test_motion_segments_and_last_valid(case, mode, golden, maps, make_checker):
    """artp_check_motions_segments: OMPL's validSegmentCount rule per edge, verdicts and lastValid.second bit-exact against
    the compiled reference (golden), with the counts given and with the counts computed from the space parameters."""
    import art_planner_b200 as ap
    name, mk, pk, n, seed, dmin, dmax = case
    m = maps(mk)
    chk = make_checker(pk, m)
    chk.setMode(mode)
    s1, s2 = synth.make_edges(m, n, seed, dmin=dmin, dmax=dmax)
    mv = ap.MotionValidator(chk)
    sp = mv.se3Space(m, cases.PARAMS[pk].reach_z)
    nd = mv.validSegmentCount(sp, s1, s2)
    assert np.array_equal(nd, golden[name + "/nd"])
    ref_v, ref_t = unpack(golden, name + "/mask", n), golden[name + "/last_t"]
    for kw in (dict(nd=nd), dict(space=sp)):
        v, t = mv.checkMotionSegments(s1, s2, **kw)
        assert np.array_equal(v, ref_v)
        assert np.array_equal(t, ref_t)
    assert 0 < ref_v.sum() < n
