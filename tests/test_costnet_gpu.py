"""GPU parity (-m gpu) of the learned motion cost (BASELINE configs[3]): feature map and costs against the fp32
evaluation of the reference module (golden file produced from the reference's own network_light.py) and against the
torch restatement in oracle/cnn_oracle.py. Tolerance from BASELINE.json: 1e-4 relative (plus 1e-5 absolute, because
two of the three outputs pass through a ReLU and are exactly 0 for part of the batch)."""
import os

import numpy as np
import pytest

import cases
from art_planner_b200 import costnet

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL, ATOL = 1e-4, 1e-5


@pytest.fixture(scope="module")
def setup():
    import art_planner_b200 as ap
    from art_planner_b200 import build, synth
    from oracle.cnn_oracle import CostNetOracle, cnn_input_from_layer
    build.build()
    m = cases.c4_map()
    chk = ap.StateValidityChecker(synth.PARAMS_YAML, device=0)
    chk.setMap(m)
    chk.updateHeightField()
    obj = ap.MotionCostObjective(chk)
    sd = costnet.make_state_dict(seed=5)
    obj.setWeights(sd)
    orc = CostNetOracle(sd)
    feat = orc.features(cnn_input_from_layer(m.elevation))
    golden = np.load(os.path.join(ROOT, "tests", "golden", "cnn_c4.npz"))
    assert abs(float(costnet.pack_blob(sd).astype(np.float64).sum()) - float(golden["blob_sum"])) < 1e-9, "weight generator drift"
    return m, obj, orc, feat, golden


@pytest.mark.parametrize("mode", [1, 0, 4, 8, 12], ids=["cuda-core", "tcgen05-two-phase", "tcgen05-single-phase", "tcgen05-multicast-pairs", "tcgen05-two-issuers"])
def test_feature_map(setup, mode):
    m, obj, orc, feat, golden = setup
    obj.setMode(mode)
    obj.updateFeatures()
    got = obj.features()                      # [Hf, Wf, 48]
    ref = feat.permute(1, 2, 0).numpy()
    assert got.shape == ref.shape == (104, 104, 48)
    scale = float(np.abs(ref).max())
    err = float(np.abs(got - ref).max()) / scale
    # 10 800-term fp32 accumulations in different orders (and the tensor core's fp32 accumulator): 1e-4 of the max
    tol = 1e-5 if mode == 1 else 1e-4
    assert err < tol, f"feature map max error {err:.3e} of max |f| = {scale:.3f}"
    assert np.allclose(got[::13, ::13].transpose(2, 0, 1), golden["feat_sample"], rtol=1e-4, atol=tol * scale)
    obj.setMode(0)


def test_costs_match_reference_module(setup):
    import torch
    m, obj, orc, feat, golden = setup
    obj.setMode(0)
    obj.updateFeatures()
    q = costnet.make_queries(m, 4096, seed=6)
    got = obj.costQuery(q)
    ref = golden["cost"]
    assert np.allclose(got, ref, rtol=RTOL, atol=ATOL), float(np.abs(got - ref).max())
    lx, ly = m.length
    assert np.allclose(got, orc.query(feat, q, m.res, lx, ly, m.cx, m.cy), rtol=RTOL, atol=ATOL)
    dev = obj.costQuery(torch.from_numpy(q).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy(), got)
    cost, feas = obj.getCost(got)
    assert np.allclose(cost, got[:, 1] * 1.0 + got[:, 2] * 5.0, rtol=1e-6)
    assert np.array_equal(feas, (got[:, 2] <= 0.5).astype(np.uint8))
    assert obj.costQuery(np.zeros((0, 6), np.float32)).shape == (0, 3)


def test_errors_without_weights_or_features():
    import art_planner_b200 as ap
    from art_planner_b200 import synth
    chk = ap.StateValidityChecker(synth.PARAMS_YAML, device=0)
    obj = ap.MotionCostObjective(chk)
    with pytest.raises(ap.ArtpError):
        obj.updateFeatures()            # no map
    chk.setMap(cases.c4_map()); chk.updateHeightField()
    with pytest.raises(ap.ArtpError):
        obj.updateFeatures()            # no weights
    obj.setWeights(costnet.make_state_dict(seed=5))
    with pytest.raises(ap.ArtpError):
        obj.costQuery(np.zeros((4, 6), np.float32))   # features not computed
