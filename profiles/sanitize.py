"""Small end-to-end run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import art_planner_b200 as ap
from art_planner_b200 import synth, costnet
import cases
for mk, pk in (("fixture", "yaml"), ("ramp", "header")):
    m = cases.MAPS[mk]()
    chk = ap.StateValidityChecker(cases.PARAMS[pk], device=0); chk.setMap(m); chk.updateHeightField()
    poses = synth.make_terrain_poses(m, 3000, seed=9)
    a = chk.isValidBatch(poses); chk.setMode(1); b = chk.isValidBatch(poses[:600]); chk.setMode(0)
    assert np.array_equal(a[:600], b)
    s1, s2 = synth.make_edges(m, 300, 5)
    ap.MotionValidator(chk, 5).checkMotionBatch(s1, s2); ap.PathLengthObjective(chk).motionCostBatch(s1, s2)
    print(mk, int(a.sum()), chk.stats())
m = cases.c4_map()
chk = ap.StateValidityChecker(synth.PARAMS_YAML, device=0); chk.setMap(m); chk.updateHeightField()
obj = ap.MotionCostObjective(chk); obj.setWeights(costnet.make_state_dict(5)); obj.updateFeatures()
print("cnn", obj.costQuery(costnet.make_queries(m, 256, 6)).sum(0))
