import sys, os, numpy as np
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import art_planner_b200 as ap
from art_planner_b200 import synth, costnet
import cases
m = cases.c4_map()
chk = ap.StateValidityChecker(synth.PARAMS_YAML, device=0); chk.setMap(m); chk.updateHeightField()
obj = ap.MotionCostObjective(chk); sd = costnet.make_state_dict(5); obj.setWeights(sd)
obj.setMode(1); obj.updateFeatures(); ref = obj.features()
obj.setMode(0)
for it in range(8):
    obj.updateFeatures(); got = obj.features()
    e = np.abs(got-ref); bad = (e.max(axis=2) > 1e-3)
    ys,xs = np.nonzero(bad)
    print(it, 'max err', float(e.max()), 'bad pixels', int(bad.sum()), 'tiles(y/16,x/8):', sorted(set(zip((ys//16).tolist(), (xs//8).tolist())))[:12], flush=True)
    if it == 3:
        q = costnet.make_queries(m, 64, 6); obj.costQuery(q)
