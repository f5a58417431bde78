import sys, os, numpy as np
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import art_planner_b200 as ap
from art_planner_b200 import synth, costnet
import cases
from oracle.cnn_oracle import CostNetOracle, cnn_input_from_layer
m = cases.c4_map()
chk = ap.StateValidityChecker(synth.PARAMS_YAML, device=0); chk.setMap(m); chk.updateHeightField()
obj = ap.MotionCostObjective(chk); sd = costnet.make_state_dict(5); obj.setWeights(sd)
orc = CostNetOracle(sd); feat = orc.features(cnn_input_from_layer(m.elevation))
q = costnet.make_queries(m, 4096, 6); lx,ly=m.length
ref = orc.query(feat, q, m.res, lx, ly, m.cx, m.cy)
for mode in (1,0):
    obj.setMode(mode); obj.updateFeatures(); got = obj.costQuery(q)
    d = np.abs(got-ref); rel = d/np.maximum(np.abs(ref),1e-30)
    print('mode',mode,'abs max per col',d.max(0),'worst rows',d.argmax(0), 'viol', int((d > 1e-4*np.abs(ref)+1e-5).sum()))
    i=d[:,0].argmax(); print(' row',i,'got',got[i],'ref',ref[i], 'q', q[i])
# also feed oracle with GPU features to isolate the head
import torch
g = torch.as_tensor(obj.features()).permute(2,0,1).contiguous()
ref2 = orc.query(g, q, m.res, lx, ly, m.cx, m.cy)
d=np.abs(got-ref2); print('head-only (oracle on GPU features): abs max per col', d.max(0))
