import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import art_planner_b200 as ap
from art_planner_b200 import synth, costnet
import cases
m = cases.c4_map()
chk = ap.StateValidityChecker(synth.PARAMS_YAML, device=0); chk.setMap(m); chk.updateHeightField()
obj = ap.MotionCostObjective(chk)
base = costnet.make_state_dict(seed=5)
def run(sd, mode):
    obj.setWeights(sd); obj.setMode(mode); obj.updateFeatures(); return obj.features()
def taps_only(taps):
    sd = dict(base); w = np.zeros_like(base['init_flatten.weight'])
    for (ky,kx) in taps: w[:,:,ky,kx] = base['init_flatten.weight'][:,:,ky,kx]
    sd['init_flatten.weight'] = w; return sd
for name,taps in [('tap00',[(0,0)]),('tap01',[(0,1)]),('tap07',[(0,7)]),('tap08',[(0,8)]),('tap10',[(1,0)]),('tap11',[(1,1)]),('tap(14,14)',[(14,14)]),('row0',[(0,k) for k in range(15)]),('all',[(a,b) for a in range(15) for b in range(15)])]:
    sd = taps_only(taps)
    ref = run(sd, 1)
    out = []
    for mode in (0, 2):
        got = run(sd, mode)
        err = np.abs(got-ref)
        out.append((mode, float(err.max()), float(np.abs(ref).max()), [int(x) for x in np.unravel_index(err.argmax(), err.shape)]))
    print(name, out, flush=True)
# pattern for tap00, mode 0
sd = taps_only([(0,0)]); ref = run(sd,1); got = run(sd,0)
e = np.abs(got-ref).max(axis=2)
print('tap00 err by pixel (first tile 16x8):'); print(np.round(e[:16,:8],3))
print('got/ref channel 0 first tile:'); print(np.round(got[:4,:8,0],3)); print(np.round(ref[:4,:8,0],3))
print('err by channel:', np.round(np.abs(got-ref).max(axis=(0,1)),3))
