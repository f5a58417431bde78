"""Stage durations (library CUDA events) of the device-resident pose check vs batch size: the fixed cost of each stage."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np, torch
import art_planner_b200 as ap
from art_planner_b200 import synth
import bench
m, poses = bench.make_inputs(0, 1_000_000)
chk = ap.StateValidityChecker(synth.PARAMS_YAML, device=0); chk.setMap(m); chk.updateHeightField()
d = torch.from_numpy(poses.astype(np.float32)).cuda()
chk.setTiming(True)
for n in (40_000, 100_000, 300_000, 600_000, 1_000_000):
    x = d[:n].contiguous(); out = torch.empty(n, dtype=torch.uint8, device="cuda")
    acc = np.zeros(5)
    for i in range(13):
        chk.isValidBatch(x, out); torch.cuda.synchronize()
        if i >= 3: acc += np.array(chk.lastStageTimesMs())
    acc /= 10
    st = chk.stats()
    print(f"n {n:8d}  classify {acc[0]:.4f}  torso {acc[1]:.4f}  reach_warp {acc[2]:.4f}  reach_groups {acc[3]:.4f}  block {acc[4]:.4f}  sum {acc.sum():.4f} ms"
          f"  queued W/F/G {st['last_queued_warp_stage']}/{st['last_queued_reach_stage']}/{st['last_reach_plane_stage']}")
