"""Minimal driver for ncu: N device-resident passes of the bench.py pose workload (1M poses, 1000x1000 map)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
import art_planner_b200 as ap
from art_planner_b200 import synth
import bench
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 4
if len(sys.argv) > 2 and sys.argv[2] == "rough":
    m = synth.make_fbm_map(bench.MAP_N, bench.MAP_N, bench.MAP_RES, seed=bench.MAP_SEED, **bench.ROUGH_MAP)
    poses = synth.make_terrain_poses(m, bench.POSES_PER_GPU, seed=bench.POSE_SEED, **bench.ROUGH_POSES)
else:
    m, poses = bench.make_inputs(0, bench.POSES_PER_GPU)
chk = ap.StateValidityChecker(synth.PARAMS_YAML, device=0); chk.setMap(m); chk.updateHeightField()
d = torch.from_numpy(poses).cuda(); v = torch.empty(len(poses), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(n_iter):
    flush.fill_(i)
    chk.isValidBatch(d, out=v)
torch.cuda.synchronize()
print("valid fraction", float(v.float().mean()), chk.stats())
