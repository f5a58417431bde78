"""Whole-call duration (CUDA events around the device-resident check, stage timing off) vs batch size, with the box kernels
serial (ARTP_FORK_ITEMS=0) or side by side (ARTP_FORK_ITEMS=huge): which batch sizes gain from the fork."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np, torch
import art_planner_b200 as ap
from art_planner_b200 import synth
import bench
m, poses = bench.make_inputs(0, 1_000_000)
d = torch.from_numpy(poses.astype(np.float32)).cuda()
ref = None
for fi in ("0", "100000000"):
    os.environ["ARTP_FORK_ITEMS"] = fi
    chk = ap.StateValidityChecker(synth.PARAMS_YAML, device=0); chk.setMap(m); chk.updateHeightField()
    for n in (4096, 16384, 65536, 262144, 524288, 1_000_000):
        x = d[:n].contiguous(); out = torch.empty(n, dtype=torch.uint8, device="cuda")
        for _ in range(5): chk.isValidBatch(x, out)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(30): chk.isValidBatch(x, out)
        b.record(); torch.cuda.synchronize()
        print(f"fork_items {fi:>9s}  n {n:8d}  {a.elapsed_time(b)/30*1e3:8.1f} us per call  valid {int(out.sum())}")
