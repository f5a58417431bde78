"""Small pose batch through the pipeline (for compute-sanitizer / debugging)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np
import art_planner_b200 as ap
from art_planner_b200 import synth
from oracle import orc
m = synth.make_fbm_map(300, 300, seed=2)
chk = ap.StateValidityChecker(synth.PARAMS_YAML, device=0); chk.setMap(m); chk.updateHeightField()
poses = synth.make_terrain_poses(m, int(sys.argv[1]) if len(sys.argv) > 1 else 3000, seed=3)
got = chk.isValidBatch(poses)
o = orc.Oracle(synth.PARAMS_YAML, "port"); o.set_map(m)
ref = o.check_poses(poses)
print("mismatches", int((got != ref).sum()), chk.stats())
