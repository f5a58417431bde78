"""compute-sanitizer driver for the kernels added after the first sanitizer pass: latency path (pose_small_kernel, poses and
single edges, both grouping modes), edge interiors, sampler / fused sample->check->compact, estimate_normals, CDF,
bit packing / bit compaction."""
import os, sys, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import art_planner_b200 as ap
from art_planner_b200 import synth
import cases
for mk in ("fixture", "flat_holes_terrace"):
    m = cases.MAPS[mk]()
    chk = ap.StateValidityChecker(cases.PARAMS["yaml"], device=0); chk.setMap(m); chk.updateHeightField()
    poses = synth.make_terrain_poses(m, 2000, seed=9)
    big = chk.isValidBatch(poses)
    for mode in (0, 1):
        chk.setMode(mode)
        for n, off in ((1, 0), (7, 10), (64, 100), (33, 300)):
            assert np.array_equal(chk.isValidBatch(poses[off:off + n]), big[off:off + n])
    chk.setMode(0)
    s1, s2 = synth.make_edges(m, 400, 5, dmin=0.05, dmax=3.0)
    mv = ap.MotionValidator(chk, 6)
    ref = mv.checkMotionBatch(s1, s2)
    assert all(mv.checkMotion(s1[i], s2[i]) == bool(ref[i]) for i in range(40))
    pref, ni = mv.checkEdgeInteriors(s1, s2)
    d1, d2 = torch.from_numpy(s1).cuda(), torch.from_numpy(s2).cuda()
    pd, _ = mv.checkEdgeInteriors(d1, d2)
    assert np.array_equal(pd.cpu().numpy(), pref)
    L = synth.make_sampler_layers(m, seed=7)
    chk.estimateNormals(0.49)
    chk.computeSampleCdf(L.sample_probability)

    class Nothing:
        pass
    for fd in (True, False):
        smp = ap.SE3FromSE2Sampler(chk, Nothing if fd else L, synth.sampler_params_for(m, fd), seed=3)
        st = smp.sampleUniformBatch(3000, first=0)
        got, nv = smp.sampleValidBatch(5000, first=0)
        assert len(got) == nv
    v = torch.from_numpy(big).cuda()
    bits = chk.packValidBits(v)
    idx, cnt = chk.compactBits(bits, len(big))
    torch.cuda.synchronize()
    assert int(cnt.item()) == int(big.sum())
    print(mk, "ok", int(big.sum()), nv)
