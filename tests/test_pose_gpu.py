"""GPU parity (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle port and the golden masks
of the reference's compiled ODE. Bit-exact valid/invalid flags are required."""
import numpy as np
import pytest

import cases
from art_planner_b200 import synth

pytestmark = pytest.mark.gpu


def unpack(golden, key, n):
    return np.unpackbits(golden[key])[:n]


@pytest.fixture(scope="module")
def checkers():
    import art_planner_b200 as ap
    from art_planner_b200 import build
    build.build()
    cache = {}

    def get(pk):
        if pk not in cache:
            cache[pk] = ap.StateValidityChecker(cases.PARAMS[pk], device=0)
        return cache[pk]
    return get


def set_map(chk, m):
    chk.setMap(m)
    chk.updateHeightField()
    assert chk.hasMap()


@pytest.mark.parametrize("mode", [0, 1], ids=["warp+group", "group-only"])
@pytest.mark.parametrize("case", cases.POSE_CASES, ids=[c[0] for c in cases.POSE_CASES])
def test_pose_masks_bit_exact(case, mode, golden, maps, port_lib, checkers):
    name, mk, pk, gen = case
    m = maps(mk)
    poses = gen(m)
    chk = checkers(pk)
    set_map(chk, m)
    chk.setMode(mode)
    try:
        got = chk.isValidBatch(poses)
        st = chk.stats()
    finally:
        chk.setMode(0)
    ref = unpack(golden, name + "/mask", len(got))
    o = port_lib.Oracle(cases.PARAMS[pk], "port")
    o.set_map(m)
    port = o.check_poses(poses)
    assert np.array_equal(port, ref)
    bad = np.nonzero(got != ref)[0]
    assert bad.size == 0, f"{bad.size} mismatches, first {bad[:8]}, deferred={st['last_deferred']}"
    if mode == 1:
        assert st["last_deferred"] == st["last_queued_boxes"] > 0


def test_device_buffers_and_host_buffers_agree(maps, checkers):
    import torch
    m = maps("fbm_rough")
    chk = checkers("yaml")
    set_map(chk, m)
    poses = synth.make_terrain_poses(m, 300000, seed=99)     # > one H2D slice: exercises the copy/compute pipeline
    host = chk.isValidBatch(poses)
    dev = chk.isValidBatch(torch.from_numpy(poses).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(host, dev.cpu().numpy())
    # float32 states (the cast Pose3FromSE3 performs first, done by the caller): identical flags
    p32 = poses.astype(np.float32)
    assert np.array_equal(chk.isValidBatch(p32), host)
    d32 = chk.isValidBatch(torch.from_numpy(p32).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(d32.cpu().numpy(), host)


def test_single_state_latency_path_and_edge_cases(maps, port_lib, checkers):
    m = maps("fixture")
    chk = checkers("yaml")
    set_map(chk, m)
    o = port_lib.Oracle(cases.PARAMS["yaml"], "port")
    o.set_map(m)
    poses = synth.make_terrain_poses(m, 64, seed=5)
    ref = o.check_poses(poses)
    for i in range(64):
        assert chk.isValid(poses[i]) == bool(ref[i])
    assert chk.isValidBatch(np.zeros((0, 7))).shape == (0,)
    # far outside the map: torso "valid", feet decided by unknown_space_untraversable (=True -> invalid)
    far = np.array([[1e3, -1e3, 0.0, 0, 0, 0, 1.0]])
    assert np.array_equal(chk.isValidBatch(far), o.check_poses(far))
    # ragged batch sizes around warp/CTA multiples
    for n in (1, 31, 33, 255, 257):
        assert np.array_equal(chk.isValidBatch(poses[:n] if n <= 64 else np.resize(poses, (n, 7))),
                              o.check_poses(poses[:n] if n <= 64 else np.resize(poses, (n, 7))))


def test_no_map_is_an_error(checkers):
    import art_planner_b200 as ap
    chk = ap.StateValidityChecker(cases.PARAMS["yaml"], device=0)
    assert not chk.hasMap()
    with pytest.raises(ap.ArtpError):
        chk.isValidBatch(np.zeros((4, 7)))


def test_compaction_is_ordered(maps, checkers):
    import torch
    m = maps("fbm_gentle")
    chk = checkers("yaml")
    set_map(chk, m)
    poses = torch.from_numpy(synth.make_terrain_poses(m, 50001, seed=17)).cuda()
    valid = chk.isValidBatch(poses)
    idx, cnt = chk.compactValid(valid, base=1000)
    torch.cuda.synchronize()
    v = valid.cpu().numpy()
    n = int(cnt.item())
    assert n == int(v.sum())
    assert np.array_equal(idx[:n].cpu().numpy(), np.nonzero(v)[0] + 1000)


def test_full_size_properties_c2(checkers):
    """BASELINE configs[1] size (1000x1000 map, 1 M samples): size-independent properties --
    idempotence, permutation equivariance, host/device agreement on a slice -- plus a strided oracle check."""
    import torch
    from oracle import orc
    m = synth.make_fbm_map(1000, 1000)
    chk = checkers("yaml")
    set_map(chk, m)
    n = 1_000_000
    poses = synth.make_terrain_poses(m, n, seed=3)
    d = torch.from_numpy(poses).cuda()
    a = chk.isValidBatch(d).clone()
    b = chk.isValidBatch(d).clone()
    perm = torch.randperm(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    c = chk.isValidBatch(d[perm].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    assert torch.equal(a[perm], c)
    o = orc.Oracle(cases.PARAMS["yaml"], "port")
    o.set_map(m)
    sel = np.arange(0, n, 50)
    assert np.array_equal(a.cpu().numpy()[sel], o.check_poses(poses[sel]))
    assert 0.05 < float(a.float().mean()) < 0.95


def test_batched_rejection_sampling(maps, checkers, port_lib):
    """sampleValidBatch keeps exactly the valid candidates, in draw order (rejection loop, prm_motion_cost.cpp:171-194)."""
    m = maps("fbm_rough")
    chk = checkers("yaml")
    set_map(chk, m)
    pool = synth.make_terrain_poses(m, 6000, seed=99)
    cur = [0]

    def sampler(k):
        out = pool[cur[0]:cur[0] + k]
        cur[0] += k
        return out
    got, drawn = chk.sampleValidBatch(sampler, 500, batch=512, max_draws=len(pool))
    o = port_lib.Oracle(cases.PARAMS["yaml"], "port")
    o.set_map(m)
    ref = o.check_poses(pool)
    want = pool[ref != 0][:500]
    assert np.array_equal(got, want)
    assert drawn % 512 == 0 or drawn == len(pool)


def test_bit_packed_mask_and_compaction(maps, checkers):
    """artp_pack_valid_bits_device / artp_compact_bits_device == the torch restatements (multi-GPU exchange format)."""
    import torch
    from art_planner_b200 import sharding
    chk = checkers("yaml")
    g = torch.Generator().manual_seed(5)
    for n in (1, 31, 32, 33, 4097, 1_000_003):
        v = (torch.rand(n, generator=g) < 0.37).to(torch.uint8)
        bits = chk.packValidBits(v.cuda())
        torch.cuda.synchronize()
        assert torch.equal(bits.cpu(), sharding.pack_bits_reference(v))
        idx, cnt = chk.compactBits(bits, n, base=1000)
        torch.cuda.synchronize()
        want = torch.nonzero(v).reshape(-1) + 1000
        assert int(cnt.item()) == want.numel() and torch.equal(idx[: want.numel()].cpu(), want)


@pytest.mark.parametrize("mk", ["fbm_rough", "flat_holes_terrace", "fixture"])
def test_latency_path_small_batches(maps, checkers, port_lib, mk):
    """n <= 16 host states take the one-launch latency path (pose_small_kernel): same verdicts as the oracle, in both
    grouping modes, through the double and the float entry points."""
    m = maps(mk)
    chk = checkers("yaml")
    set_map(chk, m)
    o = port_lib.Oracle(cases.PARAMS["yaml"], "port")
    o.set_map(m)
    poses = synth.make_terrain_poses(m, 600, seed=123)
    ref = o.check_poses(poses)
    assert 0 < ref.sum() < len(ref)
    for mode in (0, 1):
        chk.setMode(mode)
        pos = 0
        for n in list(range(1, 17)) * 2 + [17, 33, 63, 64, 64]:
            chunk = poses[pos:pos + n]
            assert np.array_equal(chk.isValidBatch(chunk), ref[pos:pos + n]), (mode, n, pos)
            assert np.array_equal(chk.isValidBatch(chunk.astype(np.float32)), ref[pos:pos + n]), (mode, n, pos)
            pos += n
    chk.setMode(0)
    far = np.array([[1e3, -1e3, 0.0, 0, 0, 0, 1.0]])           # outside the map: the outside-map rules
    assert np.array_equal(chk.isValidBatch(far), o.check_poses(far))


def test_terraces_exercise_the_deferral_path(maps, golden, checkers):
    """Normal mode on piecewise-constant terrain: the corner-candidate shortcut of the warp stage must hand boxes with an
    epsilon-mergeable earlier plane to the exact grouping stage (heightfield.cpp:1511-1556) -- assert that path runs."""
    m = maps("terraces")
    chk = checkers("yaml")
    set_map(chk, m)
    chk.setMode(0)
    name, _, _, gen = [c for c in cases.POSE_CASES if c[0] == "terraces_low_yaml"][0]
    poses = gen(m)
    got = chk.isValidBatch(poses)
    st = chk.stats()
    assert st["last_queued_boxes"] > 0 and st["last_deferred"] > 0, st
    assert np.array_equal(got, unpack(golden, name + "/mask", len(got)))


def test_group_stage_overflow_is_reported_by_the_call_that_caused_it(maps, golden):
    """A zone that does not fit the plane store: the pose is reported invalid (fail closed) and the host-buffer call itself
    returns ARTP_E_LIMIT; asynchronous calls surface it through pollError (ADVICE)."""
    import torch
    import art_planner_b200 as ap
    from art_planner_b200 import capi
    m = maps("terraces")
    chk = ap.StateValidityChecker(cases.PARAMS["yaml"], device=0)
    chk.debugSetGroupCapacity(64)          # far below a torso zone's ~2000 triangles
    set_map(chk, m)
    chk.setMode(1)                         # every in-map box goes through the grouping stage
    poses = synth.make_terrain_poses(m, 5000, seed=31)
    with pytest.raises(ap.ArtpError) as ei:
        chk.isValidBatch(poses)
    assert ei.value.code == capi.ARTP_E_LIMIT
    chk.pollError()                        # sticky word was consumed by the failing call
    d = chk.isValidBatch(torch.from_numpy(poses).cuda())
    torch.cuda.synchronize()
    with pytest.raises(ap.ArtpError):
        chk.pollError()
    ref = unpack(golden, "terraces_yaml/mask", 20000)[:5000]
    got = d.cpu().numpy()
    assert not (got & ~ref).any()          # overflow never turns an invalid pose valid
    # the latency path reports it too
    with pytest.raises(ap.ArtpError):
        for i in range(64):
            chk.isValid(poses[i])
    chk.debugSetGroupCapacity(0)
    set_map(chk, m)
    assert np.array_equal(chk.isValidBatch(poses), ref)
    chk.pollError()


def test_handle_is_thread_safe(maps, port_lib):
    """Two host threads share one handle (the reference calls its checker from the planning thread and the connection /
    cleaner threads): every call returns its own answer (staging buffer races)."""
    import threading
    import art_planner_b200 as ap
    m = maps("fbm_rough")
    chk = ap.StateValidityChecker(cases.PARAMS["yaml"], device=0)
    set_map(chk, m)
    o = port_lib.Oracle(cases.PARAMS["yaml"], "port")
    o.set_map(m)
    pa = synth.make_terrain_poses(m, 70000, seed=501)
    s1, s2 = synth.make_edges(m, 3000, seed=502)
    ref_p, ref_e = o.check_poses(pa), o.check_motions(s1, s2, 6)
    ref_c = o.path_length_cost(s1, s2)
    mv, plo = ap.MotionValidator(chk, 6), ap.PathLengthObjective(chk)
    errs = []

    def worker(kind):
        try:
            for it in range(6):
                if kind == 0:
                    assert np.array_equal(chk.isValidBatch(pa), ref_p)
                elif kind == 1:
                    assert np.array_equal(mv.checkMotionBatch(s1, s2), ref_e)
                else:
                    assert np.allclose(plo.motionCostBatch(s1, s2), ref_c, rtol=1e-12, atol=0)
                    assert chk.isValid(pa[it]) == bool(ref_p[it])
        except Exception as ex:      # noqa: BLE001
            errs.append((kind, repr(ex)))
    th = [threading.Thread(target=worker, args=(k,)) for k in (0, 1, 2, 1)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs


def test_calls_on_different_streams_share_a_handle_safely(maps, port_lib):
    """Asynchronous calls on two CUDA streams use the same per-handle queues: the library orders them."""
    import torch
    m = maps("fbm_rough")
    import art_planner_b200 as ap
    chk = ap.StateValidityChecker(cases.PARAMS["yaml"], device=0)
    set_map(chk, m)
    o = port_lib.Oracle(cases.PARAMS["yaml"], "port")
    o.set_map(m)
    pa = synth.make_terrain_poses(m, 200000, seed=601)
    pb = synth.make_terrain_poses(m, 200000, seed=602)
    ra, rb = o.check_poses_mt(pa, 8), o.check_poses_mt(pb, 8)
    da, db = torch.from_numpy(pa).cuda(), torch.from_numpy(pb).cuda()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    outs = []
    for it in range(4):
        with torch.cuda.stream(sa):
            va = chk.isValidBatch(da)
        with torch.cuda.stream(sb):
            vb = chk.isValidBatch(db)
        outs.append((va, vb))
    torch.cuda.synchronize()
    for va, vb in outs:
        assert np.array_equal(va.cpu().numpy(), ra) and np.array_equal(vb.cpu().numpy(), rb)


def test_map_window_shards_equal_the_whole_map(port_lib):
    """Spatial shards (artp_set_map_window): a handle that holds only a row slab + halo answers exactly like the whole
    map for every sample routed to it; a sample whose boxes leave the window is invalid + ARTP_E_WINDOW."""
    import art_planner_b200 as ap
    from art_planner_b200 import capi
    m = synth.make_fbm_map(600, 500, seed=11, amp=0.6)
    whole = ap.StateValidityChecker(cases.PARAMS["yaml"], device=0)
    whole.setMap(m); whole.updateHeightField()
    n = 120000
    poses = synth.make_terrain_poses(m, n, seed=12)
    ref = whole.isValidBatch(poses)
    lx, _ = m.length
    row = np.floor((m.cx + 0.5 * lx - poses[:, 0]) / m.res).astype(int)   # grid_map row of every sample
    halo, got = 40, np.full(n, 255, np.uint8)
    for s0, s1 in ((0, 200), (200, 400), (400, 600)):
        lo, hi = max(0, s0 - halo), min(m.rows, s1 + halo)
        shard = ap.StateValidityChecker(cases.PARAMS["yaml"], device=0)
        shard.setMap(m); shard.updateHeightField(window=(lo, hi - lo))
        sel = np.nonzero((row >= s0) & (row < s1))[0]
        got[sel] = shard.isValidBatch(poses[sel])
        shard.pollError()
        if s0 == 200:   # samples of the neighbouring slab far from this window: loud failure, never a wrong 'valid'
            far = np.nonzero(row < 100)[0][:500]
            with pytest.raises(ap.ArtpError) as ei:
                shard.isValidBatch(poses[far])
            assert ei.value.code == capi.ARTP_E_WINDOW
    assert np.array_equal(got, ref)
    o = port_lib.Oracle(cases.PARAMS["yaml"], "port")
    o.set_map(m)
    assert np.array_equal(ref[::40], o.check_poses(poses[::40]))
    with pytest.raises(ap.ArtpError):
        whole.updateHeightField(window=(2, 100))      # row0 must be a multiple of 4


def test_u32_compaction_and_fused_bits(maps, checkers):
    import torch
    from art_planner_b200 import sharding
    m = maps("fbm_rough")
    chk = checkers("yaml")
    set_map(chk, m)
    d = torch.from_numpy(synth.make_terrain_poses(m, 70001, seed=77)).cuda()
    v = torch.empty(70001, dtype=torch.uint8, device="cuda")
    bits = torch.empty((70001 + 31) // 32, dtype=torch.int32, device="cuda")
    chk.isValidBatchBits(d, v, bits)
    idx, cnt = chk.compactValidU32(v, base=5)
    torch.cuda.synchronize()
    assert torch.equal(v, chk.isValidBatch(d))
    assert torch.equal(bits.cpu(), sharding.pack_bits_reference(v.cpu()))
    want = torch.nonzero(v).reshape(-1).to(torch.int32) + 5
    assert int(cnt.item()) == want.numel() and torch.equal(idx[: want.numel()], want)


@pytest.mark.parametrize("mk", ["fbm_rough", "fbm_gentle", "fixture", "terraces"])
def test_both_reach_box_kernels_agree_with_the_oracle(maps, port_lib, mk):
    """Reach boxes over all-finite, merge-free zones take the 8-lane-group kernel (four boxes per warp), the others the
    one-warp-per-box kernel; ARTP_NO_GROUPS (read at artp_set_map) sends every reach box to the latter. Same masks,
    equal to the oracle, and the group kernel must actually have run in the default configuration."""
    import os
    import art_planner_b200 as ap
    m = maps(mk)
    poses = synth.make_terrain_poses(m, 30000, seed=77)
    o = port_lib.Oracle(cases.PARAMS["yaml"], "port")
    o.set_map(m)
    ref = o.check_poses(poses)
    masks, grouped = [], []
    for no_groups in (False, True):
        if no_groups:
            os.environ["ARTP_NO_GROUPS"] = "1"
        try:
            chk = ap.StateValidityChecker(cases.PARAMS["yaml"], device=0)
            set_map(chk, m)
            masks.append(chk.isValidBatch(poses))
            grouped.append(chk.stats()["last_reach_plane_stage"])
        finally:
            os.environ.pop("ARTP_NO_GROUPS", None)
    assert grouped[1] == 0
    if mk.startswith("fbm"):                  # flat / piecewise-constant terrain has mergeable planes nearly everywhere
        assert grouped[0] > 0, grouped
    for k in (0, 1):
        bad = np.nonzero(masks[k] != ref)[0]
        assert bad.size == 0, f"no_groups={k}: {bad.size} mismatches, first {bad[:8]}, grouped={grouped}"


def test_host_fed_rounds_equal_the_device_path_across_round_boundaries(maps, checkers):
    """A host-buffer call slices its batch (copy of slice i+1 under the kernels of slice i, per-slice queue segments, four
    streams) and processes more than 2^20 states in several rounds: the mask must equal the device-resident path's, for
    float and double states, from plain and from artp_host_alloc'd buffers."""
    import torch
    from art_planner_b200 import capi
    m = maps("fbm_rough")
    chk = checkers("yaml")
    set_map(chk, m)
    n = (1 << 20) + 300_001                                  # two rounds, the second one sliced as well
    poses = synth.make_terrain_poses(m, n, seed=123)
    dev = chk.isValidBatch(torch.from_numpy(poses).cuda()).cpu().numpy()
    assert 0.05 < dev.mean() < 0.95
    assert np.array_equal(chk.isValidBatch(poses), dev)
    hb_p, hb_v = capi.HostBuffer((n, 7), np.float32), capi.HostBuffer((n,), np.uint8)
    hb_p.array[:] = poses.astype(np.float32)
    hb_v.array[:] = 7
    chk.isValidHostPtr(hb_p.array.ctypes.data, n, hb_v.array.ctypes.data, f32=True)
    assert np.array_equal(hb_v.array, dev)
