"""CPU-only, world_size 2 over gloo: the N>1 host logic -- shard assignment, ordered index exchange, merge."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import cases
        from art_planner_b200 import sharding, synth
        from oracle import orc
        m = cases.MAPS["fixture"]()
        lo, hi = sharding.shard_range(n_total, rank, world)
        poses = synth.make_terrain_poses(m, hi - lo, seed=9, start=lo)       # pure function of (seed, index)
        o = orc.Oracle(cases.PARAMS["yaml"], "port")                         # the checker stands in for the GPU kernel
        o.set_map(m)
        valid = o.check_poses(poses)
        cap = (n_total + world - 1) // world
        idx = torch.zeros(cap, dtype=torch.int64)
        nz = np.nonzero(valid)[0] + lo
        idx[: len(nz)] = torch.from_numpy(nz)
        cnt = torch.tensor([len(nz)], dtype=torch.int32)
        all_idx, counts = sharding.gather_valid_indices(idx, cnt, world)
        merged = sharding.merge_gathered(all_idx, counts, cap)
        # the bit-packed exchange (what bench.py --gpus N uses): one all-gather of cap/32 words per rank
        bits = sharding.pack_bits_reference(torch.from_numpy(valid), cap)
        all_bits = sharding.gather_valid_bits(bits, world)
        from_bits = sharding.indices_from_bits(all_bits, n_total, world)
        if rank == 0:
            full = o.check_poses(synth.make_terrain_poses(m, n_total, seed=9))
            want = np.nonzero(full)[0].tolist()
            q.put((merged.numpy().tolist() == want and from_bits.numpy().tolist() == want, int(full.sum())))
    finally:
        dist.destroy_process_group()


def test_shard_ranges_partition_the_stream():
    from art_planner_b200 import sharding
    for n in (0, 1, 7, 1000, 1_000_003):
        for w in (1, 2, 3, 8):
            r = [sharding.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_two_rank_index_exchange_over_gloo(port_lib):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 4001, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    ok, n_valid = q.get(timeout=10)
    assert ok and n_valid > 0


def test_bit_packing_reference_roundtrip():
    from art_planner_b200 import sharding
    g = torch.Generator().manual_seed(3)
    for n in (1, 31, 32, 33, 1000, 4097):
        v = (torch.rand(n, generator=g) < 0.4).to(torch.uint8)
        bits = sharding.pack_bits_reference(v)
        assert bits.numel() == (n + 31) // 32 and bits.dtype == torch.int32
        assert sharding.indices_from_bits(bits, n, 1).tolist() == torch.nonzero(v).reshape(-1).tolist()


def test_spatial_slabs_cover_the_map_and_route_every_sample():
    """slab_window / rank_of_x (configs[4] layout): slabs tile the rows, windows are slab + halo with 4-aligned starts, and
    every position is routed to the rank whose slab holds its grid_map row."""
    import numpy as np
    from art_planner_b200 import sharding, synth
    for rows, world in ((4000, 8), (4000, 4), (1001, 3), (600, 7)):
        covered = np.zeros(rows, int)
        for r in range(world):
            s0, s1, lo, hi = sharding.slab_window(rows, r, world, halo=40)
            covered[s0:s1] += 1
            assert lo % 4 == 0 and lo <= max(0, s0 - 40) and hi == min(rows, s1 + 40) and lo >= 0
        assert (covered == 1).all()
        res, cx = 0.04, 1.5
        lx = rows * res
        x = cx + (synth.hash_uniform(3, 1, np.arange(20000)) - 0.5) * lx * 0.9999
        rk = sharding.rank_of_x(x, cx, lx, res, rows, world)
        row = sharding.row_of_x(x, cx, lx, res, rows)
        for r in range(world):
            s0, s1, _, _ = sharding.slab_window(rows, r, world, halo=40)
            assert ((row[rk == r] >= s0) & (row[rk == r] < s1)).all()


def test_slab_windows_cover_the_map_for_every_world_size():
    """Spatial shards (configs[4]): for every world size the slabs partition the rows, each window holds its slab plus the
    halo (clipped at the map border), starts on a multiple of 4 rows (artp_set_map_window / TMA alignment), and rank_of_x
    routes a sample to the rank whose slab holds its row."""
    from art_planner_b200 import sharding
    rng = np.random.default_rng(5)
    for rows in (4000, 1000, 1003, 257):
        res, cx = 0.04, 1.5
        length_x = rows * res
        for world in (1, 2, 3, 4, 5, 8):
            halo = 40
            covered = np.zeros(rows, dtype=np.int32)
            slabs = []
            for r in range(world):
                s0, s1, lo, hi = sharding.slab_window(rows, r, world, halo)
                assert 0 <= lo <= s0 < s1 <= hi <= rows
                assert lo % 4 == 0
                assert lo <= max(0, s0 - halo) and hi == min(rows, s1 + halo)
                covered[s0:s1] += 1
                slabs.append((s0, s1))
            assert np.all(covered == 1)
            x = cx + (rng.random(20000) - 0.5) * length_x * 1.1            # some outside the map: clamped to the edge rows
            row = sharding.row_of_x(x, cx, length_x, res, rows)
            rk = sharding.rank_of_x(x, cx, length_x, res, rows, world)
            assert rk.min() >= 0 and rk.max() < world
            for r, (s0, s1) in enumerate(slabs):
                sel = rk == r
                assert np.all((row[sel] >= s0) & (row[sel] < s1))
