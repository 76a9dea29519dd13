"""CPU, world_size 2 over gloo: the multi-GPU path shards tile fields across ranks with no data-path collective;
what is exchanged is only the timing / count reduction bench.py performs.  This test drives that host logic
with the C restatement standing in for the device (the device path itself is covered by the -m gpu tests)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from oracle import bindings
    from recastnavigation_b200 import geom
    ag = geom.Agent(cs=0.2, ch=0.2)
    verts, tris = geom.terrain_with_boxes(side_m=50.0, n_boxes=40, seed_boxes=11)
    areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
    bmin, bmax = geom.calc_bounds(verts)
    fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 64, ag.walkable_radius + 3)
    ts, tl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
    lo, hi = bench.shard_range(fields.size, rank, world, ts)   # cuts at equal (tile, triangle) pair counts
    O = bindings.COracle()
    solid = compact = 0
    for i in range(lo, hi):
        F = fields[i]
        sel = tl[ts[i]:ts[i + 1]]
        r = O.chain(int(F["width"]), int(F["height"]), F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"]), verts, tris[sel],
                    areas[sel], ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb)
        solid += r["solid"].size
        compact += r["spans"].size
    vals = torch.tensor([solid, compact, hi - lo], dtype=torch.float64)
    tmax = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(vals, op=dist.ReduceOp.SUM)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.array(vals.tolist() + tmax.tolist() + [fields.size]))
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path, oracle):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npy")
    r1 = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(r0, r1)                  # both ranks hold the reduced totals
    assert r0[2] == r0[4]                          # every tile field processed exactly once
    assert r0[3] == 2.0                            # timing is the max over ranks
    # the sharded totals equal a single-rank pass
    sys.path.insert(0, ROOT)
    from recastnavigation_b200 import geom
    ag = geom.Agent(cs=0.2, ch=0.2)
    verts, tris = geom.terrain_with_boxes(side_m=50.0, n_boxes=40, seed_boxes=11)
    areas = geom.mark_walkable_triangles(verts, tris, ag.slope)
    bmin, bmax = geom.calc_bounds(verts)
    fields, tw, th = geom.tile_fields(bmin, bmax, ag.cs, ag.ch, 64, ag.walkable_radius + 3)
    ts, tl = geom.tile_triangle_lists(verts, tris, fields, tw, th)
    solid = 0
    for i in range(fields.size):
        F = fields[i]
        sel = tl[ts[i]:ts[i + 1]]
        solid += oracle.chain(int(F["width"]), int(F["height"]), F["bmin"], F["bmax"], float(F["cs"]), float(F["ch"]), verts, tris[sel],
                              areas[sel], ag.walkable_height, ag.walkable_climb, ag.walkable_radius, ag.walkable_climb)["solid"].size
    assert solid == r0[0]
