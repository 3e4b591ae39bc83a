"""Which shard block size should an N-GPU job use?  One GPU plays every rank of an 8-way (and 2-way)
C4 job in turn; the step time of the job is the slowest rank's.  Blocks of consecutive Fibonacci
indices are latitude rings: thicker rings = more compact hit patches per warp, but fewer blocks per
rank = worse balance.  usage: probe_shard_block.py [world]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200"))
import hrt_b200 as hrt
import bench

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rx, tx = bench.c4_positions()
zr, zt = [[0.0, 0.0, 0.0]] * len(rx), [[0.0, 0.0, 0.0]] * len(tx)
ctx = hrt.Context(0)
ctx.load_scene(os.path.join(ROOT, "scenes", "simple_street_canyon_with_cars.hrt"))
P, B = bench.TOTAL_RAYS // bench.NUM_TX, bench.BOUNCES
ctx.run(rx, tx, zr, zt, bench.F_GHZ, P, B, summary=True, los=False, shard=(0, world), shard_block=1 << 16)
for lg in (14, 16, 17, 18, 19, 20):
    tot, sc, so, bo = [], [], [], []
    for rank in range(world):
        s = ctx.run(rx, tx, zr, zt, bench.F_GHZ, P, B, summary=True, los=False, shard=(rank, world), shard_block=1 << lg)["stats"]
        tot.append(s["ms_total"]); sc.append(s["ms_scatter"]); so.append(s["ms_sort"]); bo.append(s["ms_bounce"])
    print("world %d block 2^%d: step (max rank) %.2f ms, mean %.2f; scatter max %.2f mean %.2f; sort mean %.2f; bounce mean %.2f"
          % (world, lg, max(tot), sum(tot) / world, max(sc), sum(sc) / world, sum(so) / world, sum(bo) / world), flush=True)
