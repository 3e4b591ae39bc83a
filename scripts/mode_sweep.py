"""thread-per-hit vs warp-per-hit scatter mapping as a function of the ray count
(C4 scene, 4 TX / 64 RX, 5 bounces).  usage: python scripts/mode_sweep.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200")); sys.path.insert(0, ROOT)
import numpy as np
import bench, hrt_b200 as hrt
rx, tx = bench.c4_positions(); zr, zt = np.zeros_like(rx), np.zeros_like(tx)
ctx = hrt.Context(0); ctx.load_scene(bench.SCENE)
for P in (2000, 10000, 30000, 100000, 300000, 1000000):
    row = []
    for mode in ("t", "w", None):
        if mode: os.environ["HRT_SCATTER_MODE"] = mode
        else: os.environ.pop("HRT_SCATTER_MODE", None)
        for _ in range(3):
            s = ctx.run(rx, tx, zr, zt, 3.5, P, 5, summary=True, los=False)["stats"]
        row.append(s["ms_total"])
    print(f"P={P:8d} per TX: thread {row[0]:8.2f} ms  warp {row[1]:8.2f} ms  auto {row[2]:8.2f} ms")
