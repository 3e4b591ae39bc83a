"""GPU probe: per-depth cost of the C4 workload (shadow queries per ms as the
wavefront gets deeper).  usage: python scripts/probe.py [rays_per_tx]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200")); sys.path.insert(0, ROOT)
import numpy as np
import bench, hrt_b200 as hrt
P = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
rx, tx = bench.c4_positions(); zr, zt = np.zeros_like(rx), np.zeros_like(tx)
ctx = hrt.Context(0); ctx.load_scene(bench.SCENE)
prev_q, prev_ms = 0, 0.0
for B in range(1, 6):
    for _ in range(2):
        s = ctx.run(rx, tx, zr, zt, 3.5, P, B, summary=True, los=False)["stats"]
    dq, dms = s["shadow_queries"] - prev_q, s["ms_scatter"] - prev_ms
    print(f"B={B}: total ms {s['ms_total']:.1f} scatter {s['ms_scatter']:.1f} bounce {s['ms_bounce']:.1f} sort {s['ms_sort']:.1f} | depth {B-1}: {dq:.3e} shadow queries in {dms:.1f} ms = {dq/dms/1e6:.1f} Gq/s")
    prev_q, prev_ms = s["shadow_queries"], s["ms_scatter"]
