"""Soak: closest hits of the GPU BVH path against the oracle (CPU brute force over
every triangle) on random rays -- isotropic, grazing, axis-aligned and aimed at
triangle corners / edge points (tests/hrt_testlib.random_rays).  Triangle ids and
t must be bit-identical.  usage: python scripts/soak_closest.py [rays_per_seed] [seeds]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import hrt_b200 as hrt
import hrt_testlib as tl
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = hrt.Context(0)
total = bad = hits = 0
for scene in ("simple_street_canyon_with_cars", "2cars", "box", "simple_reflector"):
    ctx.load_scene(tl.scene_path(scene))
    for seed in range(100, 100 + seeds):
        rays = tl.random_rays(scene, n, seed=seed)
        t0 = time.time()
        tri_o, t_o, _ = tl.oracle_closest(scene, rays)
        tri_g, t_g, _ = ctx.closest_hits(rays)
        m = (tri_o != tri_g) | (t_o.view(np.uint32) != t_g.view(np.uint32))
        total += n; bad += int(m.sum()); hits += int((tri_o != tl.NONE).sum())
        print(f"{scene} seed {seed}: {int(m.sum())} mismatches of {n} rays ({int((tri_o != tl.NONE).sum())} hits), {time.time() - t0:.1f} s", flush=True)
print(f"TOTAL: {bad} mismatches in {total} rays ({hits} hits)")
