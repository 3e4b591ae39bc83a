"""GPU diagnostic: grazing rays, brute-force kernel / BVH kernel vs oracle (details of every mismatch)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hrt_testlib as tl
import hrt_b200 as hrt

scene = sys.argv[1] if len(sys.argv) > 1 else "simple_street_canyon_with_cars"
rays = tl.grazing_rays(scene, 200000, seed=3)
tri_o, t_o, _ = tl.oracle_closest(scene, rays)
ctx = hrt.Context(0)
ctx.load_scene(tl.scene_path(scene))
for name, kw, env in (("brute", dict(brute_force=True), {}), ("brute_nosmem", dict(brute_force=True), {"HRT_NO_SMEM": "1"}), ("bvh", {}, {})):
    os.environ.update(env)
    tri, t, _ = ctx.closest_hits(rays, **kw)
    for k in env: del os.environ[k]
    bad, dn = tl.phantom_hit_report(scene, rays, tri_o, t_o, tri, t)
    print(name, "mismatches", bad.size, "max |d.n|", dn.max() if bad.size else 0)
    for i in bad[:12]:
        print("  ray", i, rays[i].tolist(), "oracle", tri_o[i], repr(float(t_o[i])), "gpu", tri[i], repr(float(t[i])))
