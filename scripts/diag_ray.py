import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hrt_testlib as tl
import hrt_b200 as hrt
from hrt_b200 import scenes
scene = "simple_street_canyon_with_cars"
rays = tl.grazing_rays(scene, 200000, seed=3)
ctx = hrt.Context(0)
ctx.load_scene(tl.scene_path(scene))
for i in (504, 3818, 183):
    r = rays[i:i + 1].copy()
    print(i, "oracle", tl.oracle_closest(scene, r)[:2], "gpu brute single", ctx.closest_hits(r, brute_force=True)[:2], "gpu bvh single", ctx.closest_hits(r)[:2])
    r2 = np.repeat(r, 64, 0)
    print("   x64", ctx.closest_hits(r2, brute_force=True)[0][:3])
# scene with only the oracle's triangle for ray 504
meshes = scenes.read_hrt(tl.scene_path(scene))
T = tl.scene_triangles(scene)
tri = T[84].astype(np.float32)
one = [dict(vs=tri, tris=np.array([[0, 1, 2]], np.uint32), material=0, velocity=np.zeros(3, np.float32))]
scenes.write_hrt("/tmp/one.hrt", one)
ctx.load_scene("/tmp/one.hrt")
r = rays[504:505].copy()
print("one-triangle scene: oracle", tl.oracle_closest("/tmp/one.hrt", r)[:2], "gpu brute", ctx.closest_hits(r, brute_force=True)[:2], "gpu", ctx.closest_hits(r)[:2])
