"""hrt_scene_advance on the C5 scene: refit vs rebuild, several times."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200"))
import numpy as np
import hrt_b200 as hrt
from hrt_b200 import scenes
meshes, pitch = scenes.tiled_canyon(os.path.join(ROOT, "scenes", "simple_street_canyon_with_cars.hrt"), 64, 64)
for m in meshes:
    if m["material"] == scenes.MATERIAL["metal"]:
        m["velocity"] = np.array([10.0, 0.0, 0.0], np.float32)
scenes.write_hrt("/tmp/c5.hrt", meshes)
ctx = hrt.Context(0)
for k in range(2):
    t0 = time.perf_counter(); ctx.load_scene("/tmp/c5.hrt"); print("load+upload+build %.1f ms" % ((time.perf_counter() - t0) * 1e3))
for rebuild in (False, True, True, False, True):
    t0 = time.perf_counter(); ctx.advance(1e-3, rebuild=rebuild); print("advance rebuild=%s %.1f ms" % (rebuild, (time.perf_counter() - t0) * 1e3))
