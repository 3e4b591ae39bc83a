"""Writes scenes/canyon_moving.hrt: the bundled street canyon with NON-ZERO
Mesh.velocity (and real materials instead of the bundled all-"air"), so that
the mesh-velocity Doppler term of the reference (src/compute_paths.c:720-722)
is exercised -- every bundled scene has zero velocities.

  cars (20 triangles):  metal, +-14 m/s along the street (Doppler variant) + a small y/z part
  ground (2 triangles): concrete, slow drift (0.3, -0.2, 0.05)
  buildings (12):       brick / glass / marble / concrete, every second one moving (0, 0.5, 0) .. (1, 0, 0.25)

Run from the repo root: python scripts/make_moving_scene.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200"))
from hrt_b200 import scenes  # noqa: E402


def main():
    meshes = scenes.read_hrt(os.path.join(ROOT, "scenes", "simple_street_canyon_with_cars.hrt"))
    car = bld = 0
    for m in meshes:
        nt = len(m["tris"])
        if nt == 20:
            m["material"] = scenes.MATERIAL["metal"]
            m["velocity"] = np.array([14.0 if car % 2 == 0 else -14.0, 0.125 * car, -0.0625 * (car % 3)], np.float32)
            car += 1
        elif nt == 2:
            m["material"] = scenes.MATERIAL["concrete"]
            m["velocity"] = np.array([0.3, -0.2, 0.05], np.float32)
        else:
            m["material"] = (scenes.MATERIAL["brick"], scenes.MATERIAL["glass1"], scenes.MATERIAL["marble"],
                             scenes.MATERIAL["concrete"])[bld % 4]
            m["velocity"] = (np.array([0.0, 0.5, 0.0], np.float32) * (bld % 2)
                             + np.array([1.0, 0.0, 0.25], np.float32) * (bld % 3 == 2))
            bld += 1
    out = os.path.join(ROOT, "scenes", "canyon_moving.hrt")
    scenes.write_hrt(out, meshes)
    print(out, os.path.getsize(out), "bytes;", car, "cars,", bld, "buildings")


if __name__ == "__main__":
    main()
