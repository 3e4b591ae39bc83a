"""Per-call cost of the drop-in entry compute_paths() for small runs (the way
HermesPy calls it: same scene, many calls).  usage: python scripts/call_overhead.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import hrt_b200 as hrt
from hrt_b200 import abi
import hrt_testlib as tl
L = hrt.lib()
for scene, P, B, R in (("simple_reflector", 30000, 3, 1),     # BASELINE configs[0] (test/test.c:17-27)
                       ("box", 10000, 3, 1), ("simple_street_canyon_with_cars", 10000, 3, 4), ("simple_street_canyon_with_cars", 100000, 5, 16)):
    sc = L.scene_load(tl.scene_path(scene).encode())
    rx = [[0.5 * i, 1.0, 1.5] for i in range(R)]; tx = [[0, 0, 2.5]]
    if scene == "simple_reflector":
        rx, tx = [[0, 0, .5]], [[0, 0, .5]]
    zr = [[0, 0, 0]] * R; zt = [[0, 0, 0]]
    ts = []
    out = abi.alloc_outputs(R, 1, P, B, 0)
    for k in range(12):
        t0 = time.perf_counter()
        abi.call_compute_paths(L, sc, rx, tx, zr, zt, 3.0, P, B, fill=0, out=out)
        ts.append(time.perf_counter() - t0)
    abi.free_scene(sc)
    st = hrt.RunStats()
    print(f"{scene} P={P} B={B} R={R}: first {ts[0]*1e3:.1f} ms, then median {np.median(ts[2:])*1e3:.2f} ms (min {min(ts)*1e3:.2f})")
    # the same call with the scene cache off (upload + BVH build every call, as in round 1)
    os.environ["HRT_NO_SCENE_CACHE"] = "1"
    sc = L.scene_load(tl.scene_path(scene).encode())
    ts = []
    for k in range(8):
        t0 = time.perf_counter()
        abi.call_compute_paths(L, sc, rx, tx, zr, zt, 3.0, P, B, fill=0, out=out)
        ts.append(time.perf_counter() - t0)
    abi.free_scene(sc)
    del os.environ["HRT_NO_SCENE_CACHE"]
    print(f"    without the scene cache: median {np.median(ts[2:])*1e3:.2f} ms")
