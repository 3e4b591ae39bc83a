"""End-to-end timing of the drop-in C entry compute_paths() with HOST buffers
(dense reference-layout outputs incl. RaysInfo) on BASELINE configs[1], [2].
usage: python scripts/run_dense.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import hrt_b200 as hrt
from hrt_b200 import abi
import hrt_testlib as tl

def fresh_outputs(R, T, P, B):
    """like abi.alloc_outputs but the pages are never touched (what a C caller's malloc returns)"""
    o = abi.Outputs(R, T, P, B)
    for k in abi.CHAN_FIELDS:
        vec = k.startswith("directions")
        o.los[k] = np.empty((R, T, 3) if vec else (R, T), np.float32)
        o.scat[k] = np.empty((R, T, B, P, 3) if vec else (R, T, B, P), np.float32)
    o.los_rays = np.empty((R * T, 6), np.float32); o.los_active = np.empty((R * T // 8 + 1,), np.uint8)
    rows = T * (B + 1) + 1
    o.scat_rays = np.empty((rows, P, 6), np.float32); o.scat_active = np.empty((rows, P // 8 + 1), np.uint8)
    return o


L = hrt.lib()
res = []
for name, cfg, P, B in (("configs[1] box 1e6 x 3", "box_axis", 1_000_000, 3), ("configs[2] 2cars 1e7 x 5", "2cars_raised", 10_000_000, 5),
                        ("configs[0] reflector 3e4 x 3", "reflector_testc", 30000, 3)):
    scene, rx, tx, f = tl.CONFIGS[cfg]
    sc = L.scene_load(tl.scene_path(scene).encode())
    out = abi.alloc_outputs(1, 1, P, B, 0)
    best = 1e30
    for rep in range(3):
        t0 = time.perf_counter()
        abi.call_compute_paths(L, sc, rx, tx, [[0, 0, 0]], [[0, 0, 0]], f, P, B, out=out)
        best = min(best, time.perf_counter() - t0)
    act = np.unpackbits(out.scat_active[:B + 1, :], axis=1, bitorder="little")[:, :P]
    rb = int(P + act[1:B].sum())
    nbytes = sum(v.nbytes for v in out.scat.values()) + out.scat_rays[:B + 1].nbytes
    os.environ["HRT_NO_RAYSINFO"] = "1"
    best2 = 1e30
    for rep in range(2):
        t0 = time.perf_counter()
        abi.call_compute_paths(L, sc, rx, tx, [[0, 0, 0]], [[0, 0, 0]], f, P, B, out=out)
        best2 = min(best2, time.perf_counter() - t0)
    del os.environ["HRT_NO_RAYSINFO"]
    fresh = fresh_outputs(1, 1, P, B)
    t0 = time.perf_counter()
    abi.call_compute_paths(L, sc, rx, tx, [[0, 0, 0]], [[0, 0, 0]], f, P, B, out=fresh)
    t_fresh = time.perf_counter() - t0
    del fresh
    abi.free_scene(sc)
    res.append({"config": name, "seconds": best, "ray_bounces": rb, "ray_bounces_per_s": rb / best,
                "d2h_bytes": nbytes, "seconds_without_raysinfo": best2, "rb_per_s_without_raysinfo": rb / best2,
                "seconds_into_untouched_arrays": t_fresh})
print(json.dumps(res))
