"""Where a small dense call (BASELINE configs[0]) spends its time: hrt_run stats per phase."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import hrt_b200 as hrt
from hrt_b200 import abi
import hrt_testlib as tl
ctx = hrt.Context(0); ctx.load_scene(tl.scene_path("simple_reflector"))
rx, tx, z = [[0, 0, .5]], [[0, 0, .5]], [[0, 0, 0]]
P, B = 30000, 3
out = abi.alloc_outputs(1, 1, P, B, 0)
for mode in (dict(dense=True, raysinfo=True), dict(dense=True), dict(summary=True), dict(dense=True, raysinfo=True, los=False)):
    ts = []
    for k in range(10):
        t0 = time.perf_counter()
        r = ctx.run(rx, tx, z, z, 3.0, P, B, out=out if mode.get("dense") else None, **mode)
        ts.append((time.perf_counter() - t0) * 1e3)
    s = r["stats"]
    print(mode, "wall median %.3f ms | hrt_run host_total %.3f setup %.3f | gpu window %.3f (bounce %.3f scatter %.3f sort %.3f los+ %.3f) launches %d"
          % (np.median(ts[2:]), s["host_ms_total"], s["host_ms_setup"], s["ms_total"], s["ms_bounce"], s["ms_scatter"], s["ms_sort"], s["ms_other"], s["kernel_launches"]))
