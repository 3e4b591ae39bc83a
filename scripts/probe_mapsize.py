"""Which receiver-map resolution does a run get, and what does the first (map-building) call cost against the
following ones?  Canyon, 64 receivers, 4 transmitters, rays per TX from 1e3 to 1e7.
usage: python scripts/probe_mapsize.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200"))
import hrt_b200 as hrt
import bench
rx, tx = bench.c4_positions()
zr, zt = [[0.0, 0.0, 0.0]] * len(rx), [[0.0, 0.0, 0.0]] * len(tx)
for P in (1000, 10_000, 100_000, 1_000_000, 10_000_000):
    for force in (None, "0"):
        if force is None: os.environ.pop("HRT_RXMAP", None)
        else: os.environ["HRT_RXMAP"] = force
        ctx = hrt.Context(0)
        ctx.load_scene(os.path.join(ROOT, "scenes", "simple_street_canyon_with_cars.hrt"))
        ctx.run(rx, tx, zr, zt, 3.5, 100, 5, summary=True, los=False)       # context warm-up (no map: tiny)
        ts = []
        for k in range(4):
            t0 = time.perf_counter()
            s = ctx.run(rx, tx, zr, zt, 3.5, P, 5, summary=True, los=False)["stats"]
            ts.append((time.perf_counter() - t0) * 1e3)
        print(f"P={P:9d} {'auto' if force is None else 'BVH ':4s}: map G={s['rx_map_cells']:3d}  first call {ts[0]:8.2f} ms, then {min(ts[1:]):8.2f} ms  ({s['shadow_queries']:.2e} shadow queries)", flush=True)
        ctx.close()
os.environ.pop("HRT_RXMAP", None)
