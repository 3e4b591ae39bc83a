"""C4 workload: summary mode vs CIR mode vs both (time per step).  usage: python scripts/probe_cir.py [rays_per_tx]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200")); sys.path.insert(0, ROOT)
import numpy as np
import bench, hrt_b200 as hrt
P = int(float(sys.argv[1])) if len(sys.argv) > 1 else 5_000_000
rx, tx = bench.c4_positions(); zr, zt = np.zeros_like(rx), np.zeros_like(tx)
ctx = hrt.Context(0); ctx.load_scene(bench.SCENE)
for name, kw in (("summary", dict(summary=True)), ("cir", dict(cir=(0.0, 2e-9, 1024))),
                 ("summary+cir", dict(summary=True, cir=(0.0, 2e-9, 1024)))):
    for _ in range(2):
        r = ctx.run(rx, tx, zr, zt, 3.5, P, 5, los=False, **kw)
    s = r["stats"]
    extra = f" cir_dropped {s['cir_dropped']} |cir| {float(np.abs(r['cir']).sum()):.6e}" if "cir" in r else ""
    print(f"{name:12s}: total {s['ms_total']:.1f} ms scatter {s['ms_scatter']:.1f} ms -> {s['ray_bounces'] / s['ms_total'] / 1e3:.4g} M rb/s{extra}")
