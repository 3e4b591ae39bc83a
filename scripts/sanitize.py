"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck):
every kernel family once -- SAH build, octant emit, raygen, init, LoS, bounce,
hit sort, scatter (both mappings, summary + dense + CIR), scene advance (refit +
rebuild), global-memory scene.  usage: [compute-sanitizer --tool memcheck] python scripts/sanitize.py (the sanitizer is not available on the B200 pool: plain functional run there)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200")); sys.path.insert(0, ROOT)
import numpy as np
import bench, hrt_b200 as hrt
from hrt_b200 import scenes
rx, tx = bench.c4_positions(); rx, tx = rx[:12], tx[:2]
zr, zt = np.zeros_like(rx), np.zeros_like(tx)
ctx = hrt.Context(0); ctx.load_scene(bench.SCENE)
for mode in ("t", "w"):
    os.environ["HRT_SCATTER_MODE"] = mode
    r = ctx.run(rx, tx, zr, zt, 3.5, 3000, 3, dense=True, raysinfo=True, trace=True, summary=True, cir=(0.0, 5e-9, 128))
    print(mode, int(r["pair"]["n_valid"].sum()), float(np.abs(r["cir"]).sum()))
del os.environ["HRT_SCATTER_MODE"]
meshes = scenes.read_hrt(bench.SCENE)
for m in meshes:
    if len(m["tris"]) == 20: m["velocity"] = np.array([10, 0, 0], np.float32)
scenes.write_hrt("/tmp/san_moving.hrt", meshes); ctx.load_scene("/tmp/san_moving.hrt")
ctx.advance(0.1); ctx.advance(0.1, rebuild=True)
print("advance", int(ctx.run(rx, tx, zr, zt, 3.5, 2000, 2, summary=True)["pair"]["n_valid"].sum()))
big, pitch = scenes.tiled_canyon(bench.SCENE, 6, 6, block=2); scenes.write_hrt("/tmp/san_big.hrt", big)
rxb, txb = scenes.c5_positions(pitch, 6, 6, n_tx=2, n_rx=16)
for env in ({}, {"HRT_OCTANT_BYTES_MAX": "0"}, {"HRT_BVH_LBVH": "1"}):
    os.environ.update(env); ctx.load_scene("/tmp/san_big.hrt")
    r = ctx.run(rxb, txb, np.zeros_like(rxb), np.zeros_like(txb), 3.5, 2000, 3, summary=True)
    for k in env: del os.environ[k]
    print("big", env, int(r["pair"]["n_valid"].sum()), r["stats"]["scene_in_smem"])
ctx.close(); print("done")
