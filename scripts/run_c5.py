"""BASELINE configs[4] (C5): synthetic 64x64 tiled street canyon (958,464
triangles, mixed ITU materials), 16 TX / 1024 RX, 6 bounces -- at a reduced ray
count per TX (the full 1e9 rays x 1024 RX is 3e12 closest-hit queries).
usage: python scripts/run_c5.py [rays_per_tx] [num_rx] [world] [block]
With world > 1 the run is shard 0 of `world` of the job (blocks of `block` paths
dealt round-robin, as bench.py --gpus N deals them): rays_per_tx = 6.25e7 and
world = 64 is 1/64 of the full C5 job at its true ray density."""
import json, os, sys, time
os.environ.setdefault("HRT_NO_OVERLAP", "1")      # clean per-kernel event intervals (ms_bounce, ms_scatter)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200"))
import numpy as np
import hrt_b200 as hrt
from hrt_b200 import scenes

P = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
WORLD = int(sys.argv[3]) if len(sys.argv) > 3 else 1
BLOCK = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 16
SH = dict(shard=(0, WORLD), shard_block=BLOCK) if WORLD > 1 else {}
t0 = time.perf_counter()
meshes, pitch = scenes.tiled_canyon(os.path.join(ROOT, "scenes", "simple_street_canyon_with_cars.hrt"), 64, 64)
for m in meshes:                      # cars drive (Doppler + scene-advance timing); geometry at t = 0 unchanged
    if m["material"] == scenes.MATERIAL["metal"]:
        m["velocity"] = np.array([10.0, 0.0, 0.0], np.float32)
path = "/tmp/c5_tiled_canyon.hrt"
scenes.write_hrt(path, meshes)
t_gen = time.perf_counter() - t0
rx, tx = scenes.c5_positions(pitch, 64, 64, n_tx=16, n_rx=R)
zr, zt = np.zeros_like(rx), np.zeros_like(tx)
ctx = hrt.Context(0)
t0 = time.perf_counter(); ctx.load_scene(path); t_up = time.perf_counter() - t0
t0 = time.perf_counter(); ctx.load_scene(path); t_up2 = time.perf_counter() - t0
out = {}
for rep in range(1 if WORLD > 1 else 2):
    t0 = time.perf_counter()
    r = ctx.run(rx, tx, zr, zt, 3.5, P, 6, summary=True, **SH)
    wall = time.perf_counter() - t0
s = r["stats"]
out = {"config": f"C5: 64x64 tiled canyon, {s['num_tris']} triangles in {len(meshes)} meshes, 16 TX / {R} RX, {P} rays per TX, 6 bounces"
                 + (f", shard 0 of {WORLD} (blocks of {BLOCK} paths)" if WORLD > 1 else ""),
       "wall_s": wall, "bvh_build_ms": s["bvh_build_ms"], "bvh_levels": s["bvh_levels"], "bvh_sah": s["bvh_sah"], "ms_sort": s["ms_sort"],
       "scene_generate_s": t_gen, "scene_load_upload_bvh_s": t_up2, "bvh_nodes": s["num_nodes"], "scene_in_smem": s["scene_in_smem"],
       "ray_bounces": s["ray_bounces"], "shadow_queries": s["shadow_queries"], "ms_total": s["ms_total"],
       "ms_scatter": s["ms_scatter"], "ms_bounce": s["ms_bounce"],
       "ray_bounces_per_s": s["ray_bounces"] / (s["ms_total"] * 1e-3),
       "closest_hit_queries_per_s": (s["ray_bounces"] + s["shadow_queries"]) / (s["ms_total"] * 1e-3),
       "primary_queries_per_s_in_k_bounce": s["ray_bounces"] / max(s["ms_bounce"] * 1e-3, 1e-9),
       "valid_paths": int(r["pair"]["n_valid"].sum()), "occluded": int(r["pair"]["n_occluded"].sum()),
       "n_traced_per_bounce_tx0": r["bounce"]["n_traced"][0].tolist()}
c = ctx.run(rx, tx, zr, zt, 3.5, min(P, 20000), 6, summary=True, count_work=True)["stats"]
out["flops_per_shadow_query"] = (22 * c["work_scatter"][0] + 14 * c["work_scatter"][1] + 9 * c["work_scatter"][2]
                                 + 16 * c["work_scatter"][3] + 6 * c["work_scatter"][4]) / max(c["shadow_queries"], 1)
out["box_tests_per_shadow_query"] = c["work_scatter"][0] / max(c["shadow_queries"], 1)
out["tri_tests_per_shadow_query"] = c["work_scatter"][1] / max(c["shadow_queries"], 1)
out["box_tests_per_primary_query"] = c["work_bounce"][0] / max(c["ray_bounces"], 1)
# moving scene: advance by +dt and -dt (back to the original vertices up to rounding), refit vs rebuild
for name, rebuild in (("bvh_refit_ms", False), ("bvh_rebuild_ms", True)):
    t0 = time.perf_counter(); ctx.advance(1e-3, rebuild=rebuild); out[name] = (time.perf_counter() - t0) * 1e3
    ctx.advance(-1e-3, rebuild=rebuild)
print(json.dumps(out))
