"""BASELINE configs[4] at FULL size on the GPUs of one box: synthetic 64x64 tiled
street canyon (958,464 triangles, mixed ITU materials), 16 TX / 1024 RX, 1e9 rays
(6.25e7 per TX), 6 bounces.  One process per GPU (launch with torchrun; only
RANK / WORLD_SIZE / LOCAL_RANK are used), rank r traces the 65,536-path blocks
b = r (mod world) -- no data-path collective; every rank writes its summary and
device-timed duration to gpurun_out/c5_full_rank<r>.json, rank 0 waits for the
others and prints the job line (time = max over ranks).
usage: torchrun --nproc-per-node 8 scripts/run_c5_full.py [rays_per_tx] [num_rx]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200"))
import numpy as np
import hrt_b200 as hrt
from hrt_b200 import scenes

P = int(float(sys.argv[1])) if len(sys.argv) > 1 else 62_500_000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
out_dir = os.path.join(ROOT, "gpurun_out"); os.makedirs(out_dir, exist_ok=True)
mine = os.path.join(out_dir, f"c5_full_rank{rank}.json")
if os.path.exists(mine):
    os.remove(mine)
meshes, pitch = scenes.tiled_canyon(os.path.join(ROOT, "scenes", "simple_street_canyon_with_cars.hrt"), 64, 64)
path = f"/tmp/c5_tiled_canyon_{rank}.hrt"
scenes.write_hrt(path, meshes)
rx, tx = scenes.c5_positions(pitch, 64, 64, n_tx=16, n_rx=R)
zr, zt = np.zeros_like(rx), np.zeros_like(tx)
ctx = hrt.Context(local)
ctx.load_scene(path)
ctx.run(rx, tx, zr, zt, 3.5, 100_000, 6, summary=True, shard=(rank, world), shard_block=1 << 16)   # warm-up
t0 = time.perf_counter()
r = ctx.run(rx, tx, zr, zt, 3.5, P, 6, summary=True, los=(rank == 0), shard=(rank, world), shard_block=1 << 16)
wall = time.perf_counter() - t0
s = r["stats"]
res = {"rank": rank, "world": world, "wall_s": wall, "ms_total": s["ms_total"], "ms_scatter": s["ms_scatter"],
       "ray_bounces": s["ray_bounces"], "shadow_queries": s["shadow_queries"],
       "n_valid": int(r["pair"]["n_valid"].sum()), "n_occluded": int(r["pair"]["n_occluded"].sum()),
       "hit_hash": int(r["bounce"]["hit_hash"].sum(dtype=np.uint64)), "bvh_build_ms": s["bvh_build_ms"]}
json.dump(res, open(mine + ".tmp", "w")); os.replace(mine + ".tmp", mine)
ctx.close()
if rank == 0:
    parts = []
    deadline = time.time() + 1800
    while len(parts) < world and time.time() < deadline:
        parts = [json.load(open(os.path.join(out_dir, f"c5_full_rank{k}.json"))) for k in range(world)
                 if os.path.exists(os.path.join(out_dir, f"c5_full_rank{k}.json"))]
        time.sleep(0.5)
    t = max(p["ms_total"] for p in parts) * 1e-3
    rb = sum(p["ray_bounces"] for p in parts); q = sum(p["shadow_queries"] for p in parts) + rb
    print(json.dumps({"config": f"C5 FULL: 958,464 triangles, 16 TX / {R} RX, {P} rays per TX ({16 * P:.3g} rays), 6 bounces, "
                                f"{world} GPU(s), 65,536-path blocks round-robin", "ranks_reported": len(parts),
                      "seconds_max_over_ranks": t, "seconds_min_over_ranks": min(p["ms_total"] for p in parts) * 1e-3,
                      "ray_bounces": rb, "closest_hit_queries": q, "ray_bounces_per_s": rb / t,
                      "closest_hit_queries_per_s": q / t, "valid_paths": sum(p["n_valid"] for p in parts),
                      "occluded": sum(p["n_occluded"] for p in parts)}))
