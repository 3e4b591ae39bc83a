"""Soak of the receiver maps (incl. "sure" cells and the own-triangle early-out) on the GPU: random receiver and
transmitter sets in and around every bundled scene; the dense outputs and the slot states of a run whose shadow
queries go through the maps must equal, word for word, those of the brute-force run (every triangle tested for
every query: the reference's loop).  Differences are listed; the only admissible ones are phantom hits
(rays within ~1e-5 rad of a triangle's plane).
usage: python scripts/soak_maps.py [configs_per_scene] [rays_per_tx]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import hrt_b200 as hrt
import hrt_testlib as tl

K = int(sys.argv[1]) if len(sys.argv) > 1 else 6
P = int(float(sys.argv[2])) if len(sys.argv) > 2 else 20_000
rng = np.random.default_rng(2026)
ctx = hrt.Context(0)
tot_slots = tot_diff = tot_q = 0
t0 = time.time()
for scene in ("simple_street_canyon_with_cars", "canyon_moving", "2cars", "box", "simple_reflector"):
    ctx.load_scene(tl.scene_path(scene))
    T3 = tl.scene_triangles(scene).reshape(-1, 3)
    lo, hi = T3.min(0), T3.max(0)
    ext = np.maximum(hi - lo, 1.0)
    for k in range(K):
        R, T, B = int(rng.integers(8, 49)), int(rng.integers(1, 4)), int(rng.integers(2, 6))
        G = int(rng.choice([32, 64, 128, 256]))
        rx = lo + rng.random((R, 3)) * ext
        rx[: R // 4] = lo - 0.3 * ext + rng.random((R // 4, 3)) * 1.6 * ext       # some outside the bounding box
        rx[R // 4: R // 2, 2] = lo[2] + rng.random(R // 2 - R // 4) * 0.02 * ext[2] + 1e-3   # some just above the floor
        tx = lo + (0.2 + 0.6 * rng.random((T, 3))) * ext
        rxv, txv = rng.uniform(-3, 3, rx.shape), rng.uniform(-10, 10, tx.shape)
        f = float(rng.choice([3.5, 28.0, 70.0]))
        os.environ["HRT_RXMAP"] = "1"; os.environ["HRT_RXMAP_G"] = str(G)
        a = ctx.run(rx, tx, rxv, txv, f, P, B, dense=True, trace=True, summary=True)
        used = a["stats"]["rx_map"]
        os.environ["HRT_RXMAP"] = "0"; del os.environ["HRT_RXMAP_G"]
        b = ctx.run(rx, tx, rxv, txv, f, P, B, dense=True, trace=True, summary=True, brute_force=True)
        wa, wb = tl.outputs_words(a["out"]), tl.outputs_words(b["out"])
        sa, sb = a["trace"]["slot_state"], b["trace"]["slot_state"]
        diff = sa != sb
        for key in ("scat.a_te_re", "scat.a_te_im", "scat.a_tm_re", "scat.a_tm_im", "scat.tau", "scat.freq_shift"):
            diff |= (wa[key].reshape(sa.shape) != wb[key].reshape(sa.shape))
        nd = int(diff.sum())
        tot_slots += diff.size; tot_diff += nd; tot_q += int(a["stats"]["shadow_queries"])
        print(f"{scene:32s} R={R:2d} T={T} B={B} G={G:3d} f={f:4.1f}: maps used {bool(used)}, {int(a['stats']['shadow_queries']):9d} shadow queries, "
              f"{nd} differing slots", flush=True)
        if nd:
            idx = np.argwhere(diff)[:5]
            print("   first differing (rx, tx, bounce, path):", idx.tolist())
os.environ.pop("HRT_RXMAP", None)
print(f"TOTAL: {tot_q} shadow queries, {tot_slots} output slots, {tot_diff} differ ({time.time() - t0:.0f} s)")
sys.exit(1 if tot_diff else 0)
