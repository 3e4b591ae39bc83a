"""Soak: the whole path (dense ChannelInfo + RaysInfo + hit trace) of the GPU run
against the oracle at a larger size than the unit tests: street canyon, 2 moving TX,
8 moving RX, 5 bounces.  usage: python scripts/soak_paths.py [rays_per_tx]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import hrt_b200 as hrt
import hrt_testlib as tl
P = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000
B = 5
scene = "simple_street_canyon_with_cars"
rx, tx = tl.canyon_c4_positions()
rx, tx = rx[::8][:8], tx[:2]
rng = np.random.default_rng(77)
rxv, txv = rng.uniform(-3, 3, rx.shape), rng.uniform(-10, 10, tx.shape)
t0 = time.time()
a, tr = tl.run_oracle(scene, rx, tx, rxv, txv, 3.5, P, B, fill=0x00)
b, _ = tl.run_oracle(scene, rx, tx, rxv, txv, 3.5, P, B, fill=0x5A, trace=False)
mask = tl.written_mask(a, b)
t_or = time.time() - t0
ctx = hrt.Context(0); ctx.load_scene(tl.scene_path(scene))
t0 = time.time()
res = ctx.run(rx, tx, rxv, txv, 3.5, P, B, dense=True, raysinfo=True, trace=True, summary=True)
t_gpu = time.time() - t0
wr, wt = tl.outputs_words(a), tl.outputs_words(res["out"])
tl.assert_exact(wr, mask, wt)
tl.assert_gains_close(wr, mask, wt)
assert np.array_equal(tr["hit_tri"], res["trace"]["hit_tri"]) and np.array_equal(tr["slot_state"], res["trace"]["slot_state"])
pair, bounce = tl.oracle_summaries(a, tr)
tl.assert_summaries_equal(pair, bounce, res["pair"], res["bounce"])
nv = int((tr["slot_state"] == 1).sum()); no = int((tr["slot_state"] == 2).sum())
print(f"OK: {P} rays x {len(tx)} TX x {len(rx)} RX x {B} bounces: {nv} valid paths, {no} occluded slots, "
      f"{int((tr['hit_tri'] < tl.IDLE).sum())} primary queries -- every reference-written word of tau / directions / "
      f"freq_shift / RaysInfo / LoS bit-identical, hit triangles identical, gains within {tl.GAIN_RTOL} "
      f"(oracle {t_or:.0f} s on one core, GPU run incl. 1.6 GB of D2H {t_gpu:.2f} s)")
