"""Generates tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE
(oracle/_ref/libhrt_ref.so, compiled from /root/reference/src by oracle/Makefile).

Run in the authoring container only:  python tests/golden/make_golden.py
Each file holds the inputs, every output array of the reference pre-filled with
0x00, the mask of words the reference determines (two-fill technique, SURVEY
section 8c) and the hit trace of the CPU restatement, which test_oracle_vs_ref
proves bit-identical to the reference on the determined words.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", "..", "hermespy-rt_b200"))
import hrt_testlib as tl  # noqa: E402

GOLDEN = {
    # name: (config, P, B, extra_rx, moving)
    "reflector_testc": ("reflector_testc", 30000, 3, [], False),
    "reflector_testpy": ("reflector_testpy", 10000, 3, [], False),
    "box_generic": ("box_generic", 3000, 3, [[2.0, 2.0, 4.0]], True),
    "2cars_raised": ("2cars_raised", 3000, 5, [[-6.0, 4.0, 1.0]], True),
    "canyon_3rx": ("canyon_1x1", 2500, 5, [[20.0, 2.0, 1.5], [-30.0, -2.0, 1.5]], True),
    # moving meshes (Mesh.velocity != 0): reference src/compute_paths.c:720-722
    "canyon_moving": ("canyon_moving", 2500, 5, [[20.0, 2.0, 1.5], [-30.0, -2.0, 1.5]], True),
}


def main():
    only = set(sys.argv[1:])
    for name, (cfg, P, B, extra, moving) in GOLDEN.items():
        if only and name not in only:
            continue
        scene, rx, tx, f = tl.CONFIGS[cfg]
        rx = list(rx) + extra
        rxv = [[0.5 * i, -1.0, 0.25] if moving else [0, 0, 0] for i in range(len(rx))]
        txv = [[3.0, 1.0, -0.5] if moving else [0, 0, 0] for _ in tx]
        a = tl.run_ref(scene, rx, tx, rxv, txv, f, P, B, fill=0x00)
        b = tl.run_ref(scene, rx, tx, rxv, txv, f, P, B, fill=0x5A)
        mask = tl.written_mask(a, b)
        _, tr = tl.run_oracle(scene, rx, tx, rxv, txv, f, P, B)
        d = dict(scene=scene, rx=np.asarray(rx, np.float32), tx=np.asarray(tx, np.float32),
                 rxv=np.asarray(rxv, np.float32), txv=np.asarray(txv, np.float32),
                 f=np.float32(f), P=P, B=B)
        for k, v in tl.outputs_words(a).items():
            d["out." + k] = v
            d["mask." + k] = np.packbits(mask[k])
        d["trace.hit_tri"] = tr["hit_tri"]
        d["trace.slot_state"] = tr["slot_state"]
        d["trace.shadow_tri"] = tr["shadow_tri"]
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print(name, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
