"""The CPU restatement (oracle/hrt_oracle.c) against the unmodified reference
compiled in oracle/_ref -- every word the reference determines must be equal
bit for bit.  This is what pins the oracle (the reference has no golden
vectors of its own: its test/test.py:61-87 asserts shapes only)."""
import numpy as np
import pytest

import hrt_testlib as tl

pytestmark = pytest.mark.skipif(not tl.ref_available(), reason="oracle/_ref not built")

CASES = [
    # (config, P, B, extra rx, velocities)
    ("reflector_testc", 30000, 3, None, False),
    ("reflector_testpy", 10000, 3, None, False),
    ("box_axis", 20000, 3, None, False),
    ("box_generic", 20000, 3, [[2.0, 2.0, 4.0], [-4.0, -4.0, 0.5]], True),
    ("2cars_origin", 20000, 5, None, False),
    ("2cars_raised", 20000, 5, [[-6.0, 4.0, 1.0]], True),
    ("canyon_1x1", 6000, 5, [[20.0, 2.0, 1.5], [-30.0, -2.0, 1.5]], True),
    # non-zero Mesh.velocity on cars, ground and buildings (scenes/canyon_moving.hrt)
    ("canyon_moving", 6000, 5, [[20.0, 2.0, 1.5], [-30.0, -2.0, 1.5]], True),
]


@pytest.mark.parametrize("name,P,B,extra_rx,moving", CASES)
def test_oracle_bit_exact_vs_reference(name, P, B, extra_rx, moving):
    scene, rx, tx, f = tl.CONFIGS[name]
    rx = list(rx) + (extra_rx or [])
    rxv = [[0.5 * i, -1.0, 0.25] if moving else [0, 0, 0] for i in range(len(rx))]
    txv = [[3.0, 1.0, -0.5] if moving else [0, 0, 0] for _ in tx]
    a = tl.run_ref(scene, rx, tx, rxv, txv, f, P, B, fill=0x00)
    b = tl.run_ref(scene, rx, tx, rxv, txv, f, P, B, fill=0x5A)
    mask = tl.written_mask(a, b)
    o, tr = tl.run_oracle(scene, rx, tx, rxv, txv, f, P, B, fill=0x00)
    wa, wo = tl.outputs_words(a), tl.outputs_words(o)
    for k in wa:
        m = mask[k]
        assert np.array_equal(wa[k][m], wo[k][m]), k
    # the oracle must also leave untouched exactly what the reference leaves
    # untouched (it was pre-filled with 0x00 like run `a`)
    o2, _ = tl.run_oracle(scene, rx, tx, rxv, txv, f, P, B, fill=0x5A, trace=False)
    mask_o = tl.written_mask(o, o2)
    for k in mask:
        assert np.array_equal(mask[k], mask_o[k]), k
    # trace consistency: slot_state != 0  <=>  tau word is determined
    assert np.array_equal(tr["slot_state"].reshape(-1) != 0, mask["scat.tau"])
    if name == "canyon_moving":
        # the mesh-velocity term really is in play: freq_shift of valid paths is not just the TX base value
        fs = o.scat["freq_shift"].reshape(-1)
        valid = tr["slot_state"].reshape(-1) == 1
        assert mask["scat.freq_shift"][valid].all()
        assert np.unique(fs[valid]).size > 5000


def test_two_tx_layout_quirks():
    """T = 2: freq_shift / RaysInfo index algebra of the reference (SURVEY A-8,
    A-9) -- determinate words still agree."""
    scene, rx, tx, f = tl.CONFIGS["box_generic"]
    tx = list(tx) + [[-2.0, 3.0, 1.0]]
    rx = list(rx) + [[3.0, 3.0, 3.0]]
    rxv = [[0, 0, 0]] * 2
    txv = [[1.0, 2.0, 3.0], [-2.0, 0.5, 0.0]]
    a = tl.run_ref(scene, rx, tx, rxv, txv, f, 4000, 3, fill=0x00)
    b = tl.run_ref(scene, rx, tx, rxv, txv, f, 4000, 3, fill=0x5A)
    mask = tl.written_mask(a, b)
    o, _ = tl.run_oracle(scene, rx, tx, rxv, txv, f, 4000, 3, fill=0x00, trace=False)
    wa, wo = tl.outputs_words(a), tl.outputs_words(o)
    for k in wa:
        assert np.array_equal(wa[k][mask[k]], wo[k][mask[k]]), k


def test_material_table_matches_reference():
    import ctypes as C
    from hrt_b200 import abi
    ref = (abi.Material * 17).in_dll(tl.ref_lib(), "g_materials")
    ours = (abi.Material * 17).in_dll(tl.oracle_lib(), "g_materials")
    for i in range(17):
        for fld, _ in abi.Material._fields_:
            if fld == "name":
                assert ref[i].name[:ref[i].name_sz] == ours[i].name[:ours[i].name_sz]
            else:
                assert getattr(ref[i], fld) == getattr(ours[i], fld), (i, fld)
    for nm in ["air", "glass1", "ceiling_board2", "wet_ground", "metal", "nonsense", "glass"]:
        tl.ref_lib().get_material_index.argtypes = [C.c_char_p]
        tl.oracle_lib().get_material_index.argtypes = [C.c_char_p]
        assert tl.ref_lib().get_material_index(nm.encode()) == \
            tl.oracle_lib().get_material_index(nm.encode())


@pytest.mark.parametrize("scene", ["simple_reflector", "box", "2cars", "simple_street_canyon_with_cars", "canyon_moving"])
def test_normals_match_reference(scene):
    """oracle_normals() == the Mesh.ns the reference's precompute_normals leaves
    in the caller's scene (src/compute_paths.c:208-224), bit for bit."""
    from hrt_b200 import abi
    lib = tl.ref_lib()
    sc = lib.scene_load(tl.scene_path(scene).encode())
    try:
        abi.call_compute_paths(lib, sc, [[0, 0, 1.5]], [[0, 0, 3.0]], [[0, 0, 0]], [[0, 0, 0]], 3.0, 8, 1)
        ns = tl.mesh_normals(sc)
    finally:
        abi.free_scene(sc)
    assert np.array_equal(ns.view(np.uint32), tl.oracle_normals(scene).view(np.uint32))
