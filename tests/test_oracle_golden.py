"""oracle/hrt_oracle.c against the committed golden vectors (generated from the
unmodified reference by tests/golden/make_golden.py).  Runs anywhere."""
import numpy as np
import pytest

import hrt_testlib as tl


@pytest.mark.parametrize("name", tl.GOLDEN_NAMES)
def test_oracle_reproduces_golden(name):
    g = tl.load_golden(name)
    o, tr = tl.run_oracle(g["scene"], g["rx"], g["tx"], g["rxv"], g["txv"], g["f"],
                          g["P"], g["B"], fill=0x00)
    w = tl.outputs_words(o)
    for k, v in w.items():
        m = g["mask." + k]
        assert np.array_equal(v[m], g["out." + k][m]), k
    assert np.array_equal(tr["hit_tri"], g["trace.hit_tri"])
    assert np.array_equal(tr["slot_state"], g["trace.slot_state"])


def test_survey_anchor_counts():
    """Hit counts measured on the reference during the survey."""
    g = tl.load_golden("reflector_testc")
    hits0 = int((g["trace.hit_tri"][0, 0] < tl.IDLE).sum())
    assert hits0 == 4996
    assert int((g["trace.hit_tri"][0, 1] < tl.IDLE).sum()) == 0      # nothing after bounce 0
    assert int((g["trace.hit_tri"][0, 1] == tl.NONE).sum()) == 4996  # 4996 still traced
    g = tl.load_golden("reflector_testpy")
    assert int((g["trace.hit_tri"][0, 0] < tl.IDLE).sum()) == 3690


def test_los_known_answer():
    """LoS analytic value: canyon tx(0,0,10) -> rx(0,0,1.5)."""
    o, _ = tl.run_oracle("simple_street_canyon_with_cars", [[0, 0, 1.5]], [[0, 0, 10]],
                         [[0, 0, 0]], [[0, 0, 0]], 3.0, 8, 1, trace=False)
    assert abs(o.los["tau"][0, 0] - 8.5 / 299792458.0) < 1e-14
    assert abs(o.los["a_te_re"][0, 0] - 9.35558e-4) < 1e-8
    assert o.los["a_te_im"][0, 0] == 0.0


def test_metal_scatter_is_zero_gain():
    """Metal has s = 0 => scat_coefs returns exactly 0."""
    g = tl.load_golden("2cars_raised")
    T, B, P = g["trace.hit_tri"].shape
    # triangles 2.. of 2cars belong to the metal cars (ground is mesh 0: 2 tris)
    metal = (g["trace.hit_tri"] >= 2) & (g["trace.hit_tri"] < tl.IDLE)
    st = g["trace.slot_state"][0]            # rx 0
    te = tl.f32(g["out.scat.a_te_re"]).reshape(-1, T, B, P)[0]
    tau = tl.f32(g["out.scat.tau"]).reshape(-1, T, B, P)[0]
    sel = metal & (st == 1)
    assert sel.any()
    assert np.all(te[sel] == 0.0) and np.all(tau[sel] > 0.0)
