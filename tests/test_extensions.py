"""Opt-in extensions next to the path: refraction rays
with ITU-R P.2040-3 (31c)/(31d) gains and the three-lobe scattering pattern the
Material fields s1/s2/s3/s1_alpha/s3_alpha describe -- TODOs of the reference
(src/compute_paths.c:587, :726-728, :414), so there is NO reference behaviour to
match.  The definitions are stated in double precision in oracle/hrt_oracle.c
("extensions"); here they are checked for physical sanity, and the fp32 code the
kernels run (csrc/hrt_ext.cuh) is checked against them -- on the CPU through
tests/emul, on the GPU through hrt_run."""
import ctypes as C

import numpy as np
import pytest

import hrt_testlib as tl
import hrt_b200 as hrt


def _orc():
    L = tl.oracle_lib()
    L.oracle_ext_refr_coefs.argtypes = [C.c_uint32, C.c_float, C.c_double, C.POINTER(C.c_double)]
    L.oracle_ext_refract_dir.argtypes = [C.c_uint32, C.c_float, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.oracle_ext_pattern_pi.restype = C.c_double
    L.oracle_ext_pattern_pi.argtypes = [C.c_double] * 3 + [C.c_int] * 2 + [C.POINTER(C.c_double)] * 3
    L.oracle_ext_lobe_norm.restype = C.c_double
    L.oracle_ext_lobe_norm.argtypes = [C.c_int, C.c_double, C.c_double]
    return L


def _emul():
    L = tl.emul_lib()
    L.emul_ext_refr_coefs.argtypes = [C.c_uint32, C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                      C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]
    L.emul_ext_pattern_pi.restype = C.c_float
    L.emul_ext_pattern_pi.argtypes = [C.c_float] * 3 + [C.c_int] * 2 + [C.POINTER(C.c_float)] * 3
    L.emul_ext_lobe_norm.restype = C.c_float
    L.emul_ext_lobe_norm.argtypes = [C.c_int, C.c_float, C.c_float]
    return L


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def test_transmission_coefficients_are_physical():
    """(31c)/(31d): at normal incidence T_TE = T_TM = 2 / (1 + sqrt(eta)); for a lossless
    dielectric |R|^2 + Re(sqrt(eta - sin^2)) / cos(th) |T_TE|^2 = 1 with R of (31a)."""
    O = _orc()
    out = (C.c_double * 4)()
    for mat, f in ((1, 3.5), (5, 28.0), (11, 3.5)):          # concrete, glass, marble
        O.oracle_ext_refr_coefs(mat, f, 0.0, out)
        te, tm = complex(out[0], out[1]), complex(out[2], out[3])
        assert abs(te - tm) < 1e-12
    # lossless: wood-like a=1.99 with c -> use material 0 "air" (eta = 1, no interface): T = 1, no bending
    for th in (0.0, 0.4, 1.2):
        O.oracle_ext_refr_coefs(0, 3.5, th, out)
        assert abs(complex(out[0], out[1]) - 1) < 1e-3 and abs(complex(out[2], out[3]) - 1) < 1e-3
    d, n, t = _d3([np.sin(0.5), 0, -np.cos(0.5)]), _d3([0, 0, 1]), _d3([0, 0, 0])
    assert O.oracle_ext_refract_dir(1, 3.5, d, n, t) == 1
    # Snell: sin(th2) = sin(th1) / n, n = Re sqrt(eta) of concrete at 3.5 GHz ~ 2.29
    assert abs(np.hypot(t[0], t[1]) - np.sin(0.5) / np.sqrt(5.24)) < 2e-3 and t[2] < 0


@pytest.mark.parametrize("alpha", [1, 2, 4, 7, 12, 19])
def test_lobes_integrate_to_one(alpha):
    """Each lobe of the three-lobe model integrates to 1 over the hemisphere above the
    surface (what F_alpha is for), for any incidence angle."""
    O = _orc()
    nth, nph = 400, 400
    th = (np.arange(nth) + 0.5) / nth * (np.pi / 2); ph = (np.arange(nph) + 0.5) / nph * 2 * np.pi
    n = _d3([0, 0, 1])
    for th_i in (0.0, 0.5, 1.2):
        ki = np.array([np.sin(th_i), 0.0, -np.cos(th_i)])
        for lobes in ((1, 0, 0), (0, 1, 0), (0, 0, 1)):
            tot = 0.0
            for t in th[::8]:
                for p_ in ph[::8]:
                    ks = np.array([np.sin(t) * np.cos(p_), np.sin(t) * np.sin(p_), np.cos(t)])
                    tot += O.oracle_ext_pattern_pi(*lobes, alpha, alpha, _d3(ki), _d3(ks), n) / np.pi * np.sin(t)
            tot *= (np.pi / 2 / (nth / 8)) * (2 * np.pi / (nph / 8))
            assert abs(tot - 1.0) < 2e-2, (alpha, th_i, lobes, tot)


def test_fp32_extension_functions_match_the_oracle():
    """csrc/hrt_ext.cuh in fp32 (what the kernels run) against the double-precision statement."""
    O, E = _orc(), _emul()
    rng = np.random.default_rng(5)
    o4, e4 = (C.c_double * 4)(), (C.c_float * 4)()
    for _ in range(3000):
        mat = int(rng.integers(1, 17)); f = float(rng.choice([0.9, 3.5, 28.0, 70.0])); th = float(rng.uniform(0, 1.55))
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        O.oracle_ext_refr_coefs(mat, f, th, o4)
        out_d, ok = (C.c_float * 3)(), C.c_int(0)
        E.emul_ext_refr_coefs(mat, f, th, e4, _f3(d), _f3(n), out_d, C.byref(ok))
        ref = np.array(o4[:]); got = np.array(e4[:])
        if mat != 13:                                   # metal: |eta| ~ 1e8, fp32 loses the small real part of T
            for k in (0, 2):                            # as complex numbers, like the gain tolerance
                zr, zg = complex(ref[k], ref[k + 1]), complex(got[k], got[k + 1])
                assert abs(zr - zg) <= 1e-4 * abs(zr), (mat, f, th, ref, got)
        t = _d3([0, 0, 0])
        if O.oracle_ext_refract_dir(mat, f, _d3(d), _d3(n), t) and mat != 13:
            assert ok.value == 1 and np.allclose(np.array(out_d[:]), np.array(t[:]), atol=2e-5)
    for _ in range(3000):
        a1, a3 = int(rng.integers(1, 20)), int(rng.integers(1, 20))
        s = rng.dirichlet([1, 1, 1])
        n = np.array([0.0, 0.0, 1.0])
        ki = rng.normal(size=3); ki[2] = -abs(ki[2]) - 0.05; ki /= np.linalg.norm(ki)
        ks = rng.normal(size=3); ks[2] = abs(ks[2]) + 0.02; ks /= np.linalg.norm(ks)
        ref = O.oracle_ext_pattern_pi(*s, a1, a3, _d3(ki), _d3(ks), _d3(n))
        got = E.emul_ext_pattern_pi(*[float(x) for x in s], a1, a3, _f3(ki), _f3(ks), _f3(n))
        assert abs(got - ref) <= 2e-4 * abs(ref) + 1e-7, (a1, a3, ref, got)


@pytest.mark.gpu
def test_extensions_on_the_gpu_vs_oracle():
    """hrt_run with HRT_FLAG_EXT_LOBES | HRT_FLAG_EXT_REFRACT on the moving canyon (real
    materials): scatter gains and the refraction-ray list against oracle_compute_paths_ext;
    delays / directions / hit trace stay those of the plain run."""
    scene, rx, tx, f = tl.CONFIGS["canyon_moving"]
    rx = list(rx) + [[20.0, 2.0, 1.5], [-30.0, -2.0, 1.5]]
    rxv, txv = np.zeros((3, 3)), np.zeros((1, 3))
    P, B = 8000, 4
    cap = P * B
    a, tr = tl.run_oracle(scene, rx, tx, rxv, txv, f, P, B, ext=3, refract_capacity=cap)
    with hrt.Context(0) as ctx:
        ctx.load_scene(tl.scene_path(scene))
        plain = ctx.run(rx, tx, rxv, txv, f, P, B, dense=True, trace=True)
        res = ctx.run(rx, tx, rxv, txv, f, P, B, dense=True, trace=True, ext_lobes=True, refract=cap)
    assert np.array_equal(tr["hit_tri"], res["trace"]["hit_tri"]) and np.array_equal(tr["slot_state"], res["trace"]["slot_state"])
    assert np.array_equal(plain["out"].scat["tau"].view(np.uint32), res["out"].scat["tau"].view(np.uint32))
    valid = tr["slot_state"].reshape(-1) == 1
    assert valid.sum() > 20000
    worst = 0.0
    for pol in ("te", "tm"):
        ar = a.scat[f"a_{pol}_re"].reshape(-1)[valid].astype(np.float64); ai = a.scat[f"a_{pol}_im"].reshape(-1)[valid].astype(np.float64)
        br = res["out"].scat[f"a_{pol}_re"].reshape(-1)[valid].astype(np.float64); bi = res["out"].scat[f"a_{pol}_im"].reshape(-1)[valid].astype(np.float64)
        mag = np.hypot(ar, ai); err = np.hypot(ar - br, ai - bi)
        assert (err <= 3e-4 * mag + 1e-30).all(), float((err / np.maximum(mag, 1e-300)).max())
        worst = max(worst, float((err[mag > 0] / mag[mag > 0]).max()))
        # and the pattern really changes the gains with respect to the reference's placeholder
        pr = plain["out"].scat[f"a_{pol}_re"].reshape(-1)[valid].astype(np.float64)
        assert np.abs(pr - br).max() > 0
    # refraction rays: same set, same gains and geometry
    assert res["refract_found"] == tr["refract_found"] == len(res["refract"]) > 5000
    key = lambda q: np.lexsort((q["path"], q["bounce"], q["tx"]))
    ro, rg = tr["refract"][key(tr["refract"])], res["refract"][key(res["refract"])]
    for k in ("path", "tx", "bounce"):
        assert np.array_equal(ro[k], rg[k])
    metal = np.zeros(len(ro), bool)                   # cars: |eta| ~ 1e8
    to = ro["t_te_re"] + 1j * ro["t_te_im"]; tg = rg["t_te_re"] + 1j * rg["t_te_im"]
    small = np.abs(to) < 1e-6 * np.abs(to).max()
    assert (np.abs(to - tg)[~small] <= 5e-4 * np.abs(to)[~small]).all()
    assert np.allclose(ro["d"], rg["d"], atol=5e-5) and np.allclose(ro["o"], rg["o"], atol=2e-4)
