"""Test-side plumbing: loads the CPU checkers (oracle/liboracle.so and, when
present, oracle/_ref/libhrt_ref.so) and runs them through the reference ABI.
Only tests/, smoke() and bench.py's CPU-baseline legs may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from hrt_b200 import abi

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
SCENES = os.path.join(ROOT, "scenes")
ORACLE_DIR = os.path.join(ROOT, "oracle")
NONE = 0xFFFFFFFF
IDLE = 0xFFFFFFFE


def scene_path(name: str) -> str:
    return os.path.join(SCENES, name if name.endswith(".hrt") else name + ".hrt")


class OrcTrace(C.Structure):
    _fields_ = [("hit_tri", C.c_void_p), ("hit_t", C.c_void_p), ("hit_theta", C.c_void_p),
                ("slot_state", C.c_void_p), ("shadow_tri", C.c_void_p),
                ("theta_used", C.c_void_p)]


_oracle = None
_ref = None


def oracle_lib() -> C.CDLL:
    global _oracle
    if _oracle is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(so):
            subprocess.run(["make", "-C", ORACLE_DIR, "liboracle.so"], check=True,
                           capture_output=True)
        lib = C.CDLL(so)
        lib.scene_load.restype = abi.Scene
        lib.scene_load.argtypes = [C.c_char_p]
        lib.oracle_compute_paths.restype = C.c_int
        lib.oracle_compute_paths.argtypes = [
            C.POINTER(abi.Scene), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_float, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t,
            C.POINTER(abi.ChannelInfo), C.POINTER(abi.RaysInfo),
            C.POINTER(abi.ChannelInfo), C.POINTER(abi.RaysInfo), C.POINTER(OrcTrace)]
        lib.oracle_closest_hits.restype = C.c_int
        lib.oracle_closest_hits.argtypes = [C.POINTER(abi.Scene), C.c_void_p, C.c_size_t,
                                            C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_launch_dirs.restype = None
        lib.oracle_launch_dirs.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p]
        _oracle = lib
    return _oracle


def ref_available() -> bool:
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libhrt_ref.so"))


def ref_lib() -> C.CDLL:
    global _ref
    if _ref is None:
        lib = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libhrt_ref.so"))
        abi.bind_compute_paths(lib)
        _ref = lib
    return _ref


def run_ref(scene, rx, tx, rxv, txv, f_ghz, P, B, fill=0):
    """The unmodified reference, outputs pre-filled with `fill` bytes."""
    lib = ref_lib()
    sc = lib.scene_load(scene_path(scene).encode())
    try:
        return abi.call_compute_paths(lib, sc, rx, tx, rxv, txv, f_ghz, P, B, fill=fill)
    finally:
        abi.free_scene(sc)


REFRACT_DTYPE = np.dtype([("path", "<u4"), ("tx", "<u2"), ("bounce", "<u2"), ("o", "<f4", (3,)), ("d", "<f4", (3,)),
                          ("t_te_re", "<f4"), ("t_te_im", "<f4"), ("t_tm_re", "<f4"), ("t_tm_im", "<f4")])


def run_oracle(scene, rx, tx, rxv, txv, f_ghz, P, B, fill=0, trace=True, ext=0, refract_capacity=0):
    """Our CPU restatement.  Returns (Outputs, trace dict).  ext: 1 = three-lobe
    scattering pattern, 2 = refraction rays (tr["refract"]) -- the opt-in
    extensions of SURVEY 8 f4 (no reference behaviour), oracle_compute_paths_ext."""
    lib = oracle_lib()
    sc = lib.scene_load(scene_path(scene).encode())
    rx = abi.vec3_array(rx); tx = abi.vec3_array(tx)
    R, T = rx.shape[0], tx.shape[0]
    rxv = abi.vec3_array(rxv, R); txv = abi.vec3_array(txv, T)
    o = abi.alloc_outputs(R, T, P, B, fill)
    tr = {}
    ot = OrcTrace()
    if trace:
        tr["hit_tri"] = np.full((T, B, P), IDLE, np.uint32)
        tr["hit_t"] = np.zeros((T, B, P), np.float32)
        tr["hit_theta"] = np.zeros((T, B, P), np.float32)
        tr["slot_state"] = np.zeros((R, T, B, P), np.uint8)
        tr["shadow_tri"] = np.full((R, T, B, P), NONE, np.uint32)
        tr["theta_used"] = np.zeros((R, T, B, P), np.float32)
        for k in tr:
            setattr(ot, k, tr[k].ctypes.data)
    los = abi.chan_struct(o.los, 1)
    scs = abi.chan_struct(o.scat, B * P)
    rl = abi.RaysInfo(1, 1, o.los_rays.ctypes.data, o.los_active.ctypes.data)
    rs = abi.RaysInfo(B + 1, P, o.scat_rays.ctypes.data, o.scat_active.ctypes.data)
    try:
        if ext:
            lib.oracle_compute_paths_ext.restype = C.c_int
            lib.oracle_compute_paths_ext.argtypes = lib.oracle_compute_paths.argtypes + [C.c_uint, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
            rbuf = np.zeros(max(int(refract_capacity), 1), REFRACT_DTYPE)
            nref = C.c_size_t(0)
            rc = lib.oracle_compute_paths_ext(C.byref(sc), rx.ctypes.data, tx.ctypes.data, rxv.ctypes.data, txv.ctypes.data,
                                              C.c_float(f_ghz), R, T, P, B, C.byref(los), C.byref(rl), C.byref(scs), C.byref(rs),
                                              C.byref(ot) if trace else None, int(ext), rbuf.ctypes.data, int(refract_capacity),
                                              C.byref(nref))
            tr["refract_found"] = int(nref.value)
            tr["refract"] = rbuf[: min(int(nref.value), int(refract_capacity))]
        else:
            rc = lib.oracle_compute_paths(C.byref(sc), rx.ctypes.data, tx.ctypes.data,
                                          rxv.ctypes.data, txv.ctypes.data, C.c_float(f_ghz),
                                          R, T, P, B, C.byref(los), C.byref(rl),
                                          C.byref(scs), C.byref(rs),
                                          C.byref(ot) if trace else None)
        assert rc == 0
    finally:
        abi.free_scene(sc)
    return o, tr


def outputs_words(o: abi.Outputs) -> dict:
    """Every output array of a call as a flat uint32/uint8 view, keyed by name."""
    w = {}
    for k in abi.CHAN_FIELDS:
        w["los." + k] = o.los[k].view(np.uint32).reshape(-1)
        w["scat." + k] = o.scat[k].view(np.uint32).reshape(-1)
    w["los_rays"] = o.los_rays.view(np.uint32).reshape(-1)
    w["scat_rays"] = o.scat_rays.view(np.uint32).reshape(-1)
    w["los_active"] = o.los_active.reshape(-1)
    w["scat_active"] = o.scat_active.reshape(-1)
    return w


def written_mask(a: abi.Outputs, b: abi.Outputs) -> dict:
    """Words identical in two runs that differed only in the pre-fill pattern
    are the ones the callee determines ('How to compare')."""
    wa, wb = outputs_words(a), outputs_words(b)
    return {k: wa[k] == wb[k] for k in wa}


# Geometry of the named configurations ("Concrete inputs").
CONFIGS = {
    # name: (scene, rx, tx, f_GHz)
    "reflector_testc": ("simple_reflector", [[0, 0, .5]], [[0, 0, .5]], 3.0),
    "reflector_testpy": ("simple_reflector", [[0, 0, .15]], [[0, 0, .151]], 3.0),
    "box_axis": ("box", [[0, 0, 1]], [[0, 0, 2.5]], 3.0),
    "box_generic": ("box", [[-3, 1.5, 1]], [[1, -2, 2.5]], 3.0),
    "2cars_origin": ("2cars", [[0, 0, 0]], [[0, 0, 0]], 70.0),
    "2cars_raised": ("2cars", [[2, -1, 1.5]], [[0, 0, 1.5]], 70.0),
    "canyon_1x1": ("simple_street_canyon_with_cars", [[0, 0, 1.5]], [[0, 0, 10]], 3.5),
    # scripts/make_moving_scene.py: the canyon with non-zero Mesh.velocity (cars +-14 m/s,
    # drifting ground, moving buildings) and real materials -- the mesh-velocity
    # Doppler term of reference src/compute_paths.c:720-722
    "canyon_moving": ("canyon_moving", [[0, 0, 1.5]], [[0, 0, 10]], 3.5),
}


def canyon_c4_positions(num_tx=4, num_rx=64):
    """BASELINE config 4: TX i at (-45+30i, 0, 10); RX on a 16x4
    grid x=-60+8j, y in {-3,-1,1,3}, z=1.5."""
    tx = [[-45.0 + 30.0 * i, 0.0, 10.0] for i in range(num_tx)]
    rx = [[-60.0 + 8.0 * j, y, 1.5] for y in (-3.0, -1.0, 1.0, 3.0) for j in range(16)]
    return np.asarray(rx[:num_rx], np.float32), np.asarray(tx, np.float32)


# ---------------------------------------------------------------- goldens

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_NAMES = ("reflector_testc", "reflector_testpy", "box_generic", "2cars_raised",
                "canyon_3rx", "canyon_moving")


def load_golden(name: str) -> dict:
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["scene"] = str(g["scene"])
    g["P"] = int(g["P"]); g["B"] = int(g["B"]); g["f"] = float(g["f"])
    for k in list(g):
        if k.startswith("mask."):
            n = g["out." + k[5:]].size
            g[k] = np.unpackbits(g[k])[:n].astype(bool)
    return g


# Output words that must be BIT-EXACT on the GPU (they depend only on exactly
# rounded fp32 + - * / sqrt in a fixed order) ...
EXACT_KEYS = ("scat.tau", "scat.directions_rx", "scat.freq_shift", "scat_rays", "scat_active",
              "los_rays", "los_active") + tuple("los." + k for k in abi.CHAN_FIELDS)
# ... and the complex gains, which pass through sinf/cosf/expf/acosf and are
# compared with the north star's fp32 relative tolerance.
GAIN_KEYS = ("scat.a_te_re", "scat.a_te_im", "scat.a_tm_re", "scat.a_tm_im")
GAIN_RTOL = 1e-4


def f32(words: np.ndarray) -> np.ndarray:
    return words.view(np.float32)


def assert_exact(ref_words, mask, test_words, keys=EXACT_KEYS):
    """Bit equality on reference-determined words (+0 == -0 tolerated: the
    reference's `x += 0*v` Doppler no-op can flip the sign of a zero)."""
    for k in keys:
        m = mask[k]
        a, b = ref_words[k][m], test_words[k][m]
        if a.dtype == np.uint32:
            bad = (a != b) & ~(((a | b) & 0x7FFFFFFF) == 0)
        else:
            bad = a != b
        assert not bad.any(), f"{k}: {int(bad.sum())} of {a.size} words differ " \
                              f"(first at {int(np.flatnonzero(bad)[0])})"


def assert_gains_close(ref_words, mask, test_words, rtol=GAIN_RTOL):
    """|delta a| <= rtol * |a| as complex numbers, per polarisation and slot."""
    for pol in ("te", "tm"):
        m = mask[f"scat.a_{pol}_re"]
        ar = f32(ref_words[f"scat.a_{pol}_re"])[m].astype(np.float64)
        ai = f32(ref_words[f"scat.a_{pol}_im"])[m].astype(np.float64)
        br = f32(test_words[f"scat.a_{pol}_re"])[m].astype(np.float64)
        bi = f32(test_words[f"scat.a_{pol}_im"])[m].astype(np.float64)
        err = np.hypot(ar - br, ai - bi)
        mag = np.hypot(ar, ai)
        bad = err > rtol * mag + 1e-38
        assert not bad.any(), f"a_{pol}: {int(bad.sum())} of {ar.size} gains off, " \
                              f"worst rel {float((err / np.maximum(mag, 1e-300)).max()):.3e}"


# ------------------------------------------- serial CPU run of the kernel logic

_emul = None


def emul_lib() -> C.CDLL:
    """tests/emul/libhrt_emul.so: hrt_core.cuh / hrt_bvh.cuh compiled for the
    host (test-only, see tests/emul/hrt_emul.cu)."""
    global _emul
    if _emul is None:
        so = os.path.join(ROOT, "tests", "emul", "libhrt_emul.so")
        if not os.path.exists(so):
            subprocess.run(["make", "-C", os.path.join(ROOT, "hermespy-rt_b200"), "emul"],
                           check=True, capture_output=True)
        lib = C.CDLL(so)
        lib.scene_load.restype = abi.Scene
        lib.scene_load.argtypes = [C.c_char_p]
        lib.emul_closest_hits.argtypes = [C.POINTER(abi.Scene), C.c_void_p, C.c_size_t, C.c_int,
                                          C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.emul_compute_paths.argtypes = [
            C.POINTER(abi.Scene), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
            C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t,
            C.POINTER(abi.ChannelInfo), C.POINTER(abi.ChannelInfo), C.c_void_p, C.c_void_p,
            C.c_int, C.c_float, C.c_int]
        lib.emul_bvh_stats.argtypes = [C.POINTER(abi.Scene), C.c_int, C.c_void_p, C.c_void_p]
        _emul = lib
    return _emul


def run_emul(scene, rx, tx, rxv, txv, f_ghz, P, B, leaf_max=2, pad_ulps=64.0, brute=0, closed_form=False,
             rx_map=False):
    lib = emul_lib()
    sc = lib.scene_load(scene_path(scene).encode())
    rx = abi.vec3_array(rx); tx = abi.vec3_array(tx)
    R, T = rx.shape[0], tx.shape[0]
    rxv = abi.vec3_array(rxv, R); txv = abi.vec3_array(txv, T)
    o = abi.alloc_outputs(R, T, P, B, 0)
    tr = {"hit_tri": np.zeros((T, B, P), np.uint32), "slot_state": np.zeros((R, T, B, P), np.uint8)}
    los = abi.chan_struct(o.los, 1)
    scs = abi.chan_struct(o.scat, B * P)
    try:
        rc = lib.emul_compute_paths(C.byref(sc), rx.ctypes.data, tx.ctypes.data, rxv.ctypes.data,
                                    txv.ctypes.data, C.c_float(f_ghz), R, T, P, B,
                                    C.byref(los), C.byref(scs), tr["hit_tri"].ctypes.data,
                                    tr["slot_state"].ctypes.data, leaf_max, C.c_float(pad_ulps),
                                    int(brute) | (0x100 if closed_form else 0) | (0x200 if rx_map else 0))
        assert rc == 0
    finally:
        abi.free_scene(sc)
    return o, tr


def random_rays(scene_name, n, seed=0):
    """Rays with origins inside the scene's bounding box (slightly inflated) and
    isotropic directions, plus a share of axis-aligned and grazing directions."""
    lib = oracle_lib()
    sc = lib.scene_load(scene_path(scene_name).encode())
    tris, *_ = abi.scene_to_numpy(sc)
    abi.free_scene(sc)
    lo = tris.reshape(-1, 3).min(0); hi = tris.reshape(-1, 3).max(0)
    ext = np.maximum(hi - lo, 1.0)
    rng = np.random.default_rng(seed)
    o = (lo - 0.1 * ext + rng.random((n, 3)) * 1.2 * ext).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    k = n // 8
    d[:k, 2] *= 1e-3                                   # grazing w.r.t. horizontal planes
    d[k:2 * k] = np.eye(3)[rng.integers(0, 3, k)] * rng.choice([-1.0, 1.0], (k, 1))  # axis aligned
    # a share aimed exactly at triangle corners / edge midpoints (ties, edges)
    tgt = tris[rng.integers(0, len(tris), k)]
    w = rng.random((k, 3)); w[: k // 2, 2] = 0.0; w[: k // 4, 1] = 0.0
    w /= w.sum(1, keepdims=True)
    pts = (tgt * w[:, :, None]).sum(1)
    d[2 * k:3 * k] = pts - o[2 * k:3 * k]
    rays = np.concatenate([o, d.astype(np.float32)], axis=1).astype(np.float32)
    return np.ascontiguousarray(rays)


def scene_triangles(scene_name):
    """(N, 3, 3) float64 corners of every triangle in (mesh, face) order"""
    lib = oracle_lib()
    sc = lib.scene_load(scene_path(scene_name).encode())
    tris, *_ = abi.scene_to_numpy(sc)
    abi.free_scene(sc)
    return tris.astype(np.float64)


def grazing_rays(scene_name, n, seed=0, lo_exp=-7.5, hi_exp=-2.0):
    """Adversarial rays for conservative culling: each ray passes
    through a point on an EDGE of a random triangle while running almost inside
    that triangle's plane, |d.n| log-uniform in [10^lo_exp, 10^hi_exp].  Near
    |d.n| ~ 1e-7 the reference's fp32 det is rounding noise and its u, v, t can
    land inside the acceptance windows for a ray that passes metres away."""
    T = scene_triangles(scene_name)
    rng = np.random.default_rng(seed)
    tt = T[rng.integers(0, len(T), n)]
    e1 = tt[:, 1] - tt[:, 0]; e2 = tt[:, 2] - tt[:, 0]
    nrm = np.cross(e1, e2); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    ed = rng.integers(0, 3, n); w = rng.random(n)
    a = tt[np.arange(n), ed]; b = tt[np.arange(n), (ed + 1) % 3]
    pt = a + (b - a) * w[:, None]
    u = e1 / np.linalg.norm(e1, axis=1, keepdims=True)
    v = np.cross(nrm, u)
    ang = rng.uniform(0, 2 * np.pi, n)
    d = u * np.cos(ang)[:, None] + v * np.sin(ang)[:, None]
    d = d + nrm * (10.0 ** rng.uniform(lo_exp, hi_exp, n) * rng.choice([-1, 1], n))[:, None]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = pt - d * rng.uniform(0.5, 60, n)[:, None]
    return np.ascontiguousarray(np.concatenate([o, d], 1).astype(np.float32))


def phantom_hit_report(scene_name, rays, tri_ref, t_ref, tri, t):
    """Classifies the rays on which a BVH result differs from the reference loop:
    returns (indices, |d.n| of the reference's triangle).  See DESIGN.md,
    'phantom hits'."""
    bad = np.flatnonzero((tri_ref != tri) | (t_ref.view(np.uint32) != t.view(np.uint32)))
    if bad.size == 0:
        return bad, np.zeros(0)
    T = scene_triangles(scene_name)
    # the triangle the two disagree about: the reference's if it hit one, else ours
    which = np.where(tri_ref[bad] != NONE, tri_ref[bad], tri[bad])
    tt = T[which]
    nrm = np.cross(tt[:, 1] - tt[:, 0], tt[:, 2] - tt[:, 0]); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    d = rays[bad, 3:6].astype(np.float64); d /= np.linalg.norm(d, axis=1, keepdims=True)
    return bad, np.abs((d * nrm).sum(1))


def oracle_normals(scene_name):
    """(N, 3) float32 unit normals in (mesh, face) order (reference :208-224)."""
    lib = oracle_lib()
    sc = lib.scene_load(scene_path(scene_name).encode())
    n = sum(sc.meshes[m].num_triangles for m in range(sc.num_meshes))
    out = np.zeros((n, 3), np.float32)
    lib.oracle_normals.argtypes = [C.POINTER(abi.Scene), C.c_void_p]
    assert lib.oracle_normals(C.byref(sc), out.ctypes.data) == 0
    abi.free_scene(sc)
    return out


def mesh_normals(scene) -> np.ndarray:
    """Mesh.ns of a Scene struct after a compute_paths() call, concatenated."""
    parts = []
    for m in range(scene.num_meshes):
        me = scene.meshes[m]
        assert bool(me.ns), f"mesh {m}: ns not filled"
        parts.append(np.ctypeslib.as_array(C.cast(me.ns, C.POINTER(C.c_float)),
                                           (me.num_triangles, 3)).copy())
    return np.concatenate(parts)


def oracle_closest(scene_name, rays):
    lib = oracle_lib()
    sc = lib.scene_load(scene_path(scene_name).encode())
    n = rays.shape[0]
    tri = np.zeros(n, np.uint32); t = np.zeros(n, np.float32); th = np.zeros(n, np.float32)
    lib.oracle_closest_hits(C.byref(sc), rays.ctypes.data, n, tri.ctypes.data, t.ctypes.data,
                            th.ctypes.data)
    abi.free_scene(sc)
    return tri, t, th


def emul_closest(scene_name, rays, leaf_max=2, pad_ulps=64.0, brute=0):
    lib = emul_lib()
    sc = lib.scene_load(scene_path(scene_name).encode())
    n = rays.shape[0]
    tri = np.zeros(n, np.uint32); t = np.zeros(n, np.float32); th = np.zeros(n, np.float32)
    lib.emul_closest_hits(C.byref(sc), rays.ctypes.data, n, leaf_max, C.c_float(pad_ulps),
                          int(brute), tri.ctypes.data, t.ctypes.data, th.ctypes.data)
    abi.free_scene(sc)
    return tri, t, th


# ------------------------------------------------ summaries from an oracle run

def mix64(x: np.ndarray) -> np.ndarray:
    """numpy twin of hrt_mix64 (hrt_core.cuh)."""
    x = x.astype(np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def oracle_summaries(o: abi.Outputs, tr: dict, path_ids: np.ndarray | None = None):
    """HrtPairSummary / HrtBounceSummary (include/hrt_cuda.h) computed from an
    oracle run's outputs and trace.  Returns dicts of arrays."""
    R, T, B, P = o.R, o.T, o.B, o.P
    paths = (np.arange(P, dtype=np.uint64) if path_ids is None else path_ids.astype(np.uint64))
    hit = tr["hit_tri"]                                   # (T,B,P)
    is_hit = hit < IDLE
    key = mix64((paths[None, None, :] << np.uint64(32)) | hit.astype(np.uint64))
    with np.errstate(over="ignore"):
        bounce = {
            "n_traced": (hit != IDLE).sum(-1).astype(np.uint64),
            "n_hit": is_hit.sum(-1).astype(np.uint64),
            "hit_hash": np.where(is_hit, key, np.uint64(0)).sum(-1, dtype=np.uint64),
            "t_bits": np.where(is_hit, tr["hit_t"].view(np.uint32).astype(np.uint64),
                               np.uint64(0)).sum(-1, dtype=np.uint64),
        }
        st = tr["slot_state"]                             # (R,T,B,P)
        valid = st == 1
        tau_bits = o.scat["tau"].view(np.uint32).astype(np.uint64)
        pair = {
            "n_valid": valid.sum(-1).astype(np.uint64),
            "n_occluded": (st == 2).sum(-1).astype(np.uint64),
            "hit_hash": np.where(valid, key[None], np.uint64(0)).sum(-1, dtype=np.uint64),
            "tau_bits": np.where(valid, tau_bits, np.uint64(0)).sum(-1, dtype=np.uint64),
            "power_te": np.where(valid, o.scat["a_te_re"].astype(np.float64) ** 2
                                 + o.scat["a_te_im"].astype(np.float64) ** 2, 0.0).sum(-1),
            "power_tm": np.where(valid, o.scat["a_tm_re"].astype(np.float64) ** 2
                                 + o.scat["a_tm_im"].astype(np.float64) ** 2, 0.0).sum(-1),
        }
    return pair, bounce


def assert_summaries_equal(pair_ref, bounce_ref, pair, bounce, rtol=GAIN_RTOL * 2):
    for k in ("n_traced", "n_hit", "hit_hash", "t_bits"):
        assert np.array_equal(bounce_ref[k], bounce[k]), f"bounce.{k}"
    for k in ("n_valid", "n_occluded", "hit_hash", "tau_bits"):
        assert np.array_equal(pair_ref[k], pair[k]), f"pair.{k}"
    for k in ("power_te", "power_tm"):
        np.testing.assert_allclose(pair[k], pair_ref[k], rtol=rtol, atol=1e-300, err_msg=k)


# ------------------------------------------------------------ plain C caller

def build_c_caller(out_dir: str) -> str:
    """Compiles tests/c_caller/caller.c (written against include/ under the
    reference's header names) as C11 with -Wall -Wextra -Werror and links it
    against the product library.  Returns the executable's path."""
    exe = os.path.join(out_dir, "caller")
    pkg = os.path.join(ROOT, "hermespy-rt_b200")
    cmd = ["gcc", "-O2", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_caller", "caller.c"), "-L", pkg, "-lhermespy_rt",
           "-Wl,-rpath," + pkg, "-lm", "-o", exe]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return exe


C_CALLER_RX = [[0.0, 0.0, 0.5], [0.4, -0.3, 1.25]]
C_CALLER_TX = [[0.0, 0.0, 0.5]]
C_CALLER_RXV = [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0]]
C_CALLER_TXV = [[0.0, 2.0, 0.0]]
