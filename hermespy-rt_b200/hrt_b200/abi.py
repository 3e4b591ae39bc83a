"""ctypes mirror of the reference's C ABI (structs and the compute_paths call).

Struct layouts follow reference inc/vec3.h:6-8, inc/ray.h:6-9, inc/scene.h:10-66
and inc/compute_paths.h:13-30 (restated in include/hermespy_rt.h).  The helpers
here drive ANY library that exports the reference's `compute_paths` /
`scene_load` symbols -- the product library, or (from tests and the bench's CPU
baseline only) the compiled reference in oracle/_ref.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np


class Vec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class Ray(C.Structure):
    _fields_ = [("o", Vec3), ("d", Vec3)]


class Mesh(C.Structure):
    _fields_ = [
        ("num_vertices", C.c_uint32),
        ("vs", C.POINTER(Vec3)),
        ("num_triangles", C.c_uint32),
        ("is_", C.POINTER(C.c_uint32)),
        ("material_index", C.c_uint32),
        ("velocity", Vec3),
        ("ns", C.POINTER(Vec3)),
    ]


class Scene(C.Structure):
    _fields_ = [("num_meshes", C.c_uint32), ("meshes", C.POINTER(Mesh))]


class Material(C.Structure):
    _fields_ = [
        ("name_sz", C.c_uint32),
        ("name", C.c_char_p),
        ("a", C.c_float), ("b", C.c_float), ("c", C.c_float), ("d", C.c_float),
        ("s", C.c_float),
        ("s1", C.c_float), ("s2", C.c_float), ("s3", C.c_float),
        ("s1_alpha", C.c_uint8), ("s3_alpha", C.c_uint8),
    ]


class ChannelInfo(C.Structure):
    _fields_ = [
        ("num_rays", C.c_uint32),
        ("directions_rx", C.c_void_p),
        ("directions_tx", C.c_void_p),
        ("a_te_re", C.c_void_p), ("a_te_im", C.c_void_p),
        ("a_tm_re", C.c_void_p), ("a_tm_im", C.c_void_p),
        ("tau", C.c_void_p),
        ("freq_shift", C.c_void_p),
    ]


class RaysInfo(C.Structure):
    _fields_ = [
        ("num_bounces", C.c_uint32), ("num_rays", C.c_uint32),
        ("rays", C.c_void_p),
        ("rays_active", C.c_void_p),
    ]


CHAN_FIELDS = ("directions_rx", "directions_tx", "a_te_re", "a_te_im",
               "a_tm_re", "a_tm_im", "tau", "freq_shift")


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def vec3_array(x, n: int | None = None) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(-1, 3))
    if n is not None and a.shape[0] != n:
        raise ValueError(f"expected {n} positions, got {a.shape[0]}")
    return a


@dataclass
class Outputs:
    """Caller-owned output buffers of one compute_paths() call, as numpy."""
    R: int
    T: int
    P: int
    B: int
    los: dict = field(default_factory=dict)     # (R,T[,3])
    scat: dict = field(default_factory=dict)    # (R,T,B,P[,3])
    los_rays: np.ndarray | None = None          # (R*T, 6)
    los_active: np.ndarray | None = None        # bytes
    scat_rays: np.ndarray | None = None         # (T*(B+1)+1, P, 6)
    scat_active: np.ndarray | None = None       # (T*(B+1)+1, P//8+1)


def alloc_outputs(R: int, T: int, P: int, B: int, fill: int = 0) -> Outputs:
    """Allocate every output array the caller must own (sizes of the reference's
    test/test.c:29-60), each byte pre-set to `fill` so that words the callee
    never writes can be recognised by running twice with two fill patterns."""
    def buf(shape, dtype=np.float32):
        a = np.empty(shape, dtype=dtype)
        a.view(np.uint8).reshape(-1)[:] = fill
        return a

    o = Outputs(R, T, P, B)
    for k in CHAN_FIELDS:
        vec = k.startswith("directions")
        o.los[k] = buf((R, T, 3) if vec else (R, T))
        o.scat[k] = buf((R, T, B, P, 3) if vec else (R, T, B, P))
    o.los_rays = buf((R * T, 6))
    o.los_active = buf((R * T // 8 + 1,), np.uint8)
    rows = T * (B + 1) + 1
    o.scat_rays = buf((rows, P, 6))
    o.scat_active = buf((rows, P // 8 + 1), np.uint8)
    return o


def chan_struct(d: dict, num_rays: int) -> ChannelInfo:
    ci = ChannelInfo()
    ci.num_rays = num_rays
    for k in CHAN_FIELDS:
        setattr(ci, k, _ptr(d[k]))
    return ci


def bind_compute_paths(lib: C.CDLL) -> None:
    lib.scene_load.restype = Scene
    lib.scene_load.argtypes = [C.c_char_p]
    lib.compute_paths.restype = None
    lib.compute_paths.argtypes = [
        C.POINTER(Scene), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_float, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t,
        C.POINTER(ChannelInfo), C.POINTER(RaysInfo),
        C.POINTER(ChannelInfo), C.POINTER(RaysInfo),
    ]


def free_scene(scene: Scene) -> None:
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    for m in range(scene.num_meshes):
        me = scene.meshes[m]
        libc.free(C.cast(me.vs, C.c_void_p))
        libc.free(C.cast(me.is_, C.c_void_p))
        libc.free(C.cast(me.ns, C.c_void_p))
    libc.free(C.cast(scene.meshes, C.c_void_p))


def call_compute_paths(lib: C.CDLL, scene: Scene, rx, tx, rx_vel, tx_vel,
                       f_ghz: float, P: int, B: int, fill: int = 0,
                       out: Outputs | None = None) -> Outputs:
    """Run `lib.compute_paths` on an already loaded Scene struct."""
    rx = vec3_array(rx); tx = vec3_array(tx)
    R, T = rx.shape[0], tx.shape[0]
    rxv = vec3_array(rx_vel, R); txv = vec3_array(tx_vel, T)
    o = out if out is not None else alloc_outputs(R, T, P, B, fill)
    los = chan_struct(o.los, 1)
    sc = chan_struct(o.scat, B * P)
    rl = RaysInfo(1, 1, _ptr(o.los_rays), _ptr(o.los_active))
    rs = RaysInfo(B + 1, P, _ptr(o.scat_rays), _ptr(o.scat_active))
    lib.compute_paths(C.byref(scene), _ptr(rx), _ptr(tx), _ptr(rxv), _ptr(txv),
                      C.c_float(f_ghz), R, T, P, B,
                      C.byref(los), C.byref(rl), C.byref(sc), C.byref(rs))
    return o


def scene_to_numpy(scene: Scene):
    """Flatten a loaded Scene into (corners[N,3,3], mesh_of[N], material[M],
    velocity[M,3]) in the reference's (mesh, face) order."""
    tris, mesh_of, mats, vels = [], [], [], []
    for m in range(scene.num_meshes):
        me = scene.meshes[m]
        vs = np.ctypeslib.as_array(C.cast(me.vs, C.POINTER(C.c_float)),
                                   (me.num_vertices, 3)).copy()
        idx = np.ctypeslib.as_array(me.is_, (me.num_triangles * 3,)).copy().reshape(-1, 3)
        tris.append(vs[idx])
        mesh_of.append(np.full(me.num_triangles, m, np.uint32))
        mats.append(me.material_index)
        vels.append([me.velocity.x, me.velocity.y, me.velocity.z])
    return (np.concatenate(tris).astype(np.float32), np.concatenate(mesh_of),
            np.asarray(mats, np.uint32), np.asarray(vels, np.float32))
