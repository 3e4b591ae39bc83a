"""Scene tools around the .hrt format (reference src/scene.c:16-31,44-79; SURVEY
appendix B): numpy reader/writer and the synthetic tiled street canyon of
BASELINE configs[4]."""
from __future__ import annotations

import struct

import numpy as np

MATERIAL = {"air": 0, "concrete": 1, "brick": 2, "glass1": 5, "marble": 11, "metal": 13,
            "medium_dry_ground": 15, "wet_ground": 16}


def read_hrt(path):
    """-> list of dicts(vs (nv,3) f32, tris (nt,3) u32, material int, velocity (3,) f32)"""
    b = open(path, "rb").read()
    if b[:3] != b"HRT":
        raise ValueError("not an HRT file")
    off = 3
    (nm,) = struct.unpack_from("<I", b, off); off += 4
    meshes = []
    for _ in range(nm):
        (nv,) = struct.unpack_from("<I", b, off); off += 4
        vs = np.frombuffer(b, "<f4", nv * 3, off).reshape(nv, 3).copy(); off += nv * 12
        (nt,) = struct.unpack_from("<I", b, off); off += 4
        tris = np.frombuffer(b, "<u4", nt * 3, off).reshape(nt, 3).copy(); off += nt * 12
        (mat,) = struct.unpack_from("<I", b, off); off += 4
        vel = np.frombuffer(b, "<f4", 3, off).copy(); off += 12
        meshes.append(dict(vs=vs, tris=tris, material=int(mat), velocity=vel))
    return meshes


def write_hrt(path, meshes):
    if not 1 <= len(meshes) <= 1000:
        raise ValueError("the loader accepts 1..1000 meshes (reference src/scene.c:52-55)")
    with open(path, "wb") as f:
        f.write(b"HRT" + struct.pack("<I", len(meshes)))
        for m in meshes:
            vs = np.ascontiguousarray(m["vs"], "<f4"); tris = np.ascontiguousarray(m["tris"], "<u4")
            f.write(struct.pack("<I", len(vs))); f.write(vs.tobytes())
            f.write(struct.pack("<I", len(tris))); f.write(tris.tobytes())
            f.write(struct.pack("<I", int(m["material"])))
            f.write(np.asarray(m["velocity"], "<f4").tobytes())


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return x ^ (x >> 31)


def tiled_canyon(base_path, nx=64, ny=64, block=8, seed=0x48525431):
    """nx x ny copies of the street-canyon tile, merged into <= 1000 meshes grouped
    by (block x block tile group, material).  Materials: buildings (12-triangle
    meshes) from {concrete, brick, glass1, marble}, cars (20 triangles) metal,
    ground from {medium_dry_ground, wet_ground, concrete}, drawn per
    (tile group, source mesh) with splitmix64.  Deterministic."""
    base = read_hrt(base_path)
    allv = np.concatenate([m["vs"] for m in base])
    lo, hi = allv.min(0), allv.max(0)
    pitch = np.array([hi[0] - lo[0], hi[1] - lo[1], 0.0], np.float32)
    groups = {}
    for by in range(0, ny, block):
        for bx in range(0, nx, block):
            gid = (by // block) * ((nx + block - 1) // block) + bx // block
            for mi, m in enumerate(base):
                nt = len(m["tris"])
                h = _splitmix64(seed ^ (gid * 64 + mi))
                if nt == 12:
                    mat = MATERIAL[("concrete", "brick", "glass1", "marble")[h % 4]]
                elif nt == 20:
                    mat = MATERIAL["metal"]
                else:
                    mat = MATERIAL[("medium_dry_ground", "wet_ground", "concrete")[h % 3]]
                g = groups.setdefault((gid, mat), dict(vs=[], tris=[], nv=0))
                for ty in range(by, min(by + block, ny)):
                    for tx in range(bx, min(bx + block, nx)):
                        shift = pitch * np.array([tx - nx / 2 + 0.5, ty - ny / 2 + 0.5, 0.0], np.float32)
                        g["vs"].append(m["vs"] + shift.astype(np.float32))
                        g["tris"].append(m["tris"] + np.uint32(g["nv"]))
                        g["nv"] += len(m["vs"])
    meshes = [dict(vs=np.concatenate(g["vs"]).astype(np.float32), tris=np.concatenate(g["tris"]).astype(np.uint32),
                   material=mat, velocity=np.zeros(3, np.float32)) for (gid, mat), g in sorted(groups.items())]
    return meshes, pitch[:2]


def c5_positions(pitch, nx=64, ny=64, n_tx=16, n_rx=1024):
    """16 TX on a 4x4 grid at z=10 and 1024 RX on a 32x32 grid at z=1.5 over the
    central quarter of the tiled footprint."""
    ext = pitch * np.array([nx, ny]) * 0.25
    def grid(k, z, dy):
        s = int(round(np.sqrt(k)))
        xs = (np.arange(s) + 0.5) / s * 2 - 1
        # y snapped to the nearest street centre line (tile centres sit at (k+1/2)*pitch)
        snap = lambda y: (np.floor(y / pitch[1]) + 0.5) * pitch[1] + dy
        return np.array([[x * ext[0], snap(y * ext[1]), z] for y in xs for x in xs], np.float32)[:k]
    return grid(n_rx, 1.5, 1.0), grid(n_tx, 10.0, 0.0)
