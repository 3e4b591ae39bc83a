"""Sionna / Mitsuba scene -> .hrt converter (the job of the reference's offline
tool src/scene_fromSionna.c:103-470), with the fixes the
survey lists:

* materials are mapped by NAME from the XML (`mat-itu_concrete` -> "concrete");
  Sionna's unnumbered `itu_glass` / `itu_ceiling_board` map to the first of the
  two table rows; an unknown name is an error instead of silently becoming
  material 0 ("air" -- which is what happened to every mesh of the bundled
  street canyon);
* the override CSV (`<scene>.csv`: name,material_index,velocity_x,velocity_y,
  velocity_z) is parsed with all five fields (the reference compares sscanf's
  result with 4 and rejects every valid line, src/scene_fromSionna.c:226-236);
* PLY headers are parsed property by property (any extra vertex properties,
  ascii or binary_little_endian, polygons fan-triangulated) instead of assuming
  x,y,z,s,t + uchar/int faces (src/scene_fromSionna.c:139-160);
* the XML is parsed as XML.

usage: python -m hrt_b200.sionna scene.xml out.hrt
"""
from __future__ import annotations

import csv
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

from .scenes import write_hrt

# lookup keys of get_material_index (reference src/materials.c:98-116)
MATERIAL_INDEX = {n: i for i, n in enumerate(
    ["air", "concrete", "brick", "plasterboard", "wood", "glass1", "glass2", "ceiling_board1",
     "ceiling_board2", "chipboard", "plywood", "marble", "floorboard", "metal", "very_dry_ground",
     "medium_dry_ground", "wet_ground"])}
_ALIASES = {"glass": "glass1", "ceiling_board": "ceiling_board1"}

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2",
              "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
              "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def material_index(name: str) -> int:
    """'mat-itu_concrete' / 'itu_concrete' / 'concrete' -> index into g_materials"""
    n = name.strip().lower()
    for prefix in ("mat-", "itu_"):
        if n.startswith(prefix):
            n = n[len(prefix):]
    n = _ALIASES.get(n, n)
    if n not in MATERIAL_INDEX:
        raise ValueError(f"unknown radio material {name!r} (known: {', '.join(MATERIAL_INDEX)})")
    return MATERIAL_INDEX[n]


def read_ply(path: str):
    """-> (vertices (nv,3) float32, triangles (nt,3) uint32)"""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements = None, []
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: no end_header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] == "comment":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append({"name": tok[1], "count": int(tok[2]), "props": []})
            elif tok[0] == "property":
                if tok[1] == "list":
                    elements[-1]["props"].append(("list", _PLY_TYPES[tok[2]], _PLY_TYPES[tok[3]], tok[4]))
                else:
                    elements[-1]["props"].append(("scalar", _PLY_TYPES[tok[1]], None, tok[2]))
            elif tok[0] == "end_header":
                break
        if fmt not in ("binary_little_endian", "ascii"):
            raise ValueError(f"{path}: unsupported PLY format {fmt}")
        vs, faces = None, []
        ascii_tokens = f.read().split() if fmt == "ascii" else None
        apos = 0
        for el in elements:
            scalar_only = all(p[0] == "scalar" for p in el["props"])
            if fmt == "binary_little_endian" and scalar_only:
                dt = np.dtype([(p[3], "<" + p[1]) for p in el["props"]])
                rec = np.frombuffer(f.read(dt.itemsize * el["count"]), dt, el["count"])
                if el["name"] == "vertex":
                    vs = np.stack([rec["x"], rec["y"], rec["z"]], 1).astype(np.float32)
                continue
            for _ in range(el["count"]):
                row = {}
                for kind, t0, t1, name in el["props"]:
                    if fmt == "ascii":
                        if kind == "scalar":
                            row[name] = float(ascii_tokens[apos]); apos += 1
                        else:
                            k = int(ascii_tokens[apos]); apos += 1
                            row[name] = [int(x) for x in ascii_tokens[apos:apos + k]]; apos += k
                    else:
                        if kind == "scalar":
                            row[name] = np.frombuffer(f.read(np.dtype(t0).itemsize), "<" + t0)[0]
                        else:
                            k = int(np.frombuffer(f.read(np.dtype(t0).itemsize), "<" + t0)[0])
                            row[name] = np.frombuffer(f.read(np.dtype(t1).itemsize * k), "<" + t1).tolist()
                if el["name"] == "vertex":
                    faces_v = (row["x"], row["y"], row["z"])
                    vs = np.asarray([faces_v], np.float32) if vs is None else np.vstack([vs, np.asarray([faces_v], np.float32)])
                elif el["name"] == "face":
                    idx = next(v for k, v in row.items() if isinstance(v, list))
                    for k in range(1, len(idx) - 1):                 # fan triangulation
                        faces.append((idx[0], idx[k], idx[k + 1]))
    if vs is None or not faces:
        raise ValueError(f"{path}: PLY has no vertices or no faces")
    tris = np.asarray(faces, np.uint32)
    if tris.max() >= len(vs):
        raise ValueError(f"{path}: face index out of range")
    return vs, tris


def read_sionna_xml(path: str):
    """-> list of (shape name, mesh file path, material name) for every PLY shape"""
    root = ET.parse(path).getroot()
    base = os.path.dirname(os.path.abspath(path))
    out = []
    for shape in root.iter("shape"):
        if shape.get("type", "ply") != "ply":
            continue
        name = shape.get("name") or shape.get("id") or f"shape{len(out)}"
        fn = next((s.get("value") for s in shape.findall("string") if s.get("name") == "filename"), None)
        ref = next((r.get("id") for r in shape.findall("ref")), None)
        if ref is None:
            b = shape.find("bsdf")
            ref = b.get("id") if b is not None else None
        if fn is None or ref is None:
            raise ValueError(f"{path}: shape {name!r} has no mesh file or no material reference")
        out.append((name, os.path.join(base, fn), ref))
    if not out:
        raise ValueError(f"{path}: no <shape> elements")
    return out


def read_overrides(path: str):
    """<scene>.csv: name,material_index,velocity_x,velocity_y,velocity_z -> {name: (index, velocity)}"""
    if not os.path.exists(path):
        return {}
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    if not rows or [c.strip() for c in rows[0]] != ["name", "material_index", "velocity_x", "velocity_y", "velocity_z"]:
        raise ValueError(f"{path}: header must be name,material_index,velocity_x,velocity_y,velocity_z")
    out = {}
    for r in rows[1:]:
        if not r:
            continue
        if len(r) != 5:
            raise ValueError(f"{path}: expected 5 fields, got {r}")
        idx = int(r[1])
        if not 0 <= idx < len(MATERIAL_INDEX):
            raise ValueError(f"{path}: material index {idx} out of range")
        out[r[0].strip()] = (idx, np.asarray([float(r[2]), float(r[3]), float(r[4])], np.float32))
    return out


def from_sionna(xml_path: str):
    """-> meshes as scenes.write_hrt takes them"""
    over = read_overrides(os.path.splitext(xml_path)[0] + ".csv")
    meshes = []
    for name, ply, mat in read_sionna_xml(xml_path):
        vs, tris = read_ply(ply)
        idx, vel = material_index(mat), np.zeros(3, np.float32)
        if name in over:
            idx, vel = over[name]
        meshes.append(dict(name=name, vs=vs, tris=tris, material=idx, velocity=vel))
    return meshes


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 2:
        print(__doc__.strip().splitlines()[-1], file=sys.stderr)
        return 2
    meshes = from_sionna(argv[0])
    write_hrt(argv[1], meshes)
    print(f"{argv[1]}: {len(meshes)} meshes, {sum(len(m['tris']) for m in meshes)} triangles, materials "
          f"{sorted({m['material'] for m in meshes})}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
