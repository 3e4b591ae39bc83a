"""Scene tools (CPU): .hrt numpy reader/writer against the C loader, and the
synthetic tiled canyon of BASELINE configs[4]."""
import ctypes as C

import numpy as np

import hrt_testlib as tl
import hrt_b200 as hrt
from hrt_b200 import abi, scenes


def test_numpy_reader_matches_c_loader(tmp_path):
    for name in ("box", "2cars", "simple_street_canyon_with_cars"):
        meshes = scenes.read_hrt(tl.scene_path(name))
        L = hrt.lib()
        sc = L.scene_load(tl.scene_path(name).encode())
        tris, mesh_of, mats, vels = abi.scene_to_numpy(sc)
        abi.free_scene(sc)
        ours = np.concatenate([m["vs"][m["tris"]] for m in meshes])
        assert np.array_equal(ours, tris)
        assert [m["material"] for m in meshes] == mats.tolist()
        out = str(tmp_path / (name + ".hrt"))
        scenes.write_hrt(out, meshes)
        assert open(out, "rb").read() == open(tl.scene_path(name), "rb").read()


def test_tiled_canyon_shape_and_determinism(tmp_path):
    meshes, pitch = scenes.tiled_canyon(tl.scene_path("simple_street_canyon_with_cars"), 8, 8, block=4)
    assert sum(len(m["tris"]) for m in meshes) == 64 * 234
    assert 1 <= len(meshes) <= 1000
    assert {m["material"] for m in meshes} <= set(scenes.MATERIAL.values())
    assert any(m["material"] == scenes.MATERIAL["metal"] for m in meshes)
    again, _ = scenes.tiled_canyon(tl.scene_path("simple_street_canyon_with_cars"), 8, 8, block=4)
    assert all(np.array_equal(a["vs"], b["vs"]) and a["material"] == b["material"] for a, b in zip(meshes, again))
    p = str(tmp_path / "t.hrt")
    scenes.write_hrt(p, meshes)
    sc = hrt.lib().scene_load(p.encode())       # passes the C loader's validation
    assert sc.num_meshes == len(meshes)
    abi.free_scene(sc)
    # full-size C5: 64 x 64 tiles -> 958,464 triangles in <= 1000 meshes (loader limit)
    nx = ny = 64
    n_groups = (nx // 8) * (ny // 8)
    assert n_groups * 8 <= 1000


def _write_ply(path, vs, faces, binary=True):
    """a PLY as Sionna/Blender exports it: x y z + texture s t, uchar/int face lists"""
    vs = np.asarray(vs, np.float32)
    hdr = ["ply", "format %s 1.0" % ("binary_little_endian" if binary else "ascii"), "comment test",
           f"element vertex {len(vs)}", "property float x", "property float y", "property float z",
           "property float s", "property float t", f"element face {len(faces)}",
           "property list uchar int vertex_index", "end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode())
        if binary:
            for v in vs:
                f.write(np.asarray([v[0], v[1], v[2], 0.25, 0.75], "<f4").tobytes())
            for fc in faces:
                f.write(np.uint8(len(fc)).tobytes() + np.asarray(fc, "<i4").tobytes())
        else:
            for v in vs:
                f.write(("%r %r %r 0.25 0.75\n" % (float(v[0]), float(v[1]), float(v[2]))).encode())
            for fc in faces:
                f.write((" ".join(str(x) for x in [len(fc), *fc]) + "\n").encode())


def test_sionna_converter(tmp_path):
    """Sionna XML + PLY meshes + override CSV -> .hrt (reference tool
    src/scene_fromSionna.c with the survey's fixes): materials by name, CSV with
    five fields, generic PLY headers, quads triangulated; the result passes the
    C loader."""
    from hrt_b200 import sionna
    os_mesh = tmp_path / "meshes"; os_mesh.mkdir()
    ground = ([[-50, -50, 0], [50, -50, 0], [50, 50, 0], [-50, 50, 0]], [[0, 1, 2, 3]])             # one quad
    wall = ([[0, 8, 0], [30, 8, 0], [30, 8, 20], [0, 8, 20]], [[0, 1, 2], [0, 2, 3]])
    car = ([[1, 1, 0], [3, 1, 0], [3, 2, 0], [1, 2, 1.5]], [[0, 1, 2], [0, 2, 3], [0, 3, 1], [1, 3, 2]])
    _write_ply(os_mesh / "ground.ply", *ground)
    _write_ply(os_mesh / "wall.ply", *wall, binary=False)
    _write_ply(os_mesh / "car.ply", *car)
    xml = """<scene version="2.1.0">
  <bsdf type="twosided" id="mat-itu_concrete"><bsdf type="diffuse"/></bsdf>
  <bsdf type="twosided" id="mat-itu_glass"><bsdf type="diffuse"/></bsdf>
  <bsdf type="twosided" id="mat-itu_metal"><bsdf type="diffuse"/></bsdf>
  <shape type="ply" id="mesh-ground" name="ground"><string name="filename" value="meshes/ground.ply"/>
    <boolean name="face_normals" value="true"/><ref id="mat-itu_concrete" name="bsdf"/></shape>
  <shape type="ply" id="mesh-wall" name="wall"><string name="filename" value="meshes/wall.ply"/>
    <ref id="mat-itu_glass" name="bsdf"/></shape>
  <shape type="ply" id="mesh-car" name="car"><string name="filename" value="meshes/car.ply"/>
    <ref id="mat-itu_metal" name="bsdf"/></shape>
</scene>"""
    (tmp_path / "town.xml").write_text(xml)
    (tmp_path / "town.csv").write_text("name,material_index,velocity_x,velocity_y,velocity_z\ncar,13,14.5,0,-0.25\n")
    out = str(tmp_path / "town.hrt")
    assert sionna.main([str(tmp_path / "town.xml"), out]) == 0
    meshes = scenes.read_hrt(out)
    assert [m["material"] for m in meshes] == [1, 5, 13]                  # concrete, glass1, metal -- none of them "air"
    assert [len(m["tris"]) for m in meshes] == [2, 2, 4]
    assert np.array_equal(meshes[0]["vs"][meshes[0]["tris"]][1], np.asarray(ground[0], np.float32)[[0, 2, 3]])
    assert np.array_equal(meshes[1]["vs"], np.asarray(wall[0], np.float32))
    assert np.array_equal(meshes[2]["velocity"], np.asarray([14.5, 0, -0.25], np.float32))
    assert np.array_equal(meshes[0]["velocity"], np.zeros(3, np.float32))
    sc = hrt.lib().scene_load(out.encode())
    tris, mesh_of, mats, vels = abi.scene_to_numpy(sc)
    abi.free_scene(sc)
    assert len(tris) == 8 and mats.tolist() == [1, 5, 13]
    # the lookup keys are the C library's own
    L = hrt.lib()
    for name, idx in sionna.MATERIAL_INDEX.items():
        assert L.get_material_index(name.encode()) == idx
    import pytest
    with pytest.raises(ValueError):
        sionna.material_index("mat-itu_unobtainium")
