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
