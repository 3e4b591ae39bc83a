"""CPU-side checks of the boundary: the C-ABI library loads, exports every
symbol include/*.h declares, and refuses to compute without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import hrt_testlib as tl
import hrt_b200 as hrt
from hrt_b200 import abi

ROOT = tl.ROOT


def _declared_symbols():
    syms = set()
    for h in ("hermespy_rt.h", "hrt_cuda.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"^\s*(?:extern\s+)?[A-Za-z_][\w\s\*]*?\b(\w+)\s*\(", src, flags=re.M):
            name = m.group(1)
            if name.startswith(("vec3_", "free_")) or name in ("defined", "sizeof"):
                continue            # static inline helpers
            syms.add(name)
    syms.add("g_materials")
    return syms


def test_library_exports_every_declared_symbol():
    if not os.path.exists(hrt.LIB_PATH):
        subprocess.run(["make", "-C", os.path.join(ROOT, "hermespy-rt_b200")], check=True,
                       capture_output=True)
    L = hrt.lib()
    want = _declared_symbols()
    assert {"compute_paths", "compute_cir", "compute_path_list", "scene_load", "scene_save", "get_material_index", "hrt_run", "hrt_scene_advance",
            "hrt_scene_upload", "hrt_ctx_create", "hrt_closest_hits"} <= want
    for s in sorted(want):
        assert hasattr(L, s), f"libhermespy_rt.so does not export {s}"


def test_no_cpu_fallback():
    """Without a device the product must fail loudly, not compute on the CPU."""
    L = hrt.lib()
    if L.hrt_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(hrt.HrtError, match="no CUDA device|no CPU"):
        hrt.Context(0)
    with pytest.raises(hrt.HrtError):
        hrt.compute_paths(tl.scene_path("box"), [[0, 0, 1]], [[0, 0, 2]], [[0, 0, 0]], [[0, 0, 0]],
                          3.0, 1, 1, 100, 1)
    # the C entry itself: message + exit(8) (reference convention, inc/common.h:20-25)
    code = ("import sys; sys.path.insert(0, %r); import hrt_b200 as h; from hrt_b200 import abi;"
            "L = h.lib(); sc = L.scene_load(%r);"
            "abi.call_compute_paths(L, sc, [[0,0,1]], [[0,0,2]], [[0,0,0]], [[0,0,0]], 3.0, 64, 1)"
            % (os.path.join(ROOT, "hermespy-rt_b200"), tl.scene_path("box").encode()))
    p = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True)
    assert p.returncode == 8
    assert "no CUDA device" in p.stderr or "CUDA" in p.stderr


def test_product_does_not_link_oracle():
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.abspath(hrt.LIB_PATH)],
                         capture_output=True, text=True).stdout
    assert "oracle_" not in out and "orc_" not in out and "emul_" not in out


def test_scene_io_round_trip(tmp_path):
    """scene_load / scene_save of the product == the .hrt bytes on disk."""
    L = hrt.lib()
    L.scene_save.argtypes = [C.POINTER(abi.Scene), C.c_char_p]
    for name in ("simple_reflector", "box", "2cars", "simple_street_canyon_with_cars"):
        src = tl.scene_path(name)
        sc = L.scene_load(src.encode())
        dst = str(tmp_path / (name + ".hrt"))
        L.scene_save(C.byref(sc), dst.encode())
        abi.free_scene(sc)
        assert open(src, "rb").read() == open(dst, "rb").read()


def test_scene_load_rejects_garbage(tmp_path):
    bad = tmp_path / "bad.hrt"
    bad.write_bytes(b"HRX" + b"\0" * 16)
    code = ("import sys; sys.path.insert(0, %r); import hrt_b200 as h; h.lib().scene_load(%r)"
            % (os.path.join(ROOT, "hermespy-rt_b200"), str(bad).encode()))
    p = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True)
    assert p.returncode == 8            # reference convention: exit(8), src/scene.c:46-47


def test_material_derivation_matches_oracle():
    """hrt_materials_derive (host C, glibc powf) == the oracle's orc_material."""
    L = hrt.lib()
    O = tl.oracle_lib()

    class OrcMat(C.Structure):
        _fields_ = [(n, C.c_float) for n in
                    ("eta_re", "eta_im", "eta_abs", "eta_abs2", "eta_abs_inv_sqrt", "sqrt_re",
                     "sqrt_im", "inv_re", "inv_im", "inv_sqrt_re", "inv_sqrt_im", "r")]
    O.orc_material.argtypes = [C.c_uint32, C.c_float, C.POINTER(OrcMat)]
    O.orc_material.restype = None
    for f in (0.7, 3.0, 3.5, 28.0, 70.0):
        for i in range(17):
            d = hrt.MaterialDerived(); o = OrcMat()
            L.hrt_materials_derive(i, C.c_float(f), C.byref(d))
            O.orc_material(i, C.c_float(f), C.byref(o))
            for k in ("eta_abs2", "eta_abs_inv_sqrt", "sqrt_re", "sqrt_im", "inv_re", "inv_im", "r"):
                a, b = getattr(d, k), getattr(o, k)
                assert (a == b) or (np.isnan(a) and np.isnan(b)), (f, i, k, a, b)


def test_python_module_imports_and_refuses_without_gpu():
    """`import hermespy_rt` -- the module name the reference's test/test.py:5
    imports -- works (the reference's own build does not import on Linux)."""
    import importlib
    import sys
    sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200"))
    m = importlib.import_module("hermespy_rt")
    assert hasattr(m, "compute_paths") and hasattr(m, "ChannelInfo")
    if hrt.lib().hrt_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        m.compute_paths(tl.scene_path("box"), np.zeros((1, 3)), np.ones((1, 3)), np.zeros((1, 3)),
                        np.zeros((1, 3)), 3.0, 1, 1, 10, 1)
    with pytest.raises(ValueError):
        m.compute_paths(tl.scene_path("box"), np.zeros((2, 3)), np.ones((1, 3)), np.zeros((1, 3)),
                        np.zeros((1, 3)), 3.0, 1, 1, 10, 1)


def test_plain_c_caller_compiles_links_and_refuses_without_gpu(tmp_path):
    """The drop-in boundary at source level: a C11 program written against the
    reference's header names (compute_paths.h, scene.h, vec3.h, ray.h) compiles
    warning-free against include/ and links against libhermespy_rt.so.  Without a
    GPU it ends like the reference's I/O errors do: message + exit(8)."""
    import subprocess
    import torch
    exe = tl.build_c_caller(str(tmp_path))
    if torch.cuda.is_available():
        return
    p = subprocess.run([exe, tl.scene_path("simple_reflector"), "100", "2", "3.0"], capture_output=True, text=True)
    assert p.returncode == 8 and "hermespy_rt" in p.stderr
