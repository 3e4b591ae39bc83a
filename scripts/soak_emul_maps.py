"""CPU soak of the receiver-map query (tests/emul: the kernels' own element functions, serial): primary hit points
of isotropic rays from the C4 transmitters as origins (with the reference's 1e-4 offset and the triangle they lie
on, as k_scatter has it), all 64 C4 receivers, G = 256: every query through the maps -- depth bounds, "sure" cells,
own-triangle early-out -- against the brute-force loop over every triangle.
usage: python scripts/soak_emul_maps.py [scene] [n_primary_rays] [G]"""
import sys, os, ctypes as C, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200"))
import hrt_testlib as tl
from hrt_b200 import abi
os.environ["EMUL_RXMAP_VERBOSE"] = "1"
scene = sys.argv[1] if len(sys.argv) > 1 else "simple_street_canyon_with_cars"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
G = int(sys.argv[3]) if len(sys.argv) > 3 else 256
lib = tl.emul_lib()
lib.emul_rxmap_vs_brute.restype = C.c_long
lib.emul_rxmap_vs_brute.argtypes = [C.POINTER(abi.Scene), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint32,
                                    C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p]
grid, txs = tl.canyon_c4_positions()
rng = np.random.default_rng(5)
d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
o = np.asarray(txs, np.float32)[rng.integers(0, len(txs), n)]
rays = np.ascontiguousarray(np.concatenate([o, d], 1).astype(np.float32))
sc = lib.scene_load(tl.scene_path(scene).encode())
tri = np.zeros(n, np.uint32); t = np.zeros(n, np.float32); th = np.zeros(n, np.float32)
lib.emul_closest_hits(C.byref(sc), rays.ctypes.data, n, 2, C.c_float(64.0), 0, tri.ctypes.data, t.ctypes.data, th.ctypes.data)
ok = tri != 0xFFFFFFFF
T = tl.scene_triangles(scene)
tt = T[tri[ok]]
nrm = np.cross(tt[:, 1] - tt[:, 0], tt[:, 2] - tt[:, 0]); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
hit = rays[ok, :3] + t[ok, None] * rays[ok, 3:]
sgn = -np.sign((nrm * rays[ok, 3:]).sum(1))[:, None]
origins = np.ascontiguousarray((hit + nrm * sgn * 1e-4).astype(np.float32))
own = np.ascontiguousarray(tri[ok].astype(np.uint32))
rx = np.ascontiguousarray(np.asarray(grid, np.float32))
avg, lst = C.c_double(0), C.c_double(0)
t0 = time.time()
bad = lib.emul_rxmap_vs_brute(C.byref(sc), rx.ctypes.data, len(rx), origins.ctypes.data, len(origins), G,
                              C.byref(avg), C.byref(lst), own.ctypes.data)
print(f"{scene}: G={G}, {len(origins)} origins x {len(rx)} receivers = {len(origins) * len(rx)} queries: {bad} differ from brute force; "
      f"{avg.value:.2f} full triangle tests per query, {lst.value:.2f} items per cell ({time.time() - t0:.0f} s)")
sys.exit(1 if bad else 0)
