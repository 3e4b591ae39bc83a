"""Python host-side mirror of the reference interface for the compute_paths()
hot path, over the C-ABI library libhermespy_rt.so (ctypes; no torch types).

Two levels:

* ``compute_paths(mesh_filepath, rx_positions, ...) -> (los, scatter)`` -- same
  name, argument order/meaning and result attributes as the reference's
  pybind11 module (compute_paths_pybind11.cpp:99-210, test/test.py:20-87); it
  goes through the drop-in C entry ``compute_paths`` (include/hermespy_rt.h).
* ``Context`` -- the thin C-ABI of include/hrt_cuda.h (scene upload once, many
  runs, summary/streaming mode, sharding), used by bench.py and the tests.

There is no CPU fallback: if the CUDA library is missing or no device is
usable, importing works but any call raises ``HrtError`` immediately.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "..", "libhermespy_rt.so")

FLAG_DENSE, FLAG_RAYSINFO, FLAG_SUMMARY, FLAG_TRACE = 0x01, 0x02, 0x04, 0x08
FLAG_BRUTE_FORCE, FLAG_HOST_DIRS, FLAG_SUMMARY_DEV, FLAG_COUNT = 0x10, 0x20, 0x40, 0x80
FLAG_CIR, FLAG_PATHLIST, FLAG_PATHLIST_DEV = 0x100, 0x200, 0x400
FLAG_DENSE_C64, FLAG_EXT_LOBES, FLAG_EXT_REFRACT = 0x800, 0x1000, 0x2000
REFRACT_DTYPE = np.dtype([("path", "<u4"), ("tx", "<u2"), ("bounce", "<u2"), ("o", "<f4", (3,)), ("d", "<f4", (3,)),
                          ("t_te_re", "<f4"), ("t_te_im", "<f4"), ("t_tm_re", "<f4"), ("t_tm_im", "<f4")])
assert REFRACT_DTYPE.itemsize == 48
PATH_DTYPE = np.dtype([("path", "<u4"), ("rx", "<u4"), ("tx", "<u2"), ("bounce", "<u2"),
                       ("a_te_re", "<f4"), ("a_te_im", "<f4"), ("a_tm_re", "<f4"), ("a_tm_im", "<f4"),
                       ("tau", "<f4"), ("freq_shift", "<f4"), ("direction_rx", "<f4", (3,))])
assert PATH_DTYPE.itemsize == 48

PAIR_DTYPE = np.dtype([("n_valid", "<u8"), ("n_occluded", "<u8"), ("hit_hash", "<u8"),
                       ("tau_bits", "<u8"), ("power_te", "<f8"), ("power_tm", "<f8")])
BOUNCE_DTYPE = np.dtype([("n_traced", "<u8"), ("n_hit", "<u8"), ("hit_hash", "<u8"),
                         ("t_bits", "<u8")])


class HrtError(RuntimeError):
    pass


class MaterialDerived(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("eta_abs2", "eta_abs_inv_sqrt", "sqrt_re", "sqrt_im",
                                         "inv_re", "inv_im", "r", "s", "s1_alpha",
                                         "pad0", "pad1", "pad2")]


class RunParams(C.Structure):
    _fields_ = [
        ("num_rx", C.c_size_t), ("num_tx", C.c_size_t), ("num_paths", C.c_size_t),
        ("num_bounces", C.c_size_t), ("carrier_frequency_GHz", C.c_float),
        ("rx_pos", C.c_void_p), ("tx_pos", C.c_void_p), ("rx_vel", C.c_void_p), ("tx_vel", C.c_void_p),
        ("shard_rank", C.c_uint32), ("shard_world", C.c_uint32), ("shard_block", C.c_size_t),
        ("flags", C.c_uint32),
        ("los", C.POINTER(abi.ChannelInfo)), ("rays_los", C.POINTER(abi.RaysInfo)),
        ("scat", C.POINTER(abi.ChannelInfo)), ("rays_scat", C.POINTER(abi.RaysInfo)),
        ("pair_summary", C.c_void_p), ("bounce_summary", C.c_void_p),
        ("trace_hit_tri", C.c_void_p), ("trace_hit_t", C.c_void_p), ("trace_slot_state", C.c_void_p),
        ("dirs", C.c_void_p), ("stream", C.c_void_p),
        ("cir", C.c_void_p), ("cir_tau0_s", C.c_float), ("cir_dt_s", C.c_float), ("cir_bins", C.c_uint32),
        ("paths", C.c_void_p), ("paths_capacity", C.c_uint64), ("paths_count", C.POINTER(C.c_uint64)),
        ("scat_a_te_c64", C.c_void_p), ("scat_a_tm_c64", C.c_void_p),
        ("refr_rays", C.c_void_p), ("refr_capacity", C.c_uint64), ("refr_count", C.POINTER(C.c_uint64)),
    ]


class RunStats(C.Structure):
    _fields_ = [
        ("ray_bounces", C.c_uint64), ("primary_hits", C.c_uint64), ("shadow_queries", C.c_uint64),
        ("los_queries", C.c_uint64), ("ambiguous_dirs", C.c_uint64), ("kernel_launches", C.c_uint64),
        ("ms_total", C.c_float), ("ms_bounce", C.c_float), ("ms_scatter", C.c_float),
        ("ms_other", C.c_float), ("n_bounce_launches", C.c_uint32), ("n_scatter_launches", C.c_uint32),
        ("num_tris", C.c_uint32), ("num_nodes", C.c_uint32), ("scene_in_smem", C.c_uint32),
        ("box_pad", C.c_float),
        ("work_bounce", C.c_uint64 * 5), ("work_scatter", C.c_uint64 * 5),
        ("bvh_sah", C.c_uint32), ("bvh_levels", C.c_uint32), ("bvh_build_ms", C.c_float), ("ms_sort", C.c_float),
        ("cir_dropped", C.c_uint64),
        ("rx_map", C.c_uint32), ("rx_map_cells", C.c_uint32), ("rx_map_build_ms", C.c_float),
        ("host_ms_setup", C.c_float), ("host_ms_total", C.c_float),
    ]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_}
        d["work_bounce"] = list(self.work_bounce)
        d["work_scatter"] = list(self.work_scatter)
        return d


_lib = None


def lib() -> C.CDLL:
    """The product library.  Raises HrtError when it has not been built."""
    global _lib
    if _lib is None:
        path = os.path.abspath(os.environ.get("HRT_LIB", LIB_PATH))   # HRT_LIB: tuning builds
        if not os.path.exists(path):
            raise HrtError(f"{path} not found: build it with __graft_entry__.build() "
                           "(make -C hermespy-rt_b200); there is no CPU fallback")
        L = C.CDLL(path)
        abi.bind_compute_paths(L)
        L.hrt_device_count.restype = C.c_int
        L.hrt_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.hrt_ctx_destroy.argtypes = [C.c_void_p]
        L.hrt_last_error.restype = C.c_char_p
        L.hrt_last_error.argtypes = [C.c_void_p]
        L.hrt_scene_upload.argtypes = [C.c_void_p, C.POINTER(abi.Scene), C.c_void_p]
        L.hrt_materials_derive.argtypes = [C.c_uint32, C.c_float, C.POINTER(MaterialDerived)]
        L.hrt_materials_derive.restype = None
        L.hrt_materials_set.argtypes = [C.c_void_p, C.POINTER(MaterialDerived)]
        L.hrt_run.argtypes = [C.c_void_p, C.POINTER(RunParams)]
        L.hrt_get_stats.argtypes = [C.c_void_p, C.POINTER(RunStats)]
        L.hrt_scene_advance.argtypes = [C.c_void_p, C.c_float, C.c_int]
        L.hrt_closest_hits.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
        L.hrt_fp32_peak.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.hrt_multi_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
        L.hrt_multi_destroy.argtypes = [C.c_void_p]
        L.hrt_multi_last_error.restype = C.c_char_p
        L.hrt_multi_last_error.argtypes = [C.c_void_p]
        L.hrt_multi_num_devices.argtypes = [C.c_void_p]
        L.hrt_multi_ctx.restype = C.c_void_p
        L.hrt_multi_ctx.argtypes = [C.c_void_p, C.c_int]
        L.hrt_multi_scene_upload.argtypes = [C.c_void_p, C.POINTER(abi.Scene), C.c_void_p]
        L.hrt_multi_scene_advance.argtypes = [C.c_void_p, C.c_float, C.c_int]
        L.hrt_multi_materials_set.argtypes = [C.c_void_p, C.POINTER(MaterialDerived)]
        L.hrt_multi_get_stats.argtypes = [C.c_void_p, C.POINTER(RunStats)]
        L.hrt_multi_run.argtypes = [C.c_void_p, C.POINTER(RunParams)]
        L.hrt_multi_run_gathered.argtypes = [C.c_void_p, C.POINTER(RunParams), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                             C.POINTER(C.c_void_p), C.c_uint64, C.POINTER(C.c_uint64)]
        L.hrt_multi_nccl_version.argtypes = [C.c_void_p]
        L.hrt_shard_count.restype = C.c_uint64
        L.hrt_shard_count.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64]
        L.hrt_shard_path.restype = C.c_uint64
        L.hrt_shard_path.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64]
        _lib = L
    return _lib


def gather_path_lists(local: "torch.Tensor", count: int, capacity: int):
    """All-gather of per-rank path lists (torch.distributed, any backend): `local`
    is this rank's record buffer as a uint8 tensor of capacity * 48 bytes (device
    memory for NCCL), `count` its number of valid records.  Returns (records of
    all ranks as one uint8 tensor [total, 48], per-rank counts) -- every rank gets
    the whole list.  One all_gather of the counts, one of the padded buffers."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local.view(-1, 48)[:count], [count]
    cnt = torch.tensor([count], dtype=torch.int64, device=local.device)
    counts = torch.zeros(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(counts, cnt)
    allbuf = torch.empty(world * capacity * 48, dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(allbuf, local.view(-1)[: capacity * 48].contiguous())
    counts = [int(c) for c in counts.tolist()]
    parts = [allbuf.view(world, capacity, 48)[r, : min(counts[r], capacity)] for r in range(world)]
    return torch.cat(parts, 0), counts


def shard_paths(num_paths: int, rank: int, world: int, block: int) -> np.ndarray:
    """Global path indices owned by `rank` (host arithmetic of the C library)."""
    L = lib()
    n = L.hrt_shard_count(num_paths, rank, world, block)
    if world <= 1:
        return np.arange(num_paths, dtype=np.uint64)
    loc = np.arange(n, dtype=np.uint64)
    q, r = loc // np.uint64(block), loc % np.uint64(block)
    g = (q * np.uint64(world) + np.uint64(rank)) * np.uint64(block) + r
    # spot-check the closed form against the C function
    for i in (0, n // 2, n - 1):
        if n:
            assert L.hrt_shard_path(int(i), rank, world, block) == int(g[int(i)])
    return g


def device_count() -> int:
    return lib().hrt_device_count()


class Context:
    """One GPU: a scene (BVH resident in HBM) and any number of runs on it."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        L = lib()
        rc = L.hrt_ctx_create(device, C.byref(self._h))
        if rc != 0:
            raise HrtError(f"hrt_ctx_create({device}) failed ({rc}): "
                           f"{L.hrt_last_error(None).decode()}")
        self._scene = None
        self.device = device

    def close(self):
        if self._h:
            lib().hrt_ctx_destroy(self._h)
            self._h = C.c_void_p()
        if self._scene is not None:
            abi.free_scene(self._scene)
            self._scene = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            raise HrtError(f"{what} failed ({rc}): {lib().hrt_last_error(self._h).decode()}")

    # -- scene ---------------------------------------------------------------
    def load_scene(self, path: str, want_normals: bool = False):
        """scene_load() + hrt_scene_upload(): flatten, normals, BVH on the GPU."""
        L = lib()
        if self._scene is not None:
            abi.free_scene(self._scene)
        if not os.path.exists(path):
            raise HrtError(f"scene file not found: {path}")
        self._scene = L.scene_load(path.encode())
        ntri = sum(self._scene.meshes[m].num_triangles for m in range(self._scene.num_meshes))
        normals = np.zeros((max(ntri, 1), 3), np.float32) if want_normals else None
        self._check(L.hrt_scene_upload(self._h, C.byref(self._scene),
                                       normals.ctypes.data if want_normals else None),
                    "hrt_scene_upload")
        self.num_tris = ntri
        return normals[:ntri] if want_normals else None

    def set_frequency(self, f_ghz: float):
        """Material constants at f_ghz for the materials the scene uses
        (reference precompute_materials, src/compute_paths.c:171-206)."""
        L = lib()
        tab = (MaterialDerived * 17)()
        for m in range(self._scene.num_meshes):
            mi = self._scene.meshes[m].material_index
            L.hrt_materials_derive(mi, C.c_float(f_ghz), C.byref(tab[mi]))
        self._check(L.hrt_materials_set(self._h, tab), "hrt_materials_set")

    def advance(self, dt_s: float, rebuild: bool = False):
        """Move every mesh by velocity * dt_s on the GPU; refit (or rebuild) the BVH."""
        self._check(lib().hrt_scene_advance(self._h, C.c_float(dt_s), int(rebuild)), "hrt_scene_advance")

    # -- runs ----------------------------------------------------------------
    def run(self, rx, tx, rx_vel, tx_vel, f_ghz, P, B, *, dense=False, raysinfo=False,
            summary=False, trace=False, brute_force=False, count_work=False, los=True, dirs=None,
            shard=(0, 1), shard_block=1 << 20, out: abi.Outputs | None = None,
            summary_dev_ptrs=None, stream=None, cir=None, path_list=None, path_list_dev=None,
            ext_lobes=False, refract=None):
        """hrt_run().  Returns a dict with whatever was requested:
        'out' (abi.Outputs, dense), 'pair'/'bounce' (structured arrays, summary),
        'trace' (dict), 'stats'."""
        rx = abi.vec3_array(rx); tx = abi.vec3_array(tx)
        R, T = rx.shape[0], tx.shape[0]
        rxv = abi.vec3_array(rx_vel, R); txv = abi.vec3_array(tx_vel, T)
        self.set_frequency(f_ghz)
        p = RunParams()
        p.num_rx, p.num_tx, p.num_paths, p.num_bounces = R, T, P, B
        p.carrier_frequency_GHz = f_ghz
        p.rx_pos, p.tx_pos, p.rx_vel, p.tx_vel = (a.ctypes.data for a in (rx, tx, rxv, txv))
        p.shard_rank, p.shard_world = shard
        p.shard_block = shard_block
        flags = 0
        res = {}
        keep = [rx, tx, rxv, txv]
        if dense:
            flags |= FLAG_DENSE
            o = out if out is not None else abi.alloc_outputs(R, T, P, B, 0)
            res["out"] = o
            los_s = abi.chan_struct(o.los, 1); sc_s = abi.chan_struct(o.scat, B * P)
            rl = abi.RaysInfo(1, 1, o.los_rays.ctypes.data, o.los_active.ctypes.data)
            rs = abi.RaysInfo(B + 1, P, o.scat_rays.ctypes.data, o.scat_active.ctypes.data)
            keep += [los_s, sc_s, rl, rs]
            p.scat = C.pointer(sc_s)
            if los:
                p.los = C.pointer(los_s); p.rays_los = C.pointer(rl)
            if raysinfo:
                flags |= FLAG_RAYSINFO
                p.rays_scat = C.pointer(rs)
        elif los:
            o = abi.alloc_outputs(R, T, 1, 1, 0)
            res["out"] = o
            los_s = abi.chan_struct(o.los, 1)
            rl = abi.RaysInfo(1, 1, o.los_rays.ctypes.data, o.los_active.ctypes.data)
            keep += [los_s, rl]
            p.los = C.pointer(los_s); p.rays_los = C.pointer(rl)
        if summary:
            flags |= FLAG_SUMMARY
            if summary_dev_ptrs is not None:
                flags |= FLAG_SUMMARY_DEV
                p.pair_summary, p.bounce_summary = summary_dev_ptrs
            else:
                res["pair"] = np.zeros((R, T, B), PAIR_DTYPE)
                res["bounce"] = np.zeros((T, B), BOUNCE_DTYPE)
                p.pair_summary = res["pair"].ctypes.data
                p.bounce_summary = res["bounce"].ctypes.data
        if trace:
            flags |= FLAG_TRACE
            tr = {"hit_tri": np.full((T, B, P), 0xFFFFFFFE, np.uint32),
                  "hit_t": np.full((T, B, P), -1.0, np.float32),
                  "slot_state": np.zeros((R, T, B, P), np.uint8)}
            res["trace"] = tr
            p.trace_hit_tri = tr["hit_tri"].ctypes.data
            p.trace_hit_t = tr["hit_t"].ctypes.data
            p.trace_slot_state = tr["slot_state"].ctypes.data
        if brute_force:
            flags |= FLAG_BRUTE_FORCE
        if count_work:
            flags |= FLAG_COUNT
        if dirs is not None:
            dirs = abi.vec3_array(dirs, P)
            keep.append(dirs)
            flags |= FLAG_HOST_DIRS
            p.dirs = dirs.ctypes.data
        if stream is not None:
            p.stream = stream
        if cir is not None:
            # cir = (tau0_s, dt_s, bins): res["cir"] is (R, T, bins, 4) float32
            tau0, dt, bins = cir
            res["cir"] = np.zeros((R, T, int(bins), 4), np.float32)
            flags |= FLAG_CIR
            p.cir = res["cir"].ctypes.data
            p.cir_tau0_s, p.cir_dt_s, p.cir_bins = float(tau0), float(dt), int(bins)
        n_found = C.c_uint64(0)
        if path_list is not None:
            # path_list = capacity: res["paths"] is a PATH_DTYPE array of the valid scatter paths
            buf = np.zeros(int(path_list), PATH_DTYPE)
            keep.append(buf)
            flags |= FLAG_PATHLIST
            p.paths = buf.ctypes.data; p.paths_capacity = int(path_list); p.paths_count = C.pointer(n_found)
        if path_list_dev is not None:
            # path_list_dev = (device pointer, capacity in records): records stay on the GPU
            flags |= FLAG_PATHLIST | FLAG_PATHLIST_DEV
            p.paths, p.paths_capacity = int(path_list_dev[0]), int(path_list_dev[1])
            p.paths_count = C.pointer(n_found)
        n_refr = C.c_uint64(0)
        if ext_lobes:
            flags |= FLAG_EXT_LOBES
        if refract is not None:
            # refract = capacity: res["refract"] is a REFRACT_DTYPE array of the refraction rays spawned
            rbuf = np.zeros(int(refract), REFRACT_DTYPE)
            keep.append(rbuf)
            flags |= FLAG_EXT_REFRACT
            p.refr_rays = rbuf.ctypes.data; p.refr_capacity = int(refract); p.refr_count = C.pointer(n_refr)
        p.flags = flags
        self._check(self._run_call(p), "hrt_run")
        if refract is not None:
            res["refract_found"] = int(n_refr.value)
            res["refract"] = rbuf[: min(int(n_refr.value), int(refract))]
        if path_list_dev is not None:
            res["paths_found"] = int(n_found.value)
        if path_list is not None:
            res["paths_found"] = int(n_found.value)
            res["paths"] = buf[: min(int(n_found.value), int(path_list))]
        res["stats"] = self.stats()
        del keep
        return res

    def _run_call(self, p):
        return lib().hrt_run(self._h, C.byref(p))

    def stats(self) -> dict:
        s = RunStats()
        self._check(lib().hrt_get_stats(self._h, C.byref(s)), "hrt_get_stats")
        return s.as_dict()

    def fp32_peak(self):
        """(unfused FMUL+FADD, FFMA) sustained Tflop/s of this GPU, measured."""
        a, b = C.c_float(), C.c_float()
        self._check(lib().hrt_fp32_peak(self._h, C.byref(a), C.byref(b)), "hrt_fp32_peak")
        return a.value, b.value

    def closest_hits(self, rays: np.ndarray, brute_force: bool = False):
        """Batch moeller_trumbore(): rays (n,6) float32 -> (tri, t, theta)."""
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        tri = np.zeros(n, np.uint32); t = np.zeros(n, np.float32); th = np.zeros(n, np.float32)
        self._check(lib().hrt_closest_hits(self._h, rays.ctypes.data, n,
                                           FLAG_BRUTE_FORCE if brute_force else 0,
                                           tri.ctypes.data, t.ctypes.data, th.ctypes.data),
                    "hrt_closest_hits")
        return tri, t, th


class MultiContext(Context):
    """Several GPUs of one box behind one call (hrt_multi_*, include/hrt_cuda.h):
    the job is sharded by ray inside the C library, one host thread per device."""

    def __init__(self, devices=None):
        self._h = C.c_void_p()
        L = lib()
        n = 0 if devices is None else len(devices)
        arr = (C.c_int * max(n, 1))(*(devices or [0]))
        rc = L.hrt_multi_create(arr if devices is not None else None, n, C.byref(self._h))
        if rc != 0:
            raise HrtError(f"hrt_multi_create failed ({rc}): {L.hrt_multi_last_error(None).decode()}")
        self._scene = None
        self.num_devices = L.hrt_multi_num_devices(self._h)
        self.device = None

    def close(self):
        if self._h:
            lib().hrt_multi_destroy(self._h)
            self._h = C.c_void_p()
        if self._scene is not None:
            abi.free_scene(self._scene)
            self._scene = None

    def _check(self, rc, what):
        if rc != 0:
            raise HrtError(f"{what} failed ({rc}): {lib().hrt_multi_last_error(self._h).decode()}")

    def load_scene(self, path: str, want_normals: bool = False):
        L = lib()
        if self._scene is not None:
            abi.free_scene(self._scene)
        if not os.path.exists(path):
            raise HrtError(f"scene file not found: {path}")
        self._scene = L.scene_load(path.encode())
        ntri = sum(self._scene.meshes[m].num_triangles for m in range(self._scene.num_meshes))
        normals = np.zeros((max(ntri, 1), 3), np.float32) if want_normals else None
        self._check(L.hrt_multi_scene_upload(self._h, C.byref(self._scene), normals.ctypes.data if want_normals else None),
                    "hrt_multi_scene_upload")
        self.num_tris = ntri
        return normals[:ntri] if want_normals else None

    def set_frequency(self, f_ghz: float):
        L = lib()
        tab = (MaterialDerived * 17)()
        for m in range(self._scene.num_meshes):
            mi = self._scene.meshes[m].material_index
            L.hrt_materials_derive(mi, C.c_float(f_ghz), C.byref(tab[mi]))
        self._check(L.hrt_multi_materials_set(self._h, tab), "hrt_multi_materials_set")

    def advance(self, dt_s: float, rebuild: bool = False):
        self._check(lib().hrt_multi_scene_advance(self._h, C.c_float(dt_s), int(rebuild)), "hrt_multi_scene_advance")

    def _run_call(self, p):
        return lib().hrt_multi_run(self._h, C.byref(p))

    def stats(self) -> dict:
        s = RunStats()
        self._check(lib().hrt_multi_get_stats(self._h, C.byref(s)), "hrt_multi_get_stats")
        return s.as_dict()

    def run_gathered(self, rx, tx, rx_vel, tx_vel, f_ghz, P, B, pair_ptrs, bounce_ptrs, paths_ptrs=None,
                     paths_capacity_each=0, shard_block=1 << 16, los=False):
        """hrt_multi_run_gathered(): device pointers per device (lists of ints); returns per-device path counts."""
        rx = abi.vec3_array(rx); tx = abi.vec3_array(tx)
        R, T = rx.shape[0], tx.shape[0]
        rxv = abi.vec3_array(rx_vel, R); txv = abi.vec3_array(tx_vel, T)
        self.set_frequency(f_ghz)
        p = RunParams()
        p.num_rx, p.num_tx, p.num_paths, p.num_bounces = R, T, P, B
        p.carrier_frequency_GHz = f_ghz
        p.rx_pos, p.tx_pos, p.rx_vel, p.tx_vel = (a.ctypes.data for a in (rx, tx, rxv, txv))
        p.shard_block = shard_block
        n = self.num_devices
        pa = (C.c_void_p * n)(*pair_ptrs); ba = (C.c_void_p * n)(*bounce_ptrs)
        la = (C.c_void_p * n)(*paths_ptrs) if paths_ptrs else None
        counts = (C.c_uint64 * n)()
        self._check(lib().hrt_multi_run_gathered(self._h, C.byref(p), pa, ba, la, int(paths_capacity_each), counts),
                    "hrt_multi_run_gathered")
        return [int(c) for c in counts]

    def nccl_version(self) -> int:
        return lib().hrt_multi_nccl_version(self._h)


# ------------------------------------------------- reference-shaped Python API

class ChannelInfo:
    """Same read-only attributes as the reference's pybind11 ChannelInfo
    (compute_paths_pybind11.cpp:44-97, :189-196)."""

    def __init__(self, d: dict, num_paths: int):
        self.num_paths = num_paths
        R, T = d["tau"].shape[:2]
        self.directions_rx = d["directions_rx"].reshape(R, T, num_paths, 3)
        self.directions_tx = d["directions_tx"].reshape(R, T, num_paths, 3)
        self.a_te = (d["a_te_re"] + 1j * d["a_te_im"]).astype(np.complex64).reshape(R, T, num_paths)
        self.a_tm = (d["a_tm_re"] + 1j * d["a_tm_im"]).astype(np.complex64).reshape(R, T, num_paths)
        self.tau = d["tau"].reshape(R, T, num_paths)
        self.freq_shift = d["freq_shift"].reshape(R, T, num_paths)


def compute_paths(mesh_filepath, rx_positions, tx_positions, rx_velocities, tx_velocities,
                  carrier_frequency, num_rx, num_tx, num_paths, num_bounces):
    """Drop-in for ``hermespy_rt.compute_paths`` (reference
    compute_paths_pybind11.cpp:99-186): loads the scene, calls the C entry
    ``compute_paths`` and returns ``(los, scatter)`` ChannelInfo objects."""
    L = lib()
    if L.hrt_device_count() <= 0:
        raise HrtError("no CUDA device available; hermespy-rt_b200 has no CPU path")
    if not os.path.exists(mesh_filepath):
        raise HrtError(f"scene file not found: {mesh_filepath}")
    rx = abi.vec3_array(rx_positions, num_rx); tx = abi.vec3_array(tx_positions, num_tx)
    sc = L.scene_load(str(mesh_filepath).encode())
    try:
        o = abi.call_compute_paths(L, sc, rx, tx, rx_velocities, tx_velocities,
                                   float(carrier_frequency), int(num_paths), int(num_bounces))
    finally:
        abi.free_scene(sc)
    return ChannelInfo(o.los, 1), ChannelInfo(o.scat, int(num_bounces) * int(num_paths))
