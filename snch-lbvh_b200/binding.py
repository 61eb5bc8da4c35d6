"""ctypes binding of include/snch_b200.h and the host-side mirror of ``lbvh::scene<3>``.

Inputs may be numpy arrays (HOST path: the library copies H2D/D2H itself — this is what ``bench.py`` times as ``e2e``)
or torch CUDA tensors (DEVICE path: zero-copy, results are torch tensors on the same device).
"""
from __future__ import annotations

import ctypes as C
import os
from enum import IntEnum

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# every symbol include/snch_b200.h declares (checked by tests/test_abi.py against the header and the built library)
ABI_SYMBOLS = [
    "snch_last_error", "snch_abi_version", "snch_scene3_create", "snch_scene_destroy", "snch_scene_compute_silhouettes",
    "snch_scene_build", "snch_scene_stats", "snch_scene_device_repr", "snch_scene_export", "snch_closest_point_batch",
    "snch_closest_silhouette_batch", "snch_intersect_batch", "snch_sample_in_sphere_batch", "snch_scene_arena",
    "snch_scene_adopt_arena", "snch_scene_set_option", "snch_scene_counter", "snch_lbvh_build", "snch_scene_update_vertices",
    "snch_wost_step_batch", "snch_scene_save", "snch_scene_load", "snch_scene_last_kernel",
    "snch_comm_unique_id", "snch_comm_create", "snch_comm_adopt", "snch_comm_destroy", "snch_scene_broadcast", "snch_scene_rebroadcast",
    "snch_scene_replicate_local",
    "snch_scene2_create", "snch_scene2_destroy", "snch_scene2_compute_silhouettes", "snch_scene2_build", "snch_scene2_stats",
    "snch_scene2_device_repr", "snch_scene2_export", "snch_scene2_set_option", "snch_closest_point_batch2",
    "snch_closest_silhouette_batch2", "snch_intersect_batch2", "snch_sample_in_sphere_batch2",
    "snch_selftest_host_libm",
]


class SnchError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(msg)
        self.status = status


class ExportKind(IntEnum):
    NODES = 0
    AABBS = 1
    CONES = 2
    MORTON_SORTED = 3
    SORTED_INDEX = 4
    RANGES = 5
    EDGES = 6
    TRI_EDGES = 7
    TRI_OWNED = 8
    Q1_TAINT = 9


class BuildOptions(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("keep_reference_layout", C.c_uint32), ("print_collision", C.c_uint32),
                ("refit_only", C.c_uint32)]


class WostIO(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("reserved", C.c_uint32), ("points_xyz", C.c_void_p), ("flip", C.c_void_p),
                ("dirs_xyz", C.c_void_p), ("rnd_uvw", C.c_void_p), ("closest_index", C.c_void_p), ("closest_distance", C.c_void_p),
                ("silhouette_distance", C.c_void_p), ("star_radius", C.c_void_p), ("hits", C.c_void_p), ("found", C.c_void_p),
                ("sample_index", C.c_void_p), ("sample_pdf", C.c_void_p), ("sample_point_xyz", C.c_void_p),
                ("silhouette_edge", C.c_void_p), ("silhouette_point_xyz", C.c_void_p)]


class BuildStats(C.Structure):
    _fields_ = [("num_objects", C.c_uint32), ("num_nodes", C.c_uint32), ("num_edges", C.c_uint32),
                ("num_vertices", C.c_uint32), ("morton_collision", C.c_uint32), ("q1_nodes", C.c_uint32),
                ("build_ms", C.c_float), ("adjacency_ms", C.c_float), ("arena_bytes", C.c_uint64),
                ("scene_lower", C.c_float * 3), ("scene_upper", C.c_float * 3)]


class BvhDevicePod(C.Structure):
    _fields_ = [("num_nodes", C.c_uint32), ("num_objects", C.c_uint32), ("nodes", C.c_void_p), ("aabbs", C.c_void_p),
                ("cones", C.c_void_p), ("objects", C.c_void_p), ("vertices", C.c_void_p), ("silhouettes", C.c_void_p),
                ("num_vertices", C.c_uint32), ("num_silhouettes", C.c_uint32)]


HIT_DTYPE = np.dtype([("t", np.float32), ("u", np.float32), ("v", np.float32), ("prim", np.uint32)])


def lib_path() -> str:
    return os.path.join(_HERE, "libsnch_b200.so")


_lib = None


def lib():
    """Load libsnch_b200.so.  Fails loudly when it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C snch-lbvh_b200/csrc`).  snch-lbvh_b200 has no CPU or PyTorch fallback.")
    L = C.CDLL(path)
    vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
    L.snch_last_error.restype = C.c_char_p
    L.snch_abi_version.restype = C.c_int
    L.snch_scene3_create.argtypes = [vp, u32, vp, u32, C.c_int, C.POINTER(vp)]
    L.snch_scene_destroy.argtypes = [vp]
    L.snch_scene_compute_silhouettes.argtypes = [vp]
    L.snch_scene_build.argtypes = [vp, C.POINTER(BuildOptions), vp]
    L.snch_scene_stats.argtypes = [vp, C.POINTER(BuildStats)]
    L.snch_scene_device_repr.argtypes = [vp, C.POINTER(BvhDevicePod)]
    L.snch_scene_export.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.snch_closest_point_batch.argtypes = [vp, vp, u64, vp, vp, vp]
    L.snch_closest_silhouette_batch.argtypes = [vp, vp, vp, vp, u64, vp, vp, vp, vp]
    L.snch_intersect_batch.argtypes = [vp, vp, vp, vp, u64, vp, vp, C.c_int, vp]
    L.snch_sample_in_sphere_batch.argtypes = [vp, vp, vp, u64, vp, vp, vp, vp]
    L.snch_scene_arena.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
    L.snch_scene_adopt_arena.argtypes = [vp, u64, C.c_int, vp, C.POINTER(vp)]
    L.snch_scene_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.snch_scene_counter.argtypes = [vp, C.c_char_p, C.POINTER(C.c_double), C.c_int]
    L.snch_scene_update_vertices.argtypes = [vp, vp, vp]
    L.snch_wost_step_batch.argtypes = [vp, C.POINTER(WostIO), u64, vp]
    L.snch_scene_save.argtypes = [vp, C.c_char_p]
    L.snch_scene_load.argtypes = [C.c_char_p, C.c_int, vp, C.POINTER(vp)]
    L.snch_lbvh_build.argtypes = [C.c_int, u32, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_int), vp]
    L.snch_scene2_create.argtypes = [vp, u32, vp, u32, C.c_int, C.POINTER(vp)]
    L.snch_scene2_destroy.argtypes = [vp]
    L.snch_scene2_compute_silhouettes.argtypes = [vp]
    L.snch_scene2_build.argtypes = [vp, vp]
    L.snch_scene2_stats.argtypes = [vp, C.POINTER(BuildStats)]
    L.snch_scene2_device_repr.argtypes = [vp, C.POINTER(BvhDevicePod)]
    L.snch_scene2_export.argtypes = [vp, C.c_int, vp, u64]
    L.snch_scene2_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.snch_closest_point_batch2.argtypes = [vp, vp, u64, vp, vp, vp]
    L.snch_closest_silhouette_batch2.argtypes = [vp, vp, vp, vp, u64, vp, vp, vp, vp]
    L.snch_intersect_batch2.argtypes = [vp, vp, vp, vp, u64, vp, vp, C.c_int, vp]
    L.snch_sample_in_sphere_batch2.argtypes = [vp, vp, vp, u64, vp, vp, vp, vp]
    L.snch_scene_last_kernel.argtypes = [vp]
    L.snch_comm_unique_id.argtypes = [vp, u64]
    L.snch_comm_create.argtypes = [vp, u64, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.snch_comm_adopt.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.snch_comm_destroy.argtypes = [vp]
    L.snch_scene_broadcast.argtypes = [vp, C.c_int, vp, vp, C.POINTER(vp)]
    L.snch_scene_rebroadcast.argtypes = [vp, C.c_int, vp, vp]
    L.snch_scene_replicate_local.argtypes = [vp, C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    for name in ABI_SYMBOLS:
        if name not in ("snch_last_error", "snch_scene_last_kernel"):
            getattr(L, name).restype = C.c_int
    L.snch_scene_last_kernel.restype = C.c_char_p
    _lib = L
    return L


def _check(status: int):
    if status != 0:
        raise SnchError(status, lib().snch_last_error().decode("utf-8", "replace"))


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class _Arg:
    """Normalises one array argument to (pointer, keep-alive object)."""

    def __init__(self, x, dtype, shape_tail=None, allow_none=False):
        self.obj = None
        self.ptr = None
        self.torch = False
        self.n = 0
        if x is None:
            if not allow_none:
                raise ValueError("missing array argument")
            return
        if _is_torch(x):
            import torch
            tdt = {np.float32: torch.float32, np.uint8: torch.uint8, np.int32: torch.int32}[dtype]
            if not x.is_cuda:
                raise ValueError("torch tensors must live on the GPU (use numpy arrays for the host path)")
            if x.dtype == torch.bool and dtype is np.uint8:
                x = x.to(torch.uint8)
            t = x.to(tdt).contiguous()
            self.obj, self.ptr, self.torch, self.n = t, t.data_ptr(), True, t.shape[0]
        else:
            a = np.ascontiguousarray(x, dtype=dtype)
            if shape_tail is not None:
                a = a.reshape((-1,) + shape_tail)
            self.obj, self.ptr, self.n = a, a.ctypes.data, a.shape[0]


class Scene3:
    """Mirror of ``lbvh::scene<3>`` (scene.cuh:1128-1252): construct from vertices + triangle indices, then
    ``compute_silhouettes()``, ``build_bvh()``; queries are batched (one launch per call)."""

    def __init__(self, vertices, indices, device: int = 0):
        self._L = lib()
        self.vertices_h = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        self.indices_h = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1, 3)
        self.device = int(device)
        h = C.c_void_p()
        _check(self._L.snch_scene3_create(self.vertices_h.ctypes.data, len(self.vertices_h), self.indices_h.ctypes.data,
                                          len(self.indices_h), self.device, C.byref(h)))
        self._h = h
        self._keep = None

    # -- lifetime ----------------------------------------------------------------------------------------------
    @classmethod
    def _from_handle(cls, handle, device):
        self = cls.__new__(cls)
        self._L = lib()
        self._h = handle
        self.device = int(device)
        self.vertices_h = None
        self.indices_h = None
        self._keep = None
        return self

    def close(self):
        if getattr(self, "_h", None):
            self._L.snch_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- build -------------------------------------------------------------------------------------------------
    def compute_silhouettes(self):
        _check(self._L.snch_scene_compute_silhouettes(self._h))
        return self

    def build_bvh(self, stream=None, print_collision=False, refit_only=False):
        opts = BuildOptions(C.sizeof(BuildOptions), 1, int(print_collision), int(refit_only))
        _check(self._L.snch_scene_build(self._h, C.byref(opts), _stream_ptr(stream)))
        return self

    def update_vertices(self, vertices, stream=None):
        """New positions for the same topology (numpy array or torch CUDA tensor); takes effect at the next build_bvh()
        (full rebuild, or ``refit_only=True`` to keep the Morton order and hierarchy)."""
        a = _Arg(vertices, np.float32, (3,))
        if a.n != self.stats()["num_vertices"]:
            raise ValueError("update_vertices: vertex count differs from the scene's")
        _check(self._L.snch_scene_update_vertices(self._h, a.ptr, _stream_ptr(stream, a)))
        if not a.torch:
            self.vertices_h = a.obj
        return self

    def save(self, path: str):
        """Write the built scene (its pointer-free arena) to a file."""
        _check(self._L.snch_scene_save(self._h, os.fsencode(path)))
        return self

    @classmethod
    def load(cls, path: str, device: int = 0, stream=None):
        """A scene saved with save(): answers queries and exports like the original; cannot be rebuilt."""
        h = C.c_void_p()
        _check(lib().snch_scene_load(os.fsencode(path), int(device), _stream_ptr(stream), C.byref(h)))
        return cls._from_handle(h, device)

    def set_option(self, name: str, value: int):
        """Scheduling knobs of the batched kernels (include/snch_b200.h: snch_scene_set_option); results never depend on them."""
        _check(self._L.snch_scene_set_option(self._h, name.encode(), int(value)))
        return self

    def counter(self, name: str, reset: bool = False) -> float:
        """Launch accounting (include/snch_b200.h: snch_scene_counter)."""
        v = C.c_double()
        _check(self._L.snch_scene_counter(self._h, name.encode(), C.byref(v), int(reset)))
        return v.value

    def stats(self) -> dict:
        st = BuildStats()
        _check(self._L.snch_scene_stats(self._h, C.byref(st)))
        d = {k: getattr(st, k) for k, _ in BuildStats._fields_ if k not in ("scene_lower", "scene_upper")}
        d["scene_lower"] = list(st.scene_lower)
        d["scene_upper"] = list(st.scene_upper)
        return d

    def get_bvh_device_ptr(self) -> BvhDevicePod:
        """Reference-layout device pointers (``bvh_device``); raises "BVH is not built yet." like scene.cuh:1250."""
        pod = BvhDevicePod()
        _check(self._L.snch_scene_device_repr(self._h, C.byref(pod)))
        return pod

    def export(self, kind: ExportKind) -> np.ndarray:
        st = self.stats()
        n, nn, ne = st["num_objects"], st["num_nodes"], st["num_edges"]
        shape, dt = {
            ExportKind.NODES: ((nn, 4), np.uint32), ExportKind.AABBS: ((nn, 6), np.float32),
            ExportKind.CONES: ((nn, 5), np.float32), ExportKind.MORTON_SORTED: ((n,), np.uint32),
            ExportKind.SORTED_INDEX: ((n,), np.uint32), ExportKind.RANGES: ((max(n - 1, 0), 2), np.uint32),
            ExportKind.EDGES: ((ne, 4), np.int32), ExportKind.TRI_EDGES: ((n, 3), np.int32),
            ExportKind.TRI_OWNED: ((n, 3), np.int32), ExportKind.Q1_TAINT: ((nn,), np.uint8),
        }[ExportKind(kind)]
        out = np.zeros(shape, dt)
        _check(self._L.snch_scene_export(self._h, int(kind), out.ctypes.data, out.nbytes))
        return out

    # -- queries -----------------------------------------------------------------------------------------------
    def closest_point(self, points, stream=None):
        """-> (index uint32, distance float32).  query_device(bvh, nearest(p), distance_calculator())"""
        q = _Arg(points, np.float32, (3,))
        idx, ip = _out(q, (q.n,), np.uint32)
        dist, dp = _out(q, (q.n,), np.float32)
        _check(self._L.snch_closest_point_batch(self._h, q.ptr, q.n, ip, dp, _stream_ptr(stream, q)))
        return idx, dist

    def closest_silhouette(self, points, flip=None, r_max=None, stream=None, with_edge=False):
        """-> distance float32 (+inf when none).  query_device(bvh, nearest_silhouette(p, flip), ...)
        with_edge=True -> (distance, edge index uint32 (0xFFFFFFFF when none), closest point on that edge float32[n, 3])."""
        q = _Arg(points, np.float32, (3,))
        if isinstance(flip, (bool, np.bool_)):
            flip = None if not flip else (np.ones(q.n, np.uint8) if not q.torch else _torch_full(q, 1))
        f = _Arg(flip, np.uint8, None, allow_none=True)
        r = _Arg(r_max, np.float32, None, allow_none=True)
        dist, dp = _out(q, (q.n,), np.float32)
        if not with_edge:
            _check(self._L.snch_closest_silhouette_batch(self._h, q.ptr, f.ptr, r.ptr, q.n, dp, None, None, _stream_ptr(stream, q)))
            return dist
        edge, ep = _out(q, (q.n,), np.uint32)
        pt, pp = _out(q, (q.n, 3), np.float32)
        _check(self._L.snch_closest_silhouette_batch(self._h, q.ptr, f.ptr, r.ptr, q.n, dp, ep, pp, _stream_ptr(stream, q)))
        return dist, edge, pt

    def last_kernel(self) -> str:
        """Name of the traversal kernel the last batched call launched (include/snch_b200.h: snch_scene_last_kernel)."""
        return self._L.snch_scene_last_kernel(self._h).decode()

    def intersect(self, origins, directions, t_max=None, any_hit=False, stream=None):
        """-> (found uint8, hits[t,u,v,prim]).  query_device(bvh, ray_intersect<any_hit>(ray, max_dist), intersect_test())"""
        o = _Arg(origins, np.float32, (3,))
        d = _Arg(directions, np.float32, (3,))
        tm = _Arg(t_max, np.float32, None, allow_none=True)
        found, fp = _out(o, (o.n,), np.uint8)
        if any_hit:
            _check(self._L.snch_intersect_batch(self._h, o.ptr, d.ptr, tm.ptr, o.n, None, fp, 1, _stream_ptr(stream, o)))
            return found, None
        if o.torch:
            import torch
            hits = torch.empty((o.n, 4), dtype=torch.float32, device=o.obj.device)
            hp = hits.data_ptr()
        else:
            hits = np.zeros(o.n, HIT_DTYPE)
            hp = hits.ctypes.data
        _check(self._L.snch_intersect_batch(self._h, o.ptr, d.ptr, tm.ptr, o.n, hp, fp, 0, _stream_ptr(stream, o)))
        return found, hits

    def sample_in_sphere(self, spheres, rnd, stream=None):
        """-> (index int32 (-1 = miss), pdf float32, point float32[n,3]).  sample_object_in_sphere + sample_on_object"""
        s = _Arg(spheres, np.float32, (4,))
        r = _Arg(rnd, np.float32, (3,))
        idx, ip = _out(s, (s.n,), np.int32)
        pdf, pp = _out(s, (s.n,), np.float32)
        pt, tp = _out(s, (s.n, 3), np.float32)
        _check(self._L.snch_sample_in_sphere_batch(self._h, s.ptr, r.ptr, s.n, ip, pp, tp, _stream_ptr(stream, s)))
        return idx, pdf, pt

    def wost_step(self, points, directions=None, rnd=None, flip=None, stream=None, with_edge=False) -> dict:
        """One wavefront walk-on-stars step per walker (include/snch_b200.h: snch_wost_step_batch): closest point, silhouette
        within the closest distance, star radius = min of both, ray along `directions` up to the star radius, triangle sampled
        in the star sphere.  Returns a dict of arrays (numpy in -> numpy out, torch CUDA in -> torch out)."""
        q = _Arg(points, np.float32, (3,))
        d = _Arg(directions, np.float32, (3,), allow_none=True)
        r = _Arg(rnd, np.float32, (3,), allow_none=True)
        f = _Arg(flip, np.uint8, None, allow_none=True)
        out, io = {}, WostIO()
        io.struct_size = C.sizeof(WostIO)
        io.points_xyz, io.flip, io.dirs_xyz, io.rnd_uvw = q.ptr, f.ptr, d.ptr, r.ptr
        out["closest_index"], io.closest_index = _out(q, (q.n,), np.uint32)
        out["closest_distance"], io.closest_distance = _out(q, (q.n,), np.float32)
        out["silhouette_distance"], io.silhouette_distance = _out(q, (q.n,), np.float32)
        out["star_radius"], io.star_radius = _out(q, (q.n,), np.float32)
        if with_edge:
            out["silhouette_edge"], io.silhouette_edge = _out(q, (q.n,), np.uint32)
            out["silhouette_point"], io.silhouette_point_xyz = _out(q, (q.n, 3), np.float32)
        if d.ptr:
            out["found"], io.found = _out(q, (q.n,), np.uint8)
            if q.torch:
                import torch
                out["hits"] = torch.empty((q.n, 4), dtype=torch.float32, device=q.obj.device)
                io.hits = out["hits"].data_ptr()
            else:
                out["hits"] = np.zeros(q.n, HIT_DTYPE)
                io.hits = out["hits"].ctypes.data
        if r.ptr:
            out["sample_index"], io.sample_index = _out(q, (q.n,), np.int32)
            out["sample_pdf"], io.sample_pdf = _out(q, (q.n,), np.float32)
            out["sample_point"], io.sample_point_xyz = _out(q, (q.n, 3), np.float32)
        _check(self._L.snch_wost_step_batch(self._h, C.byref(io), q.n, _stream_ptr(stream, q)))
        return out

    # -- replication (multi-GPU) ---------------------------------------------------------------------------------
    def arena(self):
        """(device pointer, bytes) of the pointer-free arena that holds the whole built scene."""
        p = C.c_void_p()
        n = C.c_uint64()
        _check(self._L.snch_scene_arena(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def arena_tensor(self):
        """The arena as a torch uint8 CUDA tensor (a view, no copy) — what rank 0 hands to torch.distributed.broadcast."""
        import torch
        ptr, nbytes = self.arena()

        class _View:
            __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}

        with torch.cuda.device(self.device):
            return torch.as_tensor(_View(), device=f"cuda:{self.device}")

    def replicate_local(self, devices):
        """Replicas on other GPUs of THIS process (cudaMemcpyPeerAsync fan-out, include/snch_b200.h: snch_scene_replicate_local)."""
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        outs = (C.c_void_p * len(devices))()
        _check(self._L.snch_scene_replicate_local(self._h, devs, len(devices), outs))
        return [Scene3._from_handle(C.c_void_p(outs[i]), devices[i]) for i in range(len(devices))]

    @classmethod
    def adopt_arena(cls, arena_tensor, device: int, stream=None):
        """Create a replica on `device` from a byte-exact copy of another scene's arena (torch uint8 CUDA tensor)."""
        h = C.c_void_p()
        _check(lib().snch_scene_adopt_arena(arena_tensor.data_ptr(), arena_tensor.numel(), int(device), _stream_ptr(stream),
                                            C.byref(h)))
        return cls._from_handle(h, device)


class Comm:
    """A communicator of the library's own (libnccl.so.2, no torch): ``Comm.unique_id()`` on rank 0, ship the 128 bytes to the
    other ranks by any means, then ``Comm(id, rank, world, device)`` on every rank (collective).  ``broadcast(scene, root)``
    returns the root's scene on the root and a replica elsewhere; ``rebroadcast`` refreshes existing replicas in place."""

    ID_BYTES = 128

    def __init__(self, unique_id: bytes, rank: int, world: int, device: int):
        self._L = lib()
        self.rank, self.world, self.device = int(rank), int(world), int(device)
        buf = C.create_string_buffer(bytes(unique_id), self.ID_BYTES)
        h = C.c_void_p()
        _check(self._L.snch_comm_create(buf, self.ID_BYTES, self.rank, self.world, self.device, C.byref(h)))
        self._h = h

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(Comm.ID_BYTES)
        _check(lib().snch_comm_unique_id(buf, Comm.ID_BYTES))
        return buf.raw

    def broadcast(self, scene, root: int = 0, stream=None):
        out = C.c_void_p()
        _check(self._L.snch_scene_broadcast(scene._h if scene is not None else None, int(root), self._h, _stream_ptr(stream), C.byref(out)))
        return scene if self.rank == root else Scene3._from_handle(out, self.device)

    def rebroadcast(self, scene, root: int = 0, stream=None):
        _check(self._L.snch_scene_rebroadcast(scene._h, int(root), self._h, _stream_ptr(stream)))
        return scene

    def close(self):
        if getattr(self, "_h", None):
            self._L.snch_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scene2:
    """Mirror of ``lbvh::scene<2>`` (scene.cuh:287-703): polylines — vertices (n, 2) + segment vertex indices (m, 2) — then
    ``compute_silhouettes()``, ``build_bvh()``; batched queries like :class:`Scene3` (numpy in -> numpy out, torch CUDA in ->
    torch out).  A ray hit is (t, u = segment parameter, v = 0, prim); sampling takes circles (x, y, radius) and two uniforms."""

    def __init__(self, vertices, indices, device: int = 0):
        self._L = lib()
        self.vertices_h = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 2)
        self.indices_h = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1, 2)
        self.device = int(device)
        h = C.c_void_p()
        _check(self._L.snch_scene2_create(self.vertices_h.ctypes.data, len(self.vertices_h), self.indices_h.ctypes.data,
                                          len(self.indices_h), self.device, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.snch_scene2_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def compute_silhouettes(self):
        _check(self._L.snch_scene2_compute_silhouettes(self._h))
        return self

    def build_bvh(self, stream=None):
        _check(self._L.snch_scene2_build(self._h, _stream_ptr(stream)))
        return self

    def set_option(self, name: str, value: int):
        _check(self._L.snch_scene2_set_option(self._h, name.encode(), int(value)))
        return self

    def stats(self) -> dict:
        st = BuildStats()
        _check(self._L.snch_scene2_stats(self._h, C.byref(st)))
        return {k: getattr(st, k) for k in ("num_objects", "num_nodes", "num_edges", "num_vertices", "morton_collision", "build_ms")}

    def get_bvh_device_ptr(self) -> BvhDevicePod:
        """Reference-layout device pointers (``bvh_device<float, 2, line_segment>``); raises "BVH is not built yet." (scene.cuh:686)."""
        pod = BvhDevicePod()
        _check(self._L.snch_scene2_device_repr(self._h, C.byref(pod)))
        return pod

    def export(self, kind: ExportKind) -> np.ndarray:
        n, nv = len(self.indices_h), len(self.vertices_h)
        nn = 2 * n - 1 if n else 0
        shape, dt = {
            ExportKind.NODES: ((nn, 4), np.uint32), ExportKind.AABBS: ((nn, 4), np.float32), ExportKind.CONES: ((nn, 4), np.float32),
            ExportKind.MORTON_SORTED: ((n,), np.uint32), ExportKind.SORTED_INDEX: ((n,), np.uint32),
            ExportKind.EDGES: ((nv, 4), np.int32), ExportKind.TRI_OWNED: ((n, 2), np.int32),
        }[ExportKind(kind)]
        out = np.zeros(shape, dt)
        _check(self._L.snch_scene2_export(self._h, int(kind), out.ctypes.data, out.nbytes))
        return out

    def closest_point(self, points, stream=None):
        """-> (index uint32, distance float32).  query_device(bvh, nearest(p), scene<2>::distance_calculator())"""
        q = _Arg(points, np.float32, (2,))
        idx, ip = _out(q, (q.n,), np.uint32)
        dist, dp = _out(q, (q.n,), np.float32)
        _check(self._L.snch_closest_point_batch2(self._h, q.ptr, q.n, ip, dp, _stream_ptr(stream, q)))
        return idx, dist

    def closest_silhouette(self, points, flip=None, r_max=None, stream=None, with_vertex=False):
        """-> distance float32 (+inf when none).  query_device(bvh, nearest_silhouette(p, flip), silhouette_distance_calculator())
        with_vertex=True -> (distance, silhouette vertex index uint32 (0xFFFFFFFF when none), its position float32[n, 2])."""
        q = _Arg(points, np.float32, (2,))
        if isinstance(flip, (bool, np.bool_)):
            flip = None if not flip else (np.ones(q.n, np.uint8) if not q.torch else _torch_full(q, 1))
        f = _Arg(flip, np.uint8, None, allow_none=True)
        r = _Arg(r_max, np.float32, None, allow_none=True)
        dist, dp = _out(q, (q.n,), np.float32)
        if not with_vertex:
            _check(self._L.snch_closest_silhouette_batch2(self._h, q.ptr, f.ptr, r.ptr, q.n, dp, None, None, _stream_ptr(stream, q)))
            return dist
        vid, vp_ = _out(q, (q.n,), np.uint32)
        pt, pp = _out(q, (q.n, 2), np.float32)
        _check(self._L.snch_closest_silhouette_batch2(self._h, q.ptr, f.ptr, r.ptr, q.n, dp, vp_, pp, _stream_ptr(stream, q)))
        return dist, vid, pt

    def intersect(self, origins, directions, t_max=None, any_hit=False, stream=None):
        """-> (found uint8, hits[t,u=s,v=0,prim]).  query_device(bvh, ray_intersect<any_hit>(ray, max_dist), intersect_test())"""
        o = _Arg(origins, np.float32, (2,))
        d = _Arg(directions, np.float32, (2,))
        tm = _Arg(t_max, np.float32, None, allow_none=True)
        found, fp = _out(o, (o.n,), np.uint8)
        if any_hit:
            _check(self._L.snch_intersect_batch2(self._h, o.ptr, d.ptr, tm.ptr, o.n, None, fp, 1, _stream_ptr(stream, o)))
            return found, None
        if o.torch:
            import torch
            hits = torch.empty((o.n, 4), dtype=torch.float32, device=o.obj.device)
            hp = hits.data_ptr()
        else:
            hits = np.zeros(o.n, HIT_DTYPE)
            hp = hits.ctypes.data
        _check(self._L.snch_intersect_batch2(self._h, o.ptr, d.ptr, tm.ptr, o.n, hp, fp, 0, _stream_ptr(stream, o)))
        return found, hits

    def sample_in_sphere(self, circles, rnd, stream=None):
        """-> (index int32 (-1 = miss), pdf float32, point float32[n,2]).  sample_object_in_sphere + sample_on_object in 2-D"""
        s = _Arg(circles, np.float32, (3,))
        r = _Arg(rnd, np.float32, (2,))
        idx, ip = _out(s, (s.n,), np.int32)
        pdf, pp = _out(s, (s.n,), np.float32)
        pt, tp = _out(s, (s.n, 2), np.float32)
        _check(self._L.snch_sample_in_sphere_batch2(self._h, s.ptr, r.ptr, s.n, ip, pp, tp, _stream_ptr(stream, s)))
        return idx, pdf, pt


def _stream_ptr(stream, like=None):
    """cudaStream_t for the call.  With torch CUDA arguments and no explicit stream, torch's CURRENT stream of that device:
    inputs and outputs were produced / allocated in that stream's order (the legacy default stream would not wait for a
    non-blocking torch stream, and the caching allocator could reuse a temporary while the kernel still reads it)."""
    if stream is None:
        if like is not None and getattr(like, "torch", False):
            import torch
            return C.c_void_p(torch.cuda.current_stream(like.obj.device).cuda_stream)
        return None
    if hasattr(stream, "cuda_stream"):
        return C.c_void_p(stream.cuda_stream)
    return C.c_void_p(int(stream))


def _torch_full(like: _Arg, value: int):
    import torch
    return torch.full((like.n,), value, dtype=torch.uint8, device=like.obj.device)


def _out(like: _Arg, shape, dtype):
    """Allocate an output next to `like` (torch CUDA tensor or numpy array) and return (object, pointer)."""
    if like.torch:
        import torch
        tdt = {np.float32: torch.float32, np.uint8: torch.uint8, np.int32: torch.int32, np.uint32: torch.int32}[dtype]
        t = torch.empty(shape, dtype=tdt, device=like.obj.device)
        return t, t.data_ptr()
    a = np.zeros(shape, dtype)
    return a, a.ctypes.data
