"""ctypes front-ends for the CPU oracles (TEST INFRASTRUCTURE — see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(HERE, "_ref")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def build_oracle(with_ref: bool = True) -> None:
    """Compile liboracle.so (always) and, when /root/reference is present, oracle/_ref/*.so."""
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so", "libhostlibm.so"])
    if with_ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


_oracle_lib = None


def oracle_lib():
    global _oracle_lib
    if _oracle_lib is not None:
        return _oracle_lib
    path = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, "snch_oracle.c"), os.path.join(HERE, "snch_oracle2.c")]
    if not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(x) for x in srcs):
        build_oracle(with_ref=False)
    L = C.CDLL(path)
    vp = C.c_void_p
    L.orc_scene3_create.restype = vp
    L.orc_scene3_create.argtypes = [_f32p, C.c_int, _i32p, C.c_int]
    L.orc_scene_destroy.argtypes = [vp]
    for name in ("orc_num_objects", "orc_num_nodes", "orc_num_edges", "orc_collision"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = C.c_int
    L.orc_export_tree.argtypes = [vp, _u32p, _f32p, _f32p]
    L.orc_export_q1_taint.argtypes = [vp, _u8p]
    L.orc_export_adjacency.argtypes = [vp, _i32p, _i32p, _i32p]
    L.orc_export_morton.argtypes = [vp, _u32p, _u32p]
    L.orc_export_ranges.argtypes = [vp, _u32p]
    L.orc_closest.argtypes = [vp, _f32p, C.c_long, _u32p, _f32p, C.c_int]
    L.orc_silhouette.argtypes = [vp, _f32p, C.c_long, C.c_int, C.c_void_p, _f32p, C.c_int]
    L.orc_silhouette_ex.argtypes = [vp, _f32p, C.c_long, C.c_int, C.c_void_p, _f32p, _i32p, _f32p, C.c_int]
    L.orc_point_edge_distance.argtypes = [vp, _f32p, _i32p, C.c_long, _f32p, _f32p]
    L.orc_ray.argtypes = [vp, _f32p, _f32p, _f32p, C.c_long, C.c_int, _i32p, _f32p, _f32p, _u32p, C.c_int]
    L.orc_sample.argtypes = [vp, _f32p, _f32p, C.c_long, _i32p, _f32p, C.c_int]
    L.orc_sample_on_object.argtypes = [vp, _i32p, _f32p, _f32p, C.c_long, _f32p]
    L.orc_closest_brute.argtypes = [vp, _f32p, C.c_long, _u32p, _f32p, C.c_int]
    L.orc_ray_brute.argtypes = [vp, _f32p, _f32p, _f32p, C.c_long, _i32p, _f32p, _u32p, C.c_int]
    L.orc_point_triangle_distance.argtypes = [vp, _f32p, _u32p, C.c_long, _f32p]
    dp = C.POINTER(C.c_double)
    L.orc_closest_must_visit.argtypes = [vp, _f32p, C.c_long, dp, dp]
    L.orc_silhouette_must_visit.argtypes = [vp, _f32p, C.c_long, C.c_int, C.c_void_p, dp, dp]
    L.orc_ray_must_visit.argtypes = [vp, _f32p, _f32p, _f32p, C.c_long, dp, dp]
    L.orc_host_libm.argtypes = [C.c_int, _f32p, C.c_long, _f32p]
    L.orc_morton3.argtypes = [C.c_float, C.c_float, C.c_float]
    L.orc_morton3.restype = C.c_uint32
    L.orc_expand_bits.argtypes = [C.c_uint32]
    L.orc_expand_bits.restype = C.c_uint32
    # 2-D (snch_oracle2.c)
    L.orc_scene2_create.restype = vp
    L.orc_scene2_create.argtypes = [_f32p, C.c_int, _i32p, C.c_int]
    L.orc_scene2_destroy.argtypes = [vp]
    for name in ("orc2_num_nodes", "orc2_collision"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = C.c_int
    L.orc2_export_tree.argtypes = [vp, _u32p, _f32p, _f32p, _u8p]
    L.orc2_export_adjacency.argtypes = [vp, _i32p, _i32p]
    L.orc2_closest.argtypes = [vp, _f32p, C.c_long, _u32p, _f32p]
    L.orc2_silhouette.argtypes = [vp, _f32p, C.c_long, C.c_int, C.c_void_p, _f32p]
    L.orc2_silhouette_ex.argtypes = [vp, _f32p, C.c_long, C.c_int, C.c_void_p, _f32p, _i32p, _f32p]
    L.orc2_ray.argtypes = [vp, _f32p, _f32p, C.c_void_p, C.c_long, _i32p, _f32p, _f32p, _u32p]
    L.orc2_sample.argtypes = [vp, _f32p, _f32p, C.c_long, _i32p, _f32p]
    _oracle_lib = L
    return L


class _TreeExports:
    """Shared result containers."""

    def tree(self):
        raise NotImplementedError


class OracleScene:
    """The plain-C restatement (oracle/snch_oracle.c)."""

    def __init__(self, verts, tris):
        self.L = oracle_lib()
        self.verts = _f32(verts).reshape(-1, 3)
        self.tris = _i32(tris).reshape(-1, 3)
        self.h = self.L.orc_scene3_create(self.verts, len(self.verts), self.tris, len(self.tris))
        self.n = len(self.tris)
        self.num_nodes = self.L.orc_num_nodes(self.h)
        self.num_edges = self.L.orc_num_edges(self.h)
        self.collision = bool(self.L.orc_collision(self.h))

    def close(self):
        if self.h:
            self.L.orc_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tree(self):
        nn = self.num_nodes
        nodes = np.zeros((nn, 4), np.uint32)
        aabbs = np.zeros((nn, 6), np.float32)
        cones = np.zeros((nn, 5), np.float32)
        self.L.orc_export_tree(self.h, nodes, aabbs, cones)
        return nodes, aabbs, cones

    def q1_taint(self):
        q = np.zeros(self.num_nodes, np.uint8)
        self.L.orc_export_q1_taint(self.h, q)
        return q.astype(bool)

    def adjacency(self):
        e = np.zeros((self.num_edges, 4), np.int32)
        te = np.zeros((self.n, 3), np.int32)
        to = np.zeros((self.n, 3), np.int32)
        self.L.orc_export_adjacency(self.h, e, te, to)
        return e, te, to

    def morton(self):
        m = np.zeros(self.n, np.uint32)
        si = np.zeros(self.n, np.uint32)
        self.L.orc_export_morton(self.h, m, si)
        return m, si

    def ranges(self):
        r = np.zeros((max(self.n - 1, 0), 2), np.uint32)
        if self.n > 1:
            self.L.orc_export_ranges(self.h, r)
        return r

    def closest(self, q, nthreads=1, brute=False):
        q = _f32(q).reshape(-1, 3)
        idx = np.zeros(len(q), np.uint32)
        dist = np.zeros(len(q), np.float32)
        (self.L.orc_closest_brute if brute else self.L.orc_closest)(self.h, q, len(q), idx, dist, nthreads)
        return idx, dist

    def silhouette(self, q, flip=False, r_max=None, nthreads=1):
        q = _f32(q).reshape(-1, 3)
        dist = np.zeros(len(q), np.float32)
        rm = None
        if r_max is not None:
            rm = _f32(r_max)
            assert len(rm) == len(q)
        self.L.orc_silhouette(self.h, q, len(q), int(flip), rm.ctypes.data if rm is not None else None, dist, nthreads)
        return dist

    def silhouette_ex(self, q, flip=False, r_max=None, nthreads=1):
        """-> (distance, edge id int32 (-1 = none), closest point on that edge): the reference's walk plus the two values it
        computes and drops (scene.cuh:796-799, query.cuh:386,411)."""
        q = _f32(q).reshape(-1, 3)
        dist = np.zeros(len(q), np.float32)
        edge = np.zeros(len(q), np.int32)
        point = np.zeros((len(q), 3), np.float32)
        rm = None
        if r_max is not None:
            rm = _f32(r_max)
            assert len(rm) == len(q)
        self.L.orc_silhouette_ex(self.h, q, len(q), int(flip), rm.ctypes.data if rm is not None else None, dist, edge, point, nthreads)
        return dist, edge, point

    def point_edge_distance(self, q, edge):
        """distance from q[i] to silhouette edge edge[i] and the closest point on it (scene.cuh:230-255)"""
        q = _f32(q).reshape(-1, 3)
        e = _i32(edge)
        d = np.zeros(len(q), np.float32)
        pt = np.zeros((len(q), 3), np.float32)
        self.L.orc_point_edge_distance(self.h, q, e, len(q), d, pt)
        return d, pt

    def ray(self, org, dirs, tmax=None, any_hit=False, nthreads=1, brute=False):
        org = _f32(org).reshape(-1, 3)
        dirs = _f32(dirs).reshape(-1, 3)
        n = len(org)
        tm = np.full(n, np.inf, np.float32) if tmax is None else _f32(tmax)
        found = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        uv = np.zeros((n, 2), np.float32)
        prim = np.zeros(n, np.uint32)
        if brute:
            self.L.orc_ray_brute(self.h, org, dirs, tm, n, found, t, prim, nthreads)
        else:
            self.L.orc_ray(self.h, org, dirs, tm, n, int(any_hit), found, t, uv, prim, nthreads)
        return found, t, uv, prim

    def sample(self, sph, u, nthreads=1):
        sph = _f32(sph).reshape(-1, 4)
        u = _f32(u)
        idx = np.zeros(len(sph), np.int32)
        pdf = np.zeros(len(sph), np.float32)
        self.L.orc_sample(self.h, sph, u, len(sph), idx, pdf, nthreads)
        return idx, pdf

    def sample_on_object(self, idx, u, v):
        idx = _i32(idx)
        out = np.zeros((len(idx), 3), np.float32)
        self.L.orc_sample_on_object(self.h, idx, _f32(u), _f32(v), len(idx), out)
        return out

    def point_triangle_distance(self, q, idx):
        q = _f32(q).reshape(-1, 3)
        idx = np.ascontiguousarray(idx, np.uint32)
        d = np.zeros(len(q), np.float32)
        self.L.orc_point_triangle_distance(self.h, q, idx, len(q), d)
        return d

    def must_visit(self, kind, q, dirs=None, tmax=None, flip=False, r_max=None):
        a = C.c_double()
        b = C.c_double()
        q = _f32(q).reshape(-1, 3)
        if kind == "closest":
            self.L.orc_closest_must_visit(self.h, q, len(q), C.byref(a), C.byref(b))
        elif kind == "silhouette":
            rm = _f32(r_max) if r_max is not None else None
            self.L.orc_silhouette_must_visit(self.h, q, len(q), int(flip), rm.ctypes.data if rm is not None else None,
                                             C.byref(a), C.byref(b))
        elif kind == "ray":
            d = _f32(dirs).reshape(-1, 3)
            tm = np.full(len(q), np.inf, np.float32) if tmax is None else _f32(tmax)
            self.L.orc_ray_must_visit(self.h, q, d, tm, len(q), C.byref(a), C.byref(b))
        else:
            raise ValueError(kind)
        return a.value, b.value


class OracleScene2:
    """Plain-C restatement of lbvh::scene<2> (oracle/snch_oracle2.c): polylines, segments / silhouette vertices."""

    def __init__(self, verts, segs):
        self.L = oracle_lib()
        self.verts = _f32(verts).reshape(-1, 2)
        self.segs = _i32(segs).reshape(-1, 2)
        self.h = self.L.orc_scene2_create(self.verts, len(self.verts), self.segs, len(self.segs))
        self.n = len(self.segs)
        self.num_nodes = self.L.orc2_num_nodes(self.h)

    def close(self):
        if self.h:
            self.L.orc_scene2_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tree(self):
        """nodes (nn, 4) u32, aabbs (nn, 4) {upper xy, lower xy}, cones (nn, 4) {axis xy, half_angle, radius}, q1 taint (nn,) u8"""
        nn = self.num_nodes
        nodes, aabbs, cones, q1 = np.zeros((nn, 4), np.uint32), np.zeros((nn, 4), np.float32), np.zeros((nn, 4), np.float32), np.zeros(nn, np.uint8)
        self.L.orc2_export_tree(self.h, nodes, aabbs, cones, q1)
        return nodes, aabbs, cones, q1

    def adjacency(self):
        v4, owned = np.zeros((len(self.verts), 4), np.int32), np.zeros((self.n, 2), np.int32)
        self.L.orc2_export_adjacency(self.h, v4, owned)
        return v4, owned

    def closest(self, q):
        q = _f32(q).reshape(-1, 2)
        idx, dist = np.zeros(len(q), np.uint32), np.zeros(len(q), np.float32)
        self.L.orc2_closest(self.h, q, len(q), idx, dist)
        return idx, dist

    def silhouette(self, q, flip=False, r_max=None):
        q = _f32(q).reshape(-1, 2)
        dist = np.zeros(len(q), np.float32)
        r = None if r_max is None else _f32(np.broadcast_to(r_max, (len(q),)))
        self.L.orc2_silhouette(self.h, q, len(q), int(flip), None if r is None else r.ctypes.data, dist)
        return dist

    def silhouette_ex(self, q, flip=False, r_max=None):
        """-> (distance, silhouette vertex id int32 (-1 = none), its position)"""
        q = _f32(q).reshape(-1, 2)
        dist, vid, pt = np.zeros(len(q), np.float32), np.zeros(len(q), np.int32), np.zeros((len(q), 2), np.float32)
        r = None if r_max is None else _f32(np.broadcast_to(r_max, (len(q),)))
        self.L.orc2_silhouette_ex(self.h, q, len(q), int(flip), None if r is None else r.ctypes.data, dist, vid, pt)
        return dist, vid, pt

    def ray(self, org, dirs, tmax=None):
        org, dirs = _f32(org).reshape(-1, 2), _f32(dirs).reshape(-1, 2)
        n = len(org)
        tm = None if tmax is None else _f32(np.broadcast_to(tmax, (n,)))
        found, t, s, prim = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.uint32)
        self.L.orc2_ray(self.h, org, dirs, None if tm is None else tm.ctypes.data, n, found, t, s, prim)
        return found, t, s, prim

    def sample(self, sph, u):
        sph, u = _f32(sph).reshape(-1, 3), _f32(u).reshape(-1)
        idx, pdf = np.zeros(len(sph), np.int32), np.zeros(len(sph), np.float32)
        self.L.orc2_sample(self.h, sph, u, len(sph), idx, pdf)
        return idx, pdf


def host_libm(which: int, x) -> np.ndarray:
    """acosf (0) / sinf (1) / cosf (2) / logf (3) of the HOST's libm, element-wise"""
    x = _f32(x)
    out = np.empty_like(x)
    oracle_lib().orc_host_libm(which, x, x.size, out)
    return out


def restated_libm(which: int, x) -> np.ndarray:
    """the HOST build of the product's restatement (include/snch_lbvh/core/host_libm.cuh `*_glibc`), element-wise"""
    path = os.path.join(HERE, "libhostlibm.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-C", HERE, "libhostlibm.so"])
    L = C.CDLL(path)
    L.orc_restated_libm.argtypes = [C.c_int, _f32p, C.c_long, _f32p]
    x = _f32(x)
    out = np.empty_like(x)
    L.orc_restated_libm(which, x, x.size, out)
    return out


def ref_available(kind: str = "cpu") -> bool:
    return os.path.exists(os.path.join(_REF_DIR, {"cpu": "libsnch_ref_cpu.so", "cuda": "libsnch_ref_cuda.so",
                                                  "fcpw": "libfcpw_cpu.so"}[kind]))


_ref_libs = {}


def _ref_lib(kind):
    if kind in _ref_libs:
        return _ref_libs[kind]
    path = os.path.join(_REF_DIR, "libsnch_ref_cpu.so" if kind == "cpu" else "libsnch_ref_cuda.so")
    L = C.CDLL(path)
    vp = C.c_void_p
    L.ref3_create.restype = vp
    L.ref3_create.argtypes = [_f32p, C.c_int, _i32p, C.c_int]
    L.ref3_destroy.argtypes = [vp]
    L.ref3_time_construct.argtypes = [vp, C.c_int]
    L.ref3_time_construct.restype = C.c_double
    L.ref3_timings.argtypes = [vp, C.POINTER(C.c_double)]
    for name in ("ref3_num_objects", "ref3_num_nodes", "ref3_num_edges"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = C.c_int
    L.ref3_export_tree.argtypes = [vp, _u32p, _f32p, _f32p]
    L.ref3_export_adjacency.argtypes = [vp, _i32p, _i32p, _i32p]
    L.ref3_export_morton.argtypes = [vp, _u32p]
    L.ref3_closest.argtypes = [vp, _f32p, C.c_long, _u32p, _f32p, C.c_int]
    L.ref3_closest.restype = C.c_double
    L.ref3_silhouette.argtypes = [vp, _f32p, C.c_long, C.c_int, _f32p, C.c_int]
    L.ref3_silhouette.restype = C.c_double
    L.ref3_ray.argtypes = [vp, _f32p, _f32p, _f32p, C.c_long, _i32p, _f32p, _f32p, _u32p, C.c_int]
    L.ref3_ray.restype = C.c_double
    L.ref3_sample.argtypes = [vp, _f32p, _f32p, C.c_long, _i32p, _f32p, C.c_int]
    L.ref3_sample.restype = C.c_double
    if hasattr(L, "ref2_create"):  # 2-D wrappers (older prebuilt libraries lack them)
        L.ref2_create.argtypes = [_f32p, C.c_int, _i32p, C.c_int]
        L.ref2_create.restype = vp
        L.ref2_destroy.argtypes = [vp]
        for name in ("ref2_num_objects", "ref2_num_nodes"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = C.c_int
        L.ref2_export_tree.argtypes = [vp, _u32p, _f32p, _f32p]
        L.ref2_export_adjacency.argtypes = [vp, _i32p, _i32p]
        L.ref2_closest.argtypes = [vp, _f32p, C.c_long, _u32p, _f32p, C.c_int]
        L.ref2_closest.restype = C.c_double
        L.ref2_silhouette.argtypes = [vp, _f32p, C.c_long, C.c_int, _f32p, C.c_int]
        L.ref2_silhouette.restype = C.c_double
        L.ref2_ray.argtypes = [vp, _f32p, _f32p, _f32p, C.c_long, _i32p, _f32p, _f32p, _u32p, C.c_int]
        L.ref2_ray.restype = C.c_double
        L.ref2_sample.argtypes = [vp, _f32p, _f32p, C.c_long, _i32p, _f32p, C.c_int]
        L.ref2_sample.restype = C.c_double
    _ref_libs[kind] = L
    return L


class RefScene:
    """The UNMODIFIED reference headers: kind='cpu' (Thrust CPP backend) or kind='cuda' (nvcc sm_100a)."""

    def __init__(self, verts, tris, kind="cpu"):
        self.kind = kind
        self.L = _ref_lib(kind)
        self.verts = _f32(verts).reshape(-1, 3)
        self.tris = _i32(tris).reshape(-1, 3)
        self.h = self.L.ref3_create(self.verts, len(self.verts), self.tris, len(self.tris))
        self.n = len(self.tris)
        self.num_nodes = self.L.ref3_num_nodes(self.h)
        self.num_edges = self.L.ref3_num_edges(self.h)
        self.last_ms = 0.0

    def close(self):
        if self.h:
            self.L.ref3_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def timings(self):
        out = (C.c_double * 3)()
        self.L.ref3_timings(self.h, out)
        return {"silhouettes_ms": out[0], "build_bvh_ms": out[1], "construct_ms": out[2]}

    def time_construct(self, reps=3):
        return self.L.ref3_time_construct(self.h, reps)

    def tree(self):
        nn = self.num_nodes
        nodes = np.zeros((nn, 4), np.uint32)
        aabbs = np.zeros((nn, 6), np.float32)
        cones = np.zeros((nn, 5), np.float32)
        self.L.ref3_export_tree(self.h, nodes, aabbs, cones)
        return nodes, aabbs, cones

    def adjacency(self):
        e = np.zeros((self.num_edges, 4), np.int32)
        te = np.zeros((self.n, 3), np.int32)
        to = np.zeros((self.n, 3), np.int32)
        self.L.ref3_export_adjacency(self.h, e, te, to)
        return e, te, to

    def morton(self):
        m = np.zeros(self.n, np.uint32)
        self.L.ref3_export_morton(self.h, m)
        nodes, _, _ = self.tree()
        return m, nodes[self.n - 1:, 3].copy()

    def closest(self, q, nthreads=1):
        q = _f32(q).reshape(-1, 3)
        idx = np.zeros(len(q), np.uint32)
        dist = np.zeros(len(q), np.float32)
        self.last_ms = self.L.ref3_closest(self.h, q, len(q), idx, dist, nthreads)
        return idx, dist

    def silhouette(self, q, flip=False, nthreads=1):
        q = _f32(q).reshape(-1, 3)
        dist = np.zeros(len(q), np.float32)
        self.last_ms = self.L.ref3_silhouette(self.h, q, len(q), int(flip), dist, nthreads)
        return dist

    def ray(self, org, dirs, tmax=None, nthreads=1):
        org = _f32(org).reshape(-1, 3)
        dirs = _f32(dirs).reshape(-1, 3)
        n = len(org)
        tm = np.full(n, np.inf, np.float32) if tmax is None else _f32(tmax)
        found = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        uv = np.zeros((n, 2), np.float32)
        prim = np.zeros(n, np.uint32)
        self.last_ms = self.L.ref3_ray(self.h, org, dirs, tm, n, found, t, uv, prim, nthreads)
        return found, t, uv, prim

    def sample(self, sph, u, nthreads=1):
        sph = _f32(sph).reshape(-1, 4)
        u = _f32(u)
        idx = np.zeros(len(sph), np.int32)
        pdf = np.zeros(len(sph), np.float32)
        self.last_ms = self.L.ref3_sample(self.h, sph, u, len(sph), idx, pdf, nthreads)
        return idx, pdf


_fcpw_lib = None


class RefScene2:
    """lbvh::scene<2> of the UNMODIFIED reference headers (polylines: segments / silhouette vertices), kind 'cpu' or 'cuda'."""

    def __init__(self, verts, segs, kind="cpu"):
        self.kind = kind
        self.L = _ref_lib(kind)
        if not hasattr(self.L, "ref2_create"):
            raise RuntimeError("oracle/_ref was built without the 2-D wrappers: rebuild it (make -C oracle ref)")
        self.verts = _f32(verts).reshape(-1, 2)
        self.segs = _i32(segs).reshape(-1, 2)
        self.h = self.L.ref2_create(self.verts, len(self.verts), self.segs, len(self.segs))
        self.n = len(self.segs)
        self.num_nodes = self.L.ref2_num_nodes(self.h)

    def close(self):
        if self.h:
            self.L.ref2_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tree(self):
        nn = self.num_nodes
        nodes = np.zeros((nn, 4), np.uint32)
        aabbs = np.zeros((nn, 4), np.float32)
        cones = np.zeros((nn, 4), np.float32)
        self.L.ref2_export_tree(self.h, nodes, aabbs, cones)
        return nodes, aabbs, cones

    def adjacency(self):
        v4 = np.zeros((len(self.verts), 4), np.int32)
        owned = np.zeros((self.n, 2), np.int32)
        self.L.ref2_export_adjacency(self.h, v4, owned)
        return v4, owned

    def closest(self, q, nthreads=1):
        q = _f32(q).reshape(-1, 2)
        idx = np.zeros(len(q), np.uint32)
        dist = np.zeros(len(q), np.float32)
        self.L.ref2_closest(self.h, q, len(q), idx, dist, nthreads)
        return idx, dist

    def silhouette(self, q, flip=False, nthreads=1):
        q = _f32(q).reshape(-1, 2)
        dist = np.zeros(len(q), np.float32)
        self.L.ref2_silhouette(self.h, q, len(q), int(flip), dist, nthreads)
        return dist

    def ray(self, org, dirs, tmax=None, nthreads=1):
        org = _f32(org).reshape(-1, 2)
        dirs = _f32(dirs).reshape(-1, 2)
        n = len(org)
        tm = np.full(n, np.inf, np.float32) if tmax is None else _f32(np.broadcast_to(tmax, (n,)))
        found = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        s = np.zeros(n, np.float32)
        prim = np.zeros(n, np.uint32)
        self.L.ref2_ray(self.h, org, dirs, tm, n, found, t, s, prim, nthreads)
        return found, t, s, prim

    def sample(self, sph, u, nthreads=1):
        sph = _f32(sph).reshape(-1, 3)
        u = _f32(u).reshape(-1)
        idx = np.zeros(len(sph), np.int32)
        pdf = np.zeros(len(sph), np.float32)
        self.L.ref2_sample(self.h, sph, u, len(sph), idx, pdf, nthreads)
        return idx, pdf


def _fcpw():
    global _fcpw_lib
    if _fcpw_lib is None:
        L = C.CDLL(os.path.join(_REF_DIR, "libfcpw_cpu.so"))
        vp = C.c_void_p
        L.fcpw3_create.restype = vp
        L.fcpw3_create.argtypes = [_f32p, C.c_int, _i32p, C.c_int, C.c_int]
        L.fcpw3_destroy.argtypes = [vp]
        L.fcpw3_build_ms.argtypes = [vp]
        L.fcpw3_build_ms.restype = C.c_double
        L.fcpw3_threads.restype = C.c_int
        L.fcpw3_closest.argtypes = [vp, _f32p, C.c_long, _f32p, _i32p]
        L.fcpw3_closest.restype = C.c_double
        L.fcpw3_silhouette.argtypes = [vp, _f32p, C.c_void_p, C.c_long, C.c_int, _f32p]
        L.fcpw3_silhouette.restype = C.c_double
        L.fcpw3_ray.argtypes = [vp, _f32p, _f32p, _f32p, C.c_long, _i32p, _f32p, _i32p]
        L.fcpw3_ray.restype = C.c_double
        _fcpw_lib = L
    return _fcpw_lib


class FcpwScene:
    """fcpw's CPU backend (bundled with the reference under ext/fcpw): the reported CPU baseline."""

    def __init__(self, verts, tris, vectorize=True):
        self.L = _fcpw()
        self.verts = _f32(verts).reshape(-1, 3)
        self.tris = _i32(tris).reshape(-1, 3)
        self.h = self.L.fcpw3_create(self.verts, len(self.verts), self.tris, len(self.tris), int(vectorize))
        self.build_ms = self.L.fcpw3_build_ms(self.h)
        self.threads = self.L.fcpw3_threads()
        self.last_ms = 0.0

    def close(self):
        if self.h:
            self.L.fcpw3_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def closest(self, q):
        q = _f32(q).reshape(-1, 3)
        d = np.zeros(len(q), np.float32)
        idx = np.zeros(len(q), np.int32)
        self.last_ms = self.L.fcpw3_closest(self.h, q, len(q), d, idx)
        return idx, d

    def silhouette(self, q, r_max=None, flip=False):
        q = _f32(q).reshape(-1, 3)
        d = np.zeros(len(q), np.float32)
        rm = _f32(r_max) if r_max is not None else None
        self.last_ms = self.L.fcpw3_silhouette(self.h, q, rm.ctypes.data if rm is not None else None, len(q), int(flip), d)
        return d

    def ray(self, org, dirs, tmax=None):
        org = _f32(org).reshape(-1, 3)
        dirs = _f32(dirs).reshape(-1, 3)
        n = len(org)
        tm = np.full(n, np.inf, np.float32) if tmax is None else _f32(tmax)
        found = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        prim = np.zeros(n, np.int32)
        self.last_ms = self.L.fcpw3_ray(self.h, org, dirs, tm, n, found, t, prim)
        return found, t, prim
