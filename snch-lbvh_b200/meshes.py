"""Deterministic synthetic meshes and query sets for the BASELINE.json configs.

Everything is generated in float64 with numpy and rounded once to float32, so the same arrays are fed to the
CUDA path, the C oracle and the reference builds.  No file I/O, no network.
"""
from __future__ import annotations

import numpy as np


def icosphere(level: int = 5, radius: float = 1.0):
    """Unit icosphere, outward CCW faces.  level 5 -> 20 480 triangles / 10 242 vertices (config C1)."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array(
        [[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
         [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array(
        [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
         [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10],
         [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(level):
        nv = len(v)
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        es = np.sort(e, axis=1)
        key = es[:, 0] * nv + es[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        a, b = uniq // nv, uniq % nv
        mid = v[a] + v[b]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        v = np.concatenate([v, mid], axis=0)
        m = inv.reshape(3, -1).T + nv  # midpoints of edges (01, 12, 20) per face
        f = np.concatenate(
            [np.stack([f[:, 0], m[:, 0], m[:, 2]], 1), np.stack([f[:, 1], m[:, 1], m[:, 0]], 1),
             np.stack([f[:, 2], m[:, 2], m[:, 1]], 1), np.stack([m[:, 0], m[:, 1], m[:, 2]], 1)], axis=0)
    return (v * radius).astype(np.float32), f.astype(np.int32)


def bumpy_torus(nu: int = 708, nv: int = 708, R: float = 1.0, r0: float = 0.4, bump: float = 0.15):
    """Closed, consistently oriented 'bumpy torus' grid: 2*nu*nv triangles, nu*nv vertices.

    nu=nv=708 -> 1 002 528 triangles (config C2/C3); 1416 -> 4 010 112 (C4); 2240 -> 10 035 200 (C5).
    """
    i = np.arange(nu, dtype=np.float64)
    j = np.arange(nv, dtype=np.float64)
    u = (2.0 * np.pi / nu) * i[:, None]
    w = (2.0 * np.pi / nv) * j[None, :]
    r = r0 * (1.0 + bump * np.sin(5.0 * u) * np.sin(7.0 * w))
    x = (R + r * np.cos(w)) * np.cos(u)
    y = (R + r * np.cos(w)) * np.sin(u)
    z = r * np.sin(w)
    verts = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float32)
    ii, jj = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    i1 = (ii + 1) % nu
    j1 = (jj + 1) % nv
    a = ii * nv + jj
    b = i1 * nv + jj
    c = i1 * nv + j1
    d = ii * nv + j1
    tris = np.stack([np.stack([a, b, c], -1), np.stack([a, c, d], -1)], axis=2).reshape(-1, 3)
    return verts, tris.astype(np.int32)


def tetrahedron():
    """4-triangle closed mesh used by the known-answer tests."""
    v = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]], dtype=np.float32)
    f = np.array([[0, 1, 2], [0, 3, 1], [0, 2, 3], [1, 3, 2]], dtype=np.int32)
    return v, f


def open_grid(n: int = 8):
    """Open (boundary-carrying) height-field patch: exercises boundary silhouette edges (half_angle = pi leaves)."""
    g = np.linspace(-1.0, 1.0, n + 1)
    x, y = np.meshgrid(g, g, indexing="ij")
    z = 0.2 * np.sin(3.0 * x) * np.cos(2.0 * y)
    v = np.stack([x, y, z], -1).reshape(-1, 3).astype(np.float32)
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    a = ii * (n + 1) + jj
    b = (ii + 1) * (n + 1) + jj
    c = (ii + 1) * (n + 1) + jj + 1
    d = ii * (n + 1) + jj + 1
    f = np.stack([np.stack([a, b, c], -1), np.stack([a, c, d], -1)], axis=2).reshape(-1, 3)
    return v, f.astype(np.int32)


def mesh_bounds(verts: np.ndarray):
    return verts.min(axis=0).astype(np.float64), verts.max(axis=0).astype(np.float64)


def points_in_box(n: int, lo, hi, scale: float = 1.1, seed: int = 2025) -> np.ndarray:
    """n points uniform in the box [lo,hi] scaled by `scale` about its centre (float32, shape (n,3))."""
    rng = np.random.default_rng(seed)
    lo = np.asarray(lo, np.float64)
    hi = np.asarray(hi, np.float64)
    c = 0.5 * (lo + hi)
    h = 0.5 * (hi - lo) * scale
    p = rng.random((n, 3), dtype=np.float32).astype(np.float64)
    return (c - h + 2.0 * h * p).astype(np.float32)


def unit_directions(n: int, seed: int = 77) -> np.ndarray:
    """n directions uniform on the unit sphere (float32)."""
    rng = np.random.default_rng(seed)
    d = rng.standard_normal((n, 3))
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)
    return d.astype(np.float32)


def uniforms(n: int, k: int = 1, seed: int = 99) -> np.ndarray:
    rng = np.random.default_rng(seed)
    u = rng.random((n, k), dtype=np.float32)
    return u if k > 1 else u[:, 0].copy()


def star_radius_scale(n: int, seed: int = 4242) -> np.ndarray:
    """s ~ U[0.5, 4): WoSt star radius r_max = s * closest-point distance (config C3)."""
    rng = np.random.default_rng(seed)
    return (0.5 + 3.5 * rng.random(n, dtype=np.float32)).astype(np.float32)


# ---- 2-D polylines (lbvh::scene<2>: vertices (n, 2), segments (m, 2)) --------------------------------------------------
def wavy_circle(n: int = 4096, lobes: int = 7, amp: float = 0.25):
    """Closed, consistently oriented polyline r(t) = 1 + amp sin(lobes t): every vertex has two adjacent segments."""
    t = np.linspace(0.0, 2.0 * np.pi, n, endpoint=False)
    r = 1.0 + amp * np.sin(lobes * t)
    v = np.stack([r * np.cos(t), r * np.sin(t)], -1).astype(np.float32)
    s = np.stack([np.arange(n), (np.arange(n) + 1) % n], -1).astype(np.int32)
    return v, s


def open_polyline(n: int = 449):
    """Open zig-zag with two free ends (boundary silhouette vertices: leaf cones of half-angle pi)."""
    x = np.linspace(-1.0, 1.0, n + 1)
    y = 0.3 * np.sin(6.0 * x) + 0.05 * np.where(np.arange(n + 1) % 2 == 0, 1.0, -1.0)
    v = np.stack([x, y], -1).astype(np.float32)
    s = np.stack([np.arange(n), np.arange(n) + 1], -1).astype(np.int32)
    return v, s


def polyline_soup(seed: int = 3, loops: int = 5, n: int = 200):
    """Several closed loops and open strands in one scene, segments shuffled (ownership depends on input order) and a
    third of them reversed (inconsistent orientation: later segments overwrite a vertex's previous / next slots)."""
    rng = np.random.default_rng(seed)
    vs, ss, base = [], [], 0
    for k in range(loops):
        c = rng.uniform(-1.0, 1.0, 2)
        m = n + 17 * k
        t = np.linspace(0.0, 2.0 * np.pi, m, endpoint=False)
        r = 0.2 + 0.1 * rng.random() + 0.04 * np.sin((3 + k) * t)
        vs.append(np.stack([c[0] + r * np.cos(t), c[1] + r * np.sin(t)], -1))
        idx = np.arange(m)
        seg = np.stack([idx, (idx + 1) % m], -1)
        if k % 2 == 1:
            seg = seg[:-7]  # open strand
        ss.append(seg + base)
        base += m
    v = np.concatenate(vs).astype(np.float32)
    s = np.concatenate(ss).astype(np.int32)
    flip = rng.random(len(s)) < 0.33
    s[flip] = s[flip][:, ::-1]
    return v, s[rng.permutation(len(s))].copy()


def points_in_box2(n: int, lo, hi, scale: float = 1.1, seed: int = 2025) -> np.ndarray:
    rng = np.random.default_rng(seed)
    lo = np.asarray(lo, np.float64)
    hi = np.asarray(hi, np.float64)
    c = 0.5 * (lo + hi)
    h = 0.5 * (hi - lo) * scale
    p = rng.random((n, 2), dtype=np.float32).astype(np.float64)
    return (c - h + 2.0 * h * p).astype(np.float32)


def unit_directions2(n: int, seed: int = 77) -> np.ndarray:
    rng = np.random.default_rng(seed)
    a = rng.random(n) * 2.0 * np.pi
    return np.stack([np.cos(a), np.sin(a)], -1).astype(np.float32)
