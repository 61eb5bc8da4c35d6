"""CPU simulation (oracle tree, numpy): how many nodes / leaves an exact closest-point walk must open if every node of up to K
leaves also carried an oriented slab (unit mean normal, min / max of n.v over its vertices) next to its box.  Result at 1M
triangles, 150 uniform queries: V 264 -> 201, L 114 -> 90 even with K unbounded: the candidates are true near-ties, not loose
boxes, so the slab records were not built.   Usage: python tools/slab_sim.py [nu=708] [queries=150]"""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import snch_lbvh_b200 as pkg
from oracle import OracleScene
m = pkg.meshes
nu = int(sys.argv[1]) if len(sys.argv) > 1 else 708
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 150
v, f = m.bumpy_torus(nu, nu)
t0 = time.time()
o = OracleScene(v, f)
print('oracle build', time.time() - t0, flush=True)
nodes, aabbs, cones = o.tree()   # aabbs layout? check
ranges = o.ranges()
mort, sidx = o.morton()
n = len(f); ni = n - 1
lo, hi = m.mesh_bounds(v)
q = m.points_in_box(NS, lo, hi, 1.0, seed=2025)
_, dcp = o.closest(q, nthreads=8)
# leaf k -> triangle sidx[k]
tv = v[f[sidx]].astype(np.float64)   # (n,3,3)
e1 = tv[:,1]-tv[:,0]; e2 = tv[:,2]-tv[:,0]
tn = np.cross(e1, e2); ln = np.linalg.norm(tn, axis=1); ok = ln > 0
tn[ok] /= ln[ok,None]; tn[~ok] = 0
# figure out aabb layout: reference RefAabb stores upper then lower
A = aabbs.astype(np.float64)
up, lw = A[:, :3], A[:, 3:]
if not np.all(up >= lw): up, lw = lw, up
size = np.ones(2*n-1, np.int64); size[:ni] = ranges[:,1].astype(np.int64) - ranges[:,0] + 1
first = np.zeros(2*n-1, np.int64); first[:ni] = ranges[:,0]; first[ni:] = np.arange(n)
# prefix sums of normals for fast range normal
cn = np.concatenate([np.zeros((1,3)), np.cumsum(tn, axis=0)])
slab_cache = {}
def slab(node, K):
    s = size[node]
    if s > K or node >= ni: return None
    key = node
    if key in slab_cache: return slab_cache[key]
    a = first[node]; b = a + s
    nn = cn[b] - cn[a]; l = np.linalg.norm(nn)
    if l < 1e-12: r = None
    else:
        nn /= l
        pr = tv[a:b].reshape(-1,3) @ nn
        r = (nn, pr.min(), pr.max())
    slab_cache[key] = r
    return r
def boxd2(node, p):
    d = np.maximum(np.maximum(lw[node]-p, p-up[node]), 0.0)
    return float(d @ d)
res = {}
for K in (0, 8, 32, 128, 1024, 1<<30):
    slab_cache.clear()
    V = L = 0
    for qi in range(NS):
        p = q[qi].astype(np.float64); d2 = float(dcp[qi])**2 * (1+1e-6)
        st = [0]
        while st:
            nd = st.pop()
            V += 1
            for ch in (nodes[nd,1], nodes[nd,2]):
                ch = int(ch)
                if boxd2(ch, p) >= d2: continue
                if K:
                    sl = slab(ch, K)
                    if sl is not None:
                        t = float(sl[0] @ p); ds = max(sl[1]-t, t-sl[2], 0.0)
                        if ds*ds >= d2: continue
                if ch >= ni: L += 1
                else: st.append(ch)
    res[K] = (V/NS, L/NS)
    print('K', K, 'V', V/NS, 'L', L/NS, flush=True)
