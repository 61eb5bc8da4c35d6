/* TEST INFRASTRUCTURE — CPU oracle for the SNCH-LBVH hot path.
 *
 * A plain-C restatement of the reference's algorithm (tyanyuy3125/snch-lbvh @ /root/reference).  Every function
 * names the reference file:line it follows.  It is deliberately scalar, serial per query and keeps the reference's
 * visiting order and tie rules so that it can be compared bit-for-bit with the reference executed on the CPU
 * (oracle/_ref/libsnch_ref_cpu.so) — that comparison is what pins it (tests/test_oracle_pinning.py).
 *
 * Built with -ffp-contract=off: the host reference build has no FMA contraction either.
 *
 * Deliberate, documented deviation: cone merge leaves half_angle uninitialised in the reference
 * when the union exceeds pi (cone.cuh:454-459); the oracle (and the product) define it as pi, which is what fcpw's
 * original does, and records the affected nodes in q1_taint so tests can mask them.
 *
 * Nothing on the product path includes, links or calls this file.
 */
#define _GNU_SOURCE
#include "snch_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI_F ((float)M_PI)
#define ORC_PI_2_F ((float)M_PI_2)
#define LEAF_NONE 0xFFFFFFFFu

typedef struct { float x, y, z; } f3;

/* std::min / std::max semantics (NaN handling differs from fminf/fmaxf) */
static inline float std_min(float a, float b) { return (b < a) ? b : a; }
static inline float std_max(float a, float b) { return (a < b) ? b : a; }

static inline f3 mk3(float x, float y, float z) { f3 r = {x, y, z}; return r; }
static inline f3 sub3(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
/* utility.cuh:238 */
static inline float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* utility.cuh:382-385 */
static inline float len3(f3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
/* utility.cuh:400-403 */
static inline float sqlen3(f3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
/* utility.cuh:427-431 */
static inline f3 normalize3(f3 v) { float n = len3(v); return mk3(v.x / n, v.y / n, v.z / n); }
/* utility.cuh:449-455 */
static inline f3 cross3(f3 u, f3 v) { return mk3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x); }
static inline f3 abs3(f3 a) { return mk3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }

/* aabb layout follows aabb.cuh:12-15: upper first, then lower */
typedef struct { f3 upper, lower; } box3;
/* cone.cuh:10-16 */
typedef struct { f3 axis; float half_angle, radius; } cone3;
/* bvh.cuh:27-33 */
typedef struct { uint32_t parent, left, right, object; } node_t;

struct orc_scene
{
    int nV, nT, nE;
    f3 *verts;
    int *tris;       /* 3 per triangle */
    int *edges;      /* 4 per edge: silhouette_edge::indices (scene.cuh:711) */
    int *tri_edges;  /* 3 per triangle: edge_indices_h */
    int *tri_owned;  /* 3 per triangle: triangle::silhouette_indices */
    node_t *nodes;   /* 2N-1 */
    box3 *aabbs;
    cone3 *cones;
    uint8_t *q1;
    uint32_t *morton_sorted, *sorted_idx, *ranges;
    int collision;
};

/* ------------------------------------------------------------------------------------------------
 * Morton codes — morton_code.cuh:19-38 (expand_bits), :61-70 (morton_code float3)
 * ---------------------------------------------------------------------------------------------- */
uint32_t orc_expand_bits(uint32_t v)
{
    v = (v | (v << 16)) & 0x070000FFu;
    v = (v | (v << 8)) & 0x0700F00Fu;
    v = (v | (v << 4)) & 0x430C30C3u;
    v = (v | (v << 2)) & 0x49249249u;
    return v;
}
uint32_t orc_morton3(float x, float y, float z)
{
    const float res = 1024.0f;
    x = fminf(fmaxf(x * res, 0.0f), res - 1.0f);
    y = fminf(fmaxf(y * res, 0.0f), res - 1.0f);
    z = fminf(fmaxf(z * res, 0.0f), res - 1.0f);
    const uint32_t xx = orc_expand_bits((uint32_t)x), yy = orc_expand_bits((uint32_t)y), zz = orc_expand_bits((uint32_t)z);
    return xx * 4 + yy * 2 + zz;
}

/* ------------------------------------------------------------------------------------------------
 * AABB algebra — aabb.cuh
 * ---------------------------------------------------------------------------------------------- */
/* aabb.cuh:21-38 (point ctor) + :94-101 (expand_to_include) as used by scene.cuh:870-885 (aabb_getter) */
static box3 tri_box(f3 a, f3 b, f3 c)
{
    const float e = FLT_EPSILON;
    box3 r;
    r.upper = mk3(a.x + e, a.y + e, a.z + e);
    r.lower = mk3(a.x - e, a.y - e, a.z - e);
    const f3 p[2] = {b, c};
    for (int i = 0; i < 2; ++i)
    {
        r.lower = mk3(fminf(r.lower.x, p[i].x - e), fminf(r.lower.y, p[i].y - e), fminf(r.lower.z, p[i].z - e));
        r.upper = mk3(fmaxf(r.upper.x, p[i].x + e), fmaxf(r.upper.y, p[i].y + e), fmaxf(r.upper.z, p[i].z + e));
    }
    return r;
}
/* aabb.cuh:113-124 */
static box3 box_merge(box3 l, box3 r)
{
    box3 m;
    m.upper = mk3(fmaxf(l.upper.x, r.upper.x), fmaxf(l.upper.y, r.upper.y), fmaxf(l.upper.z, r.upper.z));
    m.lower = mk3(fminf(l.lower.x, r.lower.x), fminf(l.lower.y, r.lower.y), fminf(l.lower.z, r.lower.z));
    return m;
}
/* aabb.cuh:271-279 (the 0.5 is a double literal there; multiplying by 0.5 is exact either way) */
static f3 box_centroid(box3 b)
{
    return mk3((float)((b.upper.x + b.lower.x) * 0.5), (float)((b.upper.y + b.lower.y) * 0.5), (float)((b.upper.z + b.lower.z) * 0.5));
}
/* aabb.cuh:144-150 — squared distance point -> box */
static float box_mindist(box3 b, f3 p)
{
    const float dx = fminf(b.upper.x, fmaxf(b.lower.x, p.x)) - p.x;
    const float dy = fminf(b.upper.y, fmaxf(b.lower.y, p.y)) - p.y;
    const float dz = fminf(b.upper.z, fmaxf(b.lower.z, p.z)) - p.z;
    return dx * dx + dy * dy + dz * dz;
}
/* aabb.cuh:182-209 */
static float box_minmaxdist(box3 b, f3 p)
{
    float rmx = (b.lower.x - p.x) * (b.lower.x - p.x), rmy = (b.lower.y - p.y) * (b.lower.y - p.y), rmz = (b.lower.z - p.z) * (b.lower.z - p.z);
    float rMx = (b.upper.x - p.x) * (b.upper.x - p.x), rMy = (b.upper.y - p.y) * (b.upper.y - p.y), rMz = (b.upper.z - p.z) * (b.upper.z - p.z);
    float t;
    if ((b.upper.x + b.lower.x) * 0.5f < p.x) { t = rmx; rmx = rMx; rMx = t; }
    if ((b.upper.y + b.lower.y) * 0.5f < p.y) { t = rmy; rmy = rMy; rMy = t; }
    if ((b.upper.z + b.lower.z) * 0.5f < p.z) { t = rmz; rmz = rMz; rMz = t; }
    const float dx = rmx + rMy + rMz, dy = rMx + rmy + rMz, dz = rMx + rMy + rmz;
    return fminf(dx, fminf(dy, dz));
}
/* aabb.cuh:397-431 — slab test; entry distance clamped to >= 0 */
static int box_ray(f3 org, f3 dinv, box3 b, float max_dist, float *dist)
{
    float t1 = (b.lower.x - org.x) * dinv.x, t2 = (b.upper.x - org.x) * dinv.x;
    float tmin = fminf(t1, t2), tmax = fmaxf(t1, t2);
    t1 = (b.lower.y - org.y) * dinv.y; t2 = (b.upper.y - org.y) * dinv.y;
    tmin = fmaxf(tmin, fminf(t1, t2)); tmax = fminf(tmax, fmaxf(t1, t2));
    t1 = (b.lower.z - org.z) * dinv.z; t2 = (b.upper.z - org.z) * dinv.z;
    tmin = fmaxf(tmin, fminf(t1, t2)); tmax = fminf(tmax, fmaxf(t1, t2));
    if (tmax >= tmin && tmax >= 0.0f && tmin <= max_dist)
    {
        *dist = (tmin >= 0.0f) ? tmin : 0.0f;
        return 1;
    }
    return 0;
}
/* aabb.cuh:433-449 */
static int box_sphere(f3 c, float radius, box3 b)
{
    const float cx = std_max(b.lower.x, std_min(c.x, b.upper.x));
    const float cy = std_max(b.lower.y, std_min(c.y, b.upper.y));
    const float dx = cx - c.x, dy = cy - c.y;
    float d2 = dx * dx + dy * dy;
    const float cz = std_max(b.lower.z, std_min(c.z, b.upper.z));
    const float dz = cz - c.z;
    d2 += dz * dz;
    return d2 <= radius * radius;
}

/* ------------------------------------------------------------------------------------------------
 * Normal-cone algebra — cone.cuh
 * ---------------------------------------------------------------------------------------------- */
static inline int cone_valid(const cone3 *c) { return c->half_angle >= 0.0f; } /* cone.cuh:18-22 */
static inline int inrange(float v, float lo, float hi) { return v >= lo && v <= hi; } /* utility.cuh:241 */

/* cone.cuh:34-42 + :58-66 */
static float project_to_plane(f3 n, f3 e)
{
    const float sign = copysignf(1.0f, n.z);
    const float a = -1.0f / (sign + n.z);
    const float b = n.x * n.y * a;
    const f3 b1 = mk3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
    const f3 b2 = mk3(b, sign + n.y * n.y * a, -n.y);
    const float r1 = dot3(e, abs3(b1)), r2 = dot3(e, abs3(b2));
    return sqrtf(r1 * r1 + r2 * r2);
}
/* cone.cuh:168-212 — can a silhouette seen from `o` exist inside this node? */
static int cone_overlap(const cone3 *bc, f3 o, box3 b, float dist2_to_box)
{
    if (bc->half_angle >= ORC_PI_2_F || dist2_to_box < FLT_EPSILON) return 1;
    const f3 c = box_centroid(b);
    f3 v = sub3(c, o);
    const float l = len3(v);
    v.x /= l; v.y /= l; v.z /= l;
    const float d_axis_angle = acosf(std_max(-1.0f, std_min(1.0f, dot3(bc->axis, v))));
    if (inrange(ORC_PI_2_F, d_axis_angle - bc->half_angle, d_axis_angle + bc->half_angle)) return 1;
    if (l > bc->radius)
    {
        const float view_half = asinf(bc->radius / l);
        const float sum = bc->half_angle + view_half;
        return sum >= ORC_PI_2_F ? 1 : inrange(ORC_PI_2_F, d_axis_angle - sum, d_axis_angle + sum);
    }
    const f3 e = sub3(b.upper, c);
    float d = dot3(e, abs3(v));
    const float s = l - d;
    if (s <= 0.0f) return 1;
    d = project_to_plane(v, e);
    const float view_half = atan2f(d, s);
    const float sum = bc->half_angle + view_half;
    return sum >= ORC_PI_2_F ? 1 : inrange(ORC_PI_2_F, d_axis_angle - sum, d_axis_angle + sum);
}
/* cone.cuh:288-302 — Rodrigues rotation of u towards v by theta */
static f3 rotate3(f3 u, f3 v, float theta)
{
    const float ct = cosf(theta), st = sinf(theta);
    const f3 w = normalize3(cross3(u, v));
    const f3 o = mk3((1.0f - ct) * w.x, (1.0f - ct) * w.y, (1.0f - ct) * w.z);
    const float R[3][3] = {{ct + o.x * w.x, o.y * w.x - st * w.z, o.z * w.x + st * w.y},
                           {o.x * w.y + st * w.z, ct + o.y * w.y, o.z * w.y - st * w.x},
                           {o.x * w.z - st * w.y, o.y * w.z + st * w.x, ct + o.z * w.z}};
    return mk3(R[0][0] * u.x + R[0][1] * u.y + R[0][2] * u.z, R[1][0] * u.x + R[1][1] * u.y + R[1][2] * u.z,
               R[2][0] * u.x + R[2][1] * u.y + R[2][2] * u.z);
}
/* cone.cuh:427-480 — cone union.  *q1 is set where the reference leaves half_angle uninitialised (Q1). */
static cone3 cone_merge(const cone3 *ca, const cone3 *cb, f3 oa, f3 ob, f3 on, int *q1)
{
    cone3 r;
    memset(&r, 0, sizeof r);
    *q1 = 0;
    if (cone_valid(ca) && cone_valid(cb))
    {
        f3 axis_a = ca->axis, axis_b = cb->axis;
        float ha = ca->half_angle, hb = cb->half_angle;
        const f3 da = sub3(on, oa), db = sub3(on, ob);
        r.radius = sqrtf(std_max(ca->radius * ca->radius + sqlen3(da), cb->radius * cb->radius + sqlen3(db)));
        if (hb > ha)
        {
            f3 t = axis_a; axis_a = axis_b; axis_b = t;
            float th = ha; ha = hb; hb = th;
        }
        const float theta = acosf(std_max(-1.0f, std_min(1.0f, dot3(axis_a, axis_b))));
        if (std_min(theta + hb, ORC_PI_F) <= ha)
        {
            r.axis = axis_a;
            r.half_angle = ha;
            return r;
        }
        const float o_theta = (ha + theta + hb) / 2.0f;
        if (o_theta >= ORC_PI_F)
        {
            r.axis = axis_a;
            r.half_angle = ORC_PI_F; /* Q1: indeterminate in the reference; defined as pi here */
            *q1 = 1;
            return r;
        }
        const float r_theta = o_theta - ha;
        r.axis = rotate3(axis_a, axis_b, r_theta);
        r.half_angle = o_theta;
    }
    else if (cone_valid(ca)) r = *ca;
    else if (cone_valid(cb)) r = *cb;
    else r.half_angle = -ORC_PI_F; /* axis / radius are never read for an invalid cone */
    return r;
}

/* ------------------------------------------------------------------------------------------------
 * Silhouette edges — scene.cuh:709-824
 * ---------------------------------------------------------------------------------------------- */
static inline int edge_has_face(const int *e, int f) { return f == 0 ? e[3] != -1 : e[0] != -1; } /* scene.cuh:743-746 */
/* scene.cuh:747-772 */
static f3 edge_face_normal(const orc_scene *s, const int *e, int f, int do_normalize)
{
    int i, j, k;
    if (f == 0) { i = 3; j = 1; k = 2; } else { i = 0; j = 2; k = 1; }
    const f3 pa = s->verts[e[j]], pb = s->verts[e[k]], pc = s->verts[e[i]];
    const f3 n = cross3(sub3(pb, pa), sub3(pc, pa));
    return do_normalize ? normalize3(n) : n;
}
/* scene.cuh:773-787 */
static f3 edge_normal(const orc_scene *s, const int *e)
{
    f3 n = mk3(0.0f, 0.0f, 0.0f);
    if (edge_has_face(e, 0)) { const f3 a = edge_face_normal(s, e, 0, 0); n = mk3(n.x + a.x, n.y + a.y, n.z + a.z); }
    if (edge_has_face(e, 1)) { const f3 a = edge_face_normal(s, e, 1, 0); n = mk3(n.x + a.x, n.y + a.y, n.z + a.z); }
    return normalize3(n);
}
/* scene.cuh:230-255 */
static float closest_point_segment(f3 pa, f3 pb, f3 x, f3 *pt)
{
    const f3 u = sub3(pb, pa), v = sub3(x, pa);
    const float c1 = dot3(u, v);
    if (c1 <= 0.0f) { *pt = pa; return len3(sub3(x, *pt)); }
    const float c2 = dot3(u, u);
    if (c2 <= c1) { *pt = pb; return len3(sub3(x, *pt)); }
    const float t = c1 / c2;
    *pt = mk3(pa.x + u.x * t, pa.y + u.y * t, pa.z + u.z * t);
    return len3(sub3(x, *pt));
}
/* scene.cuh:143-174 — note view_dir is NOT normalised in 3-D (Q2) */
static int is_silhouette_edge(f3 pa, f3 pb, f3 n0, f3 n1, f3 view, float d, int flip)
{
    const float precision = 1e-3f;
    const float sign = flip ? 1.0f : -1.0f;
    if (d <= precision)
    {
        const f3 edge_dir = normalize3(sub3(pb, pa));
        const float dihedral = atan2f(dot3(edge_dir, cross3(n0, n1)), dot3(n0, n1));
        return sign * dihedral > precision;
    }
    const float dot0 = dot3(view, n0), dot1 = dot3(view, n1);
    if (fabsf(dot0) <= precision) return sign * dot1 > precision;
    if (fabsf(dot1) <= precision) return sign * dot0 > precision;
    return dot0 * dot1 < 0.0f;
}
/* scene.cuh:788-824 */
/* `point` (optional): closest_pos, the value the reference computes at :796-799 and does not return */
static int edge_closest_silhouette(const orc_scene *s, const int *e, f3 origin, float max_r2, float *distance, int flip, float min_r2, f3 *point)
{
    if (min_r2 >= max_r2) return 0;
    const f3 pa = s->verts[e[1]], pb = s->verts[e[2]];
    f3 cp;
    const float d = closest_point_segment(pa, pb, origin, &cp);
    if (d * d > max_r2) return 0;
    int is_sil = !edge_has_face(e, 0) || !edge_has_face(e, 1);
    if (!is_sil)
    {
        const f3 n0 = edge_face_normal(s, e, 0, 1), n1 = edge_face_normal(s, e, 1, 1);
        is_sil = is_silhouette_edge(pa, pb, n0, n1, sub3(origin, cp), d, flip);
    }
    if (is_sil && d * d <= max_r2) { *distance = d; if (point) *point = cp; return 1; }
    return 0;
}
/* scene.cuh:978-1003 — silhouette_distance_calculator over the <=3 owned edges of one triangle */
/* `edge` / `point` (optional): id of the owned edge that set *distance and the closest point on it ("TODO: identify nearest
 * index", query.cuh:386,411) */
static int tri_closest_silhouette(const orc_scene *s, int tri, f3 origin, float max_r2, float *distance, int flip, float min_r2, int *edge,
                                  f3 *point)
{
    float detached = max_r2;
    int ret = 0;
    for (int i = 0; i < 3; ++i)
    {
        const int ei = s->tri_owned[3 * tri + i];
        if (ei != -1 && edge_closest_silhouette(s, s->edges + 4 * ei, origin, detached, distance, flip, min_r2, point))
        {
            ret = 1;
            detached = *distance * *distance;
            if (edge) *edge = ei;
        }
    }
    return ret;
}

/* ------------------------------------------------------------------------------------------------
 * Triangle primitives — scene.cuh:34-110, 1005-1126, 14-27
 * ---------------------------------------------------------------------------------------------- */
/* scene.cuh:34-110 (Ericson RTCD 5.1.5) */
static float closest_point_triangle(f3 pa, f3 pb, f3 pc, f3 x, f3 *pt)
{
    const f3 ab = sub3(pb, pa), ac = sub3(pc, pa), ax = sub3(x, pa);
    const float d1 = dot3(ab, ax), d2 = dot3(ac, ax);
    if (d1 <= 0.0f && d2 <= 0.0f) { *pt = pa; return len3(sub3(x, *pt)); }
    const f3 bx = sub3(x, pb);
    const float d3 = dot3(ab, bx), d4 = dot3(ac, bx);
    if (d3 >= 0.0f && d4 <= d3) { *pt = pb; return len3(sub3(x, *pt)); }
    const f3 cx = sub3(x, pc);
    const float d5 = dot3(ab, cx), d6 = dot3(ac, cx);
    if (d6 >= 0.0f && d5 <= d6) { *pt = pc; return len3(sub3(x, *pt)); }
    const float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f)
    {
        const float v = d1 / (d1 - d3);
        *pt = mk3(pa.x + ab.x * v, pa.y + ab.y * v, pa.z + ab.z * v);
        return len3(sub3(x, *pt));
    }
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f)
    {
        const float w = d2 / (d2 - d6);
        *pt = mk3(pa.x + ac.x * w, pa.y + ac.y * w, pa.z + ac.z * w);
        return len3(sub3(x, *pt));
    }
    const float va = d3 * d6 - d5 * d4;
    if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f)
    {
        const float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        *pt = mk3(pb.x + (pc.x - pb.x) * w, pb.y + (pc.y - pb.y) * w, pb.z + (pc.z - pb.z) * w);
        return len3(sub3(x, *pt));
    }
    const float denom = 1.0f / (va + vb + vc);
    const float v = vb * denom, w = vc * denom;
    *pt = mk3(pa.x + ab.x * v + ac.x * w, pa.y + ab.y * v + ac.y * w, pa.z + ab.z * v + ac.z * w);
    return len3(sub3(x, *pt));
}
static inline void tri_verts(const orc_scene *s, int tri, f3 *a, f3 *b, f3 *c)
{
    *a = s->verts[s->tris[3 * tri]]; *b = s->verts[s->tris[3 * tri + 1]]; *c = s->verts[s->tris[3 * tri + 2]];
}
/* scene.cuh:963-976 */
static float tri_distance(const orc_scene *s, int tri, f3 p)
{
    f3 a, b, c, pt;
    tri_verts(s, tri, &a, &b, &c);
    return closest_point_triangle(a, b, c, p, &pt);
}
/* scene.cuh:1005-1052 — Moeller-Trumbore; host 1.0f/det and device __frcp_rn(det) are the same IEEE value */
static int tri_ray(const orc_scene *s, int tri, f3 org, f3 dir, float *t_out, float *u_out, float *v_out)
{
    f3 v0, v1, v2;
    tri_verts(s, tri, &v0, &v1, &v2);
    const f3 e1 = sub3(v1, v0), e2 = sub3(v2, v0);
    const f3 h = mk3(dir.y * e2.z - dir.z * e2.y, dir.z * e2.x - dir.x * e2.z, dir.x * e2.y - dir.y * e2.x);
    const float det = e1.x * h.x + e1.y * h.y + e1.z * h.z;
    if (fabsf(det) < FLT_EPSILON) return 0;
    const float inv_det = 1.0f / det;
    const f3 sv = sub3(org, v0);
    const float u = (sv.x * h.x + sv.y * h.y + sv.z * h.z) * inv_det;
    if (u < 0.0f || u > 1.0f) return 0;
    const f3 q = mk3(sv.y * e1.z - sv.z * e1.y, sv.z * e1.x - sv.x * e1.z, sv.x * e1.y - sv.y * e1.x);
    const float v = (dir.x * q.x + dir.y * q.y + dir.z * q.z) * inv_det;
    if (v < 0.0f || u + v > 1.0f) return 0;
    const float t = (e2.x * q.x + e2.y * q.y + e2.z * q.z) * inv_det;
    if (t >= 0.0f) { *t_out = t; *u_out = u; *v_out = v; return 1; }
    return 0;
}
/* scene.cuh:1054-1117 — sphere/triangle overlap with the vertex-only fallback (Q10) */
static int tri_sphere(const orc_scene *s, int tri, f3 center, float radius)
{
    f3 p1, p2, p3;
    tri_verts(s, tri, &p1, &p2, &p3);
    const f3 e1 = sub3(p2, p1), e2 = sub3(p3, p1);
    f3 n = mk3(e1.y * e2.z - e1.z * e2.y, e1.z * e2.x - e1.x * e2.z, e1.x * e2.y - e1.y * e2.x);
    const float nl = sqrtf(n.x * n.x + n.y * n.y + n.z * n.z);
    n = mk3(n.x / nl, n.y / nl, n.z / nl);
    const float d = n.x * p1.x + n.y * p1.y + n.z * p1.z;
    const float dist_to_plane = n.x * center.x + n.y * center.y + n.z * center.z - d;
    const f3 proj = mk3(center.x - dist_to_plane * n.x, center.y - dist_to_plane * n.y, center.z - dist_to_plane * n.z);
    const f3 v0 = sub3(p3, p1), v1 = sub3(p2, p1), v2 = sub3(proj, p1);
    const float dot00 = v0.x * v0.x + v0.y * v0.y + v0.z * v0.z;
    const float dot01 = v0.x * v1.x + v0.y * v1.y + v0.z * v1.z;
    const float dot02 = v0.x * v2.x + v0.y * v2.y + v0.z * v2.z;
    const float dot11 = v1.x * v1.x + v1.y * v1.y + v1.z * v1.z;
    const float dot12 = v1.x * v2.x + v1.y * v2.y + v1.z * v2.z;
    const float inv = 1.0f / (dot00 * dot11 - dot01 * dot01);
    const float u = (dot11 * dot02 - dot01 * dot12) * inv;
    const float v = (dot00 * dot12 - dot01 * dot02) * inv;
    if (u >= 0 && v >= 0 && u + v <= 1) return fabsf(dist_to_plane) <= radius;
    f3 cp = proj;
    if (u < 0) cp = p1;
    else if (v < 0) cp = p3;
    else if (u + v > 1) cp = p2;
    const float dx = cp.x - center.x, dy = cp.y - center.y, dz = cp.z - center.z;
    return dx * dx + dy * dy + dz * dz <= radius * radius;
}
/* scene.cuh:855-868 */
static float tri_area(const orc_scene *s, int tri)
{
    f3 a, b, c;
    tri_verts(s, tri, &a, &b, &c);
    return len3(cross3(sub3(c, a), sub3(b, a))) / 2;
}
/* scene.cuh:1119-1126 */
static float green_weight(f3 x, f3 y)
{
    const float r = std_max(len3(sub3(x, y)), 1e-4f);
    return 1.0f / (3.14159265358979323846f * 4.0f * r);
}

/* ------------------------------------------------------------------------------------------------
 * Scene preparation (host side of the reference) — scene.cuh:1135-1229
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int lo, hi, pos; } half_edge;
static int cmp_half_edge(const void *a, const void *b)
{
    const half_edge *x = a, *y = b;
    if (x->lo != y->lo) return x->lo < y->lo ? -1 : 1;
    if (x->hi != y->hi) return x->hi < y->hi ? -1 : 1;
    return x->pos < y->pos ? -1 : (x->pos > y->pos);
}
typedef struct { int first_pos, group; } edge_first;
static int cmp_edge_first(const void *a, const void *b)
{
    const edge_first *x = a, *y = b;
    return x->first_pos < y->first_pos ? -1 : (x->first_pos > y->first_pos);
}
/* scene.cuh:1135-1166: edge ids in first-seen order (the reference walks a std::map; a sort gives the same ids) */
static void assign_edge_indices(orc_scene *s)
{
    const int n = s->nT, H = 3 * n;
    half_edge *he = malloc(sizeof(half_edge) * (size_t)(H > 0 ? H : 1));
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < 3; ++j)
        {
            int I = s->tris[3 * i + j], J = s->tris[3 * i + (j + 1) % 3];
            if (I > J) { int t = I; I = J; J = t; }
            he[3 * i + j].lo = I; he[3 * i + j].hi = J; he[3 * i + j].pos = 3 * i + j;
        }
    qsort(he, (size_t)H, sizeof(half_edge), cmp_half_edge);
    int *group_of = malloc(sizeof(int) * (size_t)(H > 0 ? H : 1));
    edge_first *ef = malloc(sizeof(edge_first) * (size_t)(H > 0 ? H : 1));
    int g = 0;
    for (int k = 0; k < H; ++k)
    {
        if (k == 0 || he[k].lo != he[k - 1].lo || he[k].hi != he[k - 1].hi)
        {
            ef[g].first_pos = he[k].pos; /* sorted by pos inside a group, so this is the first occurrence */
            ef[g].group = g;
            ++g;
        }
        group_of[he[k].pos] = g - 1;
    }
    s->nE = g;
    qsort(ef, (size_t)g, sizeof(edge_first), cmp_edge_first);
    int *id_of_group = malloc(sizeof(int) * (size_t)(g > 0 ? g : 1));
    for (int r = 0; r < g; ++r) id_of_group[ef[r].group] = r;
    s->tri_edges = malloc(sizeof(int) * (size_t)(H > 0 ? H : 1));
    for (int p = 0; p < H; ++p) s->tri_edges[p] = id_of_group[group_of[p]];
    free(he); free(group_of); free(ef); free(id_of_group);
}
/* scene.cuh:1167-1204 */
static void compute_silhouettes(orc_scene *s)
{
    assign_edge_indices(s);
    s->edges = malloc(sizeof(int) * 4 * (size_t)(s->nE > 0 ? s->nE : 1));
    for (int e = 0; e < 4 * s->nE; ++e) s->edges[e] = -1;
    for (int i = 0; i < s->nT; ++i)
    {
        const int *vi = s->tris + 3 * i;
        for (int j = 0; j < 3; ++j)
        {
            const int I = j - 1 < 0 ? 2 : j - 1;
            int J = j, K = j + 1 > 2 ? 0 : j + 1;
            int orientation = 1;
            if (vi[J] > vi[K]) { int t = J; J = K; K = t; orientation = -1; }
            int *se = s->edges + 4 * s->tri_edges[3 * i + j];
            se[orientation == 1 ? 0 : 3] = vi[I];
            se[1] = vi[J];
            se[2] = vi[K];
        }
    }
}
/* scene.cuh:1205-1225: an edge is owned by the first triangle (input order) that references it (Q17) */
static void assign_ownership(orc_scene *s)
{
    uint8_t *seen = calloc((size_t)(s->nE > 0 ? s->nE : 1), 1);
    s->tri_owned = malloc(sizeof(int) * 3 * (size_t)(s->nT > 0 ? s->nT : 1));
    for (int i = 0; i < s->nT; ++i)
    {
        int *o = s->tri_owned + 3 * i, p = 0;
        o[0] = o[1] = o[2] = -1;
        for (int j = 0; j < 3; ++j)
        {
            const int e = s->tri_edges[3 * i + j];
            if (!seen[e]) { seen[e] = 1; o[p++] = e; }
        }
    }
    free(seen);
}

/* scene.cuh:887-961 — leaf normal cone from the owned edges */
static cone3 tri_cone(const orc_scene *s, int tri)
{
    f3 a, b, c;
    tri_verts(s, tri, &a, &b, &c);
    const f3 bc = box_centroid(tri_box(a, b, c));
    cone3 r;
    r.axis = mk3(0.0f, 0.0f, 0.0f);
    r.half_angle = ORC_PI_F;
    r.radius = 0.0f;
    int any = 0, all_two = 1;
    for (int i = 0; i < 3; ++i)
    {
        const int ei = s->tri_owned[3 * tri + i];
        if (ei == -1) continue;
        const int *e = s->edges + 4 * ei;
        const f3 n = edge_normal(s, e);
        r.axis.x += n.x; r.axis.y += n.y; r.axis.z += n.z;
        const f3 pa = s->verts[e[1]], pb = s->verts[e[2]];
        const f3 ec = mk3((pa.x + pb.x) / 2, (pa.y + pb.y) / 2, (pa.z + pb.z) / 2); /* scene.cuh:736-742 */
        r.radius = std_max(r.radius, len3(sub3(ec, bc)));
        all_two = all_two && edge_has_face(e, 0) && edge_has_face(e, 1);
        any = 1;
    }
    if (!any) r.half_angle = -ORC_PI_F;
    else if (!all_two) r.half_angle = ORC_PI_F;
    else
    {
        const float an = len3(r.axis);
        if (an > FLT_EPSILON)
        {
            r.axis.x /= an; r.axis.y /= an; r.axis.z /= an;
            r.half_angle = 0.0f;
            for (int i = 0; i < 3; ++i)
            {
                const int ei = s->tri_owned[3 * tri + i];
                if (ei == -1) continue;
                const int *e = s->edges + 4 * ei;
                for (int f = 0; f < 2; ++f)
                {
                    const f3 n = edge_has_face(e, f) ? edge_face_normal(s, e, f, 1) : mk3(0.0f, 0.0f, 0.0f);
                    const float ang = acosf(std_max(-1.0f, std_min(1.0f, dot3(r.axis, n))));
                    r.half_angle = std_max(r.half_angle, ang);
                }
            }
        }
    }
    return r;
}

/* ------------------------------------------------------------------------------------------------
 * construct() — bvh.cuh:380-613
 * ---------------------------------------------------------------------------------------------- */
static inline int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline int clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }

typedef struct { const uint32_t *k32; const uint64_t *k64; int n; } keyview;
/* morton_code.cuh:144-167 with the out-of-range = -1 convention of bvh.cuh:128-141 */
static inline int kdelta(const keyview *kv, int i, int j)
{
    if (j < 0 || j >= kv->n) return -1;
    return kv->k64 ? clz64(kv->k64[i] ^ kv->k64[j]) : clz32(kv->k32[i] ^ kv->k32[j]);
}
/* bvh.cuh:109-167 */
static void determine_range(const keyview *kv, int idx, int *first, int *last)
{
    if (idx == 0) { *first = 0; *last = kv->n - 1; return; }
    const int L = kdelta(kv, idx, idx - 1), R = kdelta(kv, idx, idx + 1);
    const int d = (R > L) ? 1 : -1;
    const int dmin = L < R ? L : R;
    int lmax = 2;
    while (kdelta(kv, idx, idx + d * lmax) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t > 0; t >>= 1)
        if (kdelta(kv, idx, idx + (l + t) * d) > dmin) l += t;
    int j = idx + l * d;
    if (d < 0) { int t = idx; idx = j; j = t; }
    *first = idx; *last = j;
}
/* bvh.cuh:169-199 */
static int find_split(const keyview *kv, int first, int last)
{
    const int same = kv->k64 ? (kv->k64[first] == kv->k64[last]) : (kv->k32[first] == kv->k32[last]);
    if (same) return (first + last) >> 1;
    const int dnode = kdelta(kv, first, last);
    int split = first, stride = last - first;
    do
    {
        stride = (stride + 1) >> 1;
        const int mid = split + stride;
        if (mid < last && kdelta(kv, first, mid) > dnode) split = mid;
    } while (stride > 1);
    return split;
}
typedef struct { uint64_t key; } sortrec;
static int cmp_u64(const void *a, const void *b)
{
    const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : (x > y);
}

static void construct(orc_scene *s)
{
    const int n = s->nT;
    if (n <= 0) return; /* bvh.cuh:383-386 */
    const int ni = n - 1, nn = 2 * n - 1;
    s->nodes = malloc(sizeof(node_t) * (size_t)nn);
    s->aabbs = malloc(sizeof(box3) * (size_t)nn);
    s->cones = malloc(sizeof(cone3) * (size_t)nn);
    s->q1 = calloc((size_t)nn, 1);
    s->morton_sorted = malloc(sizeof(uint32_t) * (size_t)n);
    s->sorted_idx = malloc(sizeof(uint32_t) * (size_t)n);
    s->ranges = malloc(sizeof(uint32_t) * 2 * (size_t)(ni > 0 ? ni : 1));

    /* leaf boxes in object order + scene box (bvh.cuh:423-432) */
    box3 *leaf = malloc(sizeof(box3) * (size_t)n);
    box3 whole;
    whole.upper = mk3(-INFINITY, -INFINITY, -INFINITY);
    whole.lower = mk3(INFINITY, INFINITY, INFINITY);
    for (int i = 0; i < n; ++i)
    {
        f3 a, b, c;
        tri_verts(s, i, &a, &b, &c);
        leaf[i] = tri_box(a, b, c);
        whole = box_merge(whole, leaf[i]);
    }
    /* Morton codes (bvh.cuh:281-304, 436-439) and stable sort by code (bvh.cuh:445-454):
       sorting (code<<32 | index) is the stable order */
    uint64_t *keys = malloc(sizeof(uint64_t) * (size_t)n);
    for (int i = 0; i < n; ++i)
    {
        f3 p = box_centroid(leaf[i]);
        p.x -= whole.lower.x; p.y -= whole.lower.y; p.z -= whole.lower.z;
        p.x /= (whole.upper.x - whole.lower.x);
        p.y /= (whole.upper.y - whole.lower.y);
        p.z /= (whole.upper.z - whole.lower.z);
        keys[i] = ((uint64_t)orc_morton3(p.x, p.y, p.z) << 32) | (uint32_t)i;
    }
    qsort(keys, (size_t)n, sizeof(uint64_t), cmp_u64);
    uint32_t *k32 = malloc(sizeof(uint32_t) * (size_t)n);
    s->collision = 0;
    for (int k = 0; k < n; ++k)
    {
        k32[k] = (uint32_t)(keys[k] >> 32);
        s->morton_sorted[k] = k32[k];
        s->sorted_idx[k] = (uint32_t)(keys[k] & 0xFFFFFFFFu);
        if (k > 0 && k32[k] == k32[k - 1]) s->collision = 1; /* bvh.cuh:459-464 */
    }
    /* leaves in sorted order (payload moved by the sort in the reference, bvh.cuh:452-454) */
    for (int i = 0; i < nn; ++i)
    {
        s->nodes[i].parent = s->nodes[i].left = s->nodes[i].right = s->nodes[i].object = LEAF_NONE;
    }
    for (int k = 0; k < n; ++k)
    {
        const int obj = (int)s->sorted_idx[k];
        s->aabbs[ni + k] = leaf[obj];
        s->cones[ni + k] = tri_cone(s, obj);
        s->nodes[ni + k].object = (uint32_t)obj; /* bvh.cuh:489-499 */
    }
    /* internal nodes (bvh.cuh:200-229); 32-bit codes when unique, else (code<<32|idx) (bvh.cuh:464-476,505-515) */
    keyview kv;
    kv.n = n;
    kv.k32 = s->collision ? NULL : k32;
    kv.k64 = s->collision ? keys : NULL;
    for (int i = 0; i < ni; ++i)
    {
        int first, last;
        determine_range(&kv, i, &first, &last);
        const int gamma = find_split(&kv, first, last);
        s->ranges[2 * i] = (uint32_t)first;
        s->ranges[2 * i + 1] = (uint32_t)last;
        uint32_t l = (uint32_t)gamma, r = (uint32_t)gamma + 1;
        if (first == gamma) l += (uint32_t)ni;
        if (last == gamma + 1) r += (uint32_t)ni;
        s->nodes[i].left = l;
        s->nodes[i].right = r;
        s->nodes[l].parent = (uint32_t)i;
        s->nodes[r].parent = (uint32_t)i;
    }
    /* bottom-up AABB refit (bvh.cuh:520-554), serial emulation of the visit-flag protocol */
    int *flags = calloc((size_t)(ni > 0 ? ni : 1), sizeof(int));
    for (int k = 0; k < n; ++k)
    {
        uint32_t p = s->nodes[ni + k].parent;
        while (p != LEAF_NONE)
        {
            if (flags[p]++ == 0) break;
            s->aabbs[p] = box_merge(s->aabbs[s->nodes[p].left], s->aabbs[s->nodes[p].right]);
            p = s->nodes[p].parent;
        }
    }
    /* bottom-up cone refit (bvh.cuh:556-604) */
    memset(flags, 0, sizeof(int) * (size_t)(ni > 0 ? ni : 1));
    for (int k = 0; k < n; ++k)
    {
        uint32_t p = s->nodes[ni + k].parent;
        while (p != LEAF_NONE)
        {
            if (flags[p]++ == 0) break;
            const uint32_t l = s->nodes[p].left, r = s->nodes[p].right;
            int q1;
            s->cones[p] = cone_merge(&s->cones[l], &s->cones[r], box_centroid(s->aabbs[l]), box_centroid(s->aabbs[r]),
                                     box_centroid(s->aabbs[p]), &q1);
            s->q1[p] = (uint8_t)(q1 || s->q1[l] || s->q1[r]);
            p = s->nodes[p].parent;
        }
    }
    free(flags); free(leaf); free(keys); free(k32);
}

orc_scene *orc_scene3_create(const float *xyz, int n_verts, const int *tri, int n_tris)
{
    orc_scene *s = calloc(1, sizeof(orc_scene));
    s->nV = n_verts; s->nT = n_tris;
    s->verts = malloc(sizeof(f3) * (size_t)(n_verts > 0 ? n_verts : 1));
    memcpy(s->verts, xyz, sizeof(f3) * (size_t)n_verts);
    s->tris = malloc(sizeof(int) * 3 * (size_t)(n_tris > 0 ? n_tris : 1));
    memcpy(s->tris, tri, sizeof(int) * 3 * (size_t)n_tris);
    compute_silhouettes(s);
    assign_ownership(s);
    construct(s);
    return s;
}
void orc_scene_destroy(orc_scene *s)
{
    if (!s) return;
    free(s->verts); free(s->tris); free(s->edges); free(s->tri_edges); free(s->tri_owned);
    free(s->nodes); free(s->aabbs); free(s->cones); free(s->q1); free(s->morton_sorted); free(s->sorted_idx); free(s->ranges);
    free(s);
}
int orc_num_objects(const orc_scene *s) { return s->nT; }
int orc_num_nodes(const orc_scene *s) { return s->nT ? 2 * s->nT - 1 : 0; }
int orc_num_edges(const orc_scene *s) { return s->nE; }
int orc_collision(const orc_scene *s) { return s->collision; }
void orc_export_tree(const orc_scene *s, uint32_t *nodes, float *aabbs, float *cones)
{
    const size_t nn = (size_t)orc_num_nodes(s);
    if (nodes) memcpy(nodes, s->nodes, sizeof(node_t) * nn);
    if (aabbs) memcpy(aabbs, s->aabbs, sizeof(box3) * nn);
    if (cones) memcpy(cones, s->cones, sizeof(cone3) * nn);
}
void orc_export_q1_taint(const orc_scene *s, uint8_t *q1) { memcpy(q1, s->q1, (size_t)orc_num_nodes(s)); }
void orc_export_adjacency(const orc_scene *s, int *edges4, int *tri_edges3, int *tri_owned3)
{
    if (edges4) memcpy(edges4, s->edges, sizeof(int) * 4 * (size_t)s->nE);
    if (tri_edges3) memcpy(tri_edges3, s->tri_edges, sizeof(int) * 3 * (size_t)s->nT);
    if (tri_owned3) memcpy(tri_owned3, s->tri_owned, sizeof(int) * 3 * (size_t)s->nT);
}
void orc_export_morton(const orc_scene *s, uint32_t *morton_sorted, uint32_t *sorted_idx)
{
    if (morton_sorted) memcpy(morton_sorted, s->morton_sorted, sizeof(uint32_t) * (size_t)s->nT);
    if (sorted_idx) memcpy(sorted_idx, s->sorted_idx, sizeof(uint32_t) * (size_t)s->nT);
}
void orc_export_ranges(const orc_scene *s, uint32_t *first_last)
{
    if (s->nT > 1) memcpy(first_last, s->ranges, sizeof(uint32_t) * 2 * (size_t)(s->nT - 1));
}

/* ------------------------------------------------------------------------------------------------
 * Traversals — query.cuh, sample.cuh (one query at a time; same stack discipline as the reference)
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint32_t node; float key; } stack_entry;
#define STACK_CAP 256 /* the reference uses 64 (query.cuh:250) without a guard; the oracle never overflows */

/* query.cuh:238-318 */
static void closest_one(const orc_scene *s, f3 p, uint32_t *idx, float *dist)
{
    if (s->nT == 0) { *idx = LEAF_NONE; *dist = INFINITY; return; }
    if (s->nT == 1)
    { /* Q6: the reference reads out of bounds for a single-leaf tree; defined here as "test the only leaf" */
        *idx = s->nodes[0].object; *dist = tri_distance(s, (int)s->nodes[0].object, p); *dist = sqrtf(*dist * *dist);
        return;
    }
    stack_entry st[STACK_CAP];
    int sp = 0;
    st[sp].node = 0; st[sp].key = box_mindist(s->aabbs[0], p); ++sp;
    uint32_t nearest = LEAF_NONE;
    float best = INFINITY; /* squared */
    do
    {
        const stack_entry e = st[--sp];
        if (e.key > best) continue;
        const uint32_t L = s->nodes[e.node].left, R = s->nodes[e.node].right;
        const float Lmin = box_mindist(s->aabbs[L], p), Rmin = box_mindist(s->aabbs[R], p);
        const float Lmm = box_minmaxdist(s->aabbs[L], p), Rmm = box_minmaxdist(s->aabbs[R], p);
        if (Lmin <= Rmm)
        {
            const uint32_t obj = s->nodes[L].object;
            if (obj != LEAF_NONE)
            {
                float d = tri_distance(s, (int)obj, p);
                d *= d;
                if (d <= best) { best = d; nearest = obj; }
            }
            else { st[sp].node = L; st[sp].key = Lmin; ++sp; }
        }
        if (Rmin <= Lmm)
        {
            const uint32_t obj = s->nodes[R].object;
            if (obj != LEAF_NONE)
            {
                float d = tri_distance(s, (int)obj, p);
                d *= d;
                if (d <= best) { best = d; nearest = obj; }
            }
            else { st[sp].node = R; st[sp].key = Rmin; ++sp; }
        }
    } while (sp > 0);
    *idx = nearest;
    *dist = sqrtf(best);
}

/* query.cuh:325-423 */
/* edge_out / point_out (optional): the silhouette edge that attains the returned distance and the closest point on it; -1 and
 * (0,0,0) when nothing is found.  The walk and the distance are the reference's; these are the two values it drops. */
static float silhouette_one_ex(const orc_scene *s, f3 p, int flip, float r_max, int *edge_out, f3 *point_out)
{
    int best_edge = -1;
    f3 best_pt = mk3(0.0f, 0.0f, 0.0f);
    if (edge_out) *edge_out = -1;
    if (point_out) *point_out = best_pt;
    if (s->nT == 0) return INFINITY;
    float best = r_max; /* +inf in the reference (query.cuh:343) */
    int found_any = 0;
    if (s->nT == 1)
    { /* Q6 */
        float d = INFINITY;
        if (cone_valid(&s->cones[0]) && tri_closest_silhouette(s, (int)s->nodes[0].object, p, best * best, &d, flip, 0.0f, &best_edge, &best_pt) &&
            d <= best)
        {
            if (edge_out) *edge_out = best_edge;
            if (point_out) *point_out = best_pt;
            return d;
        }
        return INFINITY;
    }
    stack_entry st[STACK_CAP];
    int sp = 0;
    st[sp].node = 0; st[sp].key = box_mindist(s->aabbs[0], p); ++sp;
    do
    {
        const stack_entry e = st[--sp];
        if (e.key > best * best) continue;
        const uint32_t ch[2] = {s->nodes[e.node].left, s->nodes[e.node].right};
        float md[2];
        int hit[2];
        for (int c = 0; c < 2; ++c)
        {
            md[c] = box_mindist(s->aabbs[ch[c]], p);
            hit[c] = cone_valid(&s->cones[ch[c]]) && cone_overlap(&s->cones[ch[c]], p, s->aabbs[ch[c]], md[c]);
        }
        for (int c = 0; c < 2; ++c)
        {
            if (!hit[c]) continue;
            const uint32_t obj = s->nodes[ch[c]].object;
            if (obj != LEAF_NONE)
            {
                float d = INFINITY;
                int e_at = -1;
                f3 p_at = mk3(0.0f, 0.0f, 0.0f);
                const int found = tri_closest_silhouette(s, (int)obj, p, best * best, &d, flip, 0.0f, &e_at, &p_at);
                if (found && d <= best) { best = d; found_any = 1; best_edge = e_at; best_pt = p_at; }
            }
            else { st[sp].node = ch[c]; st[sp].key = md[c]; ++sp; }
        }
    } while (sp > 0);
    if (found_any)
    {
        if (edge_out) *edge_out = best_edge;
        if (point_out) *point_out = best_pt;
    }
    return found_any ? best : INFINITY;
}
static float silhouette_one(const orc_scene *s, f3 p, int flip, float r_max) { return silhouette_one_ex(s, p, flip, r_max, NULL, NULL); }

/* query.cuh:79-169 */
static void ray_one(const orc_scene *s, f3 org, f3 dir, float max_dist, int any_hit, int *found, float *t, float *uv, uint32_t *prim)
{
    *found = 0; *t = INFINITY; uv[0] = uv[1] = 0.0f; *prim = LEAF_NONE;
    if (s->nT == 0) return;
    const f3 dinv = mk3(1 / dir.x, 1 / dir.y, 1 / dir.z); /* aabb.cuh:305-312 */
    stack_entry st[STACK_CAP];
    int sp = 0;
    st[sp].node = 0; st[sp].key = INFINITY; ++sp;
    float best = INFINITY;
    do
    {
        const stack_entry e = st[--sp];
        if (e.key > best) continue;
        const uint32_t obj = s->nodes[e.node].object;
        if (obj != LEAF_NONE)
        {
            float tt, u, v;
            if (tri_ray(s, (int)obj, org, dir, &tt, &u, &v) && tt < max_dist && tt < best)
            {
                best = tt; *found = 1; uv[0] = u; uv[1] = v; *prim = obj;
                if (any_hit) break;
            }
        }
        else
        {
            const uint32_t L = s->nodes[e.node].left, R = s->nodes[e.node].right;
            float Ld, Rd;
            const int Lh = box_ray(org, dinv, s->aabbs[L], max_dist, &Ld);
            const int Rh = box_ray(org, dinv, s->aabbs[R], max_dist, &Rd);
            if (Lh && Rh)
            {
                uint32_t closer = L, other = R;
                if (Rd < Ld) { float tf = Ld; Ld = Rd; Rd = tf; closer = R; other = L; }
                st[sp].node = other; st[sp].key = Rd; ++sp;
                st[sp].node = closer; st[sp].key = Ld; ++sp;
            }
            else if (Lh) { st[sp].node = L; st[sp].key = Ld; ++sp; }
            else if (Rh) { st[sp].node = R; st[sp].key = Rd; ++sp; }
        }
    } while (sp > 0);
    *t = best;
}

/* sample.cuh:23-92; pdf reported as 0 on a miss (uninitialised in the reference, Q19) */
static void sample_one(const orc_scene *s, f3 c, float radius, float u, int *idx, float *pdf)
{
    *idx = -1; *pdf = 0.0f;
    if (s->nT == 0) return;
    uint32_t node = 0;
    float path = 1.0f;
    for (;;)
    {
        const uint32_t obj = s->nodes[node].object;
        if (obj != LEAF_NONE)
        {
            if (tri_sphere(s, (int)obj, c, radius))
            {
                *idx = (int)obj;
                *pdf = path / tri_area(s, (int)obj);
            }
            return;
        }
        const uint32_t L = s->nodes[node].left, R = s->nodes[node].right;
        const float Lw = box_sphere(c, radius, s->aabbs[L]) ? green_weight(c, box_centroid(s->aabbs[L])) : 0;
        const float Rw = box_sphere(c, radius, s->aabbs[R]) ? green_weight(c, box_centroid(s->aabbs[R])) : 0;
        const float total = Lw + Rw;
        if (!(total > 0)) return;
        const float Lp = Lw / total;
        if (u < Lp) { u /= Lp; node = L; path = Lp * path; }
        else { const float Rp = 1.0f - Lp; u = (u - Lp) / Rp; node = R; path = Rp * path; }
    }
}

/* ---- threaded batch drivers ---- */
typedef struct
{
    const orc_scene *s;
    int kind, flip, any_hit;
    long lo, hi;
    const float *a, *b, *c;
    uint32_t *oidx;
    int *ofound, *oint;
    float *of0, *of1;
} job_t;
enum { K_CLOSEST, K_SIL, K_RAY, K_SAMPLE, K_CLOSEST_BRUTE, K_RAY_BRUTE };

static void *job_run(void *arg)
{
    job_t *j = arg;
    const orc_scene *s = j->s;
    for (long i = j->lo; i < j->hi; ++i)
    {
        switch (j->kind)
        {
        case K_CLOSEST:
            closest_one(s, mk3(j->a[3 * i], j->a[3 * i + 1], j->a[3 * i + 2]), &j->oidx[i], &j->of0[i]);
            break;
        case K_SIL:
        {
            int e_at;
            f3 p_at;
            j->of0[i] = silhouette_one_ex(s, mk3(j->a[3 * i], j->a[3 * i + 1], j->a[3 * i + 2]), j->flip, j->b ? j->b[i] : INFINITY, &e_at, &p_at);
            if (j->oint) j->oint[i] = e_at;
            if (j->of1) { j->of1[3 * i] = p_at.x; j->of1[3 * i + 1] = p_at.y; j->of1[3 * i + 2] = p_at.z; }
            break;
        }
        case K_RAY:
            ray_one(s, mk3(j->a[3 * i], j->a[3 * i + 1], j->a[3 * i + 2]), mk3(j->b[3 * i], j->b[3 * i + 1], j->b[3 * i + 2]), j->c[i],
                    j->any_hit, &j->ofound[i], &j->of0[i], &j->of1[2 * i], &j->oidx[i]);
            break;
        case K_SAMPLE:
            sample_one(s, mk3(j->a[4 * i], j->a[4 * i + 1], j->a[4 * i + 2]), j->a[4 * i + 3], j->b[i], &j->oint[i], &j->of0[i]);
            break;
        case K_CLOSEST_BRUTE:
        {
            const f3 p = mk3(j->a[3 * i], j->a[3 * i + 1], j->a[3 * i + 2]);
            float best = INFINITY;
            uint32_t bi = LEAF_NONE;
            for (int t = 0; t < s->nT; ++t)
            {
                const float d = tri_distance(s, t, p);
                if (d < best) { best = d; bi = (uint32_t)t; }
            }
            j->oidx[i] = bi; j->of0[i] = best;
            break;
        }
        case K_RAY_BRUTE:
        {
            const f3 o = mk3(j->a[3 * i], j->a[3 * i + 1], j->a[3 * i + 2]), d = mk3(j->b[3 * i], j->b[3 * i + 1], j->b[3 * i + 2]);
            float best = INFINITY;
            uint32_t bi = LEAF_NONE;
            for (int t = 0; t < s->nT; ++t)
            {
                float tt, u, v;
                if (tri_ray(s, t, o, d, &tt, &u, &v) && tt < j->c[i] && tt < best) { best = tt; bi = (uint32_t)t; }
            }
            j->ofound[i] = bi != LEAF_NONE; j->of0[i] = best; j->oidx[i] = bi;
            break;
        }
        }
    }
    return NULL;
}
static void run_jobs(job_t base, long n, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads == 1 || n < 2 * nthreads)
    {
        base.lo = 0; base.hi = n;
        job_run(&base);
        return;
    }
    pthread_t th[256];
    job_t jobs[256];
    for (int t = 0; t < nthreads; ++t)
    {
        jobs[t] = base;
        jobs[t].lo = n * t / nthreads;
        jobs[t].hi = n * (t + 1) / nthreads;
        pthread_create(&th[t], NULL, job_run, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
}

void orc_closest(const orc_scene *s, const float *q, long n, uint32_t *idx, float *dist, int nthreads)
{
    job_t j = {0};
    j.s = s; j.kind = K_CLOSEST; j.a = q; j.oidx = idx; j.of0 = dist;
    run_jobs(j, n, nthreads);
}
void orc_silhouette(const orc_scene *s, const float *q, long n, int flip, const float *r_max, float *dist, int nthreads)
{
    job_t j = {0};
    j.s = s; j.kind = K_SIL; j.a = q; j.b = r_max; j.flip = flip; j.of0 = dist;
    run_jobs(j, n, nthreads);
}
/* + edge[n] (silhouette edge id, -1 = none) and point[3n] (closest point on that edge) */
void orc_silhouette_ex(const orc_scene *s, const float *q, long n, int flip, const float *r_max, float *dist, int *edge, float *point, int nthreads)
{
    job_t j = {0};
    j.s = s; j.kind = K_SIL; j.a = q; j.b = r_max; j.flip = flip; j.of0 = dist; j.oint = edge; j.of1 = point;
    run_jobs(j, n, nthreads);
}
/* distance from q[i] to silhouette edge edge[i] (closest_point_segment, scene.cuh:230-255) and the closest point on it */
void orc_point_edge_distance(const orc_scene *s, const float *q, const int *edge, long n, float *dist, float *point)
{
    for (long i = 0; i < n; ++i)
    {
        const int *e = s->edges + 4 * edge[i];
        f3 cp;
        dist[i] = closest_point_segment(s->verts[e[1]], s->verts[e[2]], mk3(q[3 * i], q[3 * i + 1], q[3 * i + 2]), &cp);
        point[3 * i] = cp.x; point[3 * i + 1] = cp.y; point[3 * i + 2] = cp.z;
    }
}
void orc_ray(const orc_scene *s, const float *org, const float *dir, const float *tmax, long n, int any_hit, int *found, float *t,
             float *uv, uint32_t *prim, int nthreads)
{
    job_t j = {0};
    j.s = s; j.kind = K_RAY; j.a = org; j.b = dir; j.c = tmax; j.any_hit = any_hit; j.ofound = found; j.of0 = t; j.of1 = uv; j.oidx = prim;
    run_jobs(j, n, nthreads);
}
void orc_sample(const orc_scene *s, const float *sph4, const float *u, long n, int *idx, float *pdf, int nthreads)
{
    job_t j = {0};
    j.s = s; j.kind = K_SAMPLE; j.a = sph4; j.b = u; j.oint = idx; j.of0 = pdf;
    run_jobs(j, n, nthreads);
}
/* sample.cuh:7-21 + scene.cuh:14-27,1231-1240 */
void orc_sample_on_object(const orc_scene *s, const int *idx, const float *u, const float *v, long n, float *xyz)
{
    for (long i = 0; i < n; ++i)
    {
        if (idx[i] < 0) { xyz[3 * i] = xyz[3 * i + 1] = xyz[3 * i + 2] = 0.0f; continue; }
        f3 a, b, c;
        tri_verts(s, idx[i], &a, &b, &c);
        float uu = u[i], vv = v[i];
        if (uu + vv > 1.0f) { uu = 1.0f - uu; vv = 1.0f - vv; }
        const float w = 1.0f - uu - vv;
        xyz[3 * i] = w * a.x + uu * b.x + vv * c.x;
        xyz[3 * i + 1] = w * a.y + uu * b.y + vv * c.y;
        xyz[3 * i + 2] = w * a.z + uu * b.z + vv * c.z;
    }
}
void orc_closest_brute(const orc_scene *s, const float *q, long n, uint32_t *idx, float *dist, int nthreads)
{
    job_t j = {0};
    j.s = s; j.kind = K_CLOSEST_BRUTE; j.a = q; j.oidx = idx; j.of0 = dist;
    run_jobs(j, n, nthreads);
}
void orc_ray_brute(const orc_scene *s, const float *org, const float *dir, const float *tmax, long n, int *found, float *t,
                   uint32_t *prim, int nthreads)
{
    job_t j = {0};
    j.s = s; j.kind = K_RAY_BRUTE; j.a = org; j.b = dir; j.c = tmax; j.ofound = found; j.of0 = t; j.oidx = prim;
    run_jobs(j, n, nthreads);
}
void orc_point_triangle_distance(const orc_scene *s, const float *q, const uint32_t *idx, long n, float *dist)
{
    for (long i = 0; i < n; ++i)
        dist[i] = (idx[i] == LEAF_NONE || (int)idx[i] >= s->nT) ? INFINITY
                                                                : tri_distance(s, (int)idx[i], mk3(q[3 * i], q[3 * i + 1], q[3 * i + 2]));
}

/* ------------------------------------------------------------------------------------------------
 * Must-visit statistics: nodes any exact traversal of THIS tree has to open
 * ---------------------------------------------------------------------------------------------- */
static void must_visit_closest(const orc_scene *s, f3 p, float d2, long *vi, long *vl)
{
    uint32_t st[STACK_CAP];
    int sp = 0;
    st[sp++] = 0;
    while (sp > 0)
    {
        const uint32_t nd = st[--sp];
        ++*vi;
        const uint32_t ch[2] = {s->nodes[nd].left, s->nodes[nd].right};
        for (int c = 0; c < 2; ++c)
        {
            if (box_mindist(s->aabbs[ch[c]], p) > d2) continue;
            if (s->nodes[ch[c]].object != LEAF_NONE) ++*vl;
            else st[sp++] = ch[c];
        }
    }
}
void orc_closest_must_visit(const orc_scene *s, const float *q, long n, double *mean_internal, double *mean_leaves)
{
    long vi = 0, vl = 0;
    for (long i = 0; i < n && s->nT > 1; ++i)
    {
        const f3 p = mk3(q[3 * i], q[3 * i + 1], q[3 * i + 2]);
        uint32_t idx;
        float d;
        closest_one(s, p, &idx, &d);
        must_visit_closest(s, p, d * d, &vi, &vl);
    }
    *mean_internal = n ? (double)vi / (double)n : 0;
    *mean_leaves = n ? (double)vl / (double)n : 0;
}
void orc_silhouette_must_visit(const orc_scene *s, const float *q, long n, int flip, const float *r_max, double *mean_internal,
                               double *mean_leaves)
{
    long vi = 0, vl = 0;
    for (long i = 0; i < n && s->nT > 1; ++i)
    {
        const f3 p = mk3(q[3 * i], q[3 * i + 1], q[3 * i + 2]);
        const float rm = r_max ? r_max[i] : INFINITY;
        float d = silhouette_one(s, p, flip, rm);
        if (!(d < INFINITY)) d = rm;
        const float d2 = d * d;
        uint32_t st[STACK_CAP];
        int sp = 0;
        st[sp++] = 0;
        while (sp > 0)
        {
            const uint32_t nd = st[--sp];
            ++vi;
            const uint32_t ch[2] = {s->nodes[nd].left, s->nodes[nd].right};
            for (int c = 0; c < 2; ++c)
            {
                const float md = box_mindist(s->aabbs[ch[c]], p);
                if (md > d2) continue;
                if (!(cone_valid(&s->cones[ch[c]]) && cone_overlap(&s->cones[ch[c]], p, s->aabbs[ch[c]], md))) continue;
                if (s->nodes[ch[c]].object != LEAF_NONE) ++vl;
                else st[sp++] = ch[c];
            }
        }
    }
    *mean_internal = n ? (double)vi / (double)n : 0;
    *mean_leaves = n ? (double)vl / (double)n : 0;
}
void orc_ray_must_visit(const orc_scene *s, const float *org, const float *dir, const float *tmax, long n, double *mean_internal,
                        double *mean_leaves)
{
    long vi = 0, vl = 0;
    for (long i = 0; i < n && s->nT > 1; ++i)
    {
        const f3 o = mk3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), d = mk3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
        int found;
        float t, uv[2];
        uint32_t prim;
        ray_one(s, o, d, tmax[i], 0, &found, &t, uv, &prim);
        const float lim = found ? t : tmax[i];
        const f3 dinv = mk3(1 / d.x, 1 / d.y, 1 / d.z);
        uint32_t st[STACK_CAP];
        int sp = 0;
        st[sp++] = 0;
        while (sp > 0)
        {
            const uint32_t nd = st[--sp];
            ++vi;
            const uint32_t ch[2] = {s->nodes[nd].left, s->nodes[nd].right};
            for (int c = 0; c < 2; ++c)
            {
                float e;
                if (!box_ray(o, dinv, s->aabbs[ch[c]], lim, &e)) continue;
                if (s->nodes[ch[c]].object != LEAF_NONE) ++vl;
                else st[sp++] = ch[c];
            }
        }
    }
    *mean_internal = n ? (double)vi / (double)n : 0;
    *mean_leaves = n ? (double)vl / (double)n : 0;
}

/* the HOST libm, for tests/test_gpu_host_libm.py: which = 0 acosf, 1 sinf, 2 cosf, 3 logf */
void orc_host_libm(int which, const float *x, long n, float *out)
{
    for (long i = 0; i < n; ++i)
    {
        volatile float v = x[i];
        out[i] = which == 0 ? acosf(v) : which == 1 ? sinf(v) : which == 2 ? cosf(v) : logf(v);
    }
}

/* What a NEAR-FIRST walk with immediate leaf tests visits (tools/visit_sim.py): the nearer surviving child first, the other on
 * the stack with its box distance, rejected at pop time against the bound of that moment — the GPU kernels' order when a
 * leaf's edges are tested the moment the leaf is reached.  Same answer as the reference's walk; only the counts are reported:
 * internal nodes opened, leaves tested.  `defer` > 0 models the kernels' leaf queue: a leaf's result only tightens the bound
 * after `defer` further nodes have been opened. */
static long g_need_axis;
long orc_need_axis_count(int reset) { const long v = g_need_axis; if (reset) g_need_axis = 0; return v; }
static void nearfirst_walk(const orc_scene *s, const float *q, long n, int flip, const float *r_max, int defer, double *mean_internal,
                           double *mean_leaves, float *dist_out);
void orc_silhouette_nearfirst_visits(const orc_scene *s, const float *q, long n, int flip, const float *r_max, int defer, double *mean_internal,
                                     double *mean_leaves)
{
    nearfirst_walk(s, q, n, flip, r_max, defer, mean_internal, mean_leaves, NULL);
}
/* the distances such a walk returns (defer = 0): equal to the reference's whenever the answer does not depend on the order of
 * the walk — i.e. unless the ROUNDED distance to an edge is smaller than the rounded distance to a box that holds it (far from
 * the origin), see DESIGN.md "Far from the origin" */
void orc_silhouette_nearfirst(const orc_scene *s, const float *q, long n, int flip, const float *r_max, float *dist)
{
    double a, b;
    nearfirst_walk(s, q, n, flip, r_max, 0, &a, &b, dist);
}
static void nearfirst_walk(const orc_scene *s, const float *q, long n, int flip, const float *r_max, int defer, double *mean_internal,
                           double *mean_leaves, float *dist_out)
{
    long vi = 0, vl = 0;
    for (long i = 0; i < n && s->nT > 1; ++i)
    {
        const f3 p = mk3(q[3 * i], q[3 * i + 1], q[3 * i + 2]);
        float best = r_max ? r_max[i] : INFINITY;
        int found_any = 0;
        float pend_d[64];
        long pend_at[64];
        int np = 0;
        long opened = 0;
        stack_entry st[STACK_CAP];
        int sp = 0;
        uint32_t node = 0;
        for (;;)
        {
            /* deferred results that are due */
            int k = 0;
            for (int j = 0; j < np; ++j)
            {
                if (pend_at[j] <= opened) { if (pend_d[j] <= best) best = pend_d[j]; }
                else { pend_d[k] = pend_d[j]; pend_at[k] = pend_at[j]; ++k; }
            }
            np = k;
            ++vi; ++opened;
            const uint32_t ch[2] = {s->nodes[node].left, s->nodes[node].right};
            float md[2];
            int hit[2], need_axis = 0;
            for (int c = 0; c < 2; ++c)
            {
                md[c] = box_mindist(s->aabbs[ch[c]], p);
                hit[c] = md[c] <= best * best && cone_valid(&s->cones[ch[c]]) && cone_overlap(&s->cones[ch[c]], p, s->aabbs[ch[c]], md[c]);
                /* would this child's test have read the cone's axis / radius?  (cone.cuh:174: wide cones and points inside the box pass at once) */
                if (md[c] <= best * best && cone_valid(&s->cones[ch[c]]) && s->cones[ch[c]].half_angle < ORC_PI_2_F && !(md[c] < FLT_EPSILON)) need_axis = 1;
            }
            g_need_axis += need_axis;
            const int first = md[1] < md[0] ? 1 : 0;
            uint32_t next = LEAF_NONE;
            for (int t = 0; t < 2; ++t)
            {
                const int c = t == 0 ? first : 1 - first;
                if (!hit[c]) continue;
                const uint32_t obj = s->nodes[ch[c]].object;
                if (obj != LEAF_NONE)
                {
                    ++vl;
                    float d = INFINITY;
                    if (tri_closest_silhouette(s, (int)obj, p, best * best, &d, flip, 0.0f, NULL, NULL) && d <= best)
                    {
                        found_any = 1;
                        if (defer <= 0 || np >= 64) best = d;
                        else { pend_d[np] = d; pend_at[np] = opened + defer; ++np; }
                    }
                }
                else if (next == LEAF_NONE) next = ch[c];
                else { st[sp].node = ch[c]; st[sp].key = md[c]; ++sp; }
            }
            while (next == LEAF_NONE && sp > 0)
            {
                --sp;
                if (st[sp].key <= best * best) next = st[sp].node;
            }
            if (next == LEAF_NONE) break;
            node = next;
        }
        if (dist_out) dist_out[i] = found_any ? best : INFINITY;
    }
    *mean_internal = n ? (double)vi / (double)n : 0;
    *mean_leaves = n ? (double)vl / (double)n : 0;
}
