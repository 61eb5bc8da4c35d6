// snch_math.cuh — device-side geometry for the SNCH-LBVH kernels.
//
// Each routine computes the SAME quantity as the reference routine it cites, with the same operation order, so that
// results agree with the reference to rounding (<=1e-5 relative as BASELINE.json requires) and the integer pipeline
// (Morton codes -> sort order -> topology) agrees bit-for-bit.  The build-side box/Morton arithmetic uses explicit
// round-to-nearest intrinsics so no FMA contraction can change a code.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/snch_lbvh/core/host_libm.cuh" // glibc's acosf / sinf / cosf bits on the device (cone refit: parity with the reference's CPU build)

namespace snch
{

#define SNCH_DI __device__ __forceinline__

constexpr float kPi = 3.14159265358979323846f;
constexpr float kHalfPi = 1.57079632679489661923f;

struct V3
{
    float x, y, z;
};
SNCH_DI V3 v3(float x, float y, float z) { return V3{x, y, z}; }
SNCH_DI V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
SNCH_DI float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }                       // utility.cuh:238
SNCH_DI float len(V3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }                      // utility.cuh:382
SNCH_DI float sqlen(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }                           // utility.cuh:400
SNCH_DI V3 normalize(V3 v)                                                                        // utility.cuh:427
{
    const float n = len(v);
    return V3{v.x / n, v.y / n, v.z / n};
}
SNCH_DI V3 cross(V3 u, V3 v) { return V3{u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x}; } // utility.cuh:449
SNCH_DI V3 vabs(V3 a) { return V3{fabsf(a.x), fabsf(a.y), fabsf(a.z)}; }
// std::min / std::max semantics of the reference's host-style calls (NaN behaviour differs from fminf/fmaxf)
SNCH_DI float std_min(float a, float b) { return (b < a) ? b : a; }
SNCH_DI float std_max(float a, float b) { return (a < b) ? b : a; }

struct Box
{
    V3 lo, hi;
};
struct Cone
{
    V3 axis;
    float half_angle, radius;
};

// ---- boxes ------------------------------------------------------------------------------------------------------
// aabb.cuh:21-38 + 94-101 via scene.cuh:870-885: FLT_EPSILON-padded triangle box.  Contraction-proof.
SNCH_DI Box tri_box(V3 a, V3 b, V3 c)
{
    const float e = FLT_EPSILON;
    Box r;
    r.hi = V3{__fadd_rn(a.x, e), __fadd_rn(a.y, e), __fadd_rn(a.z, e)};
    r.lo = V3{__fsub_rn(a.x, e), __fsub_rn(a.y, e), __fsub_rn(a.z, e)};
    r.lo = V3{fminf(r.lo.x, __fsub_rn(b.x, e)), fminf(r.lo.y, __fsub_rn(b.y, e)), fminf(r.lo.z, __fsub_rn(b.z, e))};
    r.hi = V3{fmaxf(r.hi.x, __fadd_rn(b.x, e)), fmaxf(r.hi.y, __fadd_rn(b.y, e)), fmaxf(r.hi.z, __fadd_rn(b.z, e))};
    r.lo = V3{fminf(r.lo.x, __fsub_rn(c.x, e)), fminf(r.lo.y, __fsub_rn(c.y, e)), fminf(r.lo.z, __fsub_rn(c.z, e))};
    r.hi = V3{fmaxf(r.hi.x, __fadd_rn(c.x, e)), fmaxf(r.hi.y, __fadd_rn(c.y, e)), fmaxf(r.hi.z, __fadd_rn(c.z, e))};
    return r;
}
SNCH_DI Box box_merge(Box l, Box r) // aabb.cuh:113-124
{
    Box m;
    m.hi = V3{fmaxf(l.hi.x, r.hi.x), fmaxf(l.hi.y, r.hi.y), fmaxf(l.hi.z, r.hi.z)};
    m.lo = V3{fminf(l.lo.x, r.lo.x), fminf(l.lo.y, r.lo.y), fminf(l.lo.z, r.lo.z)};
    return m;
}
SNCH_DI V3 box_centroid(Box b) // aabb.cuh:271-279 ((u+l)*0.5 is exact in either precision)
{
    return V3{__fmul_rn(__fadd_rn(b.hi.x, b.lo.x), 0.5f), __fmul_rn(__fadd_rn(b.hi.y, b.lo.y), 0.5f),
              __fmul_rn(__fadd_rn(b.hi.z, b.lo.z), 0.5f)};
}
SNCH_DI float box_mindist2(V3 lo, V3 hi, V3 p) // aabb.cuh:144-150
{
    const float dx = fminf(hi.x, fmaxf(lo.x, p.x)) - p.x;
    const float dy = fminf(hi.y, fmaxf(lo.y, p.y)) - p.y;
    const float dz = fminf(hi.z, fmaxf(lo.z, p.z)) - p.z;
    return dx * dx + dy * dy + dz * dz;
}
// aabb.cuh:397-431: slab test, entry distance clamped to >= 0; NaN lanes are suppressed by fminf/fmaxf exactly as there
SNCH_DI bool box_ray(V3 lo, V3 hi, V3 org, V3 dinv, float max_dist, float *entry)
{
    float t1 = (lo.x - org.x) * dinv.x, t2 = (hi.x - org.x) * dinv.x;
    float tmin = fminf(t1, t2), tmax = fmaxf(t1, t2);
    t1 = (lo.y - org.y) * dinv.y;
    t2 = (hi.y - org.y) * dinv.y;
    tmin = fmaxf(tmin, fminf(t1, t2));
    tmax = fminf(tmax, fmaxf(t1, t2));
    t1 = (lo.z - org.z) * dinv.z;
    t2 = (hi.z - org.z) * dinv.z;
    tmin = fmaxf(tmin, fminf(t1, t2));
    tmax = fminf(tmax, fmaxf(t1, t2));
    *entry = (tmin >= 0.0f) ? tmin : 0.0f;
    return tmax >= tmin && tmax >= 0.0f && tmin <= max_dist;
}
SNCH_DI bool box_sphere(V3 lo, V3 hi, V3 c, float radius) // aabb.cuh:433-449
{
    const float cx = std_max(lo.x, std_min(c.x, hi.x));
    const float cy = std_max(lo.y, std_min(c.y, hi.y));
    const float dx = cx - c.x, dy = cy - c.y;
    float d2 = dx * dx + dy * dy;
    const float cz = std_max(lo.z, std_min(c.z, hi.z));
    const float dz = cz - c.z;
    d2 += dz * dz;
    return d2 <= radius * radius;
}

// ---- Morton ------------------------------------------------------------------------------------------------------
SNCH_DI uint32_t expand_bits10(uint32_t v) // morton_code.cuh:19-38
{
    v = (v | (v << 16)) & 0x070000FFu;
    v = (v | (v << 8)) & 0x0700F00Fu;
    v = (v | (v << 4)) & 0x430C30C3u;
    v = (v | (v << 2)) & 0x49249249u;
    return v;
}
// bvh.cuh:292-302 + morton_code.cuh:61-70.  IEEE division; x*1024 is exact; explicit rn intrinsics throughout.
SNCH_DI uint32_t morton30(Box leaf, V3 wlo, V3 whi)
{
    const V3 c = box_centroid(leaf);
    float x = __fdiv_rn(__fsub_rn(c.x, wlo.x), __fsub_rn(whi.x, wlo.x));
    float y = __fdiv_rn(__fsub_rn(c.y, wlo.y), __fsub_rn(whi.y, wlo.y));
    float z = __fdiv_rn(__fsub_rn(c.z, wlo.z), __fsub_rn(whi.z, wlo.z));
    x = fminf(fmaxf(__fmul_rn(x, 1024.0f), 0.0f), 1023.0f);
    y = fminf(fmaxf(__fmul_rn(y, 1024.0f), 0.0f), 1023.0f);
    z = fminf(fmaxf(__fmul_rn(z, 1024.0f), 0.0f), 1023.0f);
    return expand_bits10((uint32_t)x) * 4u + expand_bits10((uint32_t)y) * 2u + expand_bits10((uint32_t)z);
}

// ---- normal cones --------------------------------------------------------------------------------------------------
SNCH_DI bool inrange(float v, float lo, float hi) { return v >= lo && v <= hi; } // utility.cuh:241

SNCH_DI float project_to_plane(V3 n, V3 e) // cone.cuh:34-42, 58-66
{
    const float sign = copysignf(1.0f, n.z);
    const float a = -1.0f / (sign + n.z);
    const float b = n.x * n.y * a;
    const V3 b1 = V3{1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x};
    const V3 b2 = V3{b, sign + n.y * n.y * a, -n.y};
    const float r1 = dot(e, vabs(b1)), r2 = dot(e, vabs(b2));
    return sqrtf(r1 * r1 + r2 * r2);
}
// cone.cuh:168-212.  `md2` = squared distance from o to the box (already computed by the caller).
SNCH_DI bool cone_overlap(V3 axis, float half_angle, float radius, V3 o, V3 lo, V3 hi, float md2)
{
    if (half_angle >= kHalfPi || md2 < FLT_EPSILON) return true;
    const V3 c = V3{(hi.x + lo.x) * 0.5f, (hi.y + lo.y) * 0.5f, (hi.z + lo.z) * 0.5f};
    V3 v = c - o;
    const float l = len(v);
    v.x /= l;
    v.y /= l;
    v.z /= l;
    const float d_axis_angle = acosf(std_max(-1.0f, std_min(1.0f, dot(axis, v))));
    if (inrange(kHalfPi, d_axis_angle - half_angle, d_axis_angle + half_angle)) return true;
    if (l > radius)
    {
        const float view_half = asinf(radius / l);
        const float sum = half_angle + view_half;
        return sum >= kHalfPi ? true : inrange(kHalfPi, d_axis_angle - sum, d_axis_angle + sum);
    }
    const V3 e = hi - c;
    float d = dot(e, vabs(v));
    const float s = l - d;
    if (s <= 0.0f) return true;
    d = project_to_plane(v, e);
    const float view_half = atan2f(d, s);
    const float sum = half_angle + view_half;
    return sum >= kHalfPi ? true : inrange(kHalfPi, d_axis_angle - sum, d_axis_angle + sum);
}
SNCH_DI V3 rotate_towards(V3 u, V3 v, float theta) // cone.cuh:288-302 (Rodrigues)
{
    const float ct = lbvh::detail::cosf_host(theta), st = lbvh::detail::sinf_host(theta);
    const V3 w = normalize(cross(u, v));
    const V3 o = V3{(1.0f - ct) * w.x, (1.0f - ct) * w.y, (1.0f - ct) * w.z};
    const float r00 = ct + o.x * w.x, r01 = o.y * w.x - st * w.z, r02 = o.z * w.x + st * w.y;
    const float r10 = o.x * w.y + st * w.z, r11 = ct + o.y * w.y, r12 = o.z * w.y - st * w.x;
    const float r20 = o.x * w.z - st * w.y, r21 = o.y * w.z + st * w.x, r22 = ct + o.z * w.z;
    return V3{r00 * u.x + r01 * u.y + r02 * u.z, r10 * u.x + r11 * u.y + r12 * u.z, r20 * u.x + r21 * u.y + r22 * u.z};
}
// cone.cuh:427-480.  *q1 reports the branch where the reference leaves half_angle uninitialised (defined as pi here).
SNCH_DI Cone cone_merge(Cone ca, Cone cb, V3 oa, V3 ob, V3 on, bool *q1)
{
    Cone r;
    r.axis = V3{0.f, 0.f, 0.f};
    r.half_angle = 0.f;
    r.radius = 0.f;
    *q1 = false;
    const bool va = ca.half_angle >= 0.0f, vb = cb.half_angle >= 0.0f;
    if (va && vb)
    {
        V3 axis_a = ca.axis, axis_b = cb.axis;
        float ha = ca.half_angle, hb = cb.half_angle;
        const V3 da = on - oa, db = on - ob;
        r.radius = sqrtf(std_max(ca.radius * ca.radius + sqlen(da), cb.radius * cb.radius + sqlen(db)));
        if (hb > ha)
        {
            const V3 t = axis_a;
            axis_a = axis_b;
            axis_b = t;
            const float th = ha;
            ha = hb;
            hb = th;
        }
        const float theta = lbvh::detail::acosf_host(std_max(-1.0f, std_min(1.0f, dot(axis_a, axis_b))));
        if (std_min(theta + hb, kPi) <= ha)
        {
            r.axis = axis_a;
            r.half_angle = ha;
            return r;
        }
        const float o_theta = (ha + theta + hb) / 2.0f;
        if (o_theta >= kPi)
        {
            r.axis = axis_a;
            r.half_angle = kPi;
            *q1 = true;
            return r;
        }
        r.axis = rotate_towards(axis_a, axis_b, o_theta - ha);
        r.half_angle = o_theta;
    }
    else if (va) r = ca;
    else if (vb) r = cb;
    else r.half_angle = -kPi;
    return r;
}

// ---- primitives ----------------------------------------------------------------------------------------------------
// scene.cuh:34-110 (Ericson RTCD 5.1.5); returns the distance (not squared), like the reference.
// (A branch-free form — region selected by the same predicate chain, one length at the end — is bit-identical and was measured
// SLOWER: 60.4 vs 49.1 ms on the 16.7M-query packet batch, 114 vs 96 ms on the C5 shard; the four divisions and the selects cost
// more than the divergence over Voronoi regions, and 7 more registers cost a resident CTA.)
SNCH_DI float point_triangle_distance(V3 pa, V3 pb, V3 pc, V3 x)
{
    const V3 ab = pb - pa, ac = pc - pa, ax = x - pa;
    const float d1 = dot(ab, ax), d2 = dot(ac, ax);
    if (d1 <= 0.0f && d2 <= 0.0f) return len(x - pa);
    const V3 bx = x - pb;
    const float d3 = dot(ab, bx), d4 = dot(ac, bx);
    if (d3 >= 0.0f && d4 <= d3) return len(x - pb);
    const V3 cx = x - pc;
    const float d5 = dot(ab, cx), d6 = dot(ac, cx);
    if (d6 >= 0.0f && d5 <= d6) return len(x - pc);
    const float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f)
    {
        const float v = d1 / (d1 - d3);
        return len(x - V3{pa.x + ab.x * v, pa.y + ab.y * v, pa.z + ab.z * v});
    }
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f)
    {
        const float w = d2 / (d2 - d6);
        return len(x - V3{pa.x + ac.x * w, pa.y + ac.y * w, pa.z + ac.z * w});
    }
    const float va = d3 * d6 - d5 * d4;
    if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f)
    {
        const float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        return len(x - V3{pb.x + (pc.x - pb.x) * w, pb.y + (pc.y - pb.y) * w, pb.z + (pc.z - pb.z) * w});
    }
    const float denom = 1.0f / (va + vb + vc);
    const float v = vb * denom, w = vc * denom;
    return len(x - V3{pa.x + ab.x * v + ac.x * w, pa.y + ab.y * v + ac.y * w, pa.z + ab.z * v + ac.z * w});
}
// scene.cuh:230-255
SNCH_DI float point_segment_distance(V3 pa, V3 pb, V3 x, V3 *cp)
{
    const V3 u = pb - pa, v = x - pa;
    const float c1 = dot(u, v);
    if (c1 <= 0.0f)
    {
        *cp = pa;
        return len(x - pa);
    }
    const float c2 = dot(u, u);
    if (c2 <= c1)
    {
        *cp = pb;
        return len(x - pb);
    }
    const float t = c1 / c2;
    *cp = V3{pa.x + u.x * t, pa.y + u.y * t, pa.z + u.z * t};
    return len(x - *cp);
}
// scene.cuh:143-174 (view direction deliberately NOT normalised, quirk Q2)
SNCH_DI bool is_silhouette_edge(V3 pa, V3 pb, V3 n0, V3 n1, V3 view, float d, bool flip)
{
    const float precision = 1e-3f;
    const float sign = flip ? 1.0f : -1.0f;
    if (d <= precision)
    {
        const V3 edge_dir = normalize(pb - pa);
        const float dihedral = atan2f(dot(edge_dir, cross(n0, n1)), dot(n0, n1));
        return sign * dihedral > precision;
    }
    const float dot0 = dot(view, n0), dot1 = dot(view, n1);
    if (fabsf(dot0) <= precision) return sign * dot1 > precision;
    if (fabsf(dot1) <= precision) return sign * dot0 > precision;
    return dot0 * dot1 < 0.0f;
}
// scene.cuh:1005-1052 (Moeller-Trumbore with __frcp_rn)
SNCH_DI bool ray_triangle(V3 v0, V3 v1, V3 v2, V3 org, V3 dir, float *t, float *u, float *v)
{
    const V3 e1 = v1 - v0, e2 = v2 - v0;
    const V3 h = V3{dir.y * e2.z - dir.z * e2.y, dir.z * e2.x - dir.x * e2.z, dir.x * e2.y - dir.y * e2.x};
    const float det = e1.x * h.x + e1.y * h.y + e1.z * h.z;
    if (fabsf(det) < FLT_EPSILON) return false;
    const float inv_det = __frcp_rn(det);
    const V3 s = org - v0;
    const float uu = (s.x * h.x + s.y * h.y + s.z * h.z) * inv_det;
    if (uu < 0.0f || uu > 1.0f) return false;
    const V3 q = V3{s.y * e1.z - s.z * e1.y, s.z * e1.x - s.x * e1.z, s.x * e1.y - s.y * e1.x};
    const float vv = (dir.x * q.x + dir.y * q.y + dir.z * q.z) * inv_det;
    if (vv < 0.0f || uu + vv > 1.0f) return false;
    const float tt = (e2.x * q.x + e2.y * q.y + e2.z * q.z) * inv_det;
    if (tt >= 0.0f)
    {
        *t = tt;
        *u = uu;
        *v = vv;
        return true;
    }
    return false;
}
// scene.cuh:1054-1117 (vertex-only fallback outside the triangle, quirk Q10)
SNCH_DI bool sphere_triangle(V3 p1, V3 p2, V3 p3, V3 center, float radius)
{
    const V3 e1 = p2 - p1, e2 = p3 - p1;
    V3 n = V3{e1.y * e2.z - e1.z * e2.y, e1.z * e2.x - e1.x * e2.z, e1.x * e2.y - e1.y * e2.x};
    const float nl = sqrtf(n.x * n.x + n.y * n.y + n.z * n.z);
    n = V3{n.x / nl, n.y / nl, n.z / nl};
    const float d = n.x * p1.x + n.y * p1.y + n.z * p1.z;
    const float dist_to_plane = n.x * center.x + n.y * center.y + n.z * center.z - d;
    const V3 proj = V3{center.x - dist_to_plane * n.x, center.y - dist_to_plane * n.y, center.z - dist_to_plane * n.z};
    const V3 v0 = p3 - p1, v1 = p2 - p1, v2 = proj - p1;
    const float dot00 = v0.x * v0.x + v0.y * v0.y + v0.z * v0.z;
    const float dot01 = v0.x * v1.x + v0.y * v1.y + v0.z * v1.z;
    const float dot02 = v0.x * v2.x + v0.y * v2.y + v0.z * v2.z;
    const float dot11 = v1.x * v1.x + v1.y * v1.y + v1.z * v1.z;
    const float dot12 = v1.x * v2.x + v1.y * v2.y + v1.z * v2.z;
    const float inv = 1.0f / (dot00 * dot11 - dot01 * dot01);
    const float u = (dot11 * dot02 - dot01 * dot12) * inv;
    const float v = (dot00 * dot12 - dot01 * dot02) * inv;
    if (u >= 0 && v >= 0 && u + v <= 1) return fabsf(dist_to_plane) <= radius;
    V3 cp = proj;
    if (u < 0) cp = p1;
    else if (v < 0) cp = p3;
    else if (u + v > 1) cp = p2;
    const float dx = cp.x - center.x, dy = cp.y - center.y, dz = cp.z - center.z;
    return dx * dx + dy * dy + dz * dz <= radius * radius;
}
SNCH_DI float triangle_area(V3 a, V3 b, V3 c) { return len(cross(c - a, b - a)) / 2; } // scene.cuh:855-868
SNCH_DI float green_weight3(V3 x, V3 y)                                                 // scene.cuh:1119-1126
{
    const float r = std_max(len(x - y), 1e-4f);
    return 1.0f / (kPi * 4.0f * r);
}

} // namespace snch
