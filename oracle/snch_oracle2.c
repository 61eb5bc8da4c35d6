/* TEST INFRASTRUCTURE — CPU oracle for the 2-D instantiation of the SNCH-LBVH hot path: lbvh::scene<2> (polylines: line
 * segments / silhouette vertices).  Plain C restatement of the reference algorithm; every function cites the reference
 * file:line (relative to /root/reference/include/snch_lbvh/) it follows.  Only tests/ may load this.
 *
 * Parity status: PINNED — bit-identical trees and query results to the UNMODIFIED reference headers executed on the CPU
 * (oracle/_ref/libsnch_ref_cpu.so, ref2_* in oracle/ref_driver.cu) and to the golden vectors that library produced
 * (tests/golden/poly_*.npz): tests/test_scene2_cpu.py::test_oracle2_is_the_reference.
 *
 * Defined where the reference is undefined, as in the 3-D oracle: Q1 (cone union beyond pi: half_angle = pi), Q6 (a
 * one-segment tree answers from its only leaf), Q19 (pdf = 0 on a sampling miss).                      */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI_F 3.14159265358979323846f
#define ORC_PI_2_F 1.57079632679489661923f
#define NONE 0xFFFFFFFFu
#define BVH_OFFSET 1e-3f /* scene.cuh:12 */

typedef struct { float x, y; } f2;
typedef struct { f2 upper, lower; } box2;               /* aabb.cuh:12-15 */
typedef struct { f2 axis; float half_angle, radius; } cone2; /* cone.cuh:10-16 */
typedef struct { uint32_t parent, left, right, object; } node_t; /* bvh.cuh:27-33 */

typedef struct orc_scene2
{
    int nV, nS;
    f2 *verts;
    int *segs;   /* nS x 2 */
    int *vert4;  /* nV x 4: silhouette_vertex::indices */
    int *owned;  /* nS x 2: line_segment::silhouette_indices */
    node_t *nodes;
    box2 *aabbs;
    cone2 *cones;
    uint8_t *q1;
    int collision;
} orc_scene2;

static inline float std_min(float a, float b) { return (b < a) ? b : a; }
static inline float std_max(float a, float b) { return (a < b) ? b : a; }
static inline f2 mk2(float x, float y) { f2 r = {x, y}; return r; }
static inline f2 sub2(f2 a, f2 b) { return mk2(a.x - b.x, a.y - b.y); }
static inline float dot2(f2 a, f2 b) { return a.x * b.x + a.y * b.y; }
static inline float len2(f2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
static inline float sqlen2(f2 a) { return a.x * a.x + a.y * a.y; }
static inline f2 normalize2(f2 v) { const float n = len2(v); return mk2(v.x / n, v.y / n); }

/* ---- boxes: aabb.cuh:104-112, 130-135, 160-181, 263-270, 367-396, 434-449 -------------------------------------------- */
static box2 box_merge(box2 l, box2 r)
{
    box2 m;
    m.upper = mk2(fmaxf(l.upper.x, r.upper.x), fmaxf(l.upper.y, r.upper.y));
    m.lower = mk2(fminf(l.lower.x, r.lower.x), fminf(l.lower.y, r.lower.y));
    return m;
}
static f2 box_centroid(box2 b) { return mk2((b.upper.x + b.lower.x) * 0.5f, (b.upper.y + b.lower.y) * 0.5f); }
static float box_mindist(box2 b, f2 p)
{
    const float dx = fminf(b.upper.x, fmaxf(b.lower.x, p.x)) - p.x, dy = fminf(b.upper.y, fmaxf(b.lower.y, p.y)) - p.y;
    return dx * dx + dy * dy;
}
static float box_minmaxdist(box2 b, f2 p)
{
    float rmx = (b.lower.x - p.x) * (b.lower.x - p.x), rmy = (b.lower.y - p.y) * (b.lower.y - p.y);
    float rMx = (b.upper.x - p.x) * (b.upper.x - p.x), rMy = (b.upper.y - p.y) * (b.upper.y - p.y);
    if ((b.upper.x + b.lower.x) * 0.5f < p.x) { const float t = rmx; rmx = rMx; rMx = t; }
    if ((b.upper.y + b.lower.y) * 0.5f < p.y) { const float t = rmy; rmy = rMy; rMy = t; }
    return fminf(rmx + rMy, rMx + rmy);
}
static int box_ray(f2 org, f2 dinv, box2 b, float max_dist, float *dist)
{
    float t1 = (b.lower.x - org.x) * dinv.x, t2 = (b.upper.x - org.x) * dinv.x;
    float tmin = fminf(t1, t2), tmax = fmaxf(t1, t2);
    t1 = (b.lower.y - org.y) * dinv.y;
    t2 = (b.upper.y - org.y) * dinv.y;
    tmin = fmaxf(tmin, fminf(t1, t2));
    tmax = fminf(tmax, fmaxf(t1, t2));
    if (tmax >= tmin && tmax >= 0.0f && tmin <= max_dist)
    {
        *dist = tmin >= 0.0f ? tmin : 0.0f;
        return 1;
    }
    return 0;
}
static int box_sphere(f2 c, float radius, box2 b)
{
    const float cx = std_max(b.lower.x, std_min(c.x, b.upper.x)), cy = std_max(b.lower.y, std_min(c.y, b.upper.y));
    const float dx = cx - c.x, dy = cy - c.y;
    return dx * dx + dy * dy <= radius * radius;
}

/* ---- cones: cone.cuh:18-22, 48-53, 78-121, 262-271, 374-425 ------------------------------------------------------------ */
static inline int cone_valid(const cone2 *c) { return c->half_angle >= 0.0f; }
static inline int inrange(float v, float lo, float hi) { return v >= lo && v <= hi; }
static float project_to_plane(f2 n, f2 e)
{
    const f2 b = mk2(-n.y, n.x);
    return fabsf(dot2(e, mk2(fabsf(b.x), fabsf(b.y))));
}
static int cone_overlap(const cone2 *bc, f2 o, box2 b, float dist2_to_box)
{
    if (bc->half_angle >= ORC_PI_2_F || dist2_to_box < FLT_EPSILON) return 1;
    const f2 c = box_centroid(b);
    f2 v = sub2(c, o);
    const float l = len2(v);
    v.x /= l;
    v.y /= l;
    const float d_axis_angle = acosf(std_max(-1.0f, std_min(1.0f, dot2(bc->axis, v))));
    if (inrange(ORC_PI_2_F, d_axis_angle - bc->half_angle, d_axis_angle + bc->half_angle)) return 1;
    if (l > bc->radius)
    {
        const float sum = bc->half_angle + asinf(bc->radius / l);
        return sum >= ORC_PI_2_F ? 1 : inrange(ORC_PI_2_F, d_axis_angle - sum, d_axis_angle + sum);
    }
    const f2 e = sub2(b.upper, c);
    float d = dot2(e, mk2(fabsf(v.x), fabsf(v.y)));
    const float s = l - d;
    if (s <= 0.0f) return 1;
    d = project_to_plane(v, e);
    const float sum = bc->half_angle + atan2f(d, s);
    return sum >= ORC_PI_2_F ? 1 : inrange(ORC_PI_2_F, d_axis_angle - sum, d_axis_angle + sum);
}
static f2 rotate2(f2 u, f2 v, float theta)
{
    const float det = u.x * v.y - u.y * v.x;
    theta *= copysignf(1.0f, det);
    const float ct = cosf(theta), st = sinf(theta);
    return mk2(ct * u.x - st * u.y, st * u.x + ct * u.y);
}
static cone2 cone_merge(const cone2 *ca, const cone2 *cb, f2 oa, f2 ob, f2 on, int *q1)
{
    cone2 r;
    memset(&r, 0, sizeof r);
    *q1 = 0;
    if (cone_valid(ca) && cone_valid(cb))
    {
        f2 axis_a = ca->axis, axis_b = cb->axis;
        float ha = ca->half_angle, hb = cb->half_angle;
        r.radius = sqrtf(std_max(ca->radius * ca->radius + sqlen2(sub2(on, oa)), cb->radius * cb->radius + sqlen2(sub2(on, ob))));
        if (hb > ha)
        {
            const f2 t = axis_a; axis_a = axis_b; axis_b = t;
            const float th = ha; ha = hb; hb = th;
        }
        const float theta = acosf(std_max(-1.0f, std_min(1.0f, dot2(axis_a, axis_b))));
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
            r.half_angle = ORC_PI_F; /* Q1: left uninitialised by the reference (cone.cuh:401-405) */
            *q1 = 1;
            return r;
        }
        r.axis = rotate2(axis_a, axis_b, o_theta - ha);
        r.half_angle = o_theta;
    }
    else if (cone_valid(ca)) r = *ca;
    else if (cone_valid(cb)) r = *cb;
    else r.half_angle = -ORC_PI_F;
    return r;
}

/* ---- silhouette vertices and segments: scene.cuh:112-141, 176-200, 291-385, 396-619 ----------------------------------------- */
static inline int vert_has_face(const int *v4, int f) { return f == 0 ? v4[2] != -1 : v4[0] != -1; }
static f2 vert_face_normal(const orc_scene2 *s, const int *v4, int f, int do_normalize)
{
    const int i = f == 0 ? 1 : 0;
    const f2 pa = s->verts[v4[i]], pb = s->verts[v4[i + 1]];
    const f2 sg = sub2(pb, pa), n = mk2(sg.y, -sg.x);
    return do_normalize ? normalize2(n) : n;
}
static f2 vert_normal(const orc_scene2 *s, const int *v4)
{
    f2 n = mk2(0.0f, 0.0f);
    if (vert_has_face(v4, 0)) { const f2 a = vert_face_normal(s, v4, 0, 0); n = mk2(n.x + a.x, n.y + a.y); }
    if (vert_has_face(v4, 1)) { const f2 a = vert_face_normal(s, v4, 1, 0); n = mk2(n.x + a.x, n.y + a.y); }
    return normalize2(n);
}
static int is_silhouette_vertex(f2 n0, f2 n1, f2 view, float d, int flip)
{
    const float precision = 1e-3f, sign = flip ? 1.0f : -1.0f;
    if (d <= precision) return sign * (n0.x * n1.y - n0.y * n1.x) > precision;
    const f2 u = mk2(view.x / d, view.y / d);
    const float dot0 = dot2(u, n0), dot1 = dot2(u, n1);
    if (fabsf(dot0) <= precision) return sign * dot1 > precision;
    if (fabsf(dot1) <= precision) return sign * dot0 > precision;
    return dot0 * dot1 < 0.0f;
}
static int vert_closest_silhouette(const orc_scene2 *s, const int *v4, f2 origin, float max_r2, float *distance, int flip, float min_r2)
{
    if (min_r2 >= max_r2) return 0;
    const f2 view = sub2(origin, s->verts[v4[1]]);
    const float d = len2(view);
    if (d * d > max_r2) return 0;
    int is_sil = !vert_has_face(v4, 0) || !vert_has_face(v4, 1);
    if (!is_sil) is_sil = is_silhouette_vertex(vert_face_normal(s, v4, 0, 1), vert_face_normal(s, v4, 1, 1), view, d, flip);
    if (is_sil && d * d <= max_r2)
    {
        *distance = d;
        return 1;
    }
    return 0;
}
/* `vertex` (optional): the owned silhouette vertex that set *distance ("TODO: identify nearest index", query.cuh:386,411) */
static int seg_closest_silhouette(const orc_scene2 *s, int seg, f2 origin, float max_r2, float *distance, int flip, float min_r2, int *vertex)
{
    int ret = 0;
    for (int i = 0; i < 2; ++i)
    {
        const int v = s->owned[2 * seg + i];
        if (v == -1) continue;
        if (vert_closest_silhouette(s, s->vert4 + 4 * v, origin, max_r2, distance, flip, min_r2))
        {
            ret = 1;
            max_r2 = *distance * *distance;
            if (vertex) *vertex = v;
        }
    }
    return ret;
}
static float seg_distance(const orc_scene2 *s, int seg, f2 x)
{ /* find_closest_point_line_segment, scene.cuh:176-200 */
    const f2 pa = s->verts[s->segs[2 * seg]], pb = s->verts[s->segs[2 * seg + 1]];
    const f2 u = sub2(pb, pa), v = sub2(x, pa);
    const float c1 = dot2(u, v);
    if (c1 <= 0.0f) return len2(sub2(x, pa));
    const float c2 = dot2(u, u);
    if (c2 <= c1) return len2(sub2(x, pb));
    const float t = c1 / c2;
    return len2(sub2(x, mk2(pa.x + u.x * t, pa.y + u.y * t)));
}
static int seg_ray(const orc_scene2 *s, int seg, f2 org, f2 dir, float *t_out, float *s_out)
{ /* scene<2>::intersect_test, scene.cuh:543-577 (host build: 1 / D) */
    const f2 p0 = s->verts[s->segs[2 * seg]], p1 = s->verts[s->segs[2 * seg + 1]];
    const f2 sd = sub2(p1, p0);
    const float D = dir.x * (-sd.y) + dir.y * sd.x;
    if (fabsf(D) < FLT_EPSILON) return 0;
    const float inv = 1.0f / D;
    const float t = ((p0.x - org.x) * (-sd.y) - (p0.y - org.y) * (-sd.x)) * inv;
    const float ss = (dir.x * (p0.y - org.y) - dir.y * (p0.x - org.x)) * inv;
    if (ss >= -1e-3f && ss <= 1.0f + 1e-3f && t >= 0.0f)
    {
        *t_out = t;
        *s_out = ss;
        return 1;
    }
    return 0;
}
static int seg_sphere(const orc_scene2 *s, int seg, f2 c, float radius)
{ /* scene<2>::intersect_sphere, scene.cuh:579-604 */
    const f2 p1 = s->verts[s->segs[2 * seg]], p2 = s->verts[s->segs[2 * seg + 1]];
    const f2 d = sub2(p2, p1);
    float t = ((c.x - p1.x) * d.x + (c.y - p1.y) * d.y) / (d.x * d.x + d.y * d.y);
    t = std_max(0.0f, std_min(1.0f, t));
    const float dx = (p1.x + t * d.x) - c.x, dy = (p1.y + t * d.y) - c.y;
    return dx * dx + dy * dy <= radius * radius;
}
static float seg_length(const orc_scene2 *s, int seg) { return len2(sub2(s->verts[s->segs[2 * seg]], s->verts[s->segs[2 * seg + 1]])); }
static float green_weight(f2 x, f2 y)
{ /* scene.cuh:606-613 */
    const float r = std_max(len2(sub2(x, y)), 1e-2f);
    return fabsf(logf(r) / (ORC_PI_F * 2.0f));
}
static box2 seg_box(const orc_scene2 *s, int seg)
{ /* aabb_getter, scene.cuh:419-432 */
    const f2 p0 = s->verts[s->segs[2 * seg]], p1 = s->verts[s->segs[2 * seg + 1]];
    box2 b;
    b.upper = mk2(std_max(p0.x, p1.x) + BVH_OFFSET, std_max(p0.y, p1.y) + BVH_OFFSET);
    b.lower = mk2(std_min(p0.x, p1.x) - BVH_OFFSET, std_min(p0.y, p1.y) - BVH_OFFSET);
    return b;
}
static cone2 seg_cone(const orc_scene2 *s, int seg)
{ /* cone_getter, scene.cuh:434-500 */
    const f2 bc = box_centroid(seg_box(s, seg));
    cone2 r;
    r.axis = mk2(0.0f, 0.0f);
    r.half_angle = ORC_PI_F;
    r.radius = 0.0f;
    int any = 0, two_sided = 1;
    for (int i = 0; i < 2; ++i)
    {
        const int v = s->owned[2 * seg + i];
        if (v == -1) continue;
        const int *v4 = s->vert4 + 4 * v;
        const f2 n = vert_normal(s, v4);
        r.axis.x += n.x;
        r.axis.y += n.y;
        r.radius = std_max(r.radius, len2(sub2(s->verts[v4[1]], bc)));
        two_sided = two_sided && vert_has_face(v4, 0) && vert_has_face(v4, 1);
        any = 1;
    }
    if (!any) r.half_angle = -ORC_PI_F;
    else if (!two_sided) r.half_angle = ORC_PI_F;
    else
    {
        const float norm = len2(r.axis);
        if (norm > FLT_EPSILON)
        {
            r.axis.x /= norm;
            r.axis.y /= norm;
            r.half_angle = 0.0f;
            for (int i = 0; i < 2; ++i)
            {
                const int v = s->owned[2 * seg + i];
                if (v == -1) continue;
                const int *v4 = s->vert4 + 4 * v;
                for (int f = 0; f < 2; ++f)
                {
                    const f2 n = vert_has_face(v4, f) ? vert_face_normal(s, v4, f, 1) : mk2(0.0f, 0.0f);
                    r.half_angle = std_max(r.half_angle, acosf(std_max(-1.0f, std_min(1.0f, dot2(r.axis, n)))));
                }
            }
        }
    }
    return r;
}

/* ---- construct(): bvh.cuh:109-229, 380-613; morton_code.cuh:19-38, 84-92, 144-167 ------------------------------------------ */
static uint32_t expand_bits(uint32_t v)
{
    v = (v | (v << 16)) & 0x070000FFu;
    v = (v | (v << 8)) & 0x0700F00Fu;
    v = (v | (v << 4)) & 0x430C30C3u;
    v = (v | (v << 2)) & 0x49249249u;
    return v;
}
static uint32_t morton2(float x, float y)
{
    x = fminf(fmaxf(x * 1024.0f, 0.0f), 1023.0f);
    y = fminf(fmaxf(y * 1024.0f, 0.0f), 1023.0f);
    return expand_bits((uint32_t)x) * 2 + expand_bits((uint32_t)y);
}
typedef struct { const uint32_t *k32; const uint64_t *k64; int n; } keyview;
static inline int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline int clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }
static inline int kdelta(const keyview *kv, int i, int j)
{
    if (j < 0 || j >= kv->n) return -1;
    return kv->k64 ? clz64(kv->k64[i] ^ kv->k64[j]) : clz32(kv->k32[i] ^ kv->k32[j]);
}
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
    if (d < 0) { const int t = idx; idx = j; j = t; }
    *first = idx;
    *last = j;
}
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
static int cmp_u64(const void *a, const void *b)
{
    const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : (x > y);
}
static void construct(orc_scene2 *s)
{
    const int n = s->nS;
    if (n <= 0) return;
    const int ni = n - 1, nn = 2 * n - 1;
    s->nodes = malloc(sizeof(node_t) * (size_t)nn);
    s->aabbs = malloc(sizeof(box2) * (size_t)nn);
    s->cones = malloc(sizeof(cone2) * (size_t)nn);
    s->q1 = calloc((size_t)nn, 1);
    box2 *leaf = malloc(sizeof(box2) * (size_t)n);
    box2 whole;
    whole.upper = mk2(-INFINITY, -INFINITY);
    whole.lower = mk2(INFINITY, INFINITY);
    for (int i = 0; i < n; ++i)
    {
        leaf[i] = seg_box(s, i);
        whole = box_merge(whole, leaf[i]);
    }
    uint64_t *keys = malloc(sizeof(uint64_t) * (size_t)n);
    for (int i = 0; i < n; ++i)
    { /* default_morton_code_calculator<Real, 2, Object>, bvh.cuh:258-281 */
        f2 p = box_centroid(leaf[i]);
        p.x -= whole.lower.x;
        p.y -= whole.lower.y;
        p.x /= (whole.upper.x - whole.lower.x);
        p.y /= (whole.upper.y - whole.lower.y);
        keys[i] = ((uint64_t)morton2(p.x, p.y) << 32) | (uint32_t)i; /* sorting (code << 32 | index) is the stable order by code */
    }
    qsort(keys, (size_t)n, sizeof(uint64_t), cmp_u64);
    uint32_t *k32 = malloc(sizeof(uint32_t) * (size_t)n);
    s->collision = 0;
    for (int k = 0; k < n; ++k)
    {
        k32[k] = (uint32_t)(keys[k] >> 32);
        if (k > 0 && k32[k] == k32[k - 1]) s->collision = 1;
    }
    for (int i = 0; i < nn; ++i) s->nodes[i].parent = s->nodes[i].left = s->nodes[i].right = s->nodes[i].object = NONE;
    for (int k = 0; k < n; ++k)
    {
        const int obj = (int)(keys[k] & 0xFFFFFFFFu);
        s->aabbs[ni + k] = leaf[obj];
        s->cones[ni + k] = seg_cone(s, obj);
        s->nodes[ni + k].object = (uint32_t)obj;
    }
    keyview kv;
    kv.n = n;
    kv.k32 = s->collision ? NULL : k32;
    kv.k64 = s->collision ? keys : NULL;
    for (int i = 0; i < ni; ++i)
    {
        int first, last;
        determine_range(&kv, i, &first, &last);
        const int gamma = find_split(&kv, first, last);
        uint32_t l = (uint32_t)gamma, r = (uint32_t)gamma + 1;
        if (first == gamma) l += (uint32_t)ni;
        if (last == gamma + 1) r += (uint32_t)ni;
        s->nodes[i].left = l;
        s->nodes[i].right = r;
        s->nodes[l].parent = (uint32_t)i;
        s->nodes[r].parent = (uint32_t)i;
    }
    int *flags = calloc((size_t)(ni > 0 ? ni : 1), sizeof(int));
    for (int k = 0; k < n; ++k)
    { /* bvh.cuh:520-554 */
        uint32_t p = s->nodes[ni + k].parent;
        while (p != NONE)
        {
            if (flags[p]++ == 0) break;
            s->aabbs[p] = box_merge(s->aabbs[s->nodes[p].left], s->aabbs[s->nodes[p].right]);
            p = s->nodes[p].parent;
        }
    }
    memset(flags, 0, sizeof(int) * (size_t)(ni > 0 ? ni : 1));
    for (int k = 0; k < n; ++k)
    { /* bvh.cuh:556-604 */
        uint32_t p = s->nodes[ni + k].parent;
        while (p != NONE)
        {
            if (flags[p]++ == 0) break;
            const uint32_t l = s->nodes[p].left, r = s->nodes[p].right;
            int q1;
            s->cones[p] = cone_merge(&s->cones[l], &s->cones[r], box_centroid(s->aabbs[l]), box_centroid(s->aabbs[r]), box_centroid(s->aabbs[p]), &q1);
            s->q1[p] = (uint8_t)(q1 || s->q1[l] || s->q1[r]);
            p = s->nodes[p].parent;
        }
    }
    free(flags);
    free(leaf);
    free(keys);
    free(k32);
}

/* ---- API ----------------------------------------------------------------------------------------------------------------- */
orc_scene2 *orc_scene2_create(const float *xy, int n_verts, const int *seg, int n_segs)
{
    orc_scene2 *s = calloc(1, sizeof *s);
    s->nV = n_verts;
    s->nS = n_segs;
    s->verts = malloc(sizeof(f2) * (size_t)(n_verts ? n_verts : 1));
    s->segs = malloc(sizeof(int) * 2 * (size_t)(n_segs ? n_segs : 1));
    s->vert4 = malloc(sizeof(int) * 4 * (size_t)(n_verts ? n_verts : 1));
    s->owned = malloc(sizeof(int) * 2 * (size_t)(n_segs ? n_segs : 1));
    for (int i = 0; i < n_verts; ++i) s->verts[i] = mk2(xy[2 * i], xy[2 * i + 1]);
    memcpy(s->segs, seg, sizeof(int) * 2 * (size_t)n_segs);
    /* compute_silhouettes, scene.cuh:629-656: later segments overwrite a vertex's previous / next slots */
    for (int i = 0; i < 4 * n_verts; ++i) s->vert4[i] = -1;
    for (int i = 0; i < n_segs; ++i)
    {
        const int a = seg[2 * i], b = seg[2 * i + 1];
        s->vert4[4 * a + 1] = a;
        s->vert4[4 * a + 2] = b;
        s->vert4[4 * b + 0] = a;
        s->vert4[4 * b + 1] = b;
    }
    /* build_bvh, scene.cuh:658-681: a vertex is owned by the first segment in input order that touches it */
    char *seen = calloc((size_t)(n_verts ? n_verts : 1), 1);
    for (int i = 0; i < n_segs; ++i)
    {
        int k = 0;
        s->owned[2 * i] = s->owned[2 * i + 1] = -1;
        for (int j = 0; j < 2; ++j)
        {
            const int v = seg[2 * i + j];
            if (!seen[v])
            {
                seen[v] = 1;
                s->owned[2 * i + k++] = v;
            }
        }
    }
    free(seen);
    construct(s);
    return s;
}
void orc_scene2_destroy(orc_scene2 *s)
{
    if (!s) return;
    free(s->verts); free(s->segs); free(s->vert4); free(s->owned); free(s->nodes); free(s->aabbs); free(s->cones); free(s->q1);
    free(s);
}
int orc2_num_nodes(const orc_scene2 *s) { return s->nS ? 2 * s->nS - 1 : 0; }
int orc2_collision(const orc_scene2 *s) { return s->collision; }
void orc2_export_tree(const orc_scene2 *s, uint32_t *nodes, float *aabbs, float *cones, uint8_t *q1)
{
    const size_t nn = (size_t)orc2_num_nodes(s);
    if (nodes) memcpy(nodes, s->nodes, sizeof(node_t) * nn);
    if (aabbs) memcpy(aabbs, s->aabbs, sizeof(box2) * nn);
    if (cones) memcpy(cones, s->cones, sizeof(cone2) * nn);
    if (q1) memcpy(q1, s->q1, nn);
}
void orc2_export_adjacency(const orc_scene2 *s, int *vert4, int *owned2)
{
    memcpy(vert4, s->vert4, sizeof(int) * 4 * (size_t)s->nV);
    memcpy(owned2, s->owned, sizeof(int) * 2 * (size_t)s->nS);
}

typedef struct { uint32_t node; float key; } stack_entry;
#define STACK_CAP 256
/* query.cuh:238-318 */
void orc2_closest(const orc_scene2 *s, const float *q, long n, uint32_t *idx, float *dist)
{
    for (long i = 0; i < n; ++i)
    {
        const f2 p = mk2(q[2 * i], q[2 * i + 1]);
        if (s->nS == 0) { idx[i] = NONE; dist[i] = INFINITY; continue; }
        if (s->nS == 1) { idx[i] = 0; const float d = seg_distance(s, 0, p); dist[i] = sqrtf(d * d); continue; } /* Q6 */
        stack_entry st[STACK_CAP];
        int sp = 0;
        st[sp].node = 0; st[sp].key = box_mindist(s->aabbs[0], p); ++sp;
        uint32_t nearest = NONE;
        float best = INFINITY;
        do
        {
            const stack_entry e = st[--sp];
            if (e.key > best) continue;
            const uint32_t L = s->nodes[e.node].left, R = s->nodes[e.node].right;
            const float Lmin = box_mindist(s->aabbs[L], p), Rmin = box_mindist(s->aabbs[R], p);
            const float Lmm = box_minmaxdist(s->aabbs[L], p), Rmm = box_minmaxdist(s->aabbs[R], p);
            const uint32_t ch[2] = {L, R};
            const float mn[2] = {Lmin, Rmin};
            const int go[2] = {Lmin <= Rmm, Rmin <= Lmm};
            for (int c = 0; c < 2; ++c)
            {
                if (!go[c]) continue;
                const uint32_t obj = s->nodes[ch[c]].object;
                if (obj != NONE)
                {
                    float d = seg_distance(s, (int)obj, p);
                    d *= d;
                    if (d <= best) { best = d; nearest = obj; }
                }
                else { st[sp].node = ch[c]; st[sp].key = mn[c]; ++sp; }
            }
        } while (sp > 0);
        idx[i] = nearest;
        dist[i] = sqrtf(best);
    }
}
/* query.cuh:325-423; r_max may be NULL (the reference's unbounded search) */
/* vertex / point (optional): the silhouette vertex that attains dist[i] (-1 = none) and its position (silhouette_vertex::
 * find_closest_silhouette_point's `p`, scene.cuh:354-356) */
void orc2_silhouette_ex(const orc_scene2 *s, const float *q, long n, int flip, const float *r_max, float *dist, int *vertex, float *point)
{
    for (long i = 0; i < n; ++i)
    {
        const f2 p = mk2(q[2 * i], q[2 * i + 1]);
        float best = r_max ? r_max[i] : INFINITY;
        int found_any = 0, best_v = -1;
        dist[i] = INFINITY;
        if (vertex) vertex[i] = -1;
        if (point) point[2 * i] = point[2 * i + 1] = 0.0f;
        if (s->nS == 0) continue;
        if (s->nS == 1)
        { /* Q6 */
            float d = INFINITY;
            const float m = box_mindist(s->aabbs[0], p);
            if (m <= best * best && cone_valid(&s->cones[0]) && cone_overlap(&s->cones[0], p, s->aabbs[0], m) &&
                seg_closest_silhouette(s, 0, p, best * best, &d, flip, 0.0f, &best_v) && d <= best)
            {
                dist[i] = d;
                if (vertex) vertex[i] = best_v;
                if (point) { const f2 vp = s->verts[s->vert4[4 * best_v + 1]]; point[2 * i] = vp.x; point[2 * i + 1] = vp.y; }
            }
            continue;
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
                if (obj != NONE)
                {
                    float d = INFINITY;
                    int v_at = -1;
                    if (seg_closest_silhouette(s, (int)obj, p, best * best, &d, flip, 0.0f, &v_at) && d <= best) { best = d; found_any = 1; best_v = v_at; }
                }
                else { st[sp].node = ch[c]; st[sp].key = md[c]; ++sp; }
            }
        } while (sp > 0);
        if (found_any)
        {
            dist[i] = best;
            if (vertex) vertex[i] = best_v;
            if (point) { const f2 vp = s->verts[s->vert4[4 * best_v + 1]]; point[2 * i] = vp.x; point[2 * i + 1] = vp.y; }
        }
    }
}
void orc2_silhouette(const orc_scene2 *s, const float *q, long n, int flip, const float *r_max, float *dist)
{
    orc2_silhouette_ex(s, q, n, flip, r_max, dist, NULL, NULL);
}
/* query.cuh:79-169 */
void orc2_ray(const orc_scene2 *s, const float *org, const float *dir, const float *tmax, long n, int *found, float *t, float *sp_out, uint32_t *prim)
{
    for (long i = 0; i < n; ++i)
    {
        const f2 o = mk2(org[2 * i], org[2 * i + 1]), d = mk2(dir[2 * i], dir[2 * i + 1]);
        const float max_dist = tmax ? tmax[i] : INFINITY;
        found[i] = 0; t[i] = INFINITY; sp_out[i] = 0.0f; prim[i] = NONE;
        if (s->nS == 0) continue;
        const f2 dinv = mk2(1 / d.x, 1 / d.y);
        stack_entry st[STACK_CAP];
        int sp = 0;
        st[sp].node = 0; st[sp].key = INFINITY; ++sp;
        float best = INFINITY;
        do
        {
            const stack_entry e = st[--sp];
            if (e.key > best) continue;
            const uint32_t obj = s->nodes[e.node].object;
            if (obj != NONE)
            {
                float tt, ss;
                if (seg_ray(s, (int)obj, o, d, &tt, &ss) && tt < max_dist && tt < best)
                {
                    best = tt; found[i] = 1; sp_out[i] = ss; prim[i] = obj;
                }
            }
            else
            {
                const uint32_t L = s->nodes[e.node].left, R = s->nodes[e.node].right;
                float Ld, Rd;
                const int Lh = box_ray(o, dinv, s->aabbs[L], max_dist, &Ld), Rh = box_ray(o, dinv, s->aabbs[R], max_dist, &Rd);
                if (Lh && Rh)
                {
                    uint32_t closer = L, other = R;
                    if (Rd < Ld) { const float tf = Ld; Ld = Rd; Rd = tf; closer = R; other = L; }
                    st[sp].node = other; st[sp].key = Rd; ++sp;
                    st[sp].node = closer; st[sp].key = Ld; ++sp;
                }
                else if (Lh) { st[sp].node = L; st[sp].key = Ld; ++sp; }
                else if (Rh) { st[sp].node = R; st[sp].key = Rd; ++sp; }
            }
        } while (sp > 0);
        t[i] = best;
    }
}
/* sample.cuh:23-92; circles: x, y, radius */
void orc2_sample(const orc_scene2 *s, const float *sph3, const float *u_in, long n, int *idx, float *pdf)
{
    for (long i = 0; i < n; ++i)
    {
        const f2 c = mk2(sph3[3 * i], sph3[3 * i + 1]);
        const float radius = sph3[3 * i + 2];
        float u = u_in[i];
        idx[i] = -1; pdf[i] = 0.0f;
        if (s->nS == 0) continue;
        uint32_t node = 0;
        float path = 1.0f;
        for (;;)
        {
            const uint32_t obj = s->nodes[node].object;
            if (obj != NONE)
            {
                if (seg_sphere(s, (int)obj, c, radius)) { idx[i] = (int)obj; pdf[i] = path / seg_length(s, (int)obj); }
                break;
            }
            const uint32_t L = s->nodes[node].left, R = s->nodes[node].right;
            const float Lw = box_sphere(c, radius, s->aabbs[L]) ? green_weight(c, box_centroid(s->aabbs[L])) : 0;
            const float Rw = box_sphere(c, radius, s->aabbs[R]) ? green_weight(c, box_centroid(s->aabbs[R])) : 0;
            const float total = Lw + Rw;
            if (!(total > 0)) break;
            const float Lp = Lw / total;
            if (u < Lp) { u /= Lp; node = L; path = Lp * path; }
            else { const float Rp = 1.0f - Lp; u = (u - Lp) / Rp; node = R; path = Rp * path; }
        }
    }
}
