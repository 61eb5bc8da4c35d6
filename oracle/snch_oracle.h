/* TEST INFRASTRUCTURE — CPU oracle for the SNCH-LBVH hot path (plain C restatement of the reference
 * algorithm, see snch_oracle.c).  Only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() may
 * load this; the product library never links it and has no CPU fallback.
 *
 * Parity status: PINNED.  The reference ships no golden vectors, so this restatement is pinned
 * against the reference's own code executed here (oracle/_ref/libsnch_ref_cpu.so, tests/test_oracle_pinning.py)
 * and against fixtures generated from it (tests/golden/, tests/golden/make_golden.py). */
#ifndef SNCH_ORACLE_H
#define SNCH_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene orc_scene;

/* scene<3>(verts, indices) -> compute_silhouettes() -> build_bvh()   (scene.cuh:1128-1229, bvh.cuh:380-613) */
orc_scene *orc_scene3_create(const float *xyz, int n_verts, const int *tri, int n_tris);
void orc_scene_destroy(orc_scene *s);

int orc_num_objects(const orc_scene *s);
int orc_num_nodes(const orc_scene *s);
int orc_num_edges(const orc_scene *s);
int orc_collision(const orc_scene *s); /* 1 if the reference would take its 64-bit key path (bvh.cuh:464) */

/* reference-layout arrays: nodes 4 x u32 {parent,left,right,object}; aabbs 6 f32 {upper xyz, lower xyz};
 * cones 5 f32 {axis xyz, half_angle, radius}.  Any pointer may be NULL. */
void orc_export_tree(const orc_scene *s, uint32_t *nodes, float *aabbs, float *cones);
/* q1[i] = 1 where cone i depends on the reference's uninitialised half_angle; the oracle defines pi there */
void orc_export_q1_taint(const orc_scene *s, uint8_t *q1);
void orc_export_adjacency(const orc_scene *s, int *edges4, int *tri_edges3, int *tri_owned3);
void orc_export_morton(const orc_scene *s, uint32_t *morton_sorted, uint32_t *sorted_idx);
void orc_export_ranges(const orc_scene *s, uint32_t *first_last); /* 2 x (N-1): Karras [first,last] per internal node */

/* ---- queries: faithful restatements of the reference traversals (same visiting order, same tie rules) ---- */
void orc_closest(const orc_scene *s, const float *q, long n, uint32_t *idx, float *dist, int nthreads);
/* r_max may be NULL (reference behaviour: unbounded).  With r_max the search starts from best = r_max and
 * returns +inf when nothing is found inside. */
void orc_silhouette(const orc_scene *s, const float *q, long n, int flip, const float *r_max, float *dist, int nthreads);
/* the same walk, also returning the silhouette edge that attains the distance (-1 = none) and the closest point on it (the two
 * values the reference computes and drops: scene.cuh:796-799, query.cuh:386,411) */
void orc_silhouette_ex(const orc_scene *s, const float *q, long n, int flip, const float *r_max, float *dist, int *edge, float *point, int nthreads);
void orc_point_edge_distance(const orc_scene *s, const float *q, const int *edge, long n, float *dist, float *point);
void orc_ray(const orc_scene *s, const float *org, const float *dir, const float *tmax, long n, int any_hit, int *found,
             float *t, float *uv, uint32_t *prim, int nthreads);
void orc_sample(const orc_scene *s, const float *sph4, const float *u, long n, int *idx, float *pdf, int nthreads);
void orc_sample_on_object(const orc_scene *s, const int *idx, const float *u, const float *v, long n, float *xyz);

/* ---- brute-force linear scans (semantic cross-checks, fcpw's Baseline pattern) ---- */
void orc_closest_brute(const orc_scene *s, const float *q, long n, uint32_t *idx, float *dist, int nthreads);
void orc_ray_brute(const orc_scene *s, const float *org, const float *dir, const float *tmax, long n, int *found, float *t,
                   uint32_t *prim, int nthreads);
/* distance from q to triangle `idx` with the reference's distance_calculator (tie-aware index checks) */
void orc_host_libm(int which, const float *x, long n, float *out); /* the host's acosf / sinf / cosf / logf */
void orc_point_triangle_distance(const orc_scene *s, const float *q, const uint32_t *idx, long n, float *dist);

/* must-visit statistics for the roofline: mean internal nodes / leaves any exact traversal must open */
void orc_closest_must_visit(const orc_scene *s, const float *q, long n, double *mean_internal, double *mean_leaves);
void orc_silhouette_must_visit(const orc_scene *s, const float *q, long n, int flip, const float *r_max, double *mean_internal,
                               double *mean_leaves);
void orc_ray_must_visit(const orc_scene *s, const float *org, const float *dir, const float *tmax, long n, double *mean_internal,
                        double *mean_leaves);

/* scalar known-answer helpers */
uint32_t orc_morton3(float x, float y, float z);
uint32_t orc_expand_bits(uint32_t v);

#ifdef __cplusplus
}
#endif
#endif
