// Drop-in C++ API test (tests/test_gpu_cpp_dropin.py runs it on the GPU box and parses its output).
//
// Written the way a user of the reference writes code: include snch_lbvh/{lbvh,scene,scene_loader}.cuh, build a
// scene<3> / scene<2>, take get_bvh_device_ptr(), and call query_device() / sample_object_in_sphere() from a kernel with
// the scene's functors.  Checks, on an OBJ file passed on the command line (3-D) and a generated polyline (2-D):
//   * per-thread query_device results == the batched C-ABI results of the same scene (distances bit-equal or 1e-5)
//   * per-thread results == brute force over all primitives with the same functors
//   * generic lbvh::bvh<...> built from scene<3>::triangle objects has the same nodes/aabbs as the scene's own build
// Prints one "CHECK name value" line per check and "DROPIN_OK" at the end.
#include <snch_lbvh/lbvh.cuh>
#include <snch_lbvh/scene.cuh>
#include <snch_lbvh/scene_loader.cuh>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

using scene3 = lbvh::scene<3>;
using scene2 = lbvh::scene<2>;

#define CUDA_OK(x)                                                                    \
    do                                                                                \
    {                                                                                 \
        cudaError_t e = (x);                                                          \
        if (e != cudaSuccess)                                                         \
        {                                                                             \
            std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            return 2;                                                                 \
        }                                                                             \
    } while (0)

struct results3
{
    unsigned int *idx;
    float *dist, *sil, *t, *pdf;
    unsigned char *found;
    int *sidx;
};

__global__ void k_queries3(lbvh::bvh_device<float, 3, scene3::triangle> bvh, const float3 *q, const float3 *d, const float4 *sph, const float *u,
                           int n, results3 r)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const auto near = lbvh::query_device(bvh, lbvh::nearest(q[i]), scene3::distance_calculator());
    r.idx[i] = near.first;
    r.dist[i] = near.second;
    r.sil[i] = lbvh::query_device(bvh, lbvh::nearest_silhouette(q[i], false), scene3::silhouette_distance_calculator());
    const auto hit = lbvh::query_device(bvh, lbvh::ray_intersect(lbvh::ray<float, 3>(q[i], d[i]), INFINITY), scene3::intersect_test());
    r.found[i] = thrust::get<0>(hit) ? 1 : 0;
    r.t[i] = thrust::get<1>(hit);
    const bool any = lbvh::query_device(bvh, lbvh::ray_intersect<true>(lbvh::ray<float, 3>(q[i], d[i]), INFINITY), scene3::intersect_test());
    if (any != thrust::get<0>(hit)) r.found[i] |= 2;
    const auto s = lbvh::sample_object_in_sphere(bvh, lbvh::sphere_intersect(lbvh::sphere<float, 3>(make_float3(sph[i].x, sph[i].y, sph[i].z), sph[i].w)),
                                                 scene3::intersect_sphere(), scene3::measurement_getter(), scene3::green_weight(), u[i]);
    r.sidx[i] = s.first;
    r.pdf[i] = s.second;
}
// brute force with the same functors (semantic oracle that does not depend on any tree)
__global__ void k_brute3(lbvh::bvh_device<float, 3, scene3::triangle> bvh, const float3 *q, const float3 *d, int n, float *dist, float *t)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float best = INFINITY, bt = INFINITY;
    const lbvh::ray<float, 3> r(q[i], d[i]);
    for (unsigned int k = 0; k < bvh.num_objects; ++k)
    {
        best = fminf(best, scene3::distance_calculator()(q[i], bvh.objects[k]));
        const auto h = scene3::intersect_test()(r, bvh.objects[k]);
        if (thrust::get<0>(h)) bt = fminf(bt, thrust::get<1>(h));
    }
    dist[i] = best;
    t[i] = bt;
}
// element_intersects functor of query_device(bvh, line_intersect(l), f, out, cap): the infinite line through l.origin along l.dir
// against one segment; data = the line parameter of the crossing (the reference leaves the functor to the caller, query.cuh:12-77)
struct line_hits_segment
{
    __device__ thrust::pair<bool, float> operator()(const lbvh::line<float, 2> &l, const scene2::line_segment &s) const
    {
        const float2 a = s.vertices[s.vertex_indices.x], b = s.vertices[s.vertex_indices.y];
        const float2 e = make_float2(b.x - a.x, b.y - a.y), w = make_float2(a.x - l.origin.x, a.y - l.origin.y);
        const float den = l.dir.x * e.y - l.dir.y * e.x;
        if (fabsf(den) < 1e-12f) return thrust::make_pair(false, 0.0f);
        const float t = (w.x * e.y - w.y * e.x) / den, u = (w.x * l.dir.y - w.y * l.dir.x) / den;
        return thrust::make_pair(u >= 0.0f && u <= 1.0f, t);
    }
};
__global__ void k_queries2(lbvh::bvh_device<float, 2, scene2::line_segment> bvh, const float2 *q, const float2 *d, int n, float *dist, float *sil,
                           float *t, float *bdist, float *bsil, float *bt, unsigned int *overlap_count, int *sidx, unsigned int *line_count,
                           unsigned int *line_brute)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dist[i] = lbvh::query_device(bvh, lbvh::nearest(q[i]), scene2::distance_calculator()).second;
    sil[i] = lbvh::query_device(bvh, lbvh::nearest_silhouette(q[i], false), scene2::silhouette_distance_calculator());
    const lbvh::ray<float, 2> r(q[i], d[i]);
    const auto hit = lbvh::query_device(bvh, lbvh::ray_intersect(r, INFINITY), scene2::intersect_test());
    t[i] = thrust::get<1>(hit);
    float b = INFINITY, bs = INFINITY, btt = INFINITY;
    for (unsigned int k = 0; k < bvh.num_objects; ++k)
    {
        b = fminf(b, scene2::distance_calculator()(q[i], bvh.objects[k]));
        float ds = INFINITY;
        if (scene2::silhouette_distance_calculator()(q[i], bvh.objects[k], INFINITY, ds, false, 0.0f)) bs = fminf(bs, ds);
        const auto h = scene2::intersect_test()(r, bvh.objects[k]);
        if (thrust::get<0>(h)) btt = fminf(btt, thrust::get<1>(h));
    }
    bdist[i] = b;
    bsil[i] = bs;
    bt[i] = btt;
    unsigned int buf[8];
    const lbvh::aabb<float, 2> box(make_float2(q[i].x + 0.05f, q[i].y + 0.05f), make_float2(q[i].x - 0.05f, q[i].y - 0.05f));
    overlap_count[i] = lbvh::query_device(bvh, lbvh::overlaps(box), buf, 8);
    // every segment the infinite line through q[i] along d[i] crosses (query.cuh:12-77) vs the same functor over all segments
    thrust::pair<unsigned int, float> lbuf[4];
    const lbvh::line<float, 2> ln(q[i], d[i]);
    line_count[i] = lbvh::query_device(bvh, lbvh::line_intersect(ln), line_hits_segment(), lbuf, 4);
    unsigned int lb = 0;
    for (unsigned int k = 0; k < bvh.num_objects; ++k) lb += line_hits_segment()(ln, bvh.objects[k]).first ? 1u : 0u;
    line_brute[i] = lb;
    const auto s = lbvh::sample_object_in_sphere(bvh, lbvh::sphere_intersect(lbvh::sphere<float, 2>(q[i], 0.5f)), scene2::intersect_sphere(),
                                                 scene2::measurement_getter(), scene2::green_weight(), 0.37f);
    sidx[i] = s.first;
}

template <typename T> static T *dev(const std::vector<T> &h)
{
    T *p = nullptr;
    cudaMalloc(&p, h.size() * sizeof(T) + 16);
    cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return p;
}
template <typename T> static std::vector<T> host(const T *p, size_t n)
{
    std::vector<T> h(n);
    cudaMemcpy(h.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost);
    return h;
}
static double worst_rel(const std::vector<float> &a, const std::vector<float> &b)
{
    double w = 0;
    for (size_t i = 0; i < a.size(); ++i)
    {
        if (std::isinf(a[i]) && std::isinf(b[i])) continue;
        if (std::isinf(a[i]) != std::isinf(b[i])) return 1e30;
        // 1e-5 relative with an absolute floor of a few ulps of the coordinates (|x| ~ 1.5): next to the surface the distance
        // is a difference of nearly equal numbers, and this binary is compiled with FMA contraction while the library is not
        const double err = std::fabs((double)a[i] - b[i]);
        w = std::fmax(w, err <= 4e-7 ? 0.0 : err / std::fmax(1e-6, std::fabs((double)b[i])));
    }
    return w;
}
static double mismatch_frac(const std::vector<float> &a, const std::vector<float> &b, double tol)
{
    size_t bad = 0;
    for (size_t i = 0; i < a.size(); ++i)
    {
        if (std::isinf(a[i]) && std::isinf(b[i])) continue;
        if (std::isinf(a[i]) != std::isinf(b[i]) || std::fabs((double)a[i] - b[i]) > 4e-7 + tol * std::fabs((double)b[i])) ++bad;
    }
    return (double)bad / (double)a.size();
}

int main(int argc, char **argv)
{
    if (argc < 2)
    {
        std::printf("usage: dropin_test mesh.obj\n");
        return 2;
    }
    // ---------------------------------------------------------------------------------------------------------- 3-D
    try
    {
        scene3 unbuilt;
        unbuilt.get_bvh_device_ptr();
        std::printf("CHECK not_built_throws 0\n");
    }
    catch (const std::runtime_error &e)
    {
        std::printf("CHECK not_built_throws %d\n", std::strcmp(e.what(), "BVH is not built yet.") == 0 ? 1 : 0);
    }
    try
    {
        lbvh::scene_loader<3> missing("/nonexistent/file.obj");
        std::printf("CHECK loader_throws 0\n");
    }
    catch (const std::runtime_error &e)
    {
        std::printf("CHECK loader_throws %d\n", std::strcmp(e.what(), "Could not open .obj file.") == 0 ? 1 : 0);
    }
    lbvh::scene_loader<3> loader(argv[1]);
    scene3 sc(loader.get_vertices().begin(), loader.get_vertices().end(), loader.get_indices().begin(), loader.get_indices().end());
    sc.compute_silhouettes();
    sc.build_bvh();
    const auto &bvh = sc.get_bvh_device_ptr();
    std::printf("CHECK tris %u\nCHECK nodes %u\n", bvh.num_objects, bvh.num_nodes);

    const int n = 20000;
    std::mt19937 rng(7);
    std::uniform_real_distribution<float> U(-1.5f, 1.5f), U01(0.0f, 1.0f);
    std::vector<float3> q(n), d(n);
    std::vector<float4> sph(n);
    std::vector<float> u(n);
    for (int i = 0; i < n; ++i)
    {
        q[i] = make_float3(U(rng), U(rng), U(rng));
        float3 v;
        float l;
        do
        {
            v = make_float3(U(rng), U(rng), U(rng));
            l = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
        } while (l < 1e-3f);
        d[i] = make_float3(v.x / l, v.y / l, v.z / l);
        sph[i] = make_float4(q[i].x, q[i].y, q[i].z, 0.8f);
        u[i] = U01(rng);
    }
    float3 *dq = dev(q), *dd = dev(d);
    float4 *dsph = dev(sph);
    float *du = dev(u);
    results3 r;
    CUDA_OK(cudaMalloc(&r.idx, n * 4));
    CUDA_OK(cudaMalloc(&r.dist, n * 4));
    CUDA_OK(cudaMalloc(&r.sil, n * 4));
    CUDA_OK(cudaMalloc(&r.t, n * 4));
    CUDA_OK(cudaMalloc(&r.pdf, n * 4));
    CUDA_OK(cudaMalloc(&r.found, n));
    CUDA_OK(cudaMalloc(&r.sidx, n * 4));
    k_queries3<<<(n + 127) / 128, 128>>>(bvh, dq, dd, dsph, du, n, r);
    CUDA_OK(cudaDeviceSynchronize());
    float *bd, *bt;
    CUDA_OK(cudaMalloc(&bd, n * 4));
    CUDA_OK(cudaMalloc(&bt, n * 4));
    k_brute3<<<(n + 127) / 128, 128>>>(bvh, dq, dd, n, bd, bt);
    CUDA_OK(cudaDeviceSynchronize());
    // batched C-ABI on the same scene, device pointers
    unsigned int *bi;
    float *bdist, *bsil, *bpdf;
    snch_hit *bh;
    unsigned char *bf;
    int *bsi;
    float3 *rnd3;
    CUDA_OK(cudaMalloc(&bi, n * 4));
    CUDA_OK(cudaMalloc(&bdist, n * 4));
    CUDA_OK(cudaMalloc(&bsil, n * 4));
    CUDA_OK(cudaMalloc(&bpdf, n * 4));
    CUDA_OK(cudaMalloc(&bh, n * sizeof(snch_hit)));
    CUDA_OK(cudaMalloc(&bf, n));
    CUDA_OK(cudaMalloc(&bsi, n * 4));
    std::vector<float3> rnd(n);
    for (int i = 0; i < n; ++i) rnd[i] = make_float3(u[i], 0.3f, 0.4f);
    rnd3 = dev(rnd);
    sc.closest_points(dq, n, bi, bdist);
    unsigned int *bedge;
    float3 *bpoint;
    CUDA_OK(cudaMalloc(&bedge, n * 4));
    CUDA_OK(cudaMalloc(&bpoint, n * 12));
    sc.closest_silhouettes(dq, nullptr, nullptr, n, bsil, nullptr, bedge, bpoint);
    sc.intersect(dq, dd, nullptr, n, bh, bf);
    sc.sample_in_spheres(dsph, rnd3, n, bsi, bpdf, nullptr);
    CUDA_OK(cudaDeviceSynchronize());

    const auto h_dist = host(r.dist, n), h_sil = host(r.sil, n), h_t = host(r.t, n), h_pdf = host(r.pdf, n);
    const auto h_found = host(r.found, n);
    const auto h_sidx = host(r.sidx, n);
    const auto b_dist = host(bdist, n), b_sil = host(bsil, n), b_pdf = host(bpdf, n), brute_d = host(bd, n), brute_t = host(bt, n);
    const auto b_hits = host(bh, n);
    const auto b_found = host(bf, n);
    const auto b_sidx = host(bsi, n);
    std::vector<float> b_t(n), h_t_inf(n);
    size_t found_diff = 0, anyhit_diff = 0, sidx_diff = 0;
    for (int i = 0; i < n; ++i)
    {
        b_t[i] = b_hits[i].t;
        h_t_inf[i] = (h_found[i] & 1) ? h_t[i] : INFINITY;
        found_diff += ((h_found[i] & 1) != b_found[i]);
        anyhit_diff += (h_found[i] & 2) ? 1 : 0;
        sidx_diff += (h_sidx[i] != b_sidx[i]);
    }
    std::printf("CHECK closest_vs_batched_worst_rel %.3e\n", worst_rel(h_dist, b_dist));
    std::printf("CHECK closest_vs_brute_worst_rel %.3e\n", worst_rel(h_dist, brute_d));
    std::printf("CHECK silhouette_vs_batched_mismatch_frac %.3e\n", mismatch_frac(h_sil, b_sil, 1e-5));
    { // out_edge / out_point: a valid edge id and a point at exactly the reported distance whenever the distance is finite
        const auto h_edge = host(bedge, n);
        const auto h_point = host(bpoint, n);
        size_t bad = 0;
        for (int i = 0; i < n; ++i)
        {
            if (std::isinf(b_sil[i]))
            {
                bad += h_edge[i] != 0xFFFFFFFFu;
                continue;
            }
            const double dx = (double)q[i].x - h_point[i].x, dy = (double)q[i].y - h_point[i].y, dz = (double)q[i].z - h_point[i].z;
            const double len = std::sqrt(dx * dx + dy * dy + dz * dz);
            bad += h_edge[i] >= sc.silhouettes_h.size() || std::fabs(len - b_sil[i]) > 1e-5 * b_sil[i] + 1e-6;
        }
        std::printf("CHECK silhouette_edge_point_bad %zu\n", bad);
    }
    std::printf("CHECK ray_found_diff %zu\nCHECK anyhit_diff %zu\n", found_diff, anyhit_diff);
    std::printf("CHECK ray_t_vs_batched_mismatch_frac %.3e\n", mismatch_frac(h_t_inf, b_t, 1e-5));
    std::printf("CHECK ray_t_vs_brute_mismatch_frac %.3e\n", mismatch_frac(h_t_inf, brute_t, 1e-5));
    std::printf("CHECK sample_idx_diff %zu\n", sidx_diff);
    std::printf("CHECK sample_pdf_mismatch_frac %.3e\n", mismatch_frac(h_pdf, b_pdf, 2e-5));

    // generic lbvh::bvh over the scene's own triangle objects: same topology and boxes as the fused scene build
    {
        lbvh::bvh<float, 3, scene3::triangle, scene3::aabb_getter, scene3::cone_getter> generic(sc.triangles.begin(), sc.triangles.end(), true);
        const auto &gn = generic.nodes_host();
        const auto &ga = generic.aabbs_host();
        const auto &gc = generic.cones_host();
        const auto &sn = sc.p_bvh->nodes_host();
        const auto &sa = sc.p_bvh->aabbs_host();
        const auto &scn = sc.p_bvh->cones_host();
        size_t node_diff = (gn.size() != sn.size()), box_diff = 0, cone_bad = 0;
        for (size_t i = 0; i < gn.size() && i < sn.size(); ++i)
        {
            node_diff += std::memcmp(&gn[i], &sn[i], sizeof(gn[i])) != 0;
            box_diff += std::memcmp(&ga[i], &sa[i], sizeof(ga[i])) != 0;
            const bool both_invalid = gc[i].half_angle < 0 && scn[i].half_angle < 0;
            if (!both_invalid && !(std::fabs(gc[i].half_angle - scn[i].half_angle) <= 1e-4f && std::fabs(gc[i].radius - scn[i].radius) <= 1e-5f * (1 + scn[i].radius)))
                ++cone_bad;
        }
        std::printf("CHECK generic_node_diff %zu\nCHECK generic_aabb_diff %zu\nCHECK generic_cone_bad %zu\n", node_diff, box_diff, cone_bad);
        std::printf("CHECK generic_collision %d\n", generic.morton_collision() ? 1 : 0);
    }

    // ---------------------------------------------------------------------------------------------------------- 2-D
    {
        const int m = 400; // closed wavy loop + an open polyline (boundary vertices)
        std::vector<float2> v;
        std::vector<int2> seg;
        for (int i = 0; i < m; ++i)
        {
            const float a = 6.2831853f * i / m, rad = 1.0f + 0.2f * std::sin(7 * a);
            v.push_back(make_float2(rad * std::cos(a), rad * std::sin(a)));
            seg.push_back(make_int2(i, (i + 1) % m));
        }
        for (int i = 0; i < 50; ++i)
        {
            v.push_back(make_float2(-0.5f + 0.02f * i, 0.1f * std::sin(0.4f * i)));
            if (i) seg.push_back(make_int2(m + i - 1, m + i));
        }
        scene2 s2(v.begin(), v.end(), seg.begin(), seg.end());
        s2.compute_silhouettes();
        s2.build_bvh();
        const auto &b2 = s2.get_bvh_device_ptr();
        const int n2 = 5000;
        std::vector<float2> q2(n2), d2(n2);
        for (int i = 0; i < n2; ++i)
        {
            q2[i] = make_float2(U(rng), U(rng));
            const float a = 6.2831853f * U01(rng);
            d2[i] = make_float2(std::cos(a), std::sin(a));
        }
        float2 *dq2 = dev(q2), *dd2 = dev(d2);
        float *o[6];
        for (auto &p : o) CUDA_OK(cudaMalloc(&p, n2 * 4));
        unsigned int *oc, *lc, *lbr;
        int *os;
        CUDA_OK(cudaMalloc(&oc, n2 * 4));
        CUDA_OK(cudaMalloc(&os, n2 * 4));
        CUDA_OK(cudaMalloc(&lc, n2 * 4));
        CUDA_OK(cudaMalloc(&lbr, n2 * 4));
        k_queries2<<<(n2 + 127) / 128, 128>>>(b2, dq2, dd2, n2, o[0], o[1], o[2], o[3], o[4], o[5], oc, os, lc, lbr);
        CUDA_OK(cudaDeviceSynchronize());
        {
            const auto a = host(lc, n2), b = host(lbr, n2);
            size_t diff = 0, total = 0;
            for (int i = 0; i < n2; ++i)
            {
                diff += a[i] != b[i];
                total += b[i];
            }
            std::printf("CHECK line2d_count_diff %zu\nCHECK line2d_crossings %zu\n", diff, total);
        }
        std::printf("CHECK seg2d %u\n", b2.num_objects);
        std::printf("CHECK closest2d_vs_brute_worst_rel %.3e\n", worst_rel(host(o[0], n2), host(o[3], n2)));
        // the cone-pruned silhouette search may only ever miss what brute force finds if a cone test is wrong
        std::printf("CHECK silhouette2d_vs_brute_mismatch_frac %.3e\n", mismatch_frac(host(o[1], n2), host(o[4], n2), 1e-5));
        std::printf("CHECK ray2d_vs_brute_mismatch_frac %.3e\n", mismatch_frac(host(o[2], n2), host(o[5], n2), 1e-5));
        const auto hs = host(os, n2);
        size_t sampled = 0;
        for (int i = 0; i < n2; ++i) sampled += hs[i] >= 0;
        std::printf("CHECK sample2d_hits %zu\n", sampled);
        const auto &nodes2 = s2.p_bvh->nodes_host();
        std::printf("CHECK nodes2d %zu\n", nodes2.size());
        // batched 2-D entry points of the same scene (device pointers) vs the per-thread calls above
        float *bd, *bs;
        snch_hit *bh;
        unsigned char *bf;
        CUDA_OK(cudaMalloc(&bd, n2 * 4));
        CUDA_OK(cudaMalloc(&bs, n2 * 4));
        CUDA_OK(cudaMalloc(&bh, n2 * sizeof(snch_hit)));
        CUDA_OK(cudaMalloc(&bf, n2));
        s2.closest_points(dq2, n2, oc, bd);
        s2.closest_silhouettes(dq2, nullptr, nullptr, n2, bs);
        s2.intersect(dq2, dd2, nullptr, n2, bh, bf);
        CUDA_OK(cudaDeviceSynchronize());
        std::vector<snch_hit> hh(n2);
        CUDA_OK(cudaMemcpy(hh.data(), bh, n2 * sizeof(snch_hit), cudaMemcpyDeviceToHost));
        std::vector<float> bt(n2);
        for (int i = 0; i < n2; ++i) bt[i] = hh[i].t;
        std::printf("CHECK closest2d_vs_batched_worst_rel %.3e\n", worst_rel(host(o[0], n2), host(bd, n2)));
        std::printf("CHECK silhouette2d_vs_batched_mismatch_frac %.3e\n", mismatch_frac(host(o[1], n2), host(bs, n2), 1e-5));
        std::printf("CHECK ray2d_vs_batched_mismatch_frac %.3e\n", mismatch_frac(host(o[2], n2), bt, 1e-5));
    }
    { // replication behind the C-ABI from C++ (snch_scene_replicate_local): a replica on this device and, when the box has one,
      // on a second GPU answers a batched query bit-identically
        int ndev = 0;
        CUDA_OK(cudaGetDeviceCount(&ndev));
        const int devs[2] = {0, 1};
        const int nrep = ndev > 1 ? 2 : 1;
        snch_scene *reps[2] = {nullptr, nullptr};
        if (snch_scene_replicate_local(sc.native_handle(), devs, nrep, reps) != 0)
        {
            std::printf("replicate failed: %s\n", snch_last_error());
            return 1;
        }
        size_t diff = 0;
        std::vector<float> want(n);
        CUDA_OK(cudaMemcpy(want.data(), bsil, n * 4, cudaMemcpyDeviceToHost));
        for (int r = 0; r < nrep; ++r)
        {
            std::vector<float> got(n);
            std::vector<float3> hq(q.begin(), q.end());
            // host pointers: the library stages them on the replica's own device
            if (snch_closest_silhouette_batch(reps[r], &hq[0].x, nullptr, nullptr, (uint64_t)n, got.data(), nullptr, nullptr, nullptr) != 0)
            {
                std::printf("replica query failed: %s\n", snch_last_error());
                return 1;
            }
            for (int i = 0; i < n; ++i) diff += std::memcmp(&got[i], &want[i], 4) != 0;
            snch_scene_destroy(reps[r]);
        }
        CUDA_OK(cudaSetDevice(0));
        std::printf("CHECK replicas %d\nCHECK replica_silhouette_diff %zu\n", nrep, diff);
    }
    std::printf("DROPIN_OK\n");
    return 0;
}
