/* TEST / BASELINE INFRASTRUCTURE ONLY — never linked into the product library.
 *
 * Thin C wrapper around the UNMODIFIED reference headers (included from
 * /root/reference/include at build time, never copied into this repo).  It is
 * compiled two ways by oracle/Makefile, outputs only into oracle/_ref/:
 *
 *   libsnch_ref_cpu.so   g++ -x c++, Thrust CPP backend + oracle/ref_shim.h
 *                        -> the reference's own construct()/query_device()
 *                           executed on host cores ("Oracle B").
 *   libsnch_ref_cuda.so  nvcc sm_100a -> the reference's CUDA path ("Oracle A")
 *                        used for GPU parity and as the reference-CUDA perf bar.
 *
 * Only tests/, bench.py's baseline legs and __graft_entry__.smoke() load these.
 */
#include <snch_lbvh/lbvh.cuh>
#include <snch_lbvh/scene.cuh>

#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

using scene3 = lbvh::scene<3>;
using tri_t = scene3::triangle;
using refdev_t = lbvh::bvh_device<float, 3, tri_t>;

#ifdef __CUDACC__
#define REF_IS_CUDA 1
static void copy_out(void *dst, const void *src, size_t bytes) { cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost); }
#else
#define REF_IS_CUDA 0
static void copy_out(void *dst, const void *src, size_t bytes) { std::memcpy(dst, src, bytes); }
#endif

struct ref3
{
    scene3 *sc;
    double silhouette_ms, build_ms, construct_ms;
};

static double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

#if REF_IS_CUDA
__global__ void k_closest(refdev_t bvh, const float3 *q, long n, unsigned *idx, float *dist)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto r = lbvh::query_device(bvh, lbvh::nearest(q[i]), scene3::distance_calculator());
    idx[i] = r.first;
    dist[i] = r.second;
}
__global__ void k_silhouette(refdev_t bvh, const float3 *q, long n, bool flip, float *dist)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    dist[i] = lbvh::query_device(bvh, lbvh::nearest_silhouette(q[i], flip), scene3::silhouette_distance_calculator());
}
__global__ void k_ray(refdev_t bvh, const float3 *o, const float3 *d, const float *tmax, long n, int *found, float *t,
                      float2 *uv, unsigned *prim)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    lbvh::ray<float, 3> r(o[i], d[i]);
    auto h = lbvh::query_device(bvh, lbvh::ray_intersect(r, tmax[i]), scene3::intersect_test());
    found[i] = thrust::get<0>(h) ? 1 : 0;
    t[i] = thrust::get<1>(h);
    uv[i] = thrust::get<0>(h) ? thrust::get<2>(h) : make_float2(0.f, 0.f);
    prim[i] = thrust::get<3>(h);
}
__global__ void k_sample(refdev_t bvh, const float4 *sph, const float *u, long n, int *idx, float *pdf)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    lbvh::sphere<float, 3> s(make_float3(sph[i].x, sph[i].y, sph[i].z), sph[i].w);
    auto r = lbvh::sample_object_in_sphere(bvh, lbvh::sphere_intersect(s), scene3::intersect_sphere(),
                                           scene3::measurement_getter(), scene3::green_weight(), u[i]);
    idx[i] = r.first;
    pdf[i] = r.first >= 0 ? r.second : 0.0f;
}
template <typename T> struct dbuf
{
    T *p = nullptr;
    size_t n;
    dbuf(size_t n_) : n(n_) { cudaMalloc(&p, n * sizeof(T)); }
    dbuf(const T *h, size_t n_) : n(n_)
    {
        cudaMalloc(&p, n * sizeof(T));
        cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice);
    }
    void to(T *h) { cudaMemcpy(h, p, n * sizeof(T), cudaMemcpyDeviceToHost); }
    ~dbuf() { cudaFree(p); }
};
struct evtimer
{
    cudaEvent_t a, b;
    evtimer()
    {
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a);
    }
    double stop()
    {
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        cudaEventDestroy(a);
        cudaEventDestroy(b);
        return ms;
    }
};
#else
template <typename F> static void par_for(long n, int nthreads, F f)
{
    if (nthreads <= 1)
    {
        for (long i = 0; i < n; ++i) f(i);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back(
            [=]()
            {
                long lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
                for (long i = lo; i < hi; ++i) f(i);
            });
    for (auto &x : th) x.join();
}
#endif

extern "C"
{
    int ref3_is_cuda() { return REF_IS_CUDA; }

    ref3 *ref3_create(const float *xyz, int nV, const int *tri, int nT)
    {
        std::vector<float3> v(nV);
        std::vector<int3> idx(nT);
        for (int i = 0; i < nV; ++i) v[i] = make_float3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        for (int i = 0; i < nT; ++i) idx[i] = make_int3(tri[3 * i], tri[3 * i + 1], tri[3 * i + 2]);
        ref3 *r = new ref3();
        r->sc = new scene3(v.begin(), v.end(), idx.begin(), idx.end());
        double t0 = now_ms();
        r->sc->compute_silhouettes();
        double t1 = now_ms();
        r->sc->build_bvh();
#if REF_IS_CUDA
        cudaDeviceSynchronize();
#endif
        double t2 = now_ms();
        r->silhouette_ms = t1 - t0;
        r->build_ms = t2 - t1;
        r->construct_ms = -1.0;
        return r;
    }
    void ref3_destroy(ref3 *r)
    {
        delete r->sc;
        delete r;
    }
    /* re-run lbvh::bvh::construct() (bvh.cuh:380-613) `reps` times; returns best wall ms (device synchronised) */
    double ref3_time_construct(ref3 *r, int reps)
    {
        double best = 1e30;
        for (int i = 0; i < reps; ++i)
        {
#if REF_IS_CUDA
            cudaDeviceSynchronize();
#endif
            double t0 = now_ms();
            r->sc->p_bvh->construct();
#if REF_IS_CUDA
            cudaDeviceSynchronize();
#endif
            double t1 = now_ms();
            if (t1 - t0 < best) best = t1 - t0;
        }
        r->sc->bvh_dev = r->sc->p_bvh->get_device_repr();
        r->construct_ms = best;
        return best;
    }
    void ref3_timings(ref3 *r, double *out3)
    {
        out3[0] = r->silhouette_ms;
        out3[1] = r->build_ms;
        out3[2] = r->construct_ms;
    }
    int ref3_num_objects(ref3 *r) { return (int)r->sc->bvh_dev.num_objects; }
    int ref3_num_nodes(ref3 *r) { return (int)r->sc->bvh_dev.num_nodes; }
    int ref3_num_edges(ref3 *r) { return (int)r->sc->silhouettes_h.size(); }
    /* nodes: 4 x u32 per node {parent,left,right,object}; aabbs: 6 floats {upper xyz, lower xyz}; cones: 5 floats */
    void ref3_export_tree(ref3 *r, unsigned *nodes, float *aabbs, float *cones)
    {
        const refdev_t &d = r->sc->bvh_dev;
        static_assert(sizeof(lbvh::detail::node) == 16, "");
        static_assert(sizeof(lbvh::aabb<float, 3>) == 24, "");
        static_assert(sizeof(lbvh::cone<float, 3>) == 20, "");
        if (nodes) copy_out(nodes, d.nodes, sizeof(lbvh::detail::node) * d.num_nodes);
        if (aabbs) copy_out(aabbs, d.aabbs, sizeof(lbvh::aabb<float, 3>) * d.num_nodes);
        if (cones) copy_out(cones, d.cones, sizeof(lbvh::cone<float, 3>) * d.num_nodes);
    }
    /* silhouette_edge::indices (int4 per edge), per-triangle edge ids and owned-edge lists (int3 per triangle) */
    void ref3_export_adjacency(ref3 *r, int *edges4, int *tri_edges3, int *tri_owned3)
    {
        scene3 *s = r->sc;
        for (size_t e = 0; e < s->silhouettes_h.size(); ++e)
        {
            int4 id = static_cast<const scene3::silhouette_edge &>(s->silhouettes_h[e]).indices;
            edges4[4 * e + 0] = id.x;
            edges4[4 * e + 1] = id.y;
            edges4[4 * e + 2] = id.z;
            edges4[4 * e + 3] = id.w;
        }
        for (size_t i = 0; i < s->triangles.size(); ++i)
        {
            int3 ei = s->edge_indices_h[i];
            tri_edges3[3 * i + 0] = ei.x;
            tri_edges3[3 * i + 1] = ei.y;
            tri_edges3[3 * i + 2] = ei.z;
            int3 oi = s->triangles[i].silhouette_indices;
            tri_owned3[3 * i + 0] = oi.x;
            tri_owned3[3 * i + 1] = oi.y;
            tri_owned3[3 * i + 2] = oi.z;
        }
    }
    /* Morton codes in SORTED leaf order, recomputed with the reference's own calculator from the finished tree
     * (scene box == aabbs[0]) */
    void ref3_export_morton(ref3 *r, unsigned *morton_sorted)
    {
        const refdev_t &d = r->sc->bvh_dev;
        std::vector<lbvh::aabb<float, 3>> boxes(d.num_nodes);
        copy_out(boxes.data(), d.aabbs, sizeof(lbvh::aabb<float, 3>) * d.num_nodes);
        lbvh::default_morton_code_calculator<float, 3, tri_t> calc(boxes[0]);
        const unsigned n = d.num_objects;
        for (unsigned k = 0; k < n; ++k) morton_sorted[k] = calc(r->sc->triangles[0], boxes[n - 1 + k]);
    }

    /* ---- queries: host in / host out; returns milliseconds of the traversal alone ---- */
    double ref3_closest(ref3 *r, const float *q, long n, unsigned *idx, float *dist, int nthreads)
    {
        const refdev_t bvh = r->sc->bvh_dev;
#if REF_IS_CUDA
        dbuf<float3> dq((const float3 *)q, n);
        dbuf<unsigned> di(n);
        dbuf<float> dd(n);
        cudaDeviceSynchronize();
        evtimer t;
        k_closest<<<(unsigned)((n + 255) / 256), 256>>>(bvh, dq.p, n, di.p, dd.p);
        double ms = t.stop();
        di.to(idx);
        dd.to(dist);
        return ms;
#else
        double t0 = now_ms();
        par_for(n, nthreads,
                [=](long i)
                {
                    auto res = lbvh::query_device(bvh, lbvh::nearest(make_float3(q[3 * i], q[3 * i + 1], q[3 * i + 2])),
                                                  scene3::distance_calculator());
                    idx[i] = res.first;
                    dist[i] = res.second;
                });
        return now_ms() - t0;
#endif
    }
    double ref3_silhouette(ref3 *r, const float *q, long n, int flip, float *dist, int nthreads)
    {
        const refdev_t bvh = r->sc->bvh_dev;
#if REF_IS_CUDA
        dbuf<float3> dq((const float3 *)q, n);
        dbuf<float> dd(n);
        cudaDeviceSynchronize();
        evtimer t;
        k_silhouette<<<(unsigned)((n + 255) / 256), 256>>>(bvh, dq.p, n, flip != 0, dd.p);
        double ms = t.stop();
        dd.to(dist);
        return ms;
#else
        double t0 = now_ms();
        par_for(n, nthreads,
                [=](long i)
                {
                    dist[i] = lbvh::query_device(
                        bvh, lbvh::nearest_silhouette(make_float3(q[3 * i], q[3 * i + 1], q[3 * i + 2]), flip != 0),
                        scene3::silhouette_distance_calculator());
                });
        return now_ms() - t0;
#endif
    }
    double ref3_ray(ref3 *r, const float *o, const float *d, const float *tmax, long n, int *found, float *t, float *uv,
                    unsigned *prim, int nthreads)
    {
        const refdev_t bvh = r->sc->bvh_dev;
#if REF_IS_CUDA
        dbuf<float3> dorg((const float3 *)o, n), ddir((const float3 *)d, n);
        dbuf<float> dtm(tmax, n), dt(n);
        dbuf<int> df(n);
        dbuf<float2> duv(n);
        dbuf<unsigned> dp(n);
        cudaDeviceSynchronize();
        evtimer tm;
        k_ray<<<(unsigned)((n + 255) / 256), 256>>>(bvh, dorg.p, ddir.p, dtm.p, n, df.p, dt.p, duv.p, dp.p);
        double ms = tm.stop();
        df.to(found);
        dt.to(t);
        duv.to((float2 *)uv);
        dp.to(prim);
        return ms;
#else
        double t0 = now_ms();
        par_for(n, nthreads,
                [=](long i)
                {
                    lbvh::ray<float, 3> ry(make_float3(o[3 * i], o[3 * i + 1], o[3 * i + 2]),
                                           make_float3(d[3 * i], d[3 * i + 1], d[3 * i + 2]));
                    auto h = lbvh::query_device(bvh, lbvh::ray_intersect(ry, tmax[i]), scene3::intersect_test());
                    found[i] = thrust::get<0>(h) ? 1 : 0;
                    t[i] = thrust::get<1>(h);
                    uv[2 * i] = thrust::get<0>(h) ? thrust::get<2>(h).x : 0.f;
                    uv[2 * i + 1] = thrust::get<0>(h) ? thrust::get<2>(h).y : 0.f;
                    prim[i] = thrust::get<3>(h);
                });
        return now_ms() - t0;
#endif
    }
    /* sph: x,y,z,radius per query.  pdf is reported as 0 on a miss (the reference leaves it uninitialised, Q19) */
    double ref3_sample(ref3 *r, const float *sph, const float *u, long n, int *idx, float *pdf, int nthreads)
    {
        const refdev_t bvh = r->sc->bvh_dev;
#if REF_IS_CUDA
        dbuf<float4> ds((const float4 *)sph, n);
        dbuf<float> du(u, n), dp(n);
        dbuf<int> di(n);
        cudaDeviceSynchronize();
        evtimer tm;
        k_sample<<<(unsigned)((n + 255) / 256), 256>>>(bvh, ds.p, du.p, n, di.p, dp.p);
        double ms = tm.stop();
        di.to(idx);
        dp.to(pdf);
        return ms;
#else
        double t0 = now_ms();
        par_for(n, nthreads,
                [=](long i)
                {
                    lbvh::sphere<float, 3> s(make_float3(sph[4 * i], sph[4 * i + 1], sph[4 * i + 2]), sph[4 * i + 3]);
                    auto res = lbvh::sample_object_in_sphere(bvh, lbvh::sphere_intersect(s), scene3::intersect_sphere(),
                                                             scene3::measurement_getter(), scene3::green_weight(), u[i]);
                    idx[i] = res.first;
                    pdf[i] = res.first >= 0 ? res.second : 0.0f;
                });
        return now_ms() - t0;
#endif
    }
}

/* =====================================================================================================================
 * 2-D: lbvh::scene<2> (line segments / silhouette vertices), scene.cuh:287-703.  Same conventions as ref3_*.
 * ===================================================================================================================== */
using scene2 = lbvh::scene<2>;
using seg_t = scene2::line_segment;
using refdev2_t = lbvh::bvh_device<float, 2, seg_t>;
struct ref2
{
    scene2 *sc;
};
#if REF_IS_CUDA
__global__ void k2_closest(refdev2_t bvh, const float2 *q, long n, unsigned *idx, float *dist)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto r = lbvh::query_device(bvh, lbvh::nearest(q[i]), scene2::distance_calculator());
    idx[i] = r.first;
    dist[i] = r.second;
}
__global__ void k2_silhouette(refdev2_t bvh, const float2 *q, long n, bool flip, float *dist)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    dist[i] = lbvh::query_device(bvh, lbvh::nearest_silhouette(q[i], flip), scene2::silhouette_distance_calculator());
}
__global__ void k2_ray(refdev2_t bvh, const float2 *o, const float2 *d, const float *tmax, long n, int *found, float *t, float *s,
                       unsigned *prim)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    lbvh::ray<float, 2> r(o[i], d[i]);
    auto h = lbvh::query_device(bvh, lbvh::ray_intersect(r, tmax[i]), scene2::intersect_test());
    found[i] = thrust::get<0>(h) ? 1 : 0;
    t[i] = thrust::get<1>(h);
    s[i] = thrust::get<0>(h) ? thrust::get<2>(h) : 0.f;
    prim[i] = thrust::get<3>(h);
}
__global__ void k2_sample(refdev2_t bvh, const float3 *sph, const float *u, long n, int *idx, float *pdf)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    lbvh::sphere<float, 2> s(make_float2(sph[i].x, sph[i].y), sph[i].z);
    auto r = lbvh::sample_object_in_sphere(bvh, lbvh::sphere_intersect(s), scene2::intersect_sphere(), scene2::measurement_getter(),
                                           scene2::green_weight(), u[i]);
    idx[i] = r.first;
    pdf[i] = r.first >= 0 ? r.second : 0.0f;
}
#endif

extern "C"
{
    ref2 *ref2_create(const float *xy, int nV, const int *seg, int nS)
    {
        std::vector<float2> v(nV);
        std::vector<int2> idx(nS);
        for (int i = 0; i < nV; ++i) v[i] = make_float2(xy[2 * i], xy[2 * i + 1]);
        for (int i = 0; i < nS; ++i) idx[i] = make_int2(seg[2 * i], seg[2 * i + 1]);
        ref2 *r = new ref2();
        r->sc = new scene2(v.begin(), v.end(), idx.begin(), idx.end());
        r->sc->compute_silhouettes();
        r->sc->build_bvh();
#if REF_IS_CUDA
        cudaDeviceSynchronize();
#endif
        return r;
    }
    void ref2_destroy(ref2 *r)
    {
        delete r->sc;
        delete r;
    }
    int ref2_num_objects(ref2 *r) { return (int)r->sc->bvh_dev.num_objects; }
    int ref2_num_nodes(ref2 *r) { return (int)r->sc->bvh_dev.num_nodes; }
    /* nodes: 4 x u32 {parent,left,right,object}; aabbs: 4 floats {upper xy, lower xy}; cones: 4 floats {axis xy, half_angle, radius} */
    void ref2_export_tree(ref2 *r, unsigned *nodes, float *aabbs, float *cones)
    {
        const refdev2_t &d = r->sc->bvh_dev;
        static_assert(sizeof(lbvh::aabb<float, 2>) == 16, "");
        static_assert(sizeof(lbvh::cone<float, 2>) == 16, "");
        if (nodes) copy_out(nodes, d.nodes, sizeof(lbvh::detail::node) * d.num_nodes);
        if (aabbs) copy_out(aabbs, d.aabbs, sizeof(lbvh::aabb<float, 2>) * d.num_nodes);
        if (cones) copy_out(cones, d.cones, sizeof(lbvh::cone<float, 2>) * d.num_nodes);
    }
    /* silhouette_vertex::indices (int4 per vertex) and the owned-vertex pair of every segment (scene.cuh:634-681) */
    void ref2_export_adjacency(ref2 *r, int *vert4, int *seg_owned2)
    {
        scene2 *s = r->sc;
        for (size_t v = 0; v < s->silhouettes_h.size(); ++v)
        {
            int4 id = static_cast<const scene2::silhouette_vertex &>(s->silhouettes_h[v]).indices;
            vert4[4 * v + 0] = id.x;
            vert4[4 * v + 1] = id.y;
            vert4[4 * v + 2] = id.z;
            vert4[4 * v + 3] = id.w;
        }
        for (size_t i = 0; i < s->lines.size(); ++i)
        {
            int2 oi = s->lines[i].silhouette_indices;
            seg_owned2[2 * i + 0] = oi.x;
            seg_owned2[2 * i + 1] = oi.y;
        }
    }
    double ref2_closest(ref2 *r, const float *q, long n, unsigned *idx, float *dist, int nthreads)
    {
        const refdev2_t bvh = r->sc->bvh_dev;
#if REF_IS_CUDA
        dbuf<float2> dq((const float2 *)q, n);
        dbuf<unsigned> di(n);
        dbuf<float> dd(n);
        cudaDeviceSynchronize();
        evtimer t;
        k2_closest<<<(unsigned)((n + 255) / 256), 256>>>(bvh, dq.p, n, di.p, dd.p);
        double ms = t.stop();
        di.to(idx);
        dd.to(dist);
        return ms;
#else
        double t0 = now_ms();
        par_for(n, nthreads,
                [=](long i)
                {
                    auto res = lbvh::query_device(bvh, lbvh::nearest(make_float2(q[2 * i], q[2 * i + 1])), scene2::distance_calculator());
                    idx[i] = res.first;
                    dist[i] = res.second;
                });
        return now_ms() - t0;
#endif
    }
    double ref2_silhouette(ref2 *r, const float *q, long n, int flip, float *dist, int nthreads)
    {
        const refdev2_t bvh = r->sc->bvh_dev;
#if REF_IS_CUDA
        dbuf<float2> dq((const float2 *)q, n);
        dbuf<float> dd(n);
        cudaDeviceSynchronize();
        evtimer t;
        k2_silhouette<<<(unsigned)((n + 255) / 256), 256>>>(bvh, dq.p, n, flip != 0, dd.p);
        double ms = t.stop();
        dd.to(dist);
        return ms;
#else
        double t0 = now_ms();
        par_for(n, nthreads,
                [=](long i)
                {
                    dist[i] = lbvh::query_device(bvh, lbvh::nearest_silhouette(make_float2(q[2 * i], q[2 * i + 1]), flip != 0),
                                                 scene2::silhouette_distance_calculator());
                });
        return now_ms() - t0;
#endif
    }
    double ref2_ray(ref2 *r, const float *o, const float *d, const float *tmax, long n, int *found, float *t, float *s, unsigned *prim,
                    int nthreads)
    {
        const refdev2_t bvh = r->sc->bvh_dev;
#if REF_IS_CUDA
        dbuf<float2> dorg((const float2 *)o, n), ddir((const float2 *)d, n);
        dbuf<float> dtm(tmax, n), dt(n), ds(n);
        dbuf<int> df(n);
        dbuf<unsigned> dp(n);
        cudaDeviceSynchronize();
        evtimer tm;
        k2_ray<<<(unsigned)((n + 255) / 256), 256>>>(bvh, dorg.p, ddir.p, dtm.p, n, df.p, dt.p, ds.p, dp.p);
        double ms = tm.stop();
        df.to(found);
        dt.to(t);
        ds.to(s);
        dp.to(prim);
        return ms;
#else
        double t0 = now_ms();
        par_for(n, nthreads,
                [=](long i)
                {
                    lbvh::ray<float, 2> ry(make_float2(o[2 * i], o[2 * i + 1]), make_float2(d[2 * i], d[2 * i + 1]));
                    auto h = lbvh::query_device(bvh, lbvh::ray_intersect(ry, tmax[i]), scene2::intersect_test());
                    found[i] = thrust::get<0>(h) ? 1 : 0;
                    t[i] = thrust::get<1>(h);
                    s[i] = thrust::get<0>(h) ? thrust::get<2>(h) : 0.f;
                    prim[i] = thrust::get<3>(h);
                });
        return now_ms() - t0;
#endif
    }
    /* sph: x,y,radius per query.  pdf is reported as 0 on a miss (Q19) */
    double ref2_sample(ref2 *r, const float *sph, const float *u, long n, int *idx, float *pdf, int nthreads)
    {
        const refdev2_t bvh = r->sc->bvh_dev;
#if REF_IS_CUDA
        dbuf<float3> ds((const float3 *)sph, n);
        dbuf<float> du(u, n), dp(n);
        dbuf<int> di(n);
        cudaDeviceSynchronize();
        evtimer tm;
        k2_sample<<<(unsigned)((n + 255) / 256), 256>>>(bvh, ds.p, du.p, n, di.p, dp.p);
        double ms = tm.stop();
        di.to(idx);
        dp.to(pdf);
        return ms;
#else
        double t0 = now_ms();
        par_for(n, nthreads,
                [=](long i)
                {
                    lbvh::sphere<float, 2> s(make_float2(sph[3 * i], sph[3 * i + 1]), sph[3 * i + 2]);
                    auto res = lbvh::sample_object_in_sphere(bvh, lbvh::sphere_intersect(s), scene2::intersect_sphere(),
                                                             scene2::measurement_getter(), scene2::green_weight(), u[i]);
                    idx[i] = res.first;
                    pdf[i] = res.first >= 0 ? res.second : 0.0f;
                });
        return now_ms() - t0;
#endif
    }
}
