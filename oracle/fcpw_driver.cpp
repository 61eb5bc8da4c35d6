/* BASELINE INFRASTRUCTURE ONLY — never linked into the product library.
 *
 * C wrapper around fcpw's CPU backend (bundled with the reference under ext/fcpw; compiled where it lies by
 * oracle/Makefile into oracle/_ref/libfcpw_cpu.so).  BASELINE.json's north_star names fcpw's CPU backend as the
 * reported CPU baseline because the reference's own host query path is broken.  Protocol follows
 * Scene<3> -> computeSilhouettes() -> build(Bvh_SurfaceArea, vectorize) -> bundled queries, which fan
 * out over std::thread::hardware_concurrency() threads (ext/fcpw/include/fcpw/fcpw.inl:1236-1278).
 */
#include <fcpw/fcpw.h>

#include <chrono>
#include <thread>
#include <vector>

using namespace fcpw;

struct fcpw3
{
    Scene<3> scene;
    double build_ms;
};

static double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

extern "C"
{
    int fcpw3_threads() { return (int)std::thread::hardware_concurrency(); }

    fcpw3 *fcpw3_create(const float *xyz, int nV, const int *tri, int nT, int vectorize)
    {
        fcpw3 *f = new fcpw3();
        std::vector<Vector<3>> pos(nV);
        std::vector<Vector3i> idx(nT);
        for (int i = 0; i < nV; ++i) pos[i] = Vector<3>(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        for (int i = 0; i < nT; ++i) idx[i] = Vector3i(tri[3 * i], tri[3 * i + 1], tri[3 * i + 2]);
        double t0 = now_ms();
        f->scene.setObjectCount(1);
        f->scene.setObjectVertices(pos, 0);
        f->scene.setObjectTriangles(idx, 0);
        f->scene.computeSilhouettes();
        f->scene.build(AggregateType::Bvh_SurfaceArea, vectorize != 0);
        f->build_ms = now_ms() - t0;
        return f;
    }
    void fcpw3_destroy(fcpw3 *f) { delete f; }
    double fcpw3_build_ms(fcpw3 *f) { return f->build_ms; }

    double fcpw3_closest(fcpw3 *f, const float *q, long n, float *dist, int *idx)
    {
        std::vector<BoundingSphere<3>> sph;
        sph.reserve(n);
        for (long i = 0; i < n; ++i) sph.emplace_back(Vector<3>(q[3 * i], q[3 * i + 1], q[3 * i + 2]), maxFloat);
        std::vector<Interaction<3>> out;
        double t0 = now_ms();
        f->scene.findClosestPoints(sph, out);
        double ms = now_ms() - t0;
        for (long i = 0; i < n; ++i)
        {
            dist[i] = out[i].d;
            idx[i] = out[i].primitiveIndex;
        }
        return ms;
    }
    /* r_max may be NULL (unbounded); distance is +inf where no silhouette point was found */
    double fcpw3_silhouette(fcpw3 *f, const float *q, const float *r_max, long n, int flip, float *dist)
    {
        std::vector<BoundingSphere<3>> sph;
        sph.reserve(n);
        for (long i = 0; i < n; ++i)
            sph.emplace_back(Vector<3>(q[3 * i], q[3 * i + 1], q[3 * i + 2]), r_max ? r_max[i] * r_max[i] : maxFloat);
        std::vector<uint32_t> flips(n, flip ? 1u : 0u);
        std::vector<Interaction<3>> out;
        double t0 = now_ms();
        f->scene.findClosestSilhouettePoints(sph, out, flips);
        double ms = now_ms() - t0;
        for (long i = 0; i < n; ++i) dist[i] = out[i].primitiveIndex >= 0 ? out[i].d : INFINITY;
        return ms;
    }
    double fcpw3_ray(fcpw3 *f, const float *o, const float *d, const float *tmax, long n, int *found, float *t, int *prim)
    {
        std::vector<Ray<3>> rays;
        rays.reserve(n);
        for (long i = 0; i < n; ++i)
            rays.emplace_back(Vector<3>(o[3 * i], o[3 * i + 1], o[3 * i + 2]), Vector<3>(d[3 * i], d[3 * i + 1], d[3 * i + 2]),
                              std::isfinite(tmax[i]) ? tmax[i] : maxFloat);
        std::vector<Interaction<3>> out;
        double t0 = now_ms();
        f->scene.intersect(rays, out);
        double ms = now_ms() - t0;
        for (long i = 0; i < n; ++i)
        {
            found[i] = out[i].primitiveIndex >= 0 ? 1 : 0;
            t[i] = found[i] ? out[i].d : INFINITY;
            prim[i] = out[i].primitiveIndex;
        }
        return ms;
    }
}
