// snch_lbvh/scene_loader.cuh — minimal Wavefront OBJ reader of the drop-in C++ API (host only; not a performance path).
//
// Same interface and accepted subset as the reference's scene_loader.cuh: `v x y z` vertex lines (z dropped in 2-D),
// `l a b` segments for scene_loader<2>, `f a b c` triangles with plain 1-based indices for scene_loader<3> (no
// `a/b/c` forms, no polygons).  Anything else is skipped.  Throws std::runtime_error("Could not open .obj file.").
#ifndef SNCH_LBVH_B200_SCENE_LOADER_CUH
#define SNCH_LBVH_B200_SCENE_LOADER_CUH
#include "core/utility.cuh"

#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace lbvh
{
template <unsigned int dim> class scene_loader
{
    static_assert(dim == 2 || dim == 3, "scene_loader<2> reads polylines, scene_loader<3> triangle meshes");

public:
    using vertex_type = vector_of_t<float, dim>;
    using index_type = std::conditional_t<dim == 2, int2, int3>;

    explicit scene_loader(const std::string &filename) { load(filename); }
    const std::vector<vertex_type> &get_vertices() const { return vertices; }
    const std::vector<index_type> &get_indices() const { return indices; }
    std::size_t vertices_size() const { return vertices.size(); }
    std::size_t primitives_size() const { return indices.size(); }

private:
    std::vector<vertex_type> vertices;
    std::vector<index_type> indices;

    void load(const std::string &filename)
    {
        std::FILE *f = std::fopen(filename.c_str(), "r");
        if (!f) throw std::runtime_error("Could not open .obj file.");
        const char primitive_tag = dim == 2 ? 'l' : 'f';
        char line[1024];
        while (std::fgets(line, sizeof line, f))
        {
            const char *p = line;
            while (*p == ' ' || *p == '\t') ++p;
            const bool one_letter_tag = p[0] && (p[1] == ' ' || p[1] == '\t');
            if (!one_letter_tag) continue;
            if (p[0] == 'v')
            {
                float c[3] = {0.0f, 0.0f, 0.0f};
                std::sscanf(p + 1, "%f %f %f", &c[0], &c[1], &c[2]);
                vertex_type v;
                for (unsigned int i = 0; i < dim; ++i) detail::at(v, i) = c[i];
                vertices.push_back(v);
            }
            else if (p[0] == primitive_tag)
            {
                int id[3] = {0, 0, 0};
                std::sscanf(p + 1, "%d %d %d", &id[0], &id[1], &id[2]);
                index_type prim;
                for (unsigned int i = 0; i < dim; ++i) detail::at(prim, i) = id[i] - 1; // OBJ indices are 1-based
                indices.push_back(prim);
            }
        }
        std::fclose(f);
    }
};
} // namespace lbvh
#endif // SNCH_LBVH_B200_SCENE_LOADER_CUH
