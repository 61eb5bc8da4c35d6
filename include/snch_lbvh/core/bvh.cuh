// snch_lbvh/core/bvh.cuh — the LBVH + SNCH container of the drop-in C++ API.
//
// lbvh::bvh<Real, dim, Object, AABBGetter, ConeGetter, MortonCodeCalculator> keeps the reference's interface
// (core/bvh.cuh:311-634: constructors, assign, clear, construct, get_device_repr, *_host accessors) and its device view
// lbvh::bvh_device / cbvh_device keeps the reference's public members (bvh.cuh:36-107), because user kernels and
// query_device() read them directly.  Construction is NOT a Thrust pipeline here: the user's getters are evaluated per
// object by one small kernel of this header (user types cannot cross a C-ABI) and the tree is built by libsnch_b200.so
// (snch_lbvh_build: hand-written sm_100a kernels — Morton, radix sort of (key, index), Karras hierarchy, one fused
// box + cone refit).  Tree topology equals the reference's bit for bit; see DESIGN.md "Parity rules".
#ifndef SNCH_LBVH_B200_BVH_CUH
#define SNCH_LBVH_B200_BVH_CUH
#include "aabb.cuh"
#include "cone.cuh"
#include "morton_code.cuh"
#include "../../snch_b200.h"

#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>
#include <thrust/host_vector.h>
#include <thrust/pair.h>
#include <thrust/tuple.h>

namespace lbvh
{
namespace detail
{
struct node
{
    std::uint32_t parent_idx; // 0xFFFFFFFF for the root
    std::uint32_t left_idx;   // internal nodes are 0 .. N-2 (root 0), leaves N-1 .. 2N-2 in Morton order
    std::uint32_t right_idx;
    std::uint32_t object_idx; // 0xFFFFFFFF for internal nodes, else the index into `objects`
};

// non-owning device view, passed by value into kernels
template <typename Real, unsigned int dim, typename Object, bool IsConst> struct basic_device_bvh
{
    using real_type = Real;
    using aabb_type = aabb<real_type, dim>;
    using cone_type = cone<real_type, dim>;
    using node_type = detail::node;
    using index_type = std::uint32_t;
    using object_type = Object;
    template <typename T> using ptr = std::conditional_t<IsConst, T const *, T *>;

    unsigned int num_nodes;   // 2N - 1
    unsigned int num_objects; // N
    ptr<node_type> nodes;
    ptr<aabb_type> aabbs;
    ptr<cone_type> cones;
    ptr<object_type> objects;

    SNCH_LBVH_CALLABLE basic_device_bvh() : num_nodes(0), num_objects(0), nodes(nullptr), aabbs(nullptr), cones(nullptr), objects(nullptr) {}
    SNCH_LBVH_CALLABLE basic_device_bvh(unsigned int num_nodes_, unsigned int num_objects_, ptr<node_type> nodes_, ptr<aabb_type> aabbs_,
                                        ptr<cone_type> cones_, ptr<object_type> objects_)
        : num_nodes(num_nodes_), num_objects(num_objects_), nodes(nodes_), aabbs(aabbs_), cones(cones_), objects(objects_)
    {
    }
    // a mutable view converts to a const one
    template <bool C = IsConst, std::enable_if_t<C, int> = 0>
    SNCH_LBVH_CALLABLE basic_device_bvh(const basic_device_bvh<Real, dim, Object, false> &o)
        : num_nodes(o.num_nodes), num_objects(o.num_objects), nodes(o.nodes), aabbs(o.aabbs), cones(o.cones), objects(o.objects)
    {
    }
};

// owning device array without any library kernel behind it (allocation, copies and nothing else)
template <typename T> class device_buffer
{
public:
    device_buffer() = default;
    device_buffer(const device_buffer &o) { assign(o.data_, o.size_, cudaMemcpyDeviceToDevice); }
    device_buffer(device_buffer &&o) noexcept : data_(o.data_), size_(o.size_) { o.data_ = nullptr, o.size_ = 0; }
    device_buffer &operator=(device_buffer o) noexcept
    {
        std::swap(data_, o.data_);
        std::swap(size_, o.size_);
        return *this;
    }
    ~device_buffer() { release(); }
    void resize(std::size_t n)
    {
        if (n == size_) return;
        release();
        if (n && cudaMalloc(reinterpret_cast<void **>(&data_), n * sizeof(T)) != cudaSuccess)
        {
            data_ = nullptr;
            throw std::bad_alloc();
        }
        size_ = n;
    }
    void clear() { release(); }
    void upload(const T *host, std::size_t n) { assign(host, n, cudaMemcpyHostToDevice); }
    void download(T *host) const
    {
        if (size_ && cudaMemcpy(host, data_, size_ * sizeof(T), cudaMemcpyDeviceToHost) != cudaSuccess)
            throw std::runtime_error("snch_lbvh: device to host copy failed");
    }
    T *data() noexcept { return data_; }
    const T *data() const noexcept { return data_; }
    std::size_t size() const noexcept { return size_; }

private:
    void assign(const T *src, std::size_t n, cudaMemcpyKind kind)
    {
        resize(n);
        if (n && cudaMemcpy(data_, src, n * sizeof(T), kind) != cudaSuccess) throw std::runtime_error("snch_lbvh: copy to device failed");
    }
    void release() noexcept
    {
        if (data_) cudaFree(data_);
        data_ = nullptr;
        size_ = 0;
    }
    T *data_ = nullptr;
    std::size_t size_ = 0;
};

inline void check_status(int status)
{
    if (status != SNCH_OK) throw std::runtime_error(std::string("snch_lbvh: ") + snch_last_error());
}

// leaf pass of the generic builder: one thread per object evaluates the user's getters
template <typename Object, typename AABBGetter, typename ConeGetter, typename Box, typename Cone>
__global__ void k_leaf_arrays(const Object *objects, unsigned int n, AABBGetter box_of, ConeGetter cone_of, Box *boxes, Cone *cones)
{
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    boxes[i] = box_of(objects[i]);
    cones[i] = cone_of(objects[i]);
}
template <typename Object, typename Calc, typename Box>
__global__ void k_leaf_codes(const Object *objects, unsigned int n, Calc calc, const Box *boxes, unsigned int *codes)
{
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) codes[i] = calc(objects[i], boxes[i]);
}
} // namespace detail

template <typename Real, unsigned int dim, typename Object> using bvh_device = detail::basic_device_bvh<Real, dim, Object, false>;
template <typename Real, unsigned int dim, typename Object> using cbvh_device = detail::basic_device_bvh<Real, dim, Object, true>;

// Morton code of the leaf-box centroid normalised by the box of the whole scene (bvh.cuh:232-304).  When a bvh uses this
// (default) calculator the codes are computed inside the library; any other calculator type is evaluated by this header.
template <typename Real, unsigned int dim, typename Object> struct default_morton_code_calculator
{
    default_morton_code_calculator(aabb<Real, dim> w) : whole(w) {}
    default_morton_code_calculator() = default;
    SNCH_LBVH_CALLABLE unsigned int operator()(const Object &, const aabb<Real, dim> &box) noexcept
    {
        vector_of_t<Real, dim> p = centroid(box);
        for (unsigned int i = 0; i < dim; ++i)
            detail::at(p, i) = (detail::at(p, i) - detail::at(whole.lower, i)) / (detail::at(whole.upper, i) - detail::at(whole.lower, i));
        return morton_code(p);
    }
    aabb<Real, dim> whole;
};

template <typename Real, unsigned int dim, typename Object, typename AABBGetter, typename ConeGetter,
          typename MortonCodeCalculator = default_morton_code_calculator<Real, dim, Object>>
class bvh
{
    static_assert(std::is_same<Real, float>::value, "snch-lbvh_b200 builds float trees only (the reference's scenes are float-only too)");
    static_assert(dim == 2 || dim == 3, "2-D or 3-D");

public:
    using real_type = Real;
    using index_type = std::uint32_t;
    using object_type = Object;
    using aabb_type = aabb<real_type, dim>;
    using cone_type = cone<real_type, dim>;
    using node_type = detail::node;
    using aabb_getter_type = AABBGetter;
    using cone_getter_type = ConeGetter;
    using morton_code_calculator_type = MortonCodeCalculator;

    template <typename InputIterator>
    bvh(InputIterator first, InputIterator last, bool query_host_enabled = false) : objects_h_(first, last), query_host_enabled_(query_host_enabled)
    {
        this->construct();
    }
    bvh() = default;

    bool query_host_enabled() const noexcept { return query_host_enabled_; }
    bool &query_host_enabled() noexcept { return query_host_enabled_; }
    bool morton_collision() const noexcept { return collision_; } // the reference prints a line instead (bvh.cuh:466)

    void clear()
    {
        objects_h_.clear();
        objects_d_.clear();
        aabbs_h_.clear();
        aabbs_.clear();
        cones_h_.clear();
        cones_.clear();
        nodes_h_.clear();
        nodes_.clear();
    }
    template <typename InputIterator> void assign(InputIterator first, InputIterator last)
    {
        objects_h_.assign(first, last);
        this->construct();
    }
    bvh_device<real_type, dim, object_type> get_device_repr() noexcept
    {
        return {static_cast<unsigned int>(nodes_.size()), static_cast<unsigned int>(objects_d_.size()), nodes_.data(), aabbs_.data(),
                cones_.data(), objects_d_.data()};
    }
    cbvh_device<real_type, dim, object_type> get_device_repr() const noexcept
    {
        return {static_cast<unsigned int>(nodes_.size()), static_cast<unsigned int>(objects_d_.size()), nodes_.data(), aabbs_.data(),
                cones_.data(), objects_d_.data()};
    }

    // (re)build from objects_host(): upload, leaf pass, library build, optional host mirrors          bvh.cuh:380-613
    void construct()
    {
        const std::size_t n = objects_h_.size();
        objects_d_.upload(n ? &objects_h_[0] : nullptr, n);
        collision_ = false;
        if (n == 0)
        { // bvh.cuh:383-386: an empty tree has no arrays at all
            nodes_.clear(), aabbs_.clear(), cones_.clear();
            nodes_h_.clear(), aabbs_h_.clear(), cones_h_.clear();
            return;
        }
        nodes_.resize(2 * n - 1);
        aabbs_.resize(2 * n - 1);
        cones_.resize(2 * n - 1);
        detail::device_buffer<aabb_type> leaf_boxes;
        detail::device_buffer<cone_type> leaf_cones;
        leaf_boxes.resize(n);
        leaf_cones.resize(n);
        const unsigned int un = static_cast<unsigned int>(n), grid = (un + 127u) / 128u;
        detail::k_leaf_arrays<<<grid, 128>>>(objects_d_.data(), un, AABBGetter(), ConeGetter(), leaf_boxes.data(), leaf_cones.data());
        detail::device_buffer<unsigned int> codes;
        if (!std::is_same<MortonCodeCalculator, default_morton_code_calculator<Real, dim, Object>>::value)
        { // custom calculator: it wants the whole-scene box at construction, so reduce the leaf boxes (rare path, host side)
            std::vector<aabb_type> hb(n);
            leaf_boxes.download(hb.data());
            aabb_type whole = hb[0];
            for (std::size_t i = 1; i < n; ++i) whole = merge(whole, hb[i]);
            codes.resize(n);
            detail::k_leaf_codes<<<grid, 128>>>(objects_d_.data(), un, MortonCodeCalculator(whole), leaf_boxes.data(), codes.data());
        }
        int collision = 0;
        detail::check_status(snch_lbvh_build(static_cast<int>(dim), un, leaf_boxes.data(), leaf_cones.data(), codes.size() ? codes.data() : nullptr,
                                             nodes_.data(), aabbs_.data(), cones_.data(), nullptr, nullptr, &collision, nullptr));
        collision_ = collision != 0;
        if (query_host_enabled_)
        {
            nodes_h_.resize(2 * n - 1), aabbs_h_.resize(2 * n - 1), cones_h_.resize(2 * n - 1);
            nodes_.download(&nodes_h_[0]);
            aabbs_.download(&aabbs_h_[0]);
            cones_.download(&cones_h_[0]);
        }
    }

    thrust::host_vector<object_type> const &objects_host() const noexcept { return objects_h_; }
    thrust::host_vector<object_type> &objects_host() noexcept { return objects_h_; }
    thrust::host_vector<node_type> const &nodes_host() const noexcept { return nodes_h_; }
    thrust::host_vector<node_type> &nodes_host() noexcept { return nodes_h_; }
    thrust::host_vector<aabb_type> const &aabbs_host() const noexcept { return aabbs_h_; }
    thrust::host_vector<aabb_type> &aabbs_host() noexcept { return aabbs_h_; }
    thrust::host_vector<cone_type> const &cones_host() const noexcept { return cones_h_; } // typed as cones (the reference says aabb, Q8)
    thrust::host_vector<cone_type> &cones_host() noexcept { return cones_h_; }

private:
    thrust::host_vector<object_type> objects_h_;
    detail::device_buffer<object_type> objects_d_;
    thrust::host_vector<aabb_type> aabbs_h_;
    detail::device_buffer<aabb_type> aabbs_;
    thrust::host_vector<cone_type> cones_h_;
    detail::device_buffer<cone_type> cones_;
    thrust::host_vector<node_type> nodes_h_;
    detail::device_buffer<node_type> nodes_;
    bool query_host_enabled_ = false;
    bool collision_ = false;
};
} // namespace lbvh
#endif // SNCH_LBVH_B200_BVH_CUH
